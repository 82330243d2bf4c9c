"""Synthetic scenario inputs (roadnet / flow / config JSON) in the reference's file formats.

The engine consumes the *same* JSON the reference does (roadnet: roadnet.cpp:42-325,
flow: engine.cpp:106-164, config: engine.cpp:37-84).  BASELINE.json's configs are grids made by
the reference's ``tools/generator``; that tool is not available on the GPU box, so this module
re-creates the same grid scenario from first principles (a regular lattice of 4-arm
signalised intersections with one left / straight / right lane per approach and Hermite
laneLink curves).  ``tests/test_cpu.py::test_grid_generator_matches_reference_tool`` checks that the
JSON produced here is identical, value for value, to what ``tools/generator/generate_grid_scenario.py``
writes (when the reference tree is present).

It also implements the dense "random walk" flow recipe from SURVEY.md Appendix A that reaches
the ~1e5-vehicle operating point on the 30x30 grid.
"""
from __future__ import annotations

import json
import math
import os
import random
from typing import Dict, List, Optional, Sequence

# direction index -> lattice step (east, north, west, south), as in the reference tool
_DX = (1, 0, -1, 0)
_DY = (0, 1, 0, -1)

DEFAULT_VEHICLE = {
    "length": 5.0,
    "width": 2.0,
    "maxPosAcc": 2.0,
    "maxNegAcc": 4.5,
    "usualPosAcc": 2.0,
    "usualNegAcc": 4.5,
    "minGap": 2.5,
    "maxSpeed": 16.67,
    "headwayTime": 1.5,
}


def _unit(road):
    a, b = road["points"][0], road["points"][-1]
    ex, ey = b["x"] - a["x"], b["y"] - a["y"]
    n = math.sqrt(ex * ex + ey * ey)
    return ex / n, ey / n


def _lane_shift(road, lane_index):
    s = 0.0
    for i in range(lane_index):
        s += road["lanes"][i]["width"]
    return s + road["lanes"][lane_index]["width"] * 0.5


def _exit_point(road, width, lane_index):
    ux, uy = _unit(road)
    sh = _lane_shift(road, lane_index)
    p = road["points"][-1]
    x, y = p["x"] - ux * width, p["y"] - uy * width
    return x + uy * sh, y - ux * sh


def _entry_point(road, width, lane_index):
    ux, uy = _unit(road)
    sh = _lane_shift(road, lane_index)
    p = road["points"][0]
    x, y = p["x"] + ux * width, p["y"] + uy * width
    return x + uy * sh, y - ux * sh


def _hermite(road_a, lane_a, road_b, lane_b, width, mid):
    """Cubic Hermite curve from the exit of (road_a, lane_a) to the entry of (road_b, lane_b)."""
    ax, ay = _unit(road_a)
    bx, by = _unit(road_b)
    pax, pay = _exit_point(road_a, width, lane_a)
    pbx, pby = _entry_point(road_b, width, lane_b)
    ax, ay, bx, by = ax * width, ay * width, bx * width, by * width
    pts = []
    for i in range(mid + 1):
        t = i / mid
        t3 = t * t * t
        t2 = t * t
        h00 = 2 * t3 - 3 * t2 + 1
        h10 = t3 - 2 * t2 + t
        h01 = -2 * t3 + 3 * t2
        h11 = t3 - t2
        x = h00 * pax + h10 * ax + h01 * pbx + h11 * bx
        y = h00 * pay + h10 * ay + h01 * pby + h11 * by
        pts.append({"x": x, "y": y})
    return pts


def _turn_type(da, db):
    if (da + 1) % 4 == db:
        return "turn_left"
    if (db + 1) % 4 == da:
        return "turn_right"
    if da == db:
        return "go_straight"
    return None


def grid_roadnet(rows: int, cols: int, row_distance=300, col_distance=300, intersection_width=30,
                 n_left=1, n_straight=1, n_right=1, lane_max_speed=16.67, lane_width=4,
                 tl_plan=True, mid_points=10) -> dict:
    """``rows x cols`` signalised lattice surrounded by a ring of virtual (source/sink) nodes."""
    R, C = rows + 2, cols + 2
    n_lanes = n_left + n_straight + n_right

    def inside(i, j):
        return 0 <= i < R and 0 <= j < C

    def interior(i, j):
        return 0 < i < R - 1 and 0 < j < C - 1

    widths = [[intersection_width if interior(i, j) else 0 for j in range(C)] for i in range(R)]
    # lattice coordinates (the reference accumulates along rows first, then columns)
    xs = [[None] * C for _ in range(R)]
    ys = [[None] * C for _ in range(R)]
    for i in range(R):
        for j in range(C):
            if j > 0:
                xs[i][j] = xs[i][j - 1] + row_distance
                ys[i][j] = ys[i][j - 1]
            elif i > 0:
                xs[i][j] = xs[i - 1][j]
                ys[i][j] = ys[i - 1][j] + col_distance
            else:
                xs[i][j] = -row_distance
                ys[i][j] = -col_distance

    def make_road(i, j, k):
        ni, nj = i + _DY[k], j + _DX[k]
        if not inside(ni, nj):
            return None
        return {
            "id": "road_%d_%d_%d" % (j, i, k),
            "_dir": k,
            "_keep": interior(i, j) or interior(ni, nj),
            "points": [{"x": xs[i][j], "y": ys[i][j]}, {"x": xs[ni][nj], "y": ys[ni][nj]}],
            "lanes": [{"width": lane_width, "maxSpeed": lane_max_speed}] * n_lanes,
            "startIntersection": "intersection_%d_%d" % (j, i),
            "endIntersection": "intersection_%d_%d" % (nj, ni),
        }

    roads = [[[make_road(i, j, k) for k in range(4)] for j in range(C)] for i in range(R)]

    def lane_kind_ok(kind, c):
        if kind == "turn_left":
            return 0 <= c < n_left
        if kind == "go_straight":
            return n_left <= c < n_left + n_straight
        return n_left + n_straight <= c < n_lanes

    intersections = []
    for i in range(R):
        for j in range(C):
            corner = (i in (0, R - 1)) and (j in (0, C - 1))
            width = widths[i][j]
            out_roads = [r for r in roads[i][j] if r is not None and r["_keep"]]
            in_roads = [roads[i - _DY[k]][j - _DX[k]][k] for k in range(4) if inside(i - _DY[k], j - _DX[k])]
            in_roads = [r for r in in_roads if r["_keep"]]
            road_links = []
            for ra in in_roads:
                for rb in out_roads:
                    kind = _turn_type(ra["_dir"], rb["_dir"])
                    if kind is None:
                        continue
                    lane_links = []
                    for c in range(n_lanes):
                        if not lane_kind_ok(kind, c):
                            continue
                        for d in range(n_lanes):
                            lane_links.append({"startLaneIndex": c, "endLaneIndex": d,
                                               "points": _hermite(ra, c, rb, d, width, mid_points)})
                    if lane_links:
                        road_links.append({"type": kind, "startRoad": ra["id"], "endRoad": rb["id"],
                                           "direction": ra["_dir"], "laneLinks": lane_links})
            idx = range(len(road_links))
            left = {x for x in idx if road_links[x]["type"] == "turn_left"}
            right = {x for x in idx if road_links[x]["type"] == "turn_right"}
            straight = {x for x in idx if road_links[x]["type"] == "go_straight"}
            by_dir = [{x for x in idx if road_links[x]["direction"] == k} for k in range(4)]
            we, ns, ew, sn = by_dir
            if tl_plan:
                phases = [(30, ((ew | we) & straight) | right), (5, right)]
                if n_left:
                    phases += [(30, ((ew | we) & left) | right), (5, right)]
                phases += [(30, ((ns | sn) & straight) | right), (5, right)]
                if n_left:
                    phases += [(30, ((sn | ns) & left) | right), (5, right)]
            else:
                phases = [(5, right),
                          (30, ((ew | we) & straight) | right), (30, ((ns | sn) & straight) | right),
                          (30, ((ew | we) & left) | right), (30, ((sn | ns) & left) | right),
                          (30, we | right), (30, ew | right), (30, ns | right), (30, sn | right)]
            if corner:
                continue
            intersections.append({
                "id": "intersection_%d_%d" % (j, i),
                "point": {"x": xs[i][j], "y": ys[i][j]},
                "width": width,
                "roads": [r["id"] for r in in_roads + out_roads],
                "roadLinks": road_links,
                "trafficLight": {
                    "roadLinkIndices": list(idx),
                    "lightphases": [{"time": t, "availableRoadLinks": sorted(s)} for t, s in phases],
                },
                "virtual": not interior(i, j),
            })

    final_roads = []
    for i in range(R):
        for j in range(C):
            for k in range(4):
                r = roads[i][j][k]
                if r is not None and r["_keep"]:
                    final_roads.append({k2: v for k2, v in r.items() if not k2.startswith("_")})
    return {"intersections": intersections, "roads": final_roads}


def grid_straight_flows(rows: int, cols: int, interval=2.0, vehicle: Optional[dict] = None) -> list:
    """The reference generator's default demand: one straight flow per border road, both ways."""
    vehicle = dict(vehicle or DEFAULT_VEHICLE)

    def line(x, y, k, n):
        out = []
        for _ in range(n):
            out.append("road_%d_%d_%d" % (x, y, k))
            x += _DX[k]
            y += _DY[k]
        return out

    routes = []
    for i in range(1, rows + 1):
        routes.append(line(0, i, 0, cols + 1))
        routes.append(line(cols + 1, i, 2, cols + 1))
    for i in range(1, cols + 1):
        routes.append(line(i, 0, 1, rows + 1))
        routes.append(line(i, rows + 1, 3, rows + 1))
    return [{"vehicle": vehicle, "route": r, "interval": interval, "startTime": 0, "endTime": -1}
            for r in routes]


def random_walk_flows(roadnet: dict, frac=0.5, interval=10.0, max_len=12, seed=1,
                      vehicle: Optional[dict] = None, fleet_spread=0.0) -> list:
    """Dense demand: for every road with a successor, with probability
    ``frac`` one flow following a seeded random walk over roadLink successors (<= max_len roads).

    ``fleet_spread`` > 0 gives every flow its own vehicle parameters, drawn (from a second seeded stream, so
    the routes do not change) within +-fleet_spread of the template: a heterogeneous fleet.  With identical
    vehicles two queue heads that start from rest in the same step enter a lane with bit-equal distances now
    and then, and the reference orders such a pair by thread timing (engine.cpp:247-249, :480) -- its own result is
    then not defined; distinct parameters make that coincidence vanish."""
    vehicle = dict(vehicle or DEFAULT_VEHICLE)
    fleet = random.Random(seed * 7919 + 17)
    nxt: Dict[str, set] = {}
    for inter in roadnet["intersections"]:
        for rl in inter.get("roadLinks", []):
            nxt.setdefault(rl["startRoad"], set()).add(rl["endRoad"])
    rng = random.Random(seed)
    flows = []
    for road in roadnet["roads"]:
        rid = road["id"]
        if rid not in nxt:
            continue
        if rng.random() > frac:
            continue
        route = [rid]
        while len(route) < max_len and route[-1] in nxt:
            route.append(rng.choice(sorted(nxt[route[-1]])))
        if len(route) >= 2:
            veh = vehicle
            if fleet_spread > 0:
                veh = dict(vehicle)
                for k in ("length", "maxPosAcc", "maxNegAcc", "usualPosAcc", "usualNegAcc", "minGap", "maxSpeed", "headwayTime"):
                    veh[k] = vehicle[k] * (1.0 + fleet_spread * (2.0 * fleet.random() - 1.0))
            flows.append({"vehicle": veh, "route": route, "interval": float(interval),
                          "startTime": 0, "endTime": -1})
    return flows


def write_scenario(directory: str, roadnet: dict, flows: list, *, interval=1.0, seed=0,
                   rl_traffic_light=False, lane_change=False, save_replay=False, name="") -> str:
    """Write roadnet / flow / config JSON into ``directory``; returns the config path."""
    os.makedirs(directory, exist_ok=True)
    sfx = ("_" + name) if name else ""
    rn, fl, cf = "roadnet%s.json" % sfx, "flow%s.json" % sfx, "config%s.json" % sfx
    with open(os.path.join(directory, rn), "w") as f:
        json.dump(roadnet, f)
    with open(os.path.join(directory, fl), "w") as f:
        json.dump(flows, f)
    d = os.path.abspath(directory)
    if not d.endswith("/"):
        d += "/"  # the engine concatenates dir + file (engine.cpp:60,65)
    cfg = {"interval": interval, "seed": seed, "dir": d, "roadnetFile": rn, "flowFile": fl,
           "rlTrafficLight": rl_traffic_light, "laneChange": lane_change, "saveReplay": save_replay,
           "roadnetLogFile": "replay_roadnet%s.json" % sfx, "replayLogFile": "replay%s.txt" % sfx}
    path = os.path.join(directory, cf)
    with open(path, "w") as f:
        json.dump(cfg, f)
    return path


def make_grid_scenario(directory: str, rows: int, cols: int, *, dense: Optional[dict] = None,
                       flow_interval=2.0, name="", **cfg) -> str:
    """Convenience: grid roadnet + (default straight | dense random-walk) flows + config."""
    net = grid_roadnet(rows, cols)
    if dense is None:
        flows = grid_straight_flows(rows, cols, interval=flow_interval)
    else:
        flows = random_walk_flows(net, **dense)
    return write_scenario(directory, net, flows, name=name, **cfg)

"""Random road networks in the reference's JSON format, for fuzzing the loader and the oracle
against the compiled reference: jittered lattice with missing edges, 1-3 lanes per direction
(different per direction), bent roads, unequal lane widths and speed limits, arbitrary turn
geometry, laneLinks with and without explicit points, intersections with random phase plans
(including single-phase = "implicit" ones), flows with start/end windows, intervals down to one second (flow.h:34 asserts >= 1),
assorted vehicles and sparse anchor lists that need the router's Dijkstra."""
import math
import random


def random_roadnet(seed: int, rows: int = 3, cols: int = 4) -> dict:
    rng = random.Random(seed)
    R, C = rows + 2, cols + 2

    def interior(i, j):
        return 0 < i < R - 1 and 0 < j < C - 1

    pos = {}
    for i in range(R):
        for j in range(C):
            if (i in (0, R - 1)) and (j in (0, C - 1)):
                continue
            pos[(i, j)] = (j * 260.0 + rng.uniform(-45, 45), i * 240.0 + rng.uniform(-45, 45))
    edges = []
    for (i, j) in sorted(pos):
        for (di, dj) in ((0, 1), (1, 0)):
            a, b = (i, j), (i + di, j + dj)
            if b not in pos or not (interior(*a) or interior(*b)):
                continue
            if interior(*a) and interior(*b) and rng.random() < 0.15:
                continue                                   # a missing street
            edges.append((a, b))
    roads, out_of, in_to = [], {}, {}
    for a, b in edges:
        for s, e in ((a, b), (b, a)):
            n = rng.choice([1, 2, 2, 3])
            pts = [{"x": pos[s][0], "y": pos[s][1]}]
            if rng.random() < 0.4:
                mx, my = (pos[s][0] + pos[e][0]) / 2, (pos[s][1] + pos[e][1]) / 2
                dx, dy = pos[e][0] - pos[s][0], pos[e][1] - pos[s][1]
                k = rng.uniform(-0.1, 0.1)
                pts.append({"x": mx - dy * k, "y": my + dx * k})
            pts.append({"x": pos[e][0], "y": pos[e][1]})
            rid = "r_%d_%d__%d_%d" % (s + e)
            road = {"id": rid, "points": pts,
                    "lanes": [{"width": rng.choice([3.0, 3.5, 4.0]), "maxSpeed": rng.choice([8.33, 11.11, 13.89, 16.67])} for _ in range(n)],
                    "startIntersection": "n_%d_%d" % s, "endIntersection": "n_%d_%d" % e}
            roads.append(road)
            out_of.setdefault(s, []).append(road)
            in_to.setdefault(e, []).append(road)

    def heading(road, at_end):
        p = road["points"]
        a, b = (p[-2], p[-1]) if at_end else (p[0], p[1])
        return math.atan2(b["y"] - a["y"], b["x"] - a["x"])

    intersections = []
    for node in sorted(pos):
        ins, outs = in_to.get(node, []), out_of.get(node, [])
        virtual = not interior(*node)
        inter = {"id": "n_%d_%d" % node, "point": {"x": pos[node][0], "y": pos[node][1]},
                 "width": 0 if virtual else rng.choice([12.0, 15.0, 20.0, 25.0]),
                 "roads": [r["id"] for r in ins + outs], "roadLinks": [],
                 "trafficLight": {"roadLinkIndices": [], "lightphases": []}, "virtual": virtual}
        if not virtual:
            links = []
            for ra in ins:
                for rb in outs:
                    if rb["endIntersection"] == ra["startIntersection"]:
                        continue                            # no U-turns
                    turn = (heading(rb, False) - heading(ra, True) + math.pi) % (2 * math.pi) - math.pi
                    kind = "go_straight" if abs(turn) < 0.6 else ("turn_left" if turn > 0 else "turn_right")
                    na, nb = len(ra["lanes"]), len(rb["lanes"])
                    pairs = [(c, d) for c in range(na) for d in range(nb) if rng.random() < 0.7]
                    if not pairs:
                        pairs = [(rng.randrange(na), rng.randrange(nb))]
                    lane_links = []
                    for c, d in pairs:
                        ll = {"startLaneIndex": c, "endLaneIndex": d}
                        r = rng.random()
                        if r < 0.15:
                            ll["points"] = []               # empty list: the loader's default curve
                        elif r < 0.3:                       # explicit polyline through a random interior point
                            pa, pb = ra["points"][-1], rb["points"][0]
                            ll["points"] = [{"x": pa["x"] + rng.uniform(-6, 6), "y": pa["y"] + rng.uniform(-6, 6)},
                                            {"x": pos[node][0] + rng.uniform(-3, 3), "y": pos[node][1] + rng.uniform(-3, 3)},
                                            {"x": pb["x"] + rng.uniform(-6, 6), "y": pb["y"] + rng.uniform(-6, 6)}]
                        lane_links.append(ll)
                    links.append({"type": kind, "startRoad": ra["id"], "endRoad": rb["id"], "direction": 0, "laneLinks": lane_links})
            inter["roadLinks"] = links
            n = len(links)
            n_ph = rng.choice([1, 2, 3, 4, 5])
            phases = []
            for k in range(n_ph):
                avail = sorted(x for x in range(n) if n_ph == 1 or x % n_ph == k or rng.random() < 0.35)
                phases.append({"time": rng.choice([5, 8, 12.5, 20, 30]), "availableRoadLinks": avail})
            inter["trafficLight"] = {"roadLinkIndices": list(range(n)), "lightphases": phases}
        intersections.append(inter)
    return {"intersections": intersections, "roads": roads}


def random_flows(net: dict, seed: int, n_flows: int = 60) -> list:
    rng = random.Random(seed)
    nxt = {}
    for inter in net["intersections"]:
        for rl in inter["roadLinks"]:
            nxt.setdefault(rl["startRoad"], set()).add(rl["endRoad"])
    starts = sorted(nxt)
    flows = []
    while len(flows) < n_flows:
        route = [rng.choice(starts)]
        while len(route) < rng.randint(2, 9) and route[-1] in nxt:
            route.append(rng.choice(sorted(nxt[route[-1]])))
        if len(route) < 2:
            continue
        r = rng.random()
        if r < 0.3:
            route = [route[0], route[-1]]                   # sparse anchors: the engine's Dijkstra fills the gap
        elif r < 0.4 and len(route) > 3:
            route = [route[0], route[len(route) // 2], route[-1]]
        elif r < 0.45:
            route = [route[0], route[0], route[-1]]         # repeated anchor (router.cpp:234)
        veh = {"length": rng.choice([4.0, 5.0, 7.5, 12.0]), "width": 2.0,
               "maxPosAcc": rng.choice([1.5, 2.0, 3.0]), "maxNegAcc": rng.choice([3.5, 4.5, 6.0]),
               "usualPosAcc": rng.choice([1.5, 2.0]), "usualNegAcc": rng.choice([2.5, 3.5, 4.5]),
               "minGap": rng.choice([1.5, 2.5]), "maxSpeed": rng.choice([8.0, 11.11, 16.67]),
               "headwayTime": rng.choice([1.0, 1.5, 2.0])}
        start = rng.choice([0, 0, 0, 17, 60])
        flows.append({"vehicle": veh, "route": route, "interval": rng.choice([1.0, 1.0, 2.0, 3.5, 6.0]),
                      "startTime": start, "endTime": rng.choice([-1, -1, start + 150, start + 400])})
    return flows

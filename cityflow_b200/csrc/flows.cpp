#include "flows.h"

#include <climits>
#include <cstdlib>
#include <iostream>
#include <queue>
#include <tuple>

#include "json_min.h"

namespace cfb {

bool VehicleTemplate::operator<(const VehicleTemplate &o) const {
    auto t = [](const VehicleTemplate &v) {
        return std::make_tuple(v.speed, v.len, v.width, v.maxPosAcc, v.maxNegAcc, v.usualPosAcc, v.usualNegAcc,
                               v.minGap, v.maxSpeed, v.headwayTime, v.yieldDistance, v.turnSpeed);
    };
    return t(*this) < t(o);
}

// ------------------------------------------------------------------------------------------
// Road-level shortest path, LENGTH metric (router.cpp:160-226).  Same container algorithms
// (std::priority_queue ordered on distance only) and the same relaxation order as the
// reference, so equal-cost alternatives resolve identically.
bool Routing::dijkstra(int start, int end, std::vector<int> &buffer) const {
    const int n = net_.nRoads();
    std::vector<double> dis(n, 0.0);
    std::vector<char> hasDis(n, 0), visited(n, 0);
    std::vector<int> from(n, -1);
    using Item = std::pair<int, double>;
    auto cmp = [](const Item &a, const Item &b) { return a.second > b.second; };
    std::priority_queue<Item, std::vector<Item>, decltype(cmp)> queue(cmp);
    bool success = false;
    dis[start] = 0;
    hasDis[start] = 1;
    queue.push(std::make_pair(start, 0));
    while (!queue.empty()) {
        int cur = queue.top().first;
        if (cur == end) {
            success = true;
            break;
        }
        queue.pop();
        if (visited[cur]) continue;
        visited[cur] = 1;
        double curDis = dis[cur];
        for (int adj : net_.interRoads[net_.roadEndInter[cur]]) {
            if (!net_.roadConnected(cur, adj)) continue;
            double newDis = curDis + net_.roadAverageLength(adj);
            if (!hasDis[adj] || newDis < dis[adj]) {
                from[adj] = cur;
                dis[adj] = newDis;
                hasDis[adj] = 1;
                queue.emplace(std::make_pair(adj, newDis));
            }
        }
    }
    std::vector<int> path;
    path.push_back(end);
    int it = from[end];
    while (it != -1 && it != start) {
        path.push_back(it);
        it = from[it];
    }
    buffer.insert(buffer.end(), path.rbegin(), path.rend());
    return success;
}

bool Routing::resolve(const std::vector<int> &anchors, std::vector<int> &roads) const {
    roads.clear();
    if (anchors.empty()) return false;
    roads.push_back(anchors[0]);
    for (size_t i = 1; i < anchors.size(); ++i) {
        if (anchors[i - 1] == anchors[i]) continue;            // router.cpp:234
        if (!dijkstra(anchors[i - 1], anchors[i], roads)) return false;
    }
    return roads.size() > 1;                                    // router.cpp:239
}

// Router::getNextDrivable(lane) for a lane on roads[r] (router.cpp:49-76) with
// selectLaneIndex's "closest lane index, first wins" rule (router.cpp:96-111).
int Routing::chooseLink(int lane, const std::vector<int> &roads, int r) const {
    const int last = (int) roads.size() - 1;
    if (r >= last) return PLAN_END;
    int best = -1, bestDiff = INT_MAX;
    for (int ll : net_.laneOutLinks[lane]) {
        const int endLane = net_.llEndLane[ll];
        if (net_.laneRoad[endLane] != roads[r + 1]) continue;
        if (r + 2 <= last) {  // the end lane must be able to continue to the road after next
            bool cont = false;
            for (int l2 : net_.laneOutLinks[endLane])
                if (net_.laneRoad[net_.llEndLane[l2]] == roads[r + 2]) { cont = true; break; }
            if (!cont) continue;
        }
        int diff = std::abs(net_.laneIdx[endLane] - net_.laneIdx[lane]);
        if (diff < bestDiff) { bestDiff = diff; best = ll; }
    }
    return best < 0 ? PLAN_DEAD : best;
}

int Routing::buildPlan(const std::vector<int> &roads, int startLane, int roadPos, int routeId, const std::vector<int> &prefix) {
    const int id = (int) planBeg_.size();
    planBeg_.push_back((int) planData_.size());
    planRoute_.push_back(routeId);
    planRoadPos_.push_back(roadPos);
    planData_.insert(planData_.end(), prefix.begin(), prefix.end());   // planFrom(): drivables before the sequential part
    int lane = startLane;
    for (int r = roadPos;; ++r) {
        planData_.push_back(lane);
        int ll = chooseLink(lane, roads, r);
        // Router::isLastRoad compares the ROAD with route.back() (router.cpp:131-134), not the position: on a
        // route that returns to its first road a lane that cannot continue counts as "on the last road" --
        // the vehicle is not stopped (router.h:66-68) and leaves the network at the end of that lane
        // (vehicle.cpp:58-61).  Found by the fuzz tests on the emulated device step.
        if (ll == PLAN_DEAD && roads[r] == roads.back()) ll = PLAN_END;
        if (ll < 0) {
            planData_.push_back(ll);  // PLAN_END or PLAN_DEAD
            break;
        }
        planData_.push_back(net_.nLanes() + ll);
        lane = net_.llEndLane[ll];
    }
    return id;
}

int Routing::planFrom(int routeId, int drivable) {
    const auto key = std::make_pair(routeId, drivable);
    auto it = planFrom_.find(key);
    if (it != planFrom_.end()) return it->second;
    const Route &rt = routes_[routeId];
    if (!rt.valid) return -1;
    const std::vector<int> roads = rt.roads;   // (buildPlan grows the tables: no references into them)
    auto firstVisit = [&roads](int road) {
        for (size_t r = 0; r < roads.size(); ++r)
            if (roads[r] == road) return (int) r;
        return -1;
    };
    const int nL = net_.nLanes();
    int plan = -1;
    if (drivable >= nL) {   // on a laneLink: its end lane, looked up from the start of the route
        const int lane = net_.llEndLane[drivable - nL];
        const int pos = firstVisit(net_.laneRoad[lane]);
        if (pos < 0) return -1;
        plan = buildPlan(roads, lane, pos, routeId, {drivable});
    } else {
        const int pos = firstVisit(net_.laneRoad[drivable]);
        if (pos < 0) return -1;
        const int ll = chooseLink(drivable, roads, pos);
        const int pos2 = ll >= 0 ? firstVisit(roads[pos + 1]) : -1;
        if (ll >= 0 && pos2 != pos + 1) {
            // the next road was visited before: the router takes that visit for the current one once the vehicle is there
            plan = buildPlan(roads, net_.llEndLane[ll], pos2, routeId, {drivable, nL + ll});
        } else {
            if (pos == 0)
                for (size_t k = 0; k < rt.startLanes.size(); ++k)
                    if (rt.startLanes[k] == drivable) plan = rt.planOfStartLane[k];
            if (plan < 0) plan = buildPlan(roads, drivable, pos, routeId);
        }
    }
    planFrom_[key] = plan;
    return plan;
}

int Routing::intern(const std::vector<int> &anchors) {
    auto it = byAnchors_.find(anchors);
    if (it != byAnchors_.end()) return it->second;
    Route rt;
    rt.valid = resolve(anchors, rt.roads);
    if (rt.valid) {
        // Router::getFirstDrivable: lanes of the first road that link to the second road
        const int r0 = rt.roads[0];
        for (int l = net_.roadLaneBeg[r0]; l < net_.roadLaneBeg[r0 + 1]; ++l) {
            bool ok = false;
            for (int ll : net_.laneOutLinks[l])
                if (net_.laneRoad[net_.llEndLane[ll]] == rt.roads[1]) { ok = true; break; }
            if (ok) rt.startLanes.push_back(l);
        }
    }
    const int id = (int) routes_.size();
    if (rt.valid)
        for (int l : rt.startLanes) rt.planOfStartLane.push_back(buildPlan(rt.roads, l, 0, id));
    lanePlanRoad_.push_back((int) lanePlanBeg_.size());
    if (rt.valid && lanePlans_) {
        for (int r = 0; r < (int) rt.roads.size(); ++r) {
            lanePlanBeg_.push_back((int) lanePlanId_.size());
            const int road = rt.roads[r];
            for (int l = net_.roadLaneBeg[road]; l < net_.roadLaneBeg[road + 1]; ++l) {
                int plan = -1;
                if (r == 0)
                    for (size_t k = 0; k < rt.startLanes.size(); ++k)
                        if (rt.startLanes[k] == l) plan = rt.planOfStartLane[k];
                if (plan < 0) plan = buildPlan(rt.roads, l, r, id);
                lanePlanId_.push_back(plan);
            }
        }
    }
    routes_.push_back(std::move(rt));
    byAnchors_[anchors] = id;
    anchorsOf_.push_back(anchors);
    return id;
}

// ------------------------------------------------------------------------------------------
namespace {
struct FlowFormatError : std::runtime_error {
    explicit FlowFormatError(const std::string &m) : std::runtime_error(m) {}
};
double reqDouble(const Json &o, const char *name) {
    const Json *v = o.find(name);
    if (!v) throw FlowFormatError(std::string(name) + " is required but missing in json file");
    if (!v->isNumber()) throw FlowFormatError(std::string(name) + ": expected type d");
    return v->asDouble();
}
int optInt(const Json &o, const char *name, int dflt) {
    const Json *v = o.find(name);
    if (!v || !v->isInt()) return dflt;  // utility.h:129-136
    return v->asInt();
}
}  // namespace

bool loadFlows(const std::string &path, const RoadNet &net, std::vector<FlowDef> &out) {
    bool opened = false;
    Json root = Json::parseFile(path, &opened);
    if (!opened) {
        std::cerr << "cannot open flow file!" << std::endl;
        return false;
    }
    std::string where;
    try {
        if (!root.isArray()) throw FlowFormatError("flow file: expected type array");
        for (size_t i = 0; i < root.arr.size(); ++i) {
            where = "flow[" + std::to_string(i) + "]";
            const Json &fj = root.arr[i];
            FlowDef f;
            const Json *rj = fj.find("route");
            if (!rj) throw FlowFormatError("route is required but missing in json file");
            if (!rj->isArray()) throw FlowFormatError("route: expected type array");
            for (const Json &r : rj->arr) {
                if (!r.isString()) throw FlowFormatError("route: expected type string");
                auto it = net.roadIndex.find(r.s);
                if (it == net.roadIndex.end()) throw FlowFormatError("No such road: " + r.s);
                f.anchors.push_back(it->second);
            }
            const Json *vj = fj.find("vehicle");
            if (!vj) throw FlowFormatError("vehicle is required but missing in json file");
            if (!vj->isObject()) throw FlowFormatError("vehicle: expected type object");
            f.tmpl.len = reqDouble(*vj, "length");
            f.tmpl.width = reqDouble(*vj, "width");
            f.tmpl.maxPosAcc = reqDouble(*vj, "maxPosAcc");
            f.tmpl.maxNegAcc = reqDouble(*vj, "maxNegAcc");
            f.tmpl.usualPosAcc = reqDouble(*vj, "usualPosAcc");
            f.tmpl.usualNegAcc = reqDouble(*vj, "usualNegAcc");
            f.tmpl.minGap = reqDouble(*vj, "minGap");
            f.tmpl.maxSpeed = reqDouble(*vj, "maxSpeed");
            f.tmpl.headwayTime = reqDouble(*vj, "headwayTime");
            f.startTime = optInt(fj, "startTime", 0);
            f.endTime = optInt(fj, "endTime", -1);
            f.interval = reqDouble(fj, "interval");
            f.id = "flow_" + std::to_string(i);
            if (f.anchors.empty()) throw FlowFormatError("route must not be empty");
            // the reference aborts on this (live assert, flow.h:34)
            if (!(f.interval >= 1 || f.startTime == f.endTime))
                throw FlowFormatError("flow interval must be >= 1 (or startTime == endTime)");
            out.push_back(std::move(f));
        }
    } catch (const FlowFormatError &e) {
        std::cerr << "Error occurred when reading flow file" << std::endl;
        std::cerr << "/" << where << " " << e.what() << std::endl;
        return false;
    }
    return true;
}

}  // namespace cfb

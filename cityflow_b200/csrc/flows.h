// Host-side demand description: flow file parsing (engine.cpp:106-164), route resolution
// (Router::updateShortestPath / dijkstra, router.cpp:160-243) and the per-(route, start lane)
// drivable sequences ("plans") that replace the reference's lazy per-vehicle lookahead deque
// (Router::getNextDrivable, router.cpp:39-76) on the device.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "roadnet.h"

namespace cfb {

// VehicleInfo defaults, vehicle.h:31-45
struct VehicleTemplate {
    double speed = 0;
    double len = 5;
    double width = 2;
    double maxPosAcc = 4.5;
    double maxNegAcc = 4.5;
    double usualPosAcc = 2.5;
    double usualNegAcc = 2.5;
    double minGap = 2;
    double maxSpeed = 16.66667;
    double headwayTime = 1;
    double yieldDistance = 5;
    double turnSpeed = 8.3333;
    bool operator<(const VehicleTemplate &o) const;
};

struct FlowDef {
    VehicleTemplate tmpl;
    std::vector<int> anchors;   // road ids as listed in "route"
    double interval = 0;
    int startTime = 0, endTime = -1;
    std::string id;             // "flow_<i>"
};

// A resolved road-level route (after Dijkstra between anchors) plus, for every admissible
// start lane, the full drivable sequence lane,link,lane,...  Drivable ids: lane l -> l,
// laneLink k -> nLanes + k.  PLAN_END terminates a sequence; PLAN_DEAD marks "lane cannot reach
// the next road" (the reference asserts there, vehicle.cpp:60).
constexpr int PLAN_END = -1;
constexpr int PLAN_DEAD = -2;

struct Route {
    bool valid = false;
    std::vector<int> roads;
    std::vector<int> startLanes;          // candidate first lanes (Router::getFirstDrivable, router.cpp:23-37)
    std::vector<int> planOfStartLane;     // plan id per candidate
};

class Routing {
public:
    explicit Routing(const RoadNet &net) : net_(net) {}
    // Shortest road path between consecutive anchors (LENGTH metric); false = invalid route.
    bool resolve(const std::vector<int> &anchors, std::vector<int> &roads) const;
    // Interns the route; builds plans for every candidate start lane. Returns route id.
    int intern(const std::vector<int> &anchors);
    const Route &route(int id) const { return routes_[id]; }
    const std::vector<int> &anchorsOf(int id) const { return anchorsOf_[id]; }   // what intern() was called with (archives)
    int numRoutes() const { return (int) routes_.size(); }
    // plan storage (flat, PLAN_END terminated)
    const std::vector<int> &planData() const { return planData_; }
    const std::vector<int> &planBeg() const { return planBeg_; }
    int numPlans() const { return (int) planBeg_.size(); }
    // next laneLink for a vehicle on `lane` whose route continues with roads[r+1] (, roads[r+2])
    int chooseLink(int lane, const std::vector<int> &roads, int r) const;
    // Lane change: a shadow vehicle continues its
    // parent's route from the lane it was inserted into, so every lane of every road of a route gets
    // a plan.  Must be switched on before the first intern().
    void enableLanePlans() { lanePlans_ = true; }
    bool lanePlansEnabled() const { return lanePlans_; }
    // plan that starts on lane `laneIdx` of the route's road number `roadPos`
    int lanePlan(int routeId, int roadPos, int laneIdx) const {
        return lanePlanId_[lanePlanBeg_[lanePlanRoad_[routeId] + roadPos] + laneIdx];
    }
    // Plan of a vehicle picked up mid-route on `drivable` (a reference-schema archive, archive.cpp:378-385 + :433): the
    // sequence the reference's router yields from there.  A loaded Router starts with iCurRoad = route.begin()
    // (router.cpp:16-21), so on a route that visits a road twice it takes the FIRST visit for the current one: the
    // current lane's road is looked up from the start of the route (router.cpp:49-57), and so is the road of the first
    // lane the vehicle moves on to (Router::update advances iCurRoad from where it stood, router.cpp:78-94) -- only
    // from there on the pointer follows the route.  A vehicle on a laneLink continues on that link's end lane
    // (router.cpp:44-45).  -1: the drivable's road is not on the route.
    int planFrom(int routeId, int drivable);
    int planRoute(int plan) const { return planRoute_[plan]; }      // route a plan belongs to
    int planRoadPos(int plan) const { return planRoadPos_[plan]; }  // position in that route of the plan's first road
    const std::vector<int> &lanePlanRoadTable() const { return lanePlanRoad_; }
    const std::vector<int> &lanePlanBegTable() const { return lanePlanBeg_; }
    const std::vector<int> &lanePlanIdTable() const { return lanePlanId_; }
    const std::vector<int> &planRouteTable() const { return planRoute_; }
    const std::vector<int> &planRoadPosTable() const { return planRoadPos_; }

private:
    bool dijkstra(int start, int end, std::vector<int> &buffer) const;
    int buildPlan(const std::vector<int> &roads, int startLane, int roadPos = 0, int routeId = -1,
                  const std::vector<int> &prefix = std::vector<int>());
    std::map<std::pair<int, int>, int> planFrom_;
    const RoadNet &net_;
    std::vector<Route> routes_;
    std::map<std::vector<int>, int> byAnchors_;
    std::vector<std::vector<int>> anchorsOf_;
    std::vector<int> planData_, planBeg_;
    bool lanePlans_ = false;
    std::vector<int> planRoute_, planRoadPos_;            // per plan
    std::vector<int> lanePlanRoad_;                       // per route: first entry of its roads in lanePlanBeg_
    std::vector<int> lanePlanBeg_, lanePlanId_;           // per (route, road position): first entry of its lanes in lanePlanId_
};

// Parses the flow file; returns false with a message on stderr on format errors.
bool loadFlows(const std::string &path, const RoadNet &net, std::vector<FlowDef> &out);

}  // namespace cfb

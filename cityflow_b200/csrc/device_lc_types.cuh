// device_lc_types.cuh -- data layout of the lane-change path (see device_lc.cuh); included by device_view.cuh before the
// View struct.
#pragma once

namespace cfb {

constexpr int LC_PRIO_PENDING = INT_MIN;   // a shadow's priority until the host has drawn it
constexpr int LC_MAX_CAND = 4096;          // candidates / involved vehicles per step (capacity, checked)

// Per-slot lane-change state: LaneChangeInfo (vehicle.h:74-79) + LaneChange (lanechange.h:26-44).
// Signals are epoch-stamped instead of cleared every step (LaneChange::clearSignal, engine.cpp:424).
struct __align__(16) LcSlot {
    int partner;        // slot of the shadow / parent, -1 none
    int type;           // partnerType: 0 none, 1 has a shadow, 2 is a shadow
    int changing;       // LaneChange::changing
    int finished;       // LaneChange::finished (replaced by its shadow: not a "finished vehicle")
    int sendTarget;     // signalSend->target (lane id, -1 none)
    int sendDir;        // signalSend->direction
    int sendEpoch;      // step the signal was made in; a changing vehicle's signal persists
    int recvSrc;        // signalRecv->source (slot)
    int recvEpoch;
    int tgtLeader;      // targetLeader / targetFollower (slots, -1 none), valid while tgtEpoch == epoch
    int tgtFollower;
    int tgtEpoch;
    int lastDir;
    int plan;           // plan the vehicle follows (route and road position derive from it)
    int speedEpoch;     // buffer.isSpeedSet for this step (set by the partner), value in bufSpeed
    int pad0;
    double offset, waiting, lastChange;
    double leaderGap, followerGap;
    double gap;         // ControllerInfo::gap as the VEHICLE keeps it (V.gap[] is per position and is not
                        // carried along when a head has no leader; makeSignal reads the stale value)
    double head;        // order-independent part of the next speed (k_control), for the sequential tail
    double bufSpeed;
    int headBlocker;    // blocker found by the head computation (-1 none)
    int pad1;
};

struct LcCtrl {
    int nCand, nInvolved, nShadows, spareUsed;
    int error, pad[3];
};

struct LcView {
    LcSlot *slot;              // per slot
    int *segIdx;               // per position: Segment index (roadnet.cpp:863-875), valid after k_lc_segments
    const int *posDrv;         // per position: its drivable (static)
    const int *segBeg;         // per lane: first entry of its segments in segStart (nLanes + 1)
    const double *segStart;    // Segment::startPos
    const int *laneRoad;       // per lane: its road
    const int *routeLastRoad;  // per route: route.back() -- Router::isLastRoad is an identity test (router.cpp:131-134)
    const int *laneIdx;        // per lane: index in its road
    const int *laneRoadN;      // per lane: number of lanes of its road
    const double *laneWidth;
    const int *planRoute, *planRoadPos;                 // per plan
    const int *lanePlanRoad, *lanePlanBeg, *lanePlanId; // Routing::lanePlan tables
    int *cand, *involved;      // slots
    int *scratchA, *scratchB, *scratchC, *scratchD;   // LC_MAX_CAND ints each: keys / sorted copy / road of entry / shadow created
    const int *spare;          // free slots the host lent for this step's shadows
    int nSpare;
    int2 *shadowLog;           // (parent slot, shadow slot), schedule order = RNG order of their priorities
    LcCtrl *ctrl;
};

__device__ __forceinline__ bool lcSendValid(const LcSlot &L, int epoch) { return L.changing || L.sendEpoch == epoch; }
__device__ __forceinline__ bool lcRecvValid(const LcSlot &L, int epoch) { return L.recvEpoch == epoch; }
// LaneChange::planChange lanechange.cpp:23-25
__device__ __forceinline__ bool lcPlanChange(const LcSlot &L, int drivable, int epoch) {
    return (lcSendValid(L, epoch) && L.sendTarget >= 0 && L.sendTarget != drivable) || L.changing;
}
__device__ __forceinline__ void lcResetSlot(LcSlot &L, int plan) {
    L.partner = -1; L.type = 0; L.changing = 0; L.finished = 0;
    L.sendTarget = -1; L.sendDir = 0; L.sendEpoch = -1; L.recvSrc = -1; L.recvEpoch = -1;
    L.tgtLeader = L.tgtFollower = -1; L.tgtEpoch = -1; L.lastDir = 0; L.plan = plan; L.speedEpoch = -1; L.pad0 = 0;
    L.offset = 0; L.waiting = 0; L.lastChange = 0; L.leaderGap = 0; L.followerGap = 0; L.gap = 0; L.head = 0;
    L.bufSpeed = 0; L.headBlocker = -1; L.pad1 = 0;
}

}  // namespace cfb

// CPU probe: the WHOLE device step on the host.  TEST INFRASTRUCTURE ONLY.
//
// The bodies of all five step kernels (csrc/device_phases_a.cuh: k_ingest, k_notify; device_control.cuh:
// k_control; device_phases_b.cuh: k_move, k_leader) and, with laneChange, the kernels of the lane-change
// draft (csrc/device_lc.cuh) run on an emulated warp (tests/device_emu.h: 32 lock-stepped fibers) over
// persistent lane-bucket arrays laid out like DeviceSim's.  The emulated engine gets the same spawn events as
// the restatement (oracle/cityflow_oracle.cpp) and evolves ON ITS OWN; after every step its full state --
// list order, distance, speed, leader, gap, blocker, enterLaneLinkTime of every vehicle, vehicle counts,
// finished vehicles, and with laneChange every shadow, partner link, offset, waiting time -- must equal the
// restatement's.  That is the GPU parity test (tests/test_gpu_parity.py) minus the GPU: it checks the
// kernels' logic and their interplay, not launch configuration, streams or the host engine.
//
//   g++ -std=c++17 -O1 -ffp-contract=off -I/usr/local/cuda/include -Icityflow_b200/csrc tests/device_step_probe.cpp \
//       cityflow_b200/csrc/roadnet.cpp cityflow_b200/csrc/flows.cpp -o probe && ./probe config.json steps
#include <algorithm>
#include <climits>
#include <cmath>
#include <map>
#include <vector>

#include "device_emu.h"
static struct { unsigned x = 0, y = 0, z = 0; } blockIdx;
static struct { unsigned x = 1, y = 1, z = 1; } gridDim;
#define CFB_LANE_CHANGE 1
#define CFB_DEAD_END_STOP 1
using std::max;
using std::min;
namespace cfb { namespace cg = cooperative_groups; }

#include "device_sim.h"
#include "device_view.cuh"
#include "device_phases_a.cuh"
#include "device_control.cuh"   // (+ device_lc.cuh, kernels included: blockIdx / gridDim are globals here)
#include "device_phases_b.cuh"

#include "../oracle/cityflow_oracle.cpp"

namespace {

using namespace cfb;

template <class T> struct Buf : std::vector<T> { T *p() { return this->data(); } };

struct HostSim {   // DeviceSim's arrays (device_sim.cu constructor), on the host
    View V{};
    int P = 0, slotCap = 0;
    Buf<double> drvLength, drvMaxSpeed, lcDist, phaseTime, gap, remain, cust, slotCust, segStart, laneWidth;
    Buf<int> off, laneOutBeg, laneOutLinks, llStartLane, llEndLane, llRoadLink, llCrossBeg, lcIdx, csLink, lcPeer, interPhaseBeg,
        interRLBeg, phaseAvailBeg, rlInter, planBeg, planData, leader, count, pos, waitHead, waitTail, waitNext, curPhase, entCnt, ent,
        act0, act1, extra, blk, delStep, segIdx, posDrv, segBeg, laneIdx, laneRoadN, planRoute, planRoadPos, lpRoad, lpBeg, lpId, cand,
        involved, spare, prio, routeLast, laneRoadT;
    Buf<unsigned char> phaseAvail, interVirtual, inserted, rlAvail;
    Buf<DTmpl> tmpl;
    Buf<double2> kin, nkin, mkin;
    Buf<int4> ids, nav, slotInfo, mids, mnav, linkInfo;
    Buf<int2> nbuf, finSlots, veh0, veh1, shadowLog;
    Buf<Notify> notify;
    Buf<Tail> tail;
    Buf<unsigned> foeMask;
    Buf<LcSlot> lcSlot;
    Buf<SpawnRec> spawn;
    Ctrl ctrl{};
    LcCtrl lcCtrl{};
    long long steps = 0;

    void init(const RoadNet &net, double interval, bool rl, bool laneChange, const std::vector<std::vector<double>> &segStartPerLane) {
        const int nL = net.nLanes(), nK = net.nLinks(), nD = nL + nK;
        V.nLanes = nL; V.nLinks = nK; V.nDrv = nD; V.nInter = net.nInter(); V.nRL = net.nRoadLinks(); V.nCross = net.nCross();
        V.dt = interval; V.rl = rl ? 1 : 0;
        drvLength.resize(nD); drvMaxSpeed.resize(nD); off.assign(nD + 1, 0);
        for (int d = 0; d < nD; ++d) {
            drvLength[d] = d < nL ? net.laneLength[d] : net.llLength[d - nL];
            drvMaxSpeed[d] = d < nL ? net.laneMaxSpeed[d] : 10000.0;
            int cap = (int) (drvLength[d] / 2.5) + 8;
            cap = (cap + 3) & ~3;
            off[d + 1] = off[d] + cap;
        }
        P = off[nD];
        laneOutBeg.assign(nL + 1, 0);
        for (int l = 0; l < nL; ++l) { for (int ll : net.laneOutLinks[l]) laneOutLinks.push_back(ll); laneOutBeg[l + 1] = (int) laneOutLinks.size(); }
        if (laneOutLinks.empty()) laneOutLinks.push_back(0);
        llCrossBeg.assign(nK + 1, 0);
        for (int k = 0; k < nK; ++k) {
            for (const CrossRef &c : net.llCrosses[k]) { lcIdx.push_back(c.cross * 2 + c.side); lcDist.push_back(net.crossDist[c.side][c.cross]); }
            llCrossBeg[k + 1] = (int) lcIdx.size();
        }
        csLink.assign(std::max(2 * net.nCross(), 1), 0);
        for (int c = 0; c < net.nCross(); ++c) { csLink[2 * c] = net.crossLink[0][c]; csLink[2 * c + 1] = net.crossLink[1][c]; }
        lcPeer.assign(lcIdx.size(), 0);
        std::vector<int> flatOf(std::max(2 * net.nCross(), 1), 0);
        int maxCross = 1;
        for (int k = 0; k < nK; ++k) {
            maxCross = std::max(maxCross, llCrossBeg[k + 1] - llCrossBeg[k]);
            for (int q = llCrossBeg[k]; q < llCrossBeg[k + 1]; ++q) flatOf[lcIdx[q]] = q;
        }
        for (size_t q = 0; q < lcIdx.size(); ++q) lcPeer[q] = flatOf[lcIdx[q] ^ 1];
        V.maskWords = (maxCross + 31) / 32;
        if (lcIdx.empty()) { lcIdx.push_back(0); lcDist.push_back(0); lcPeer.push_back(0); }
        auto cp = [](Buf<int> &b, const std::vector<int> &v) { b.assign(v.begin(), v.end()); if (b.empty()) b.push_back(0); };
        cp(llStartLane, net.llStartLane); cp(llEndLane, net.llEndLane); cp(llRoadLink, net.llRoadLink);
        linkInfo.assign(std::max(nK, 1), make_int4(0, 0, 0, 0));
        for (int k = 0; k < nK; ++k)
            linkInfo[k] = make_int4(net.llRoadLink[k], net.llEndLane[k], llCrossBeg[k], (net.linkIsTurn(k) ? 1 : 0) | (net.rlType[net.llRoadLink[k]] << 8));
        cp(interPhaseBeg, net.interPhaseBeg); cp(interRLBeg, net.interRoadLinkBeg); cp(phaseAvailBeg, net.phaseAvailBeg); cp(rlInter, net.rlInter);
        phaseTime.assign(net.phaseTime.begin(), net.phaseTime.end()); if (phaseTime.empty()) phaseTime.push_back(0);
        phaseAvail.assign(net.phaseAvail.begin(), net.phaseAvail.end()); if (phaseAvail.empty()) phaseAvail.push_back(0);
        interVirtual.assign(net.interVirtual.begin(), net.interVirtual.end());
        // dynamic
        kin.assign(P, make_double2(0, 0)); nkin.assign(P, make_double2(0, 0)); gap.assign(P, 0); leader.assign(P, -1);
        ids.assign(P, make_int4(0, 0, 0, 0)); nav.assign(P, make_int4(0, 0, 0, 0)); nbuf.assign(P, make_int2(0, 0));
        count.assign(nD, 0); entCnt.assign(nD, 0); ent.assign((size_t) nD * ENT_CAP, 0);
        waitHead.assign(std::max(nL, 1), -1); waitTail.assign(std::max(nL, 1), -1); inserted.assign(std::max(nL, 1), 0);
        notify.assign(std::max(2 * net.nCross(), 1), Notify{0.0, 0, 0});
        curPhase.assign(net.nInter(), 0); remain.assign(net.nInter(), 0.0); rlAvail.assign(std::max(net.nRoadLinks(), 1), 0);
        for (int i = 0; i < net.nInter(); ++i) if (!net.interVirtual[i]) remain[i] = net.phaseTime[net.interPhaseBeg[i]];
        V.moverCap = P;
        mkin.assign(P, make_double2(0, 0)); mids.assign(P, make_int4(0, 0, 0, 0)); mnav.assign(P, make_int4(0, 0, 0, 0));
        cust.assign(P, NAN);
        Tail empty{}; empty.pos = -1; empty.prev = -1;
        tail.assign(nD, empty);
        foeMask.assign((size_t) std::max(nK, 1) * V.maskWords, 0u);
        V.vehCap = P;
        veh0.assign(P, make_int2(0, 0)); veh1.assign(P, make_int2(0, 0)); act0.assign(nD, 0); act1.assign(nD, 0); extra.assign(nD, 0);
        V.finCap = 1 << 16;
        finSlots.assign(V.finCap, make_int2(0, 0));
        slotCap = 1 << 17;
        pos.assign(slotCap, -1); waitNext.assign(slotCap, -1); slotInfo.assign(slotCap, make_int4(0, 0, 0, 0)); slotCust.assign(slotCap, NAN);
        blk.assign(slotCap, -1); delStep.assign(slotCap, INT_MIN);
        tmpl.assign(slotCap, DTmpl{});
        spawn.assign(1 << 14, SpawnRec{});
        // lane change
        segBeg.assign(nL + 1, 0); laneIdx.resize(nL); laneRoadN.resize(nL); laneWidth.resize(nL);
        for (int l = 0; l < nL; ++l) {
            segBeg[l] = (int) segStart.size();
            for (double x : segStartPerLane[l]) segStart.push_back(x);
            laneIdx[l] = net.laneIdx[l]; laneRoadN[l] = net.roadNumLanes(net.laneRoad[l]); laneWidth[l] = net.laneWidth[l];
        }
        segBeg[nL] = (int) segStart.size();
        posDrv.resize(P);
        for (int d = 0; d < nD; ++d) for (int p = off[d]; p < off[d + 1]; ++p) posDrv[p] = d;
        segIdx.assign(P, 0); cand.assign(LC_MAX_CAND, 0); involved.assign(LC_MAX_CAND, 0); shadowLog.assign(LC_MAX_CAND, make_int2(0, 0));
        prio.assign(LC_MAX_CAND, 0); lcSlot.assign(slotCap, LcSlot{}); spare.assign(256, 0);
        // pointers
        V.drvLength = drvLength.p(); V.drvMaxSpeed = drvMaxSpeed.p(); V.off = off.p(); V.laneOutBeg = laneOutBeg.p(); V.laneOutLinks = laneOutLinks.p();
        V.llStartLane = llStartLane.p(); V.llEndLane = llEndLane.p(); V.llRoadLink = llRoadLink.p(); V.linkInfo = linkInfo.p();
        V.llCrossBeg = llCrossBeg.p(); V.lcIdx = lcIdx.p(); V.lcDist = lcDist.p(); V.csLink = csLink.p(); V.lcPeer = lcPeer.p();
        V.interPhaseBeg = interPhaseBeg.p(); V.interRLBeg = interRLBeg.p(); V.phaseAvailBeg = phaseAvailBeg.p(); V.rlInter = rlInter.p();
        V.phaseTime = phaseTime.p(); V.phaseAvail = phaseAvail.p(); V.interVirtual = interVirtual.p(); V.tmpl = tmpl.p();
        V.kin = kin.p(); V.nkin = nkin.p(); V.gap = gap.p(); V.leader = leader.p(); V.ids = ids.p(); V.nav = nav.p(); V.nbuf = nbuf.p();
        V.count = count.p(); V.pos = pos.p(); V.waitHead = waitHead.p(); V.waitTail = waitTail.p(); V.waitNext = waitNext.p();
        V.slotInfo = slotInfo.p(); V.inserted = inserted.p(); V.notify = notify.p(); V.tail = tail.p(); V.foeMask = foeMask.p();
        V.curPhase = curPhase.p(); V.remain = remain.p(); V.rlAvail = rlAvail.p(); V.entCnt = entCnt.p(); V.ent = ent.p();
        V.mkin = mkin.p(); V.mids = mids.p(); V.mnav = mnav.p(); V.finSlots = finSlots.p();
        V.vehList[0] = veh0.p(); V.vehList[1] = veh1.p(); V.actList[0] = act0.p(); V.actList[1] = act1.p(); V.extraList = extra.p();
        V.cust = cust.p(); V.slotCust = slotCust.p(); V.blk = blk.p(); V.delStep = delStep.p(); V.ctrl = &ctrl;
        V.lcOn = laneChange ? 1 : 0;
        LcView &C = V.lc;
        C.slot = lcSlot.p(); C.segIdx = segIdx.p(); C.posDrv = posDrv.p(); C.segBeg = segBeg.p(); C.segStart = segStart.p();
        laneRoadT.assign(net.laneRoad.begin(), net.laneRoad.end()); C.laneRoad = laneRoadT.p();
        C.laneIdx = laneIdx.p(); C.laneRoadN = laneRoadN.p(); C.laneWidth = laneWidth.p(); C.cand = cand.p(); C.involved = involved.p();
        C.spare = spare.p(); C.nSpare = 0; C.shadowLog = shadowLog.p(); C.ctrl = &lcCtrl;
    }
    void setPlans(const Routing &R) {
        planBeg.assign(R.planBeg().begin(), R.planBeg().end()); planData.assign(R.planData().begin(), R.planData().end());
        planRoute.assign(R.planRouteTable().begin(), R.planRouteTable().end()); planRoadPos.assign(R.planRoadPosTable().begin(), R.planRoadPosTable().end());
        lpRoad.assign(R.lanePlanRoadTable().begin(), R.lanePlanRoadTable().end()); lpBeg.assign(R.lanePlanBegTable().begin(), R.lanePlanBegTable().end());
        lpId.assign(R.lanePlanIdTable().begin(), R.lanePlanIdTable().end());
        routeLast.assign(std::max(R.numRoutes(), 1), -1);
        for (int r = 0; r < R.numRoutes(); ++r) if (R.route(r).valid) routeLast[r] = R.route(r).roads.back();
        V.lc.routeLastRoad = routeLast.p();
        V.planBeg = planBeg.p(); V.planData = planData.p();
        V.lc.planRoute = planRoute.p(); V.lc.planRoadPos = planRoadPos.p(); V.lc.lanePlanRoad = lpRoad.p(); V.lc.lanePlanBeg = lpBeg.p(); V.lc.lanePlanId = lpId.p();
    }
    template <class F> void run(int nBlocks, F f) {
        gridDim.x = (unsigned) nBlocks;
        emu::launch(nBlocks, [&](int b, int n) { blockIdx.x = (unsigned) b; f(b, n); });
    }
};

HostSim *S = nullptr;
std::map<Veh *, int> slotOf;
std::vector<Veh *> vehOfSlot;
int g_nextSlot = 0, g_fail = 0;
long long g_checked = 0, g_shadows = 0, g_finished = 0;

#define CHECK(cond, ...)                                                                            \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            if (g_fail < 12) { printf("FAIL step %zu: ", o.step); printf(__VA_ARGS__); printf("\n"); } \
            ++g_fail;                                                                               \
        }                                                                                           \
    } while (0)

int newSlot(Veh *v) {
    const int s = g_nextSlot++;
    if (s >= S->slotCap) { printf("probe: out of slots\n"); exit(3); }
    slotOf[v] = s;
    if ((int) vehOfSlot.size() <= s) vehOfSlot.resize(s + 1, nullptr);
    vehOfSlot[s] = v;
    return s;
}

// one step of the emulated engine, fed with what the restatement spawned in its own (already executed) step
void deviceStep(Oracle &o) {
    HostSim &H = *S;
    View &V = H.V;
    Routing &R = *o.routing;
    // ---- host part: spawn records (HostEngine::prepareStep), lane-sorted, stable ----
    std::vector<SpawnRec> recs;
    for (Veh *v : o.spawnedThisStep) {
        const int s = newSlot(v);
        const int route = R.intern(v->route);
        if (R.route(route).roads != v->route) { printf("probe: interning the resolved route changed it (prio %d)\n", v->priority); ++g_fail; }
        SpawnRec r{};
        r.slot = s; r.lane = v->firstLaneForProbe; r.tmpl = s; r.priority = v->priority;
        r.plan = R.lanePlan(route, 0, o.net.laneIdx[r.lane]);
        recs.push_back(r);
        DTmpl t{};   // toDevice() of device_sim.cu
        t.len = v->t.len; t.maxPosAcc = v->t.maxPosAcc; t.maxNegAcc = v->t.maxNegAcc; t.usualPosAcc = v->t.usualPosAcc;
        t.usualNegAcc = v->t.usualNegAcc; t.minGap = v->t.minGap; t.maxSpeed = v->t.maxSpeed; t.headwayTime = v->t.headwayTime;
        t.yieldDistance = v->t.yieldDistance; t.turnSpeed = v->t.turnSpeed;
        t.approachDist = v->t.maxSpeed * v->t.maxSpeed / v->t.usualNegAcc / 2 + v->t.maxSpeed * o.interval * 2;
        t.speed0 = v->initialSpeedForProbe;
        H.tmpl[s] = t;
    }
    std::stable_sort(recs.begin(), recs.end(), [](const SpawnRec &a, const SpawnRec &b) { return a.lane < b.lane; });
    H.setPlans(R);   // (routes are interned lazily: tables may have grown)
    H.spawn[0].slot = (int) recs.size();
    for (size_t k = 0; k < recs.size(); ++k) H.spawn[k + 1] = recs[k];
    V.spawn = H.spawn.p() + 1;
    V.par = (int) (H.steps & 1);
    const int G = 3;   // blocks per launch: more than one, so the grid-stride loops are exercised
    H.run(G, [&](int b, int n) { phase_ingest(V, b, n); });
    if (V.lcOn) {
        for (int k = 0; k < 64; ++k) H.spare[k] = g_nextSlot + k;   // slots lent for this step's shadows
        V.lc.nSpare = 64;
        H.lcCtrl.nCand = H.lcCtrl.nInvolved = H.lcCtrl.nShadows = H.lcCtrl.spareUsed = 0;   // k_lc_begin
        H.run(G, [&](int, int) { k_lc_admitted(V, V.lc); });
        H.run(G, [&](int, int) { k_lc_segments(V, V.lc); });
        H.run(G, [&](int, int) { k_lc_signal(V, V.lc); });
        if (const char *tp = getenv("PROBE_TRACE_PRIO")) {
            const int want = atoi(tp);
            for (auto &kv : o.pool) if (kv.first == want && slotOf.count(kv.second)) {
                const int s = slotOf.at(kv.second), p = H.pos[s];
                if (p < 0) continue;
                const LcSlot &L = H.lcSlot[s];
                const int dd = H.posDrv[p];
                const int rp = lcRoadPos(V, V.lc, L.plan, H.nav[p].x);
                printf("  signal step %zu: drivable %d laneIdx %d/%d dis %.4f own gap %.6f len %.2f seg %d plan %d planIdx %d roadPos %d next %d -> target %d epoch %d\n", o.step, dd, H.laneIdx[dd],
                       H.laneRoadN[dd], H.kin[p].x, L.gap, H.tmpl[H.ids[p].y].len, H.segIdx[p], L.plan, H.nav[p].x, rp, H.ids[p].w, L.sendTarget, L.sendEpoch);
                {
                    Veh *v = kv.second;
                    printf("    restatement: iCur %d route size %zu onLastRoad %d; routerNextOf(outer) %d; plan's route %d roads:", v->iCur, v->route.size(), (int) o.onLastRoad(*v),
                           dd + 1 < V.nLanes ? o.routerNextOf(*v, dd + 1) : -9, H.planRoute[L.plan]);
                    for (int r2 : v->route) printf(" %d", r2);
                    printf(" | lane road %d; lanePlan ids for roadPos %d:", o.net.laneRoad[dd], rp);
                    for (int k2 = 0; k2 < H.laneRoadN[dd]; ++k2) { const int np = lcLanePlan(V.lc, L.plan, rp, dd - H.laneIdx[dd] + k2); printf(" %d(next %d)", np, H.planData[H.planBeg[np] + 1]); }
                    printf("\n");
                }
                for (int adj : {dd + 1, dd - 1}) {
                    if (adj < 0 || adj >= V.nLanes || o.net.laneRoad[adj] != o.net.laneRoad[dd]) continue;
                    const int lp = lcVehicleAfter(V, V.lc, adj, H.kin[p].x, H.segIdx[p]);
                    printf("    lane %d: continues %d, vehicle ahead pos %d est %.4f (count %d)\n", adj, (int) lcLaneContinues(V, V.lc, L.plan, rp, adj, H.ids[p].w == -1), lp,
                           lp < 0 ? H.drvLength[adj] - H.kin[p].x : H.kin[lp].x - H.kin[p].x - H.tmpl[H.ids[lp].y].len, H.count[adj]);
                }
            }
        }
        H.run(1, [&](int, int) { k_lc_schedule(V, V.lc); });
        // host round trip: the shadows' priorities, in schedule order (HostEngine::nextStepLaneChange); here they
        // come from the restatement's own draws, matched by parent
        const int ns = H.lcCtrl.nShadows;
        for (int k = 0; k < ns; ++k) {
            Veh *parent = vehOfSlot[H.shadowLog[k].x];
            int pr = INT_MIN;
            for (auto &sp : o.shadowsThisStep) if (sp.first == parent) pr = sp.second;
            CHECK(pr != INT_MIN, "the draft created a shadow for prio %d, the restatement did not", parent ? parent->priority : 0);
            H.prio[k] = pr;
            // the restatement's shadow object, if it survived its first step (it may have aborted at once)
            auto it = pr != INT_MIN ? o.pool.find(pr) : o.pool.end();
            if (it != o.pool.end()) {
                const int sh = H.shadowLog[k].y;
                slotOf[it->second] = sh;
                if ((int) vehOfSlot.size() <= sh) vehOfSlot.resize(sh + 1, nullptr);
                vehOfSlot[sh] = it->second;
            }
        }
        CHECK(ns == (int) o.shadowsThisStep.size(), "shadows created: restatement %zu, draft %d", o.shadowsThisStep.size(), ns);
        for (auto &sp : o.shadowsThisStep) {   // diagnostics: a parent the draft did not serve
            Veh *pv = sp.first;
            const int ps = slotOf.at(pv);
            bool found = false;
            for (int k = 0; k < ns; ++k) found |= H.shadowLog[k].x == ps;
            if (!found) {
                const LcSlot &L = H.lcSlot[ps];
                const int p = H.pos[ps];
                printf("  missing shadow of prio %d: device send valid %d target %d (restatement target %d) recv valid %d tgtL %d tgtF %d gaps %g %g | stale gap %g vs %g | dis %g seg %d | cand %d\n",
                       pv->priority, (int) lcSendValid(L, (int) o.step), L.sendTarget, pv->sigSend ? pv->sigSend->target : -9, (int) lcRecvValid(L, (int) o.step), L.tgtLeader, L.tgtFollower,
                       L.leaderGap, L.followerGap, L.gap, pv->gap, H.kin[p].x, H.segIdx[p], H.lcCtrl.nCand);
            }
        }
        g_nextSlot += H.lcCtrl.spareUsed;
        g_shadows += ns;
        if (ns > 0) H.run(1, [&](int, int) { k_lc_priorities(V, V.lc, H.prio.p(), ns); });
        H.run(G, [&](int, int) { k_lc_leader(V, V.lc); });
        if (const char *tp = getenv("PROBE_TRACE_PRIO")) {
            const int want = atoi(tp);
            for (int k = 0; k < ns; ++k) if (H.prio[k] == want) {
                const int sh = H.shadowLog[k].y, p = H.pos[sh];
                printf("  mid-step: shadow slot %d at pos %d (drivable %d, count %d) next %d planIdx %d leader %d V.gap %.6f own %.6f; nAct %d; parent tgtLeader %d\n", sh, p, H.posDrv[p],
                       H.count[H.posDrv[p]], H.ids[p].w, H.nav[p].x, H.leader[p], H.gap[p], H.lcSlot[sh].gap, H.ctrl.nAct[V.par], H.lcSlot[H.shadowLog[k].x].tgtLeader);
                bool inAct = false;
                for (int a2 = 0; a2 < H.ctrl.nAct[V.par]; ++a2) inAct |= V.actList[V.par][a2] == H.posDrv[p];
                printf("  drivable in actList: %d\n", (int) inAct);
            }
        }
    }
    H.run(G, [&](int b, int n) { phase_notify(V, b, n); });
    H.run(G, [&](int b, int n) { phase_control(V, b, n); });
    if (V.lcOn) H.run(1, [&](int, int) { k_lc_control_tail(V, V.lc); });
    H.run(G, [&](int b, int n) { phase_move(V, b, n); });
    H.run(G, [&](int b, int n) { phase_leader(V, b, n); });
    H.steps += 1;
    CHECK(H.ctrl.error == 0 && H.lcCtrl.error == 0, "device error flags %d / %d", H.ctrl.error, H.lcCtrl.error);
}

void compare(Oracle &o, std::vector<Veh *> &removedThisStep) {
    HostSim &H = *S;
    View &V = H.V;
    // shadows created this step: map the restatement's object to the draft's slot (through the parent)
    for (auto &kv : o.pool) {
        Veh *v = kv.second;
        if (!v->running || slotOf.count(v)) continue;
        CHECK(v->partnerType == 2 && v->partner && slotOf.count(v->partner), "unknown running vehicle prio %d", v->priority);
        if (v->partner && slotOf.count(v->partner)) {
            const int sh = H.lcSlot[slotOf.at(v->partner)].partner;
            CHECK(sh >= 0, "parent prio %d has no shadow on the device", v->partner->priority);
            if (sh >= 0) { slotOf[v] = sh; if ((int) vehOfSlot.size() <= sh) vehOfSlot.resize(sh + 1, nullptr); vehOfSlot[sh] = v; }
        }
    }
    if (const char *tp = getenv("PROBE_TRACE_PRIO")) {
        const int want = atoi(tp);
        for (auto &kv : o.pool) if (kv.first == want && kv.second->running && slotOf.count(kv.second)) {
            Veh *v = kv.second; const int s = slotOf.at(v); const LcSlot &L = H.lcSlot[s]; const int p = H.pos[s];
            printf("  trace step %zu prio %d: type %d/%d drivable %d/%d dis %.4f/%.4f leader %d/%d own gap %.6f/%.6f V.gap %.6f changing %d/%d\n", o.step, want, v->partnerType, L.type,
                   v->drivable, H.posDrv[p], v->dis, H.kin[p].x, v->leader ? v->leader->priority : 0, H.leader[p] >= 0 ? H.ids[H.leader[p]].z : 0, v->gap, L.gap, H.gap[p], v->changing, L.changing);
        }
    }
    CHECK((int) o.activeCount == H.ctrl.active, "active vehicles %zu vs %d", o.activeCount, H.ctrl.active);
    for (int d = 0; d < V.nDrv; ++d) {
        CHECK((int) o.lists[d].size() == H.count[d], "drivable %d holds %zu vs %d vehicles", d, o.lists[d].size(), H.count[d]);
        if ((int) o.lists[d].size() != H.count[d]) continue;
        int k = 0;
        for (Veh *v : o.lists[d]) {
            const int p = H.off[d] + k++;
            const int s = slotOf.count(v) ? slotOf.at(v) : -1;
            CHECK(H.ids[p].x == s, "drivable %d position %d: slot %d, expected %d (prio %d)", d, k - 1, H.ids[p].x, s, v->priority);
            if (H.ids[p].x != s) continue;
            CHECK(H.pos[s] == p && H.ids[p].z == v->priority, "prio %d: pos / priority record", v->priority);
            CHECK(H.kin[p].x == v->dis && H.kin[p].y == v->t.speed, "prio %d: dis %.17g/%.17g speed %.17g/%.17g", v->priority, v->dis, H.kin[p].x,
                  v->t.speed, H.kin[p].y);
            const int lp = H.leader[p];
            CHECK((v->leader ? slotOf.at(v->leader) : -1) == (lp >= 0 ? H.ids[lp].x : -1), "prio %d: leader", v->priority);
            if (v->leader && lp >= 0) CHECK(H.gap[p] == v->gap, "prio %d: gap %.17g vs %.17g", v->priority, v->gap, H.gap[p]);
            int devBlocker = H.nav[p].z;   // a blocker that left in the step just done is dropped lazily on the device (DeviceSim::debugDump)
            if (devBlocker >= 0 && H.delStep[devBlocker] == H.ctrl.step - 1) devBlocker = -1;
            CHECK((v->blocker ? slotOf.at(v->blocker) : -1) == devBlocker, "prio %d (type %d) on drivable %d: blocker %d (prio %d) vs slot %d", v->priority,
                  v->partnerType, d, v->blocker ? slotOf.at(v->blocker) : -1, v->blocker ? v->blocker->priority : 0, devBlocker);
            CHECK((int) v->enterLaneLinkTime == H.nav[p].w && v->prevDrivable == H.nav[p].y, "prio %d: enterLaneLinkTime / prevDrivable", v->priority);
            if (V.lcOn) {
                const LcSlot &L = H.lcSlot[s];
                CHECK(L.type == v->partnerType && (v->partner ? slotOf.at(v->partner) : -1) == L.partner, "prio %d: partner type %d/%d", v->priority, v->partnerType, L.type);
                CHECK(L.gap == v->gap, "prio %d (drivable %d, list index %d of %d, leader %d): the vehicle's own gap value %.17g vs %.17g", v->priority, d, k - 1,
                      H.count[d], v->leader ? v->leader->priority : 0, v->gap, L.gap);
                CHECK(L.changing == (int) v->changing && L.offset == v->offset, "prio %d: changing %d/%d offset %g/%g", v->priority, v->changing, L.changing, v->offset, L.offset);
                CHECK(L.waiting == v->waitingTime && L.lastChange == v->lastChangeTime && L.lastDir == v->lastDir, "prio %d: waiting %g/%g lastChange %g/%g lastDir %d/%d",
                      v->priority, v->waitingTime, L.waiting, v->lastChangeTime, L.lastChange, v->lastDir, L.lastDir);
            }
            ++g_checked;
        }
    }
    // waiting queues
    for (int l = 0; l < V.nLanes; ++l) {
        int n = 0;
        for (int s = H.waitHead[l]; s >= 0; s = H.waitNext[s]) ++n;
        CHECK(n == (int) o.waiting[l].size(), "lane %d: waiting queue %zu vs %d", l, o.waiting[l].size(), n);
    }
    // finished ring (HostEngine::drain)
    int fin = 0;
    for (int k = 0; k < H.ctrl.finCount; ++k) {
        const int tag = H.finSlots[k].x, s = tag & ~0x40000000;
        if (!(tag & 0x40000000)) ++fin;
        Veh *v = vehOfSlot[s];
        if (v) { slotOf.erase(v); vehOfSlot[s] = nullptr; }
    }
    H.ctrl.finCount = 0;
    g_finished += fin;
    CHECK(g_finished == o.finishedCnt, "finished vehicles %d vs %lld", o.finishedCnt, g_finished);
    (void) removedThisStep;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: device_step_probe config.json steps\n"); return 64; }
    Oracle o;
    if (!o.load(argv[1])) { fprintf(stderr, "cannot load %s\n", argv[1]); return 2; }
    o.routing->enableLanePlans();
    HostSim H;
    H.init(o.net, o.interval, o.rlTrafficLight, o.laneChange, o.segStart);
    S = &H;
    const int steps = atoi(argv[2]);
    std::vector<Veh *> removed;
    for (int s = 0; s < steps && g_fail == 0; ++s) {
        o.nextStep();
        deviceStep(o);
        compare(o, removed);
    }
    printf("%s %lld steps, %lld vehicle states, %lld finished, %lld shadows, %d failures\n", g_fail ? "FAIL" : "OK", H.steps, g_checked, g_finished, g_shadows, g_fail);
    return g_fail ? 1 : 0;
}

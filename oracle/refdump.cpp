// TEST INFRASTRUCTURE ONLY -- white-box driver around the UNMODIFIED reference engine.
//
// Built by oracle/Makefile against the reference sources where they lie under
// /root/reference (never copied into this repository); the binary lands in
// oracle/_ref/ (git-ignored, travels to the GPU box).  It exposes three modes:
//
//   refdump static <config.json> <out.bin>
//       static tables the hot path consumes (lane / laneLink lengths, cross
//       ordering and distances) -> used to validate our host loader bit-for-bit.
//   refdump run <config.json> <steps> <threads> <out.bin> [every]
//       per-step dynamic state of every running vehicle (raw IEEE-754 bits) ->
//       the parity oracle for tests/ and for generating tests/golden/.
//   refdump archive <config.json> <steps> <threads> <out.json>
//       the reference's JSON archive (Archive::dump) after <steps> steps;
//   refdump resume <config.json> <archive.json> <steps> <threads> <out.bin> [every]
//       Engine::loadFromFile(<archive.json>), then as `run` -> archives interchange with cityflow_b200's in both directions.
//   refdump runlc <config.json> <steps> <threads> <out.bin> [every]
//       the same for laneChange=true runs: every running vehicle including shadows, keyed by
//       priority, with the lane-change state (partner, offset, changing, waiting time).
//   refdump counts <config.json> <steps> <threads> <out.bin> [every]
//       light-weight observables for long / large runs (the multi-GPU parity tests): the vehicle count
//       after EVERY step, and every `every` steps the per-lane vehicle count, per-lane waiting count
//       (speed < 0.1) and per-lane sum of speeds.
//   refdump sweep <config.json> <steps> <warmup> <t1,t2,...>
//       thread sweep at ONE operating point: the first thread count runs the warm-up
//       steps, its state goes through the reference's own Archive (dump / loadFromFile) into one engine per
//       further thread count, each is timed over <steps> steps; prints one JSON line.
//   refdump bench <config.json> <steps> <threads> [warmup] [counts.bin]
//       timing loop in the shape of tools/debug/simple_run.cpp:42-57, prints one
//       JSON line -> the `--impl reference` arm of bench.py.
//
// Access to Engine internals uses the access-specifier trick described in
// SURVEY.md Appendix A (class layout is unaffected).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <condition_variable>
#include <vector>
#include <unistd.h>

#define private public
#define protected public
#include "engine/engine.h"
#undef private
#undef protected

using namespace CityFlow;

namespace {

struct Out {
    FILE *fp;
    explicit Out(const char *path) : fp(fopen(path, "wb")) {
        if (!fp) { perror(path); exit(2); }
    }
    ~Out() { if (fp) fclose(fp); }
    void close() { if (fp) { fclose(fp); fp = nullptr; } }
    void i32(int32_t v) { fwrite(&v, 4, 1, fp); }
    void i64(int64_t v) { fwrite(&v, 8, 1, fp); }
    void f64(double v) { fwrite(&v, 8, 1, fp); }
};

struct Index {
    std::map<const Drivable *, int> drivable;
    std::map<const RoadLink *, int> roadLink;
    std::map<const Cross *, int> cross;
    std::map<const Road *, int> road;
    std::map<const Intersection *, int> inter;
    explicit Index(Engine &e) {
        int k = 0;
        for (Lane *l : e.roadnet.getLanes()) drivable[l] = k++;
        for (LaneLink *l : e.roadnet.getLaneLinks()) drivable[l] = k++;
        int r = 0, c = 0, rd = 0, it = 0;
        for (Road &road_ : e.roadnet.getRoads()) road[&road_] = rd++;
        for (Intersection &in : e.roadnet.getIntersections()) {
            inter[&in] = it++;
            for (RoadLink &rl : in.getRoadLinks()) roadLink[&rl] = r++;
            for (Cross &cr : in.getCrosses()) cross[&cr] = c++;
        }
    }
};

// "flow_<k>_<n>" -> (k, n); anything else -> (-1, hash-free -1)
void parseId(const std::string &id, int32_t &flow, int32_t &cnt) {
    flow = -1; cnt = -1;
    if (id.compare(0, 5, "flow_") == 0) {
        size_t us = id.find('_', 5);
        if (us != std::string::npos) {
            flow = atoi(id.substr(5, us - 5).c_str());
            cnt = atoi(id.substr(us + 1).c_str());
        }
    } else if (id.compare(0, 16, "manually_pushed_") == 0) {
        flow = -2;
        cnt = atoi(id.substr(16).c_str());
    }
}

int dumpStatic(const char *cfg, const char *outPath) {
    Engine e(cfg, 1);
    Index ix(e);
    Out o(outPath);
    const auto &lanes = e.roadnet.getLanes();
    const auto &links = e.roadnet.getLaneLinks();
    o.i32(0x43465331);  // 'CFS1'
    o.i32((int32_t) e.roadnet.getRoads().size());
    o.i32((int32_t) e.roadnet.getIntersections().size());
    o.i32((int32_t) lanes.size());
    o.i32((int32_t) links.size());
    o.i32((int32_t) ix.roadLink.size());
    o.i32((int32_t) ix.cross.size());
    for (Lane *l : lanes) {
        o.f64(l->getLength());
        o.f64(l->getMaxSpeed());
        o.i32(ix.road[l->getBelongRoad()]);
        o.i32((int32_t) l->getLaneIndex());
        o.i32((int32_t) l->getLaneLinks().size());
        for (LaneLink *ll : l->getLaneLinks()) o.i32(ix.drivable[ll]);
    }
    for (LaneLink *l : links) {
        o.f64(l->getLength());
        o.i32(ix.drivable[l->getStartLane()]);
        o.i32(ix.drivable[l->getEndLane()]);
        o.i32(ix.roadLink[l->getRoadLink()]);
        o.i32((int32_t) l->getRoadLinkType());
        o.i32((int32_t) l->getCrosses().size());
        for (Cross *c : l->getCrosses()) {
            int side = (c->getLaneLink(0) == l) ? 0 : 1;
            o.i32(ix.cross[c]);
            o.i32(side);
            o.i32(ix.drivable[c->getLaneLink(1 - side)]);
            o.f64(c->distanceOnLane[side]);
            o.f64(c->distanceOnLane[1 - side]);
        }
    }
    // traffic lights: per intersection, phases (time, availability per roadLink)
    for (Intersection &in : e.roadnet.getIntersections()) {
        o.i32(in.isVirtualIntersection() ? 1 : 0);
        o.i32((int32_t) in.getRoadLinks().size());
        auto &phases = in.getTrafficLight().getPhases();
        o.i32((int32_t) phases.size());
        for (auto &ph : phases) {
            o.f64(ph.time);
            for (size_t k = 0; k < in.getRoadLinks().size(); ++k) o.i32(ph.roadLinkAvailable[k] ? 1 : 0);
        }
    }
    o.close();
    fflush(stdout);
    _exit(0);   // (no ~Engine, see dumpRun)
    return 0;
}

int dumpRun(const char *cfg, int steps, int threads, const char *outPath, int every, const char *resumeFrom = nullptr) {
    Engine e(cfg, threads);
    if (resumeFrom) e.loadFromFile(resumeFrom);   // Engine::loadFromFile engine.cpp:822-825: a JSON archive (its own or one written by cityflow_b200)
    Index ix(e);
    Out o(outPath);
    const auto &lanes = e.roadnet.getLanes();
    o.i32(0x43464431);  // 'CFD1'
    o.i32((int32_t) lanes.size());
    for (int s = 0; s < steps; ++s) {
        e.nextStep();
        if ((s + 1) % every != 0 && s + 1 != steps) continue;
        std::vector<const Vehicle *> run = e.getRunningVehicles(false);
        o.i32(s + 1);
        o.i32((int32_t) e.getVehicleCount());
        o.i32((int32_t) run.size());
        o.i32((int32_t) e.vehiclePool.size());
        o.i32(e.finishedVehicleCnt);
        o.f64(e.cumulativeTravelTime);
        for (Lane *l : lanes) o.i32((int32_t) l->getVehicleCount());
        for (Lane *l : lanes) {
            int c = 0;
            for (Vehicle *v : l->getVehicles()) c += v->getSpeed() < 0.1;
            o.i32(c);
        }
        for (Lane *l : lanes) o.i32((int32_t) l->getWaitingBuffer().size());
        for (Intersection &in : e.roadnet.getIntersections()) {
            o.i32(in.isVirtualIntersection() ? -1 : in.getTrafficLight().getCurrentPhaseIndex());
        }
        for (const Vehicle *v : run) {
            int32_t f, c, lf = -1, lc = -1, bf = -1, bc = -1;
            parseId(v->getId(), f, c);
            if (v->getLeader()) parseId(v->getLeader()->getId(), lf, lc);
            if (v->getBlocker()) parseId(v->getBlocker()->getId(), bf, bc);
            o.i32(f); o.i32(c);
            o.i32(v->getPriority());
            o.i32(ix.drivable[v->getCurDrivable()]);
            o.i32(lf); o.i32(lc);
            o.i32(bf); o.i32(bc);
            o.f64(v->getDistance());
            o.f64(v->getSpeed());
            o.f64(v->getLeader() ? v->getGap() : 0.0);
            o.i64((int64_t) v->controllerInfo.enterLaneLinkTime);
        }
        // order inside every drivable (front -> back) as (flow,cnt) pairs
        for (Drivable *d : e.roadnet.getDrivables()) {
            o.i32((int32_t) d->getVehicles().size());
            for (Vehicle *v : d->getVehicles()) {
                int32_t f, c;
                parseId(v->getId(), f, c);
                o.i32(f); o.i32(c);
            }
        }
    }
    // The reference's ~Engine (engine.cpp:762-771) can fail to return when worker threads are still parked
    // on its barriers (seen with 8 threads): the dump is complete, so close it and leave without the destructor.
    o.close();
    fflush(stdout);
    _exit(0);
    return 0;
}

// The reference's own JSON archive after `steps` steps (Engine::snapshot + Archive::dump, archive.cpp:153-177).
// Two members of Vehicle::ControllerInfo have no initialiser (vehicle.h:85-86) and are dumped as they are: `gap` until the
// vehicle first has a leader, `enterLaneLinkTime` until it first changes drivable.  Neither is read in that state, but a
// `gap` that happens to be NaN / infinite makes rapidjson's writer stop (the file is cut at a 64 KiB flush) and an
// `enterLaneLinkTime` above INT_MAX makes the reference refuse its own file -- both seen here, depending on what the heap
// held.  So the harness (not the engine) gives the two members defined values first: what cityflow_b200 writes for them.
int writeArchive(const char *cfg, int steps, int threads, const char *outPath) {
    Engine e(cfg, threads);
    for (int s = 0; s < steps; ++s) e.nextStep();
    for (auto &vp : e.vehiclePool) {
        Vehicle *v = vp.second.first;
        if (!v->controllerInfo.leader) v->controllerInfo.gap = 0.0;
        if (!v->controllerInfo.prevDrivable) v->controllerInfo.enterLaneLinkTime = (size_t) std::numeric_limits<int>::max();
    }
    e.snapshot().dump(outPath);
    fflush(stdout);
    _exit(0);   // (no ~Engine, see dumpRun)
    return 0;
}

// Lane-change runs (laneChange=true): every running vehicle INCLUDING shadows, identified by
// priority (a shadow shares its parent's name until the change completes).  Meant for the build with
// the priority-ordered worker set (oracle/lc_order_patch.sh -> refdump_lcorder); the unmodified
// build produces the same record layout for the statistical comparison.
int dumpRunLC(const char *cfg, int steps, int threads, const char *outPath, int every) {
    Engine e(cfg, threads);
    Index ix(e);
    Out o(outPath);
    const auto &lanes = e.roadnet.getLanes();
    o.i32(0x43464C31);  // 'CFL1'
    o.i32((int32_t) lanes.size());
    for (int s = 0; s < steps; ++s) {
        e.nextStep();
        if ((s + 1) % every != 0 && s + 1 != steps) continue;
        std::vector<const Vehicle *> run;
        for (const auto &vp : e.vehiclePool)
            if (vp.second.first->isRunning()) run.push_back(vp.second.first);
        o.i32(s + 1);
        o.i32((int32_t) e.getVehicleCount());
        o.i32((int32_t) run.size());
        o.i32((int32_t) e.vehiclePool.size());
        o.i32(e.finishedVehicleCnt);
        o.f64(e.cumulativeTravelTime);
        for (Lane *l : lanes) o.i32((int32_t) l->getVehicleCount());
        for (Intersection &in : e.roadnet.getIntersections())
            o.i32(in.isVirtualIntersection() ? -1 : in.getTrafficLight().getCurrentPhaseIndex());
        for (const Vehicle *v : run) {
            int32_t f, c;
            parseId(v->getId(), f, c);
            const LaneChange &lc = *v->laneChange;
            o.i32(f); o.i32(c);
            o.i32(v->getPriority());
            o.i32(v->laneChangeInfo.partnerType);
            o.i32(v->getPartner() ? v->getPartner()->getPriority() : -1);
            o.i32(ix.drivable[v->getCurDrivable()]);
            o.i32(v->getLeader() ? v->getLeader()->getPriority() : -1);
            o.i32(v->getBlocker() ? v->getBlocker()->getPriority() : -1);
            o.i32((int32_t) lc.changing | ((int32_t) lc.finished << 1));
            o.i32(lc.lastDir);
            o.f64(v->getDistance());
            o.f64(v->getSpeed());
            o.f64(v->getLeader() ? v->getGap() : 0.0);
            o.f64(v->laneChangeInfo.offset);
            o.f64(lc.waitingTime);
            o.f64(lc.lastChangeTime);
        }
        for (Drivable *d : e.roadnet.getDrivables()) {  // list order (front -> back) as priorities
            o.i32((int32_t) d->getVehicles().size());
            for (Vehicle *v : d->getVehicles()) o.i32(v->getPriority());
        }
    }
    // The reference's ~Engine (engine.cpp:762-771) can fail to return when worker threads are still parked
    // on its barriers (seen with 8 threads): the dump is complete, so close it and leave without the destructor.
    o.close();
    fflush(stdout);
    _exit(0);
    return 0;
}

// Observables only (engine.cpp:615-648 as the Python getters compute them), cheap enough for 1e6 vehicles.
int dumpCounts(const char *cfg, int steps, int threads, const char *outPath, int every) {
    Engine e(cfg, threads);
    Out o(outPath);
    const auto &lanes = e.roadnet.getLanes();
    o.i32(0x43464E31);  // 'CFN1'
    o.i32((int32_t) lanes.size());
    o.i32(steps);
    o.i32(every);
    for (int s = 0; s < steps; ++s) {
        e.nextStep();
        o.i32((int32_t) e.getVehicleCount());
        if ((s + 1) % every != 0 && s + 1 != steps) continue;
        o.i32(s + 1);
        for (Lane *l : lanes) o.i32((int32_t) l->getVehicleCount());
        for (Lane *l : lanes) {
            int c = 0;
            for (Vehicle *v : l->getVehicles()) c += v->getSpeed() < 0.1;
            o.i32(c);
        }
        for (Lane *l : lanes) {
            double sum = 0;
            for (Vehicle *v : l->getVehicles()) sum += v->getSpeed();
            o.f64(sum);
        }
    }
    o.close();
    fflush(stdout);
    _exit(0);   // (no ~Engine, see dumpRun)
    return 0;
}

int sweep(const char *cfg, int steps, int warmup, const char *list) {
    std::vector<int> threads;
    for (const char *p = list; *p;) {
        threads.push_back(atoi(p));
        while (*p && *p != ',') ++p;
        if (*p == ',') ++p;
    }
    if (threads.empty()) return 64;
    char path[] = "/tmp/refdump_sweep_XXXXXX";
    int fd = mkstemp(path);
    if (fd >= 0) close(fd);
    std::string out = "{\"steps\": " + std::to_string(steps) + ", \"warmup\": " + std::to_string(warmup) + ", \"sweep\": {";
    size_t vehicles = 0;
    for (size_t k = 0; k < threads.size(); ++k) {
        Engine *e = new Engine(cfg, threads[k]);   // never destroyed (see dumpRun)
        if (k == 0) {
            for (int s = 0; s < warmup; ++s) e->nextStep();
            if (threads.size() > 1) e->snapshot().dump(path);
        } else {
            e->loadFromFile(path);
        }
        long long vs = 0;
        auto t2 = std::chrono::steady_clock::now();
        for (int s = 0; s < steps; ++s) {
            e->nextStep();
            vs += (long long) e->getVehicleCount();
        }
        auto t3 = std::chrono::steady_clock::now();
        const double sec = std::chrono::duration<double>(t3 - t2).count();
        vehicles = e->getVehicleCount();
        char buf[160];
        snprintf(buf, sizeof buf, "%s\"%d\": {\"seconds\": %.6f, \"vehicle_steps\": %lld, \"vehicle_steps_per_s\": %.3f}", k ? ", " : "", threads[k], sec, vs,
                 sec > 0 ? vs / sec : 0.0);
        out += buf;
    }
    unlink(path);
    printf("%s}, \"final_vehicles\": %zu}\n", out.c_str(), vehicles);
    fflush(stdout);
    _exit(0);
    return 0;
}

int bench(const char *cfg, int steps, int threads, int warmup, const char *countsPath) {
    auto t0 = std::chrono::steady_clock::now();
    Engine e(cfg, threads);
    auto t1 = std::chrono::steady_clock::now();
    std::vector<int32_t> counts;   // get_vehicle_count() after every step, warm-up included (parity_check of bench.py)
    counts.reserve((size_t) warmup + steps);
    for (int s = 0; s < warmup; ++s) { e.nextStep(); counts.push_back((int32_t) e.getVehicleCount()); }
    long long vs = 0;
    auto t2 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; ++s) {
        e.nextStep();
        const long long c = (long long) e.getVehicleCount();
        vs += c;
        counts.push_back((int32_t) c);
    }
    auto t3 = std::chrono::steady_clock::now();
    if (countsPath) {
        Out o(countsPath);
        for (int32_t c : counts) o.i32(c);
        o.close();
    }
    double load = std::chrono::duration<double>(t1 - t0).count();
    double sec = std::chrono::duration<double>(t3 - t2).count();
    printf("{\"steps\": %d, \"warmup\": %d, \"threads\": %d, \"load_s\": %.6f, \"seconds\": %.6f, "
           "\"vehicle_steps\": %lld, \"final_vehicles\": %zu, \"vehicle_steps_per_s\": %.3f}\n",
           steps, warmup, threads, load, sec, vs, e.getVehicleCount(), sec > 0 ? vs / sec : 0.0);
    // The reference's ~Engine (engine.cpp:762-771) occasionally never returns with many worker
    // threads (observed: 8 threads, 30x30); the measurement is complete, so leave without it.
    fflush(stdout);
    _exit(0);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "static")) return dumpStatic(argv[2], argv[3]);
    if (argc >= 6 && !strcmp(argv[1], "run"))
        return dumpRun(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5], argc >= 7 ? atoi(argv[6]) : 1);
    if (argc >= 6 && !strcmp(argv[1], "archive")) return writeArchive(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
    if (argc >= 7 && !strcmp(argv[1], "resume"))
        return dumpRun(argv[2], atoi(argv[4]), atoi(argv[5]), argv[6], argc >= 8 ? atoi(argv[7]) : 1, argv[3]);
    if (argc >= 6 && !strcmp(argv[1], "runlc"))
        return dumpRunLC(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5], argc >= 7 ? atoi(argv[6]) : 1);
    if (argc >= 6 && !strcmp(argv[1], "counts"))
        return dumpCounts(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5], argc >= 7 ? atoi(argv[6]) : 1);
    if (argc >= 5 && !strcmp(argv[1], "bench"))
        return bench(argv[2], atoi(argv[3]), atoi(argv[4]), argc >= 6 ? atoi(argv[5]) : 0, argc >= 7 ? argv[6] : nullptr);
    if (argc >= 6 && !strcmp(argv[1], "sweep")) return sweep(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
    fprintf(stderr, "usage: refdump static|run|runlc|counts|sweep|bench|archive|resume ...\n");
    return 64;
}

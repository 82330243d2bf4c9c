/* cityflow_b200_lc_draft.h -- DRAFT, only in libraries built with `make EXTRA=-DCFB_LANE_CHANGE`.
 *
 * With that build "laneChange": true is accepted (Engine::scheduleLaneChange engine.cpp:792-809,
 * lanechange.cpp, engine.cpp:195-244 on the device: cityflow_b200/csrc/device_lc.cuh) and this one
 * extra call exists for the parity tests.  The path has not run on a GPU yet; the default library
 * neither contains it nor declares this symbol. */
#pragma once
#include "cityflow_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* every running vehicle INCLUDING shadows in vehiclePool (priority) order; partner / leader / blocker
 * are given as priorities (-1 none).  Same layout as oracle/harness.py LC_DTYPE (refdump runlc). */
typedef struct {
    int32_t flow, cnt, priority, partner_type, partner, drivable, leader, blocker, flags, last_dir;
    double dis, speed, gap, offset, waiting_time, last_change_time;
} cfb_lc_vehicle;
int64_t cfb_debug_lc_vehicles(cfb_engine *e, cfb_lc_vehicle *out, int64_t cap);
#ifdef __cplusplus
}
#endif

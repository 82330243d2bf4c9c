#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Applied by oracle/Makefile to a THROW-AWAY copy of the reference's
# src/ tree (under mktemp, deleted after the build) to build oracle/_ref/refdump_lcorder.
#
# Why: with laneChange=true the reference walks each worker's vehicles in the iteration order of a
# std::set<Vehicle *> (engine.h:27), i.e. by HEAP ADDRESS.  That order decides which lane-change
# candidate is scheduled first and which shadow draws which priority, so the
# unmodified engine's trajectories depend on the allocator -- no restatement can be pinned to them.
# This patch defines the two things the unmodified engine leaves to the allocator, and nothing else:
#  1. that set is ordered by vehicle priority (unique per live vehicle, engine.cpp:601) instead of
#     by address;
#  2. ControllerInfo::gap (vehicle.h:86) and LaneChange::lastDir (lanechange.h:26) start at 0.  The
#     reference never initialises them; SimpleLaneChange::makeSignal (lanechange.cpp:162-165) reads
#     `gap` of a vehicle that has not had a leader yet, i.e. whatever the heap block held before.
# With laneChange=false neither matters (tests check that the patched and the unmodified build dump
# identical states there).
# vehicleRemoveBuffer (engine.h:40) stays address-ordered: it is only ever searched, and it is
# searched with pointers to vehicles that may already be deleted (engine.cpp:419).
set -e
SRC="$1"
sed -i \
    -e 's/std::set<Vehicle \*> vehicleRemoveBuffer/STD_SET_KEEP vehicleRemoveBuffer/' \
    -e 's/std::set<Vehicle \*>/VehicleSet/g' \
    -e 's/STD_SET_KEEP/std::set<Vehicle *>/' \
    -e 's/^    class Engine {$/    struct VehiclePriorityLess { bool operator()(const Vehicle *a, const Vehicle *b) const; };\n    typedef std::set<Vehicle *, VehiclePriorityLess> VehicleSet;\n\n    class Engine {/' \
    "$SRC/engine/engine.h"
sed -i \
    -e 's/std::set<CityFlow::Vehicle \*>/VehicleSet/g' \
    -e 's/std::set<Vehicle \*>/VehicleSet/g' \
    -e '0,/^namespace CityFlow {$/s//namespace CityFlow {\n    bool VehiclePriorityLess::operator()(const Vehicle *a, const Vehicle *b) const { return a->getPriority() < b->getPriority(); }/' \
    "$SRC/engine/engine.cpp"
sed -i -e 's/^            double gap;$/            double gap = 0;/' "$SRC/vehicle/vehicle.h"
sed -i -e 's/^        int lastDir;$/        int lastDir = 0;/' "$SRC/vehicle/lanechange.h"
grep -q "double gap = 0;" "$SRC/vehicle/vehicle.h"
grep -q "int lastDir = 0;" "$SRC/vehicle/lanechange.h"
grep -q "typedef std::set<Vehicle \*, VehiclePriorityLess> VehicleSet;" "$SRC/engine/engine.h"
grep -q "VehiclePriorityLess::operator()" "$SRC/engine/engine.cpp"
! grep -q "std::set<Vehicle \*> &" "$SRC/engine/engine.h"

"""Archives: snapshot / load in memory, and the two file forms of Archive.dump -- "x.json" is the reference's JSON
schema (a file the reference's load_from_file reads, and the other way round), any other name this engine's exact
binary image."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cityflow  # noqa: E402
from cityflow_b200 import scenario  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    cfg = scenario.make_grid_scenario(d, 4, 4)
    eng = cityflow.Engine(cfg, thread_num=1)
    for _ in range(200):
        eng.next_step()
    arc = eng.snapshot()                               # device-resident copy of the state
    for _ in range(100):
        eng.next_step()
    record = (eng.get_lane_vehicle_count(), eng.get_average_travel_time())
    eng.load(arc)                                      # back to step 200
    for _ in range(100):
        eng.next_step()
    assert record == (eng.get_lane_vehicle_count(), eng.get_average_travel_time())

    arc.dump(os.path.join(d, "save.bin"))              # exact binary image
    arc.dump(os.path.join(d, "save.json"))             # the reference's schema
    doc = json.load(open(os.path.join(d, "save.json")))
    print("save.json: step %d, %d vehicles (%d running), keys of a vehicle: %s ..." % (
        doc["step"], len(doc["vehicles"]), doc["activeVehicleCount"], ", ".join(sorted(doc["vehicles"][0])[:6])))
    print("sizes: save.bin %d KB, save.json %d KB" % (os.path.getsize(os.path.join(d, "save.bin")) // 1024,
                                                      os.path.getsize(os.path.join(d, "save.json")) // 1024))
    other = cityflow.Engine(cfg, thread_num=1)
    other.load_from_file(os.path.join(d, "save.json"))  # would equally read a file written by the reference
    print("loaded: t = %.0f s, %d vehicles running" % (other.get_current_time(), other.get_vehicle_count()))

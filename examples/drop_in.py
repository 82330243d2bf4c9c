"""A traffic-signal script written for the reference runs unchanged: only `import cityflow`
resolves to this repository (cityflow/__init__.py -> cityflow_b200)."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cityflow  # noqa: E402
from cityflow_b200 import scenario  # noqa: E402  (only to write a synthetic grid scenario)

with tempfile.TemporaryDirectory() as d:
    cfg = scenario.make_grid_scenario(d, 6, 6, dense=dict(frac=1.0, interval=5.0, seed=1), rl_traffic_light=True)
    eng = cityflow.Engine(cfg, thread_num=4)            # thread_num is accepted and ignored
    roadnet = json.load(open(os.path.join(d, "roadnet.json")))
    lights = [i for i in roadnet["intersections"] if not i["virtual"]]
    for t in range(600):
        if t % 30 == 0:
            for i in lights:
                eng.set_tl_phase(i["id"], (t // 30) % len(i["trafficLight"]["lightphases"]))
        eng.next_step()
    counts = eng.get_lane_vehicle_count()
    waiting = eng.get_lane_waiting_vehicle_count()
    speeds = eng.get_vehicle_speed()
    print("t=%.0fs  vehicles=%d  waiting=%d  mean speed=%.2f m/s  average travel time=%.1f s" % (
        eng.get_current_time(), eng.get_vehicle_count(), sum(waiting.values()),
        sum(speeds.values()) / max(1, len(speeds)), eng.get_average_travel_time()))
    busiest = max(counts, key=counts.get)
    print("busiest lane:", busiest, counts[busiest], "vehicles:", eng.get_lane_vehicles()[busiest][:5])

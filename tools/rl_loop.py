"""BASELINE.json configs[4] shape: the traffic-signal RL inner loop on a 6x6 grid --
per step: set_tl_phase on all 36 signalised intersections, next_step, get_lane_vehicle_count,
get_lane_waiting_vehicle_count.  Prints env-steps/s and vehicle-steps/s for this engine and, when
the compiled reference module travelled (oracle/_ref/cityflow*.so), for the reference."""
import glob
import importlib.machinery
import importlib.util
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cityflow_b200 import scenario  # noqa: E402


def load_reference_module():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cityflow*.so"))
    if not so:
        return None
    loader = importlib.machinery.ExtensionFileLoader("cityflow", so[0])
    spec = importlib.util.spec_from_loader("cityflow", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def run(mod, cfg, inters, steps, threads=1):
    eng = mod.Engine(cfg, thread_num=threads)
    for _ in range(100):
        eng.next_step()
    vs = 0
    t0 = time.perf_counter()
    for t in range(steps):
        ph = (t // 30) % 8
        for i in inters:
            eng.set_tl_phase(i, ph)
        eng.next_step()
        c = eng.get_lane_vehicle_count()
        w = eng.get_lane_waiting_vehicle_count()
        vs += eng.get_vehicle_count()
    sec = time.perf_counter() - t0
    return {"env_steps_per_s": steps / sec, "vehicle_steps_per_s": vs / sec, "mean_vehicles": vs / steps,
            "lanes": len(c), "waiting_sum_last": sum(w.values())}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    d = tempfile.mkdtemp()
    cfg = scenario.make_grid_scenario(d, 6, 6, dense=dict(frac=1.0, interval=5.0, seed=2), name="rl", rl_traffic_light=True)
    net = json.load(open(os.path.join(d, "roadnet_rl.json")))
    inters = [i["id"] for i in net["intersections"] if not i["virtual"]]
    import cityflow as ours
    out = {"ours": run(ours, cfg, inters, steps)}
    ref = load_reference_module()
    if ref is not None:
        out["reference_1_thread"] = run(ref, cfg, inters, steps, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

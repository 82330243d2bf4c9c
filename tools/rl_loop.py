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


def run_device(cfg, steps):
    """Same loop with actions and observations staying on the GPU (cityflow_b200 extras):
    a stand-in policy (arg-max over per-intersection waiting counts, a few torch ops) picks the
    phases from the observation tensors; nothing is copied to the host until the loop ends."""
    import torch
    import cityflow
    import cityflow_b200
    eng = cityflow.Engine(cfg, thread_num=1)
    for _ in range(100):
        eng.next_step()
    n_int = eng.num_intersections()
    obs = cityflow_b200.LaneObservations(eng)
    cnt, wait = obs.vehicle_count, obs.waiting_count
    n_lanes = len(obs.lane_ids)
    bucket = (torch.arange(n_lanes, device=cnt.device) % 8).to(torch.int64)          # lane -> phase it "votes" for
    owner = (torch.arange(n_lanes, device=cnt.device) * n_int // n_lanes).to(torch.int64)
    votes = torch.zeros(n_int * 8, device=cnt.device, dtype=torch.int32)
    key = owner * 8 + bucket
    vs = torch.zeros((), device=cnt.device, dtype=torch.int64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(steps):
        obs.refresh()
        votes.zero_()
        votes.index_add_(0, key, wait)
        act = votes.view(n_int, 8).argmax(1).to(torch.int32)
        cityflow_b200.set_tl_phases_tensor(eng, act)
        eng.next_step()
        vs += cnt.sum()
    torch.cuda.synchronize()
    eng.synchronize()
    sec = time.perf_counter() - t0
    return {"env_steps_per_s": steps / sec, "vehicle_steps_per_s": int(vs) / sec, "mean_vehicles": int(vs) / steps,
            "lanes": n_lanes, "policy": "torch index_add + argmax on the observation tensors, actions via set_tl_phases_tensor"}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    d = tempfile.mkdtemp()
    cfg = scenario.make_grid_scenario(d, 6, 6, dense=dict(frac=1.0, interval=5.0, seed=2), name="rl", rl_traffic_light=True)
    net = json.load(open(os.path.join(d, "roadnet_rl.json")))
    inters = [i["id"] for i in net["intersections"] if not i["virtual"]]
    import cityflow as ours
    out = {"ours": run(ours, cfg, inters, steps)}
    try:
        out["ours_device_resident"] = run_device(cfg, steps)
    except Exception as e:  # noqa: BLE001  (report, the host-API numbers above still stand)
        out["ours_device_resident"] = {"error": repr(e)}
    ref = load_reference_module()
    if ref is not None:
        out["reference_1_thread"] = run(ref, cfg, inters, steps, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

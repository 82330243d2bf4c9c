"""The RL inner loop with nothing crossing to the host: observations are torch CUDA tensors over
the engine's own buffers, actions go back as one int32 tensor."""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cityflow  # noqa: E402
import cityflow_b200  # noqa: E402
from cityflow_b200 import scenario  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    cfg = scenario.make_grid_scenario(d, 10, 10, dense=dict(frac=1.0, interval=5.0, seed=1), rl_traffic_light=True)
    eng = cityflow.Engine(cfg)
    obs = cityflow_b200.LaneObservations(eng)                     # tensors are created once
    n_int, n_lanes = eng.num_intersections(), len(obs.lane_ids)
    policy = torch.nn.Sequential(torch.nn.Linear(2 * n_lanes, 256), torch.nn.ReLU(), torch.nn.Linear(256, n_int * 8)).cuda()
    steps = 500
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(steps):
            obs.refresh()                                           # k_lane_obs, ordered against torch's stream
            x = torch.cat([obs.vehicle_count, obs.waiting_count]).float()
            act = policy(x).view(n_int, 8).argmax(1).to(torch.int32)
            cityflow_b200.set_tl_phases_tensor(eng, act)            # k_set_phases reads `act` on the engine's stream
            eng.next_step()
    torch.cuda.synchronize()
    eng.synchronize()
    dt = time.perf_counter() - t0
    print("%d env steps in %.2f s (%.0f steps/s), %d vehicles, mean lane speed %.2f m/s" % (
        steps, dt, steps / dt, eng.get_vehicle_count(),
        float((obs.refresh().speed_sum.sum() / obs.vehicle_count.sum().clamp(min=1)))))

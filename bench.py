#!/usr/bin/env python
"""bench.py -- vehicle-steps/second of the CityFlow step engine on B200 (BASELINE.json metric).

A "step" is one Engine::nextStep over the whole road network.  Workload at N=1 (and per rank at
N>1, weak scaling -- independent replicas, the path has no exchange step in this round): the
30x30 grid of BASELINE.json configs[2] with the dense random-walk demand of SURVEY.md
Appendix A (frac=0.5, interval=10 s, ~1e5 concurrent vehicles).  Both arms first advance the
simulation `--prefill` steps (untimed scenario preparation: the network starts empty) so that the
timed steps run at the ~1e5-vehicle operating point.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA engine
  python bench.py --impl reference [...]                         # reference CPU engine, host cores

Prints ONE JSON line on rank 0 (see the task contract): value / e2e / roofline / cpu_baseline /
clocks / gpu_launches.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vehicle_steps_per_sec"
UNIT = "vehicle-steps/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=1000)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--prefill", type=int, default=1200, help="untimed steps that fill the network before timing")
    p.add_argument("--rows", type=int, default=30)
    p.add_argument("--cols", type=int, default=30)
    p.add_argument("--frac", type=float, default=0.5)
    p.add_argument("--flow-interval", type=float, default=10.0)
    p.add_argument("--threads", type=int, default=0, help="reference arm: thread_num (default nproc)")
    p.add_argument("--cpu-steps", type=int, default=100, help="cpu_baseline sample: timed steps after the prefill")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--multi", default="sharded", choices=["sharded", "replicas"],
                   help="N>1: 'sharded' = ONE simulation of a rows x (cols*N) grid cut into N column strips with NCCL seam "
                        "exchange (weak scaling); 'replicas' = N independent copies of the N=1 workload")
    p.add_argument("--clock-ms", type=int, default=200, help="nvidia-smi sampling period (0 = off)")
    p.add_argument("--profile-steps", type=int, default=0,
                   help="ncu mode: after prefill+warmup run this many steps between cudaProfilerStart/Stop and exit "
                        "(use with ncu --profile-from-start off)")
    return p.parse_args()


def grid_cols(args):
    """Weak scaling of the sharded engine: one strip of `cols` columns per GPU."""
    n = max(args.gpus, 1)
    return args.cols * n if (n > 1 and args.multi == "sharded") else args.cols


def make_scenario(args, directory):
    from cityflow_b200 import scenario
    return scenario.make_grid_scenario(
        directory, args.rows, grid_cols(args), name="bench",
        dense=dict(frac=args.frac, interval=args.flow_interval, seed=1))


def workload_name(args):
    return "%dx%d grid (tools/generator layout), random-walk flows frac=%g interval=%gs seed=1, interval=1.0s, seed=0" % (
        args.rows, grid_cols(args), args.frac, args.flow_interval)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0, period_ms=50):
        self.device = device
        self.period_ms = period_ms
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak_gbs():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------
def reference_bench(cfg, steps, warmup_total, threads):
    """The reference's own CPU engine (unmodified sources compiled into oracle/_ref) or, when that
    build did not travel, the CPU restatement (single thread)."""
    from oracle import harness as H
    if H.have_ref():
        r = H.RefDump.bench(cfg, steps, threads, warmup_total)
        return dict(kind="reference", cores=threads, value=r["vehicle_steps_per_s"], seconds=r["seconds"],
                    vehicle_steps=r["vehicle_steps"], final_vehicles=r["final_vehicles"])
    if not H.have_port():
        H.build(ref=False)
    o = H.PortOracle(cfg)
    o.next_step(warmup_total)
    vs = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        o.next_step()
        vs += o.vehicle_count()
    sec = time.perf_counter() - t0
    return dict(kind="port", cores=1, value=vs / sec, seconds=sec, vehicle_steps=vs, final_vehicles=o.vehicle_count())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.threads or (os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as d:
        cfg = make_scenario(args, d)
        r = reference_bench(cfg, args.steps, args.prefill + args.warmup, threads)
    sample = "%d timed steps after %d untimed steps, thread_num=%d" % (args.steps, args.prefill + args.warmup, r["cores"])
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "prefill_steps": args.prefill,
                   "mean_vehicles": r["vehicle_steps"] / max(args.steps, 1), "host_cores": os.cpu_count(),
                   "note": "reference CPU engine runs once on rank 0 regardless of --gpus"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from cityflow_b200.distutil import Ranks
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    ranks = Ranks("nccl", torch.device("cuda", local))
    world, rank, dist = ranks.world, ranks.rank, ranks.dist

    import cityflow  # our drop-in module (repo root)
    import cityflow_b200
    sharded = world > 1 and args.multi == "sharded"
    tmp = tempfile.TemporaryDirectory()
    cfg = make_scenario(args, tmp.name)
    t0 = time.perf_counter()
    if sharded:
        ids = [cityflow_b200.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng = cityflow.Engine(cfg, thread_num=1, device=local, shard_rank=rank, shard_world=world, nccl_id=ids[0])
    else:
        eng = cityflow.Engine(cfg, thread_num=1, device=local)
    load_s = time.perf_counter() - t0

    def barrier():
        ranks.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    reduce_max, reduce_sum = ranks.max, ranks.sum

    # ---- scenario preparation + warm-up (clock sampling starts here: the timed regions are sub-second) ----
    sampler = ClockSampler(local, args.clock_ms)
    if rank == 0 and args.clock_ms > 0:
        sampler.start()
    eng.next_steps(args.prefill)
    eng.synchronize()
    eng.timed_steps(max(args.warmup, 3), True)
    n_start = eng.get_vehicle_count()
    if args.profile_steps > 0:  # profiling run: numbers printed under a profiler are never bench values
        cudart = torch.cuda.cudart()
        cudart.cudaProfilerStart()
        eng.timed_steps(args.profile_steps, True)
        cudart.cudaProfilerStop()
        if rank == 0:
            sampler.stop()
            print(json.dumps({"profiled_steps": args.profile_steps, "vehicles": n_start}))
        return

    # ---- value: K steps, device time (CUDA events on the engine stream), L2 flushed between steps ----
    barrier()
    launches0 = eng.gpu_launches()
    ms_flush, vs_flush = eng.timed_steps(args.steps, True)
    barrier()
    launches = eng.gpu_launches() - launches0
    ms_flush_max = reduce_max(ms_flush)
    vs_total = reduce_sum(vs_flush)
    value = vs_total / (ms_flush_max / 1e3)

    # same, back to back without flush (state stays L2 resident, as in real stepping)
    barrier()
    ms_warm, vs_warm = eng.timed_steps(args.steps, False)
    barrier()
    value_warm = reduce_sum(vs_warm) / (reduce_max(ms_warm) / 1e3)

    # ---- e2e: the user-visible loop through the public API, host clock, H2D + D2H every step ----
    h2d0, d2h0 = eng.transfer_bytes()
    barrier()
    t0 = time.perf_counter()
    acc = 0
    for _ in range(args.steps):
        eng.next_step()
        acc += eng.get_vehicle_count()
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d1, d2h1 = eng.transfer_bytes()
    host_gen_ms, host_enq_ms = eng.host_times()
    # sharded: get_vehicle_count() already is the network-wide count on every rank
    e2e_value = (acc if sharded else reduce_sum(acc)) / reduce_max(e2e_s)
    clocks = sampler.stop() if rank == 0 else {}

    eng_steps_total = args.prefill + max(args.warmup, 3) + 3 * args.steps
    # ---- per-kernel device time (CUDA events around every kernel; separate pass, serialised) ----
    if sharded:   # per-kernel events are only wired for the single-engine launch path
        (k_ing, k_not, k_ctl, k_mov, k_led), kn = (0.0, 0.0, 0.0, 0.0, 0.0), 1
    else:
        eng.enable_kernel_timing(True)
        ksteps = min(args.steps, 50)
        for _ in range(ksteps):
            eng.next_step()
        (k_ing, k_not, k_ctl, k_mov, k_led), kn = eng.kernel_times()
        eng.enable_kernel_timing(False)
    n_now = eng.get_vehicle_count()
    n_drv = eng.num_drivables() if hasattr(eng, "num_drivables") else 0
    kms = {"k_ingest": k_ing / kn, "k_notify": k_not / kn, "k_control": k_ctl / kn, "k_move": k_mov / kn, "k_leader": k_led / kn}
    dominant = max(kms, key=kms.get)
    peak, peak_src = measured_peak_gbs()
    # algorithmic bytes per launch: per running vehicle N, per drivable D
    alg = {
        "k_leader": 32 * n_now + 12 * n_drv,     # leader-scan kernel
        "k_control": 96 * n_now + 4 * n_drv,
        "k_move": 88 * n_now + 8 * n_drv,
        "k_notify": 16 * n_now + 24 * (n_drv),
        "k_ingest": 12 * n_drv,
    }

    try:  # DRAM bytes per launch from the committed ncu --set full capture (same workload)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01f_dram_traffic_bytes_per_launch.json")))
    except Exception:
        traffic = {}
    same_workload = (args.rows, args.cols, args.frac, args.flow_interval) == (30, 30, 0.5, 10.0)

    def roof(k):
        gbs = alg[k] / (kms[k] * 1e-3) / 1e9 if kms[k] > 0 else 0.0
        return {"kernel": k, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": traffic.get(k) if same_workload else None,
                "note": "latency-bound at this size (dependent loads + FP64 div/sqrt chains), see profiles/r01c_control_cycles.md", "algorithmic_bytes_per_launch": alg[k], "avg_launch_ms": kms[k], "peak_source": peak_src}

    line = None
    if rank == 0:
        mean_n = vs_total / max(args.steps, 1) / world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_flush_max / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args), "prefill_steps": args.prefill, "mean_vehicles_per_gpu": mean_n,
                "vehicles_at_start": n_start, "parallelism": ("sharded: %d column strips of ONE simulation, seam records over NCCL (2 send/recv groups per step: movers, tails + blocker lists)" % world)
                if sharded else "replicas x%d (one engine per GPU, no exchange)" % world,
                "l2": "value: 256 MiB memset between timed steps (state ~30 MB would otherwise stay L2-resident); "
                      "value_l2_warm and e2e: steps back to back as in real stepping",
                "load_seconds": load_s,
            },
            "value_l2_warm": value_warm,
            "ms_per_step_l2_warm": reduce_max(ms_warm) / max(args.steps, 1) if dist is None else None,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": (h2d1 - h2d0) / max(args.steps, 1),
                    "d2h_bytes_per_step": (d2h1 - d2h0) / max(args.steps, 1),
                    "ms_per_step": 1e3 * e2e_s / max(args.steps, 1),
                    "api": "cityflow.Engine.next_step() + get_vehicle_count() every step"},
            "gpu_launches": int(launches),
            "kernel_ms": kms,
            "host_ms_per_step": {"spawn_generation": host_gen_ms / max(eng_steps_total, 1), "enqueue": host_enq_ms / max(eng_steps_total, 1)},
            "roofline": roof(dominant) if not sharded else None,
            "roofline_leader_scan": roof("k_leader") if not sharded else None,
            "clocks": clocks,
        }
    if dist is not None:
        dist.barrier()
    # ---- cpu_baseline: the reference's CPU engine on this box's host cores (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = args.threads or (os.cpu_count() or 1)
        try:
            r = reference_bench(cfg, args.cpu_steps, args.prefill, threads)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                "sample": "%d timed steps after %d untimed steps of the same scenario (mean %.0f vehicles)" % (
                    args.cpu_steps, args.prefill, r["vehicle_steps"] / max(args.cpu_steps, 1))}
        except Exception as ex:  # keep the GPU numbers even if the CPU arm fails
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "failed: %r" % (ex,)}
    if rank == 0:
        print(json.dumps(line))
    del eng
    tmp.cleanup()
    ranks.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

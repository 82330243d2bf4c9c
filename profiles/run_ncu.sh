#!/bin/bash
# Profiling recipe (B200_PROFILING.md), run under gpurun from the repo root.  bench.py
# --profile-steps brackets the steps of interest with cudaProfilerStart/Stop so the ~6000
# prefill launches are not intercepted (ncu --profile-from-start off).
#   1. launch list with per-launch device time (cold-cache, serialised: compare SHARES)
#   2. one --set full capture of the five step kernels at the ~1e5-vehicle operating point
# Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
PRE=${PRE:-1200}
timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv \
    python bench.py --profile-steps 20 --warmup 3 --prefill $PRE > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_(ingest|notify|control|move|leader)' -c 10 -f -o gpurun_out/prof_step \
    python bench.py --profile-steps 2 --warmup 3 --prefill $PRE > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out

#!/bin/bash
# GPU box: the two ncu passes behind profiles/ (run only after the plain bench exited 0).
#   1. launch list of the bench command (per-launch durations, cold-cache and serialised: shares, not absolutes)
#   2. one --set full capture of the fused kernel (16 384 frames of config 3) with source correlation
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:features_kernel --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/prof_features python bench.py --frames 16384 --steps 2 --warmup 3 --no-cpu-baseline --direct-frames 0 \
    > gpurun_out/prof_ncu2.log 2>&1
for f in gpurun_out/prof_ncu1.log gpurun_out/prof_ncu2.log; do tail -n 3 $f; done

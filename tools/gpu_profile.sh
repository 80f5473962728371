#!/bin/bash
# GPU box: the ncu passes behind profiles/ (run only after the plain bench exited 0).
#   1. launch list of the bench command (per-launch durations, cold-cache and serialised: shares, not absolutes)
#   2. --set full captures, one launch each, with source correlation:
#        features_kernel  cfg3 (16 384 frames)      stage_frames_kernel (K0) cfg3
#        allpairs_kernel  cfg4 full / contact map   features_kernel cfg5 (4 096 frames)
#        cov_gemm_kernel  n = 14 000, K = 4 096
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-all-configs > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-all-configs > gpurun_out/prof_ncu1.log 2>&1
NCU="ncu --set full --import-source on --clock-control none --launch-count 1 -f"
$NCU -k regex:features_kernel --launch-skip 3 -o gpurun_out/prof_features python tools/gpu_probe.py 3 16384 > gpurun_out/prof_ncu2.log 2>&1
$NCU -k regex:stage_frames_kernel --launch-skip 2 -o gpurun_out/prof_stage python tools/gpu_probe.py 3 16384 > gpurun_out/prof_ncu3.log 2>&1
# config 4 runs the all-pairs flavour: two launches per run (fast tiles, general tiles), both captured
NCU2="ncu --set full --import-source on --clock-control none --launch-count 2 -f"
$NCU2 -k regex:allpairs_kernel --launch-skip 2 -o gpurun_out/prof_cfg4_full python tools/gpu_cfg4.py 2048 > gpurun_out/prof_ncu4.log 2>&1
$NCU2 -k regex:allpairs_kernel --launch-skip 44 -o gpurun_out/prof_cfg4_lean python tools/gpu_cfg4.py 2048 > gpurun_out/prof_ncu5.log 2>&1
$NCU -k regex:features_kernel --launch-skip 3 -o gpurun_out/prof_cfg5 python tools/gpu_probe.py 5 4096 > gpurun_out/prof_ncu6.log 2>&1
$NCU -k regex:cov_gemm_kernel --launch-skip 1 -o gpurun_out/prof_cov python tools/gpu_cov.py 14000 4096 > gpurun_out/prof_ncu7.log 2>&1
for f in gpurun_out/prof_ncu*.log; do tail -n 2 $f; done
ls -la gpurun_out/*.ncu-rep

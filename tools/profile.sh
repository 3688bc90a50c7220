#!/bin/bash
# Round profile run on the GPU box (gpurun): bench line, launch list, one full capture of the headline kernel.
# Every ncu pass runs only after the same command exited 0 without ncu.
set -e
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_short.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse_step|fine_step|coarse_adjoint|dataset_pass" \
    -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:coarse_step_staged -s 3 -c 1 -f -o gpurun_out/step_update \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_full.log 2>&1

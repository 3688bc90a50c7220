#!/bin/bash
# Round profile run on the GPU box (gpurun): bench line, launch list, full captures of the headline kernel, the adjoint
# and the fine endpoint kernel.  Every ncu pass runs only after the same command exited 0 without ncu.
set -e
R=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu --no-legs > gpurun_out/${R}_bench_short.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse_step|fine_step|coarse_adjoint|dataset_pass|ks1d|ks2d|heat1d" \
    -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-legs > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > /dev/null
ncu --set full --clock-control none --import-source on -k regex:coarse_step_staged -s 3 -c 1 -f -o gpurun_out/${R}_step_update \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_full.log 2>&1
python tools/run_adjoint.py > gpurun_out/${R}_adjoint_times.txt
ncu --set full --clock-control none --import-source on -k regex:coarse_adjoint -s 2 -c 1 -f -o gpurun_out/${R}_adjoint \
    python tools/run_adjoint.py > gpurun_out/ncu_adj.log 2>&1
python tools/run_fine.py 1025 64 as_written pos > gpurun_out/${R}_fine_times.txt
ncu --set full --clock-control none --import-source on -k regex:fine_step_ep -s 5 -c 1 -f -o gpurun_out/${R}_fine1025 \
    python tools/run_fine.py 1025 64 as_written pos > gpurun_out/ncu_fine.log 2>&1

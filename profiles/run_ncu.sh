#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/:
#   launches.csv      every kernel launch of a short bench.py run with its device time
#   prof_*.ncu-rep    --set full captures of the hot kernels (read here with ncu -i ... --page raw)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_|dispatch_' -c 4 \
    -f -o gpurun_out/prof_pi python profiles/prof_target.py pi 2 > gpurun_out/prof_pi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_|dispatch_' -c 4 \
    -f -o gpurun_out/prof_payload python profiles/prof_target.py payload 1 > gpurun_out/prof_payload.log 2>&1
ls -la gpurun_out

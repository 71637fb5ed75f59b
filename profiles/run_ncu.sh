#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/:
#   launches.csv      every kernel launch of a short bench.py run with its device time
#   prof_*.ncu-rep    --set full captures of the hot kernels (read here with ncu -i ... --page raw)
#   src_sha256.txt    digest of the sources the captures were taken from (bench.py compares it with the sources it runs)
set -x
mkdir -p gpurun_out
# the build is identified by its SOURCES (nvcc output is not bit-reproducible): same digest as bench.py's source_digest()
cat $(ls fiber_b200/csrc/*.cu fiber_b200/csrc/*.cuh include/*.h include/*.cuh | sort) | sha256sum | cut -d' ' -f1 > gpurun_out/src_sha256.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-parzen > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_|dispatch_' -c 4 \
    -f -o gpurun_out/prof_pi python profiles/prof_target.py pi 2 > gpurun_out/prof_pi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_|dispatch_' -c 6 \
    -f -o gpurun_out/prof_payload python profiles/prof_target.py payload 1 > gpurun_out/prof_payload.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dispatch_parzen' -s 1 -c 2 \
    -f -o gpurun_out/prof_parzen python profiles/prof_target.py parzen 2 > gpurun_out/prof_parzen.log 2>&1
ls -la gpurun_out

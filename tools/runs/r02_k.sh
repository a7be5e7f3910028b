#!/bin/bash
# round 2, call k: ncu capture of the faithful mode on the event-list engine (engine3.cuh), C4
mkdir -p gpurun_out
timeout 300 python bench.py --mode faithful --photons 2e5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transport4 -s 1 -c 1 -o gpurun_out/r02_k_faithful_c4 \
    python bench.py --mode faithful --photons 2e5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_k_ncu.log 2>&1
tail -2 gpurun_out/r02_k_ncu.log

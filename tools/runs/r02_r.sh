#!/bin/bash
# round 2, call r: ncu capture of the multi-detector walk (C2, 68 detectors x 5e5 packets)
mkdir -p gpurun_out
timeout 300 python bench.py --workload c2 --multi 68 --photons 5e5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_r_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transport3 -s 1 -c 1 -o gpurun_out/r02_r_multi_c2 \
    python bench.py --workload c2 --multi 68 --photons 5e5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_r_ncu.log 2>&1
tail -2 gpurun_out/r02_r_ncu.log

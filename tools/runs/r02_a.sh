#!/bin/bash
# round 2, call a: GPU test suite, C4 / C5 bench lines, launch list + full capture of the C5 kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_a_gpus.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_a_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_a_bench_c4.json 2> gpurun_out/r02_a_bench_c4.err
python bench.py --workload c5 --photons 2e6 --steps 2 --warmup 2 > gpurun_out/r02_a_bench_c5.json 2> gpurun_out/r02_a_bench_c5.err
python bench.py --workload c5 --photons 1e6 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transport3 -s 1 -c 1 -o gpurun_out/r02_a_c5 \
    python bench.py --workload c5 --photons 1e6 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_a_ncu.log 2>&1
tail -3 gpurun_out/r02_a_pytest.log
cat gpurun_out/r02_a_bench_c4.json | head -c 1500

#!/bin/bash
# round 2, call d: full GPU suite, bench lines after the marcher / watchdog changes, launch list + full capture of the C4 kernel
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02_d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_d_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_d_bench_c4.json 2> gpurun_out/r02_d_bench_c4.err
timeout 300 python bench.py --workload c5 --photons 2e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_d_bench_c5.json 2> gpurun_out/r02_d_bench_c5.err
for w in c1 c2 c3; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_d_bench_$w.json 2> gpurun_out/r02_d_bench_$w.err
done
timeout 300 python bench.py --workload c2 --multi 68 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_d_bench_c2_multi68.json 2> gpurun_out/r02_d_bench_c2_multi68.err
timeout 300 python bench.py --steps 2 --warmup 1 --photons 4e6 --no-cpu-baseline > gpurun_out/r02_d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transport3 -s 1 -c 1 -o gpurun_out/r02_d_c4 \
    python bench.py --steps 2 --warmup 1 --photons 4e6 --no-cpu-baseline > gpurun_out/r02_d_ncu.log 2>&1
tail -5 gpurun_out/r02_d_pytest.log
for f in gpurun_out/r02_d_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['shard_check']['ok'])
except Exception as e: print('ERR', e)"; done

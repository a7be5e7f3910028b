#!/bin/bash
# round 2, third session, last call: GPU suite and the bench lines that the last two changes touch (3-D pass length, hot launch index)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02_zz2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_zz2_pytest.log
tail -4 gpurun_out/r02_zz2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_zz2_smoke.log 2>&1; tail -1 gpurun_out/r02_zz2_smoke.log
timeout 600 python bench.py > gpurun_out/r02_zz2_bench_c4.json 2> gpurun_out/r02_zz2_bench_c4.err
timeout 600 python bench.py --workload c5 --photons 1.25e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_zz2_bench_c5.json 2>/dev/null
timeout 300 python bench.py --workload c2 --batch 73 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_zz2_bench_c2_batch73.json 2>/dev/null
timeout 300 python bench.py --photons 1e6 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_zz2_bench_c4_1e6.json 2>/dev/null
for f in gpurun_out/r02_zz2_bench_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read()); print('%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('roofline',{}).get('frac'), (d.get('shard_check') or {}).get('ok'), d.get('clocks'))"; done

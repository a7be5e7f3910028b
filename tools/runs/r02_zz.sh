#!/bin/bash
# round 2, third session, final call: full GPU suite, smoke, every bench line, launch list of the default bench, full captures of the C4 and C5 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r02_zz_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_zz_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_zz_pytest.log
tail -5 gpurun_out/r02_zz_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_zz_smoke.log 2>&1; tail -1 gpurun_out/r02_zz_smoke.log
timeout 600 python bench.py > gpurun_out/r02_zz_bench_c4.json 2> gpurun_out/r02_zz_bench_c4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_zz_bench_c4_reference.json 2> gpurun_out/r02_zz_bench_c4_reference.err
for w in c1 c2 c3; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r02_zz_bench_$w.json 2> gpurun_out/r02_zz_bench_$w.err
done
timeout 600 python bench.py --workload c5 --photons 1.25e7 --steps 2 --warmup 1 > gpurun_out/r02_zz_bench_c5.json 2> gpurun_out/r02_zz_bench_c5.err
timeout 300 python bench.py --workload c2 --batch 73 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_zz_bench_c2_batch73.json 2>/dev/null
timeout 300 python bench.py --workload c2 --multi 68 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_zz_bench_c2_multi68.json 2>/dev/null
timeout 300 python bench.py --photons 1e6 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_zz_bench_c4_1e6.json 2>/dev/null
timeout 600 python bench.py --mode faithful --photons 4e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_zz_bench_c4_faithful.json 2>/dev/null
for f in gpurun_out/r02_zz_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print('%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('roofline',{}).get('frac'), (d.get('shard_check') or {}).get('ok'), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('ERR', e)"; done
# launch list of the default bench command (shares, not absolutes) and full captures
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_zz_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_zz_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_zz_ncu1.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --photons 4e6 --no-cpu-baseline > gpurun_out/r02_zz_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transport3 -s 1 -c 1 -o gpurun_out/r02_zz_c4 python bench.py --steps 1 --warmup 1 --photons 4e6 --no-cpu-baseline > gpurun_out/r02_zz_ncu2.log 2>&1
timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 --photons 1e6 --no-cpu-baseline > gpurun_out/r02_zz_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transport3 -s 1 -c 1 -o gpurun_out/r02_zz_c5 python bench.py --workload c5 --steps 1 --warmup 1 --photons 1e6 --no-cpu-baseline > gpurun_out/r02_zz_ncu3.log 2>&1
ls -la gpurun_out/r02_zz_*.ncu-rep gpurun_out/r02_zz_launches_c4.csv

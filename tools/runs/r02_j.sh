#!/bin/bash
# round 2, call j: the faithful mode on the event-list engine (engine3.cuh): parity suite + throughput
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistical.py -m gpu -q -x ) > gpurun_out/r02_j_pytest.log 2>&1
tail -5 gpurun_out/r02_j_pytest.log
for w in c4 c1 c5; do
  ph="--photons 1e6"; [ $w = c5 ] && ph="--photons 2e5"
  timeout 600 python bench.py --workload $w --mode faithful $ph --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_j_faithful_$w.json 2> gpurun_out/r02_j_faithful_$w.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_j_faithful_$w.json').read()); print('faithful $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])
except Exception as e: print('faithful $w ERR', e)"
done

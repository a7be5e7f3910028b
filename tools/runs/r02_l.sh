#!/bin/bash
# round 2, call l: faithful mode, persistent-lane engine (ARTES_ENGINE=1, tuning build) against the event-list engine, C4 at 4e6 packets
mkdir -p gpurun_out
export ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_tuning.so
for e in 1 2 1 2; do
  ARTES_ENGINE=$e timeout 600 python bench.py --mode faithful --photons 4e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_l_faithful_engine$e.json 2> gpurun_out/r02_l_faithful_engine$e.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_l_faithful_engine$e.json').read()); print('faithful c4 ARTES_ENGINE=$e', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])
except Exception as e: print('ERR', e)"
done
ARTES_ENGINE=1 timeout 600 python bench.py --mode faithful --workload c5 --photons 2e5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_l_faithful_c5_engine1.json 2>/dev/null
ARTES_ENGINE=2 timeout 600 python bench.py --mode faithful --workload c5 --photons 2e5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_l_faithful_c5_engine3.json 2>/dev/null
python -c "
import json
for e in ('1','3'):
    d=json.loads(open('gpurun_out/r02_l_faithful_c5_engine%s.json'%e).read()); print('faithful c5 engine',e, '%.4g'%d['value'])"

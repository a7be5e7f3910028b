#!/bin/bash
# round 2, call i: pure-global reductions in the deposit again, bounded waits on / off, multi-detector walk after the row-by-row Mueller product
mkdir -p gpurun_out
for v in cur nowd cur nowd; do
  for w in c4 c1 c2; do
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_i_${v}_$w.json 2> gpurun_out/r02_i_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_i_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])
except Exception as e: print('$v $w ERR', e)"
  done
done
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_cur.so timeout 300 python bench.py --workload c2 --multi 68 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_i_cur_c2_multi68.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_i_cur_c2_multi68.json').read()); print('cur c2 multi68', '%.4g'%d['value'])"

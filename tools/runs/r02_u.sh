#!/bin/bash
# round 2, call u: first crossings of a transport ray marched by the event that set it up (E2_HEAD_STEPS), same box A/B + parity
mkdir -p gpurun_out
for v in base head4 head8 head16 base; do
  for w in c4 c1 c2 c5 c3; do
    ph=""; [ $w = c5 ] && ph="--photons 1e6"
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w $ph --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_u_${v}_$w.json 2> gpurun_out/r02_u_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_u_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])
except Exception as e: print('$v $w ERR', e)"
  done
done
for v in base head8; do
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --photons 1e6 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_u_${v}_c4_1e6.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_u_${v}_c4_1e6.json').read()); print('$v c4 1e6', '%.4g'%d['value'])"
done
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_head8.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistical.py -m gpu -q -x > gpurun_out/r02_u_pytest.log 2>&1; tail -4 gpurun_out/r02_u_pytest.log

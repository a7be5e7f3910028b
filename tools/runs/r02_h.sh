#!/bin/bash
# round 2, call h: block shapes after the row-by-row Mueller product (register budget 128 -> 96 / 80), same box
mkdir -p gpurun_out
for v in cur nt320np512 nt320np640 nt384np768 cur; do
  for w in c4 c5 c1 c2; do
    ph=""; [ $w = c5 ] && ph="--photons 1e6"
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w $ph --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_h_${v}_$w.json 2> gpurun_out/r02_h_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_h_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])
except Exception as e: print('$v $w ERR', e)"
  done
done
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_cur.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "same_stream or crossing or golden" > gpurun_out/r02_h_pytest.log 2>&1; tail -3 gpurun_out/r02_h_pytest.log

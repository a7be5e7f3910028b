#!/bin/bash
# round 2, call c: which of the marcher changes pays (variants of the library, one switch off each), statistical gate re-run
mkdir -p gpurun_out
for v in none all nokapn nor2s nowd; do
  for w in c4 c5; do
    ph=""; [ $w = c5 ] && ph="--photons 2e6"
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w $ph --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_c_${v}_$w.json 2> gpurun_out/r02_c_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_c_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['clocks']['sm_mhz'])
except Exception as e: print('$v $w ERR', e)"
  done
done
( time timeout 900 python -m pytest tests/test_gpu_statistical.py -m gpu -q ) > gpurun_out/r02_c_pytest.log 2>&1
tail -4 gpurun_out/r02_c_pytest.log

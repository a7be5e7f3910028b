#!/bin/bash
# round 2, call e: same-box A/B of the watchdog and of the block-private detector image, driver tests
mkdir -p gpurun_out
for v in cur nowd sdet56 nosdet cur; do
  for w in c4 c1 c2; do
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_e_${v}_$w.json 2> gpurun_out/r02_e_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_e_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['clocks']['sm_mhz'])
except Exception as e: print('$v $w ERR', e)"
  done
done
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_cur.so timeout 300 python bench.py --workload c2 --batch 73 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_e_cur_c2_batch73.json 2>/dev/null
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_nosdet.so timeout 300 python bench.py --workload c2 --batch 73 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_e_nosdet_c2_batch73.json 2>/dev/null
python -c "
import json
for v in ('cur','nosdet'):
    d=json.loads(open('gpurun_out/r02_e_%s_c2_batch73.json'%v).read()); print(v,'c2 batch73', '%.4g'%d['value'])"
( time timeout 900 python -m pytest tests/test_driver.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_e_pytest.log 2>&1
tail -4 gpurun_out/r02_e_pytest.log

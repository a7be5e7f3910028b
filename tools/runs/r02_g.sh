#!/bin/bash
# round 2, call g: same-box A/B (watchdog v3, pass length, starvation threshold, private image for multi-detector walks), ingest at scale
mkdir -p gpurun_out
free -g > gpurun_out/r02_g_mem.txt
for v in cur nowd inner1 inner3 starve16 cur; do
  for w in c4 c5 c2; do
    ph=""; [ $w = c5 ] && ph="--photons 1e6"
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w $ph --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_g_${v}_$w.json 2> gpurun_out/r02_g_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_g_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])
except Exception as e: print('$v $w ERR', e)"
  done
done
for v in cur nosdet; do
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload c2 --multi 68 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_g_${v}_c2_multi68.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_g_${v}_c2_multi68.json').read()); print('$v c2 multi68', '%.4g'%d['value'])"
done
timeout 900 python tools/gpu_ingest.py --max-gb 18 > gpurun_out/r02_g_ingest.json 2> gpurun_out/r02_g_ingest.err; cat gpurun_out/r02_g_ingest.json; tail -3 gpurun_out/r02_g_ingest.err

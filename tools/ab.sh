#!/bin/bash
# Same-box A/B of library variants built by tools/build_variants.sh.  usage (on the GPU box): bash tools/ab.sh "c4 c1" base cut ...
# prints packets/s of bench.py --workload W --steps 3 --warmup 2 per variant; W = "c4@1e6" sets --photons
wls="$1"; shift
for v in "$@"; do
  line="$v"
  for w in $wls; do
    wl="${w%%@*}"; ph=""; [[ "$w" == *@* ]] && ph="--photons ${w#*@}"
    r=$(ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so python bench.py --workload $wl $ph --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4e'%d['value'])")
    line="$line  $w $r"
  done
  echo "$line"
done

#!/bin/bash
# Evidence run of a round on one B200: bench lines of all workloads, the reference arm, the batched phase curve,
# then (only after the plain runs exited 0) the ncu launch list and one full capture of the dominant kernel.
# usage (on the GPU box): bash tools/evidence.sh r01_f
tag=${1:-r01_x}; out=gpurun_out; mkdir -p $out
python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err || exit 1
for c in c1 c2 c3 c5; do python bench.py --workload $c > $out/bench_${tag}_$c.json 2> $out/bench_${tag}_$c.err || exit 1; done
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.err || exit 1
python bench.py --workload c2 --batch 73 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_${tag}_c2_batch73.json 2> $out/bench_${tag}_c2_batch73.err || exit 1
python bench.py --workload c2 --photons 1000000 --steps 73 --warmup 3 --no-cpu-baseline > $out/bench_${tag}_c2_single_1e6.json 2> $out/bench_${tag}_c2_single_1e6.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:transport3 -s 1 -c 1 -o $out/prof_${tag} -f python tools/gpu_tune.py c4 8000000 fast > $out/ncu_full_${tag}.log 2>&1
tail -c 400 $out/bench_${tag}.json; echo; for c in c1 c2 c3 c5 c2_batch73 c2_single_1e6 ref; do python - <<PY
import json; d=json.load(open("$out/bench_${tag}_$c.json")); print("$c", "%.4e"%d["value"], "e2e %.4e"%d["e2e"]["value"], d.get("clocks"))
PY
done

#!/bin/bash
# round 2, call f2 (8 GPUs, final build): multi-GPU equality tests, C5 scale run, C4 weak / strong scaling at 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 600 python -m pytest tests/test_multirank_gpu.py tests/test_gpu_parity.py -m gpu -q -k "two_device or torchrun or dense" ) > gpurun_out/r02_f2_pytest.log 2>&1
tail -3 gpurun_out/r02_f2_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_f2_c4_n1.json 2> gpurun_out/r02_f2_c4_n1.err
timeout 600 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_f2_c4_weak_n8.json 2> gpurun_out/r02_f2_c4_weak_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --steps 3 --warmup 3 --photons 1.25e6 > gpurun_out/r02_f2_c4_strong_n8.json 2> gpurun_out/r02_f2_c4_strong_n8.err
timeout 600 python bench.py --workload c5 --photons 1.25e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_f2_c5_n1.json 2> gpurun_out/r02_f2_c5_n1.err
timeout 900 $TR --nproc-per-node 8 --master-port 29713 bench.py --gpus 8 --workload c5 --photons 1.25e7 --steps 2 --warmup 1 > gpurun_out/r02_f2_c5_n8.json 2> gpurun_out/r02_f2_c5_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29714 bench.py --gpus 8 --workload c2 --multi 68 --photons 1e6 --steps 2 --warmup 2 > gpurun_out/r02_f2_c2_multi68_n8.json 2> gpurun_out/r02_f2_c2_multi68_n8.err
for f in gpurun_out/r02_f2_c*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['n_gpus'], '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['shard_check']['ok'], d['rank_times']['kernel_ms_per_step_min'], d['rank_times']['kernel_ms_per_step_max'], d['rank_times']['reduce_ms_per_step_max'])
except Exception as e: print('ERR', e)"; done

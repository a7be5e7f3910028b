#!/bin/bash
# round 2, call f (8 GPUs): C5 scale run (multi-wavelength batch, >= 1e8 packets per GPU), weak and strong scaling of C4, multi-GPU equality tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_f_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# multi-GPU equality: two-device context (ncclCommInitAll) and two torchrun ranks (ncclCommInitRank), both against one launch
( time timeout 600 python -m pytest tests/test_multirank_gpu.py tests/test_gpu_parity.py -m gpu -q -k "two_device or torchrun" ) > gpurun_out/r02_f_pytest.log 2>&1
tail -3 gpurun_out/r02_f_pytest.log
# C5 at scale
timeout 600 python bench.py --workload c5 --photons 1.25e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_f_c5_n1.json 2> gpurun_out/r02_f_c5_n1.err
timeout 900 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --workload c5 --photons 1.25e7 --steps 2 --warmup 1 > gpurun_out/r02_f_c5_n8.json 2> gpurun_out/r02_f_c5_n8.err
# C4: weak scaling (1e7 per GPU) and strong scaling (1e7 in total)
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_f_c4_n1.json 2> gpurun_out/r02_f_c4_n1.err
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 2962$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r02_f_c4_weak_n$n.json 2> gpurun_out/r02_f_c4_weak_n$n.err
  p=$(python -c "print(1e7/$n)")
  timeout 600 $TR --nproc-per-node $n --master-port 2963$n bench.py --gpus $n --steps 3 --warmup 3 --photons $p > gpurun_out/r02_f_c4_strong_n$n.json 2> gpurun_out/r02_f_c4_strong_n$n.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --photons 1e6 --no-cpu-baseline > gpurun_out/r02_f_c4_n1_1e6.json 2> gpurun_out/r02_f_c4_n1_1e6.err
for f in gpurun_out/r02_f_c*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read()); print(d['n_gpus'], '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['shard_check']['ok'], d['rank_times']['kernel_ms_per_step_min'], d['rank_times']['kernel_ms_per_step_max'], d['rank_times']['reduce_ms_per_step_max'])
except Exception as e: print('ERR', e)"; done

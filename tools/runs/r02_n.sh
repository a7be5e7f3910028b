#!/bin/bash
# round 2, call n: multi-detector fan-out dealt over (photon, detector) pairs; persistent-lane kernel out of the product build
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_statistical.py tests/test_driver.py -m gpu -q -x ) > gpurun_out/r02_n_pytest.log 2>&1
tail -4 gpurun_out/r02_n_pytest.log
for k in 68 73 16; do
timeout 300 python bench.py --workload c2 --multi $k --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_n_c2_multi$k.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_n_c2_multi$k.json').read()); print('c2 multi $k', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])"
done
timeout 300 python bench.py --workload c4 --multi 36 --photons 1e6 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_n_c4_multi36.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_n_c4_multi36.json').read()); print('c4 multi 36 (64x64 images, global atomics)', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_n_c4.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_n_c4.json').read()); print('c4', '%.4g'%d['value'])"

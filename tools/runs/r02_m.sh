#!/bin/bash
# round 2, call m: full GPU suite after the checkpointed faithful CDF search; faithful throughput
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_m_pytest.log
tail -6 gpurun_out/r02_m_pytest.log
timeout 600 python bench.py --mode faithful --photons 4e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_m_faithful_c4.json 2> gpurun_out/r02_m_faithful_c4.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_m_faithful_c4.json').read()); print('faithful c4', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'], d['roofline']['frac'])"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_m_smoke.log 2>&1; tail -2 gpurun_out/r02_m_smoke.log

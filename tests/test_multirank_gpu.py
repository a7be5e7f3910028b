"""N > 1 on the GPU, launched the way the driver launches bench.py (torchrun, one process per rank).

Two ranks shard the photon ids and every rank walks its shard through libartes_gpu.  With two or more GPUs the ranks sit
on different devices and the library's own NCCL communicator sums the images inside artes_gpu_run (the production
path of bench.py / SCALE).  On a one-GPU box NCCL refuses two ranks on one device, so both ranks use cuda:0 and the
per-rank images are summed on the host over gloo -- the sharding, the photon-id keyed streams and the process plumbing
are still the real ones, only the transport of the sum differs.  Either way the result must equal the single launch of
all photon ids: count planes identical, Stokes sums to rounding."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import math, os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from artes_b200 import abi, host, dist as adist
from tools import atmospheres as A
ngpu = torch.cuda.device_count()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
use_nccl = ngpu >= world
dev = rank if use_nccl else 0
torch.cuda.set_device(dev)
dist.init_process_group("gloo")
atm = A.c4_mie_patches()
p = host.Params(nx=16, ny=16, det_phi=math.radians(50.0))
t = host.Transport(atm, p, devices=(dev,), mode=abi.MODE_FAST)
if use_nccl:
    adist.init_library_comm(t.gpu, dist, rank, world)
t.set_wavelength(0)
total = 60001
base, n = adist.shard(total, world, rank)
r = t.gpu.run(t.launch_struct(n, seed=51, photon_id_base=base))
det = torch.from_numpy(r["det"].copy())
cnt = torch.tensor([r["stats"][k] for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel", "n_draws")], dtype=torch.int64)
if not use_nccl:                      # one GPU: host-side sum over gloo
    dist.all_reduce(det); dist.all_reduce(cnt)
Ls = [t.launch_struct(n, seed=51, photon_id_base=10**9 + base, det_phi=math.radians(a)) for a in (0.0, 45.0, 90.0, 135.0, 180.0)]
# batched launches shard too: launch k of rank r walks ids 1e9 + base_r + k*n_r ...
rb = t.gpu.run_batch(Ls)
detb = torch.from_numpy(rb["det"].copy())
if not use_nccl:
    dist.all_reduce(detb)
if rank == 0:
    t1 = host.Transport(atm, p, devices=(dev,), mode=abi.MODE_FAST)      # no communicator
    t1.set_wavelength(0)
    full = t1.gpu.run(t1.launch_struct(total, seed=51, photon_id_base=0))
    assert np.array_equal(det[2].numpy(), full["det"][2]), "count planes differ"
    np.testing.assert_allclose(det[0].numpy(), full["det"][0], rtol=1e-9, atol=1e-12 * np.abs(full["det"][0]).max())
    assert cnt.tolist() == [full["stats"][k] for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel", "n_draws")]
    # ... reference for the batch: every (rank, launch) range walked alone
    acc = np.zeros_like(rb["det"])
    for rr in range(world):
        b_, n_ = adist.shard(total, world, rr)
        for k, a in enumerate((0.0, 45.0, 90.0, 135.0, 180.0)):
            acc[k] += t1.gpu.run(t1.launch_struct(n_, seed=51, photon_id_base=10**9 + b_ + k * n_, det_phi=math.radians(a)))["det"]
    assert np.array_equal(detb[:, 2].numpy(), acc[:, 2]), "batched count planes differ"
    np.testing.assert_allclose(detb[:, 0].numpy(), acc[:, 0], rtol=1e-9, atol=1e-12 * np.abs(acc[:, 0]).max())
    print("MULTIRANK_GPU_OK", world, "nccl" if use_nccl else "gloo-sum", cnt.tolist())
    t1.close()
t.close()
dist.barrier()
dist.destroy_process_group()
'''


def test_two_torchrun_ranks_equal_one_launch(tmp_path):
    script = tmp_path / "worker_gpu.py"
    script.write_text(WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIRANK_GPU_OK 2" in r.stdout

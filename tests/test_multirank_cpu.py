"""N > 1 on the CPU: two `gloo` ranks shard the photon ids exactly as bench.py / the driver do on GPUs, each runs its
shard (the CPU oracle stands in for the device here -- test infrastructure), the images are summed with an
all-reduce, and the result must equal the single-rank run: the invariant that makes the NCCL reduce of
libartes_gpu correct (src/ARTES.f90:959-975 is the same sum over OpenMP threads)."""
import os
import subprocess
import sys

import numpy as np

from artes_b200 import dist as adist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from artes_b200 import dist as adist
from artes_b200.abi import make_launch
from oracle_lib import Oracle
from tools import atmospheres as A
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
atm = A.c2_hg_deck()
o = Oracle(); o.set_atmosphere(atm)
xm = 1.3 * atm.rfront[-1]
total = 30001
base, n = adist.shard(total, world, rank)
kw = dict(x_max=xm, y_max=xm, seed=12, nx=8, ny=8)
r = o.run(make_launch(n_photons=n, photon_id_base=base, **kw), nthreads=2)
det = torch.from_numpy(r["det"].copy())
cnt = torch.tensor([r["stats"][k] for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel")], dtype=torch.int64)
dist.all_reduce(det); dist.all_reduce(cnt)
if rank == 0:
    full = o.run(make_launch(n_photons=total, photon_id_base=0, **kw), nthreads=2)
    assert (det[2].numpy() == full["det"][2]).all(), "counts differ"
    np.testing.assert_allclose(det[0].numpy(), full["det"][0], rtol=1e-10, atol=1e-14)
    assert cnt.tolist() == [full["stats"][k] for k in ("n_emit", "n_cell_face", "n_scatter", "n_peel")]
    print("MULTIRANK_OK", world, cnt.tolist())
dist.destroy_process_group()
'''


def test_shard_ranges_partition_the_photon_ids():
    for total in (0, 1, 7, 1000, 10**10 + 3):
        for world in (1, 2, 3, 8):
            rs = [adist.shard(total, world, r) for r in range(world)]
            assert rs[0][0] == 0 and sum(n for _, n in rs) == total
            for (b0, n0), (b1, _) in zip(rs, rs[1:]):
                assert b0 + n0 == b1
            assert max(n for _, n in rs) - min(n for _, n in rs) <= 1
    # weak-scaling steps of bench.py never reuse a photon id
    seen = set()
    for step in range(3):
        for rank in range(4):
            b = adist.step_base(step, 4, rank, 10)
            ids = set(range(b, b + 10))
            assert not (ids & seen)
            seen |= ids


def test_two_gloo_ranks_equal_one(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTIRANK_OK 2" in r.stdout

"""Multi-GPU plumbing: one process per GPU (torchrun), photons sharded by global photon id, the Stokes images and
counters summed by NCCL inside libartes_gpu (src/ARTES.f90:959-975 is the reference's thread sum).

Photons never interact (src/ARTES.f90:546: the loop body has no cross-iteration dependence apart from the
accumulators), so the path shards with no data-path collective: every rank runs an independent photon-id range
on replicated tables and ONE all-reduce per launch adds the [image | flux | flow | counters] buffers."""
from __future__ import annotations


def shard(total: int, world: int, rank: int):
    """Contiguous photon-id range of `rank`: (first id, count); ranges are disjoint and cover [0, total)."""
    per, rem = divmod(int(total), int(world))
    return rank * per + min(rank, rem), per + (1 if rank < rem else 0)


def step_base(step: int, world: int, rank: int, per_rank: int):
    """First photon id of `rank` in weak-scaling step `step` (bench.py: per_rank photons per GPU and step)."""
    return (step * world + rank) * per_rank


def init_library_comm(gpu, dist_module, rank: int, world: int):
    """Create the library's own NCCL communicator: rank 0 makes the unique id, torch.distributed carries it."""
    from . import lib
    obj = [lib.nccl_unique_id() if rank == 0 else None]
    dist_module.broadcast_object_list(obj, src=0)
    gpu.nccl_init_rank(world, rank, obj[0])

"""Sample-index partition across the GPUs of one box + the per-frame accumulator reduce.

The reference's only parallelism is data-parallel over (sample block, tile) tasks pulled from
one atomic counter (raytracer.odin:540-560).  Here every GPU holds a full scene replica and
renders a contiguous block of sample indices for ALL pixels; the counter-based RNG is keyed by
the global sample index, so any split renders the same sample set.  The partial accumulators
(planar total / total_squared / count, 8 planes x H*W f32) are combined with ONE reduce per frame
(NCCL over NVLink on GPUs; gloo in the CPU tests) — there is no other data-path collective.
"""
from typing import Tuple


def sample_partition(first_sample: int, n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of `rank`: (first, count).  Blocks differ by at most one sample and their
    union is exactly [first_sample, first_sample + n_samples)."""
    base, rem = divmod(n_samples, world)
    count = base + (1 if rank < rem else 0)
    first = first_sample + rank * base + min(rank, rem)
    return first, count


def reduce_accum(accum, dst: int = 0, group=None):
    """One sum-reduce of the planar accumulator tensor to rank `dst` (in place on dst)."""
    import torch.distributed as dist

    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum

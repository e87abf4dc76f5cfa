"""Multi-GPU sharding of the pixel-sample space (replaces the tile scheduler of src/librender/renderproc.cpp:117-182 and the
TCP workers of src/libcore/sched_remote.cpp for this path).

Every (pixel, sample index) path is independent and the film stores sum(w*L), sum(w*alpha), sum(w) separately
(include/mitsuba/render/imageblock.h:124-131), so each rank renders a contiguous range of sample indices for ALL pixels into a
private full-size film and the films are summed with ONE reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  The RNG is
keyed by (pixel, sample, vertex, seed), hence the result does not depend on the number of ranks up to fp32 summation order.
"""


def sample_range(total_spp, rank, world):
    """Contiguous, balanced split of sample indices [0, total_spp) -> [begin, end) of `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('invalid rank/world')
    base, rem = divmod(int(total_spp), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_film(film_tensor, dst=0):
    """Sum of the per-rank films on rank `dst` (one collective; no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return film_tensor


def render_sharded(render_range, total_spp, film_tensor, rank, world):
    """render_range(begin, end) must ACCUMULATE sample indices [begin, end) into film_tensor; then the films are reduced."""
    b, e = sample_range(total_spp, rank, world)
    if e > b:
        render_range(b, e)
    return reduce_film(film_tensor)

"""Multi-GPU plumbing of the path: one process per GPU, sample-sharded rendering, NCCL all-reduce of image + gradients.

The reference is single-GPU; every lane is independent until film accumulation and gradient accumulation
(src/integrator/integrator.cpp:88,117), so the only exchange is a sum of the per-rank film and of the flat gradient
vector. torch.distributed is the transport (backend nccl on GPUs, gloo in the CPU tests).
"""


def shard_samples(spp, rank, world):
    """samples [s0, s1) of every pixel owned by `rank` — must match pb_ctx_set_shard (csrc/pb_capi.cu)"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("Invalid shard")
    return (spp * rank) // world, (spp * (rank + 1)) // world


def shard_lanes(n, rank, world):
    """lane range [l0, l1) of an edge-term wavefront owned by `rank`"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("Invalid shard")
    return (n * rank) // world, (n * (rank + 1)) // world


def all_reduce_sum_(tensor, group=None):
    """in-place sum over ranks; a no-op outside an initialised process group"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def reduce_render(image, grad=None, group=None):
    """what a sharded renderC / renderD+VJP must do before returning: sum the film, sum the gradient vector"""
    all_reduce_sum_(image, group)
    if grad is not None:
        all_reduce_sum_(grad, group)
    return image, grad

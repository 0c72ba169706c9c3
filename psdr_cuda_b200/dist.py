"""Multi-GPU plumbing of the path: one process per GPU; the exchange itself lives in the library (csrc/pb_dist.cu: an NCCL communicator
per context, `pb_allreduce_image` / `pb_allreduce_grads` enqueued on the context's stream). What is Python here is only how the
ranks find each other: the 128-byte NCCL unique id travels through an initialised torch.distributed group (nccl on GPUs, gloo in
the CPU tests), and the partition helpers mirror the library's so that host code can reason about who owns what.

The reference is single-GPU; every lane is independent until film accumulation and gradient accumulation
(src/integrator/integrator.cpp:88,117), so the only exchange is the film and ONE sum of the flat gradient vector.
"""


def shard_samples(spp, rank, world):
    """samples [s0, s1) of every pixel owned by `rank` under sample sharding — must match pb_ctx_set_shard (csrc/pb_capi.cu)"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("Invalid shard")
    return (spp * rank) // world, (spp * (rank + 1)) // world


def shard_rows(height, rank, world, tile_rows=0):
    """image rows owned by `rank` under pixel sharding (PB_SHARD_PIXELS): tiles of `tile_rows` rows dealt round-robin
    (0 = ceil(height / world): one contiguous block per rank) — must match render_interior / global_lane (csrc)"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("Invalid shard")
    t = max(1, tile_rows if tile_rows > 0 else -(-height // world))
    rows = []
    r0 = rank * t
    while r0 < height:
        rows += list(range(r0, min(height, r0 + t)))
        r0 += t * world
    return rows


def shard_lanes(n, rank, world):
    """lane range [l0, l1) of an edge-term wavefront owned by `rank`"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("Invalid shard")
    return (n * rank) // world, (n * (rank + 1)) // world


def broadcast_unique_id(make_id, group=None, device="cpu"):
    """rank 0 calls `make_id()` (-> 128 bytes, pb_dist_unique_id); every rank of the torch.distributed group gets them"""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank(group) == 0:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).to(device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def init_context(ctx, group=None, mode="pixels", tile_rows=0):
    """give a capi.Context its NCCL communicator and shard for the ranks of an initialised torch.distributed group"""
    import torch.distributed as dist
    dev = "cuda:%d" % ctx.device if dist.get_backend(group) == "nccl" else "cpu"
    uid = broadcast_unique_id(ctx.dist_unique_id, group, dev)
    ctx.dist_init(uid, dist.get_rank(group), dist.get_world_size(group))
    ctx.set_shard_mode(mode, tile_rows)
    return ctx

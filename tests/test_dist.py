"""N > 1 host logic on the CPU: the partitions (samples / pixel-row tiles / edge lanes) tile the job exactly, a world_size-2 gloo run
moves the library's 128-byte communicator id through the group and reproduces the single-rank film and gradient from the shards'
contributions, and the C ABI's collective entry points behave without a GPU (no-ops on one rank, loud errors otherwise).
The NCCL path itself (two ranks rendering for real) is tests/test_dist_gpu.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from psdr_cuda_b200 import dist as pdist


def test_shards_partition_exactly():
    for spp in (1, 2, 3, 7, 8, 16, 256, 255):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                s0, s1 = pdist.shard_samples(spp, r, world)
                assert 0 <= s0 <= s1 <= spp
                got += list(range(s0, s1))
            assert got == list(range(spp))
    n = 512 * 512 * 128
    ends = [pdist.shard_lanes(n, r, 8) for r in range(8)]
    assert ends[0][0] == 0 and ends[-1][1] == n and all(ends[i][1] == ends[i + 1][0] for i in range(7))
    with pytest.raises(ValueError):
        pdist.shard_samples(8, 2, 2)


def test_pixel_row_tiles_partition_exactly():
    for height in (1, 7, 64, 100, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            for tile in (0, 1, 3, 16):
                rows = [pdist.shard_rows(height, r, world, tile) for r in range(world)]
                flat = sorted(x for r in rows for x in r)
                assert flat == list(range(height)), (height, world, tile)
                if tile == 0 and height % world == 0:      # one contiguous, equal block per rank: the film exchange is an in-place all-gather
                    for r in range(world):
                        assert rows[r] == list(range(r * height // world, (r + 1) * height // world))
    # the library's global_lane mapping (csrc/pb_wavefront.cuh) restated: local row -> global row
    def to_global(lr, tile, rank, world):
        t = lr // tile
        return (t * world + rank) * tile + (lr - t * tile)
    for height, world, tile in ((100, 3, 16), (512, 8, 64), (37, 4, 1)):
        for r in range(world):
            rows = pdist.shard_rows(height, r, world, tile)
            assert [to_global(i, tile, r, world) for i in range(len(rows))] == rows


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = pdist.broadcast_unique_id(lambda: bytes(range(128)))          # rank 0's id reaches every rank
    h, w, spp = 8, 5, 6
    rng = np.random.default_rng(0)
    per_sample = rng.normal(size=(spp, h * w, 3)).astype(np.float32)     # what each (sample, pixel) contributes to the film
    per_sample_grad = rng.normal(size=(spp, h * w, 12)).astype(np.float32)
    # sample sharding: partial sums of every pixel -> sum
    s0, s1 = pdist.shard_samples(spp, rank, world)
    img_s = torch.from_numpy(per_sample[s0:s1].sum(0) / spp)
    grad_s = torch.from_numpy(per_sample_grad[s0:s1].sum((0, 1)))
    dist.all_reduce(img_s); dist.all_reduce(grad_s)
    # pixel tiles: disjoint rows, zeros elsewhere -> the same sum is exact (x + 0), one gradient all-reduce
    rows = pdist.shard_rows(h, rank, world, 0)
    mask = np.zeros((h, w), bool); mask[rows] = True
    img_p = torch.from_numpy(np.where(mask.reshape(-1, 1), per_sample.sum(0) / spp, 0).astype(np.float32))
    grad_p = torch.from_numpy(per_sample_grad[:, mask.reshape(-1)].sum((0, 1)))
    dist.all_reduce(img_p); dist.all_reduce(grad_p)
    q.put((rank, uid, img_s.numpy(), grad_s.numpy(), img_p.numpy(), grad_p.numpy(), per_sample.sum(0) / spp, per_sample_grad.sum((0, 1))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_matches_single():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, uid, img_s, grad_s, img_p, grad_p, img_ref, grad_ref in res:
        assert uid == bytes(range(128))
        assert np.allclose(img_s, img_ref, atol=1e-6) and np.allclose(grad_s, grad_ref, atol=1e-4)
        assert np.array_equal(img_p, img_ref.astype(np.float32)) and np.allclose(grad_p, grad_ref, atol=1e-4)


def test_collective_entry_points_without_a_gpu(native_lib):
    from psdr_cuda_b200 import capi
    L = capi.lib()
    assert L.pb_dist_available() in (0, 1)
    assert L.pb_allreduce_grads(None, None, C.c_int64(0)) != 0          # no context: an error code, not a crash

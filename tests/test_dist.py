"""CPU test of the N>1 host logic: world_size-2 gloo run of the shard partition + film / gradient all-reduce."""
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


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spp, npix = 6, 50
    rng = np.random.default_rng(0)
    per_sample = rng.normal(size=(spp, npix, 3)).astype(np.float32)          # what each sample index contributes
    per_sample_grad = rng.normal(size=(spp, 12)).astype(np.float32)
    s0, s1 = pdist.shard_samples(spp, rank, world)
    img = torch.from_numpy(per_sample[s0:s1].sum(0) / spp)
    grad = torch.from_numpy(per_sample_grad[s0:s1].sum(0))
    pdist.reduce_render(img, grad)
    q.put((rank, img.numpy(), grad.numpy(), per_sample.sum(0) / spp, per_sample_grad.sum(0)))
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
    for rank, img, grad, img_ref, grad_ref in res:
        assert np.allclose(img, img_ref, atol=1e-6) and np.allclose(grad, grad_ref, atol=1e-5)

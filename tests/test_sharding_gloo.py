"""CPU, world_size 2 over gloo: the host-side logic of the clip-sharded path (shard bounds, noise-stream equivalence
with the single-process run, ordering of the final all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rohm_b200 import parallel


def test_shard_bounds_cover_and_are_contiguous():
    for n in (1, 2, 7, 32, 128, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a stand-in sampler with the same RNG call pattern as p_sample_loop: x_T, then one draw per step
        gen_full = torch.Generator().manual_seed(123)
        ref = torch.randn(n_clips, 5, 3, generator=gen_full)
        for _ in range(4):
            ref = 0.5 * ref + torch.randn(ref.shape, generator=gen_full)
        tape = parallel.ShardedNoise(n_clips, rank, world, generator=torch.Generator().manual_seed(123))
        cond = torch.arange(n_clips, dtype=torch.float32).view(-1, 1, 1).expand(n_clips, 5, 3).contiguous()
        local = parallel.shard_batch({"cond": cond, "tag": "x", "scalar": torch.tensor(3.0)}, rank, world, n_clips)
        lo, hi = parallel.shard_bounds(n_clips, rank, world)
        assert local["cond"].shape[0] == hi - lo and local["tag"] == "x" and local["scalar"].dim() == 0
        x = tape.randn(hi - lo, 5, 3)
        for _ in range(4):
            x = 0.5 * x + tape.randn_like(x)
        out = parallel.gather_clips(x + local["cond"], n_clips)
        ok = torch.equal(out, ref + cond)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7])
def test_two_rank_sharded_run_equals_single_process(n_clips):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, n_clips, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def _worker_guidance(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class M:  # the attribute protocol PoseNet.guide_skating_with_smpl reads
            pass

        m = parallel.global_guidance(M())
        sums = torch.tensor([1.0 + rank, 10.0 * (rank + 1), 0.5, 2.0])  # this shard's {sum_abs, cnt_abs, sum_smpl, cnt_smpl}
        m.guidance_sum_reducer(sums)
        ok = torch.equal(sums, torch.tensor([3.0, 30.0, 1.0, 4.0]))
        parallel.global_guidance(m, enable=False)
        ret[rank] = bool(ok) and not hasattr(m, "guidance_sum_reducer")
    finally:
        dist.destroy_process_group()


def test_global_guidance_reducer_sums_the_shard_statistics():
    """The optional exact-global guidance mode: the four loss sums of every shard are all-reduced (one 4-float collective
    per guided step), so the batch-wide normalisers of an unsharded reference run are reproduced."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker_guidance, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}

"""The N>1 path on CPU: chunk planning / assignment, ordered concatenation, and the max-over-ranks time
reduction over a real world_size-2 gloo process group (the GPU box runs the same code over nccl)."""
import os, socket
import pytest
from av1_base_b200 import sharding


def test_plan_and_assignment_cover_every_frame_once():
    for n, k, w in [(2400, 150, 8), (2400, 240, 8), (1200, 240, 1), (7, 3, 2), (0, 240, 4), (240, 240, 8)]:
        chunks = sharding.plan_chunks(n, k)
        assert sum(c[1] for c in chunks) == n
        assert all(c[1] <= k for c in chunks)
        seen = []
        for r in range(w):
            seen += sharding.chunks_of_rank(chunks, r, w)
        assert sorted(seen) == chunks
    # 2400 frames with scene_len 150 -> 16 chunks, evenly divisible over 1/2/4/8 GPUs
    chunks = sharding.plan_chunks(2400, 150)
    for w in (1, 2, 4, 8):
        assert {len(sharding.chunks_of_rank(chunks, r, w)) for r in range(w)} == {16 // w}


def test_concat_in_order():
    chunks = sharding.plan_chunks(70, 10)
    per_rank = [[("c%d" % i) for i, _ in enumerate(chunks) if i % 3 == r] for r in range(3)]
    assert sharding.concat_in_order(per_rank, len(chunks), 3) == ["c%d" % i for i in range(len(chunks))]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunks = sharding.plan_chunks(100, 10)
    mine = sharding.chunks_of_rank(chunks, rank, world)
    dist.barrier()
    elapsed = 1.0 + rank            # rank 1 is the slow one
    tmax = sharding.max_over_ranks(elapsed, dist)
    total_frames = sum(c[1] for c in mine)
    import torch
    t = torch.tensor([total_frames], dtype=torch.int64)
    dist.all_reduce(t)
    q.put((rank, tmax, int(t.item()), len(mine)))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [2.0, 2.0]          # every rank sees the maximum
    assert [r[2] for r in res] == [100, 100]          # all frames covered exactly once
    assert [r[3] for r in res] == [5, 5]


def test_host_threads_are_divided_between_ranks(monkeypatch):
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert sharding.host_threads_per_rank() == max(1, (os.cpu_count() or 1) // 8)

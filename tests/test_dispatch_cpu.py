"""Host-side partitioning: utterance sharding (incl. a world_size-2 gloo run) and
the halo chunk plan of the long-form path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_utterances_partitions(pkg):
    lengths = [800, 120, 400, 400, 640, 90, 333, 800, 12]
    for world in (1, 2, 4, 8):
        seen = []
        loads = []
        for r in range(world):
            idx = pkg.shard_utterances(lengths, world, r)
            seen += idx
            loads.append(sum(lengths[i] for i in idx))
        assert sorted(seen) == list(range(len(lengths)))
        assert max(loads) - min(loads) <= max(lengths)
    assert pkg.shard_utterances([5] * 256, 8, 3) == list(range(3, 256, 8))
    with pytest.raises(ValueError):
        pkg.shard_utterances(lengths, 2, 2)


def test_chunk_plan_covers_stream(pkg):
    for frames, core in ((12000, 1000), (700, 200), (30, 200), (2, 2), (1001 * 2, 500)):
        plan = pkg.chunk_plan(frames, core)
        assert plan[0][2] == 0 and plan[-1][3] == frames
        for (lo, hi, klo, khi), nxt in zip(plan, plan[1:] + [None]):
            assert lo % 2 == 0 and klo % 2 == 0 and 0 <= lo <= klo < khi <= hi <= frames
            assert klo - lo in (0, 24) or lo == 0
            assert hi - khi == 24 or hi == frames
            if nxt:
                assert nxt[2] == khi
    with pytest.raises(ValueError):
        pkg.chunk_plan(100, 3)


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [400 + 7 * i for i in range(37)]
    mine = pkg.shard_utterances(lengths, world, rank)
    owner = torch.full((len(lengths),), -1, dtype=torch.int64)
    owner[mine] = rank
    gathered = [torch.empty_like(owner) for _ in range(world)]
    dist.all_gather(gathered, owner)          # test-only exchange; the data path has none
    stack = torch.stack(gathered)
    q.put((rank, bool(((stack >= 0).sum(0) == 1).all()), len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sorted(n for _, _, n in res) == [18, 19]

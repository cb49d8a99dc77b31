"""N>1 host logic on CPU: 2 gloo ranks exercise the frame sharding, the max-over-ranks timing
reduction and the reconstructed-reference exchange used by bench.py --gpus N."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svt_av1_psy_b200 import sharding
    mine = sharding.frames_for_rank(12, rank, world, gop=2)
    # a mini-GOP batch of 2 pictures per rank: picture k of rank r is filled with 100 r + k
    like = torch.zeros(7)
    ex = sharding.ReconExchange(dist, rank, world, like, batch=2)
    frames = [torch.full((7,), float(100 * rank + k)) for k in range(2)]
    ex.wait(ex.post(frames))
    inbox = [[float(t[0]) for t in slot] for slot in ex.inbox]
    ms = sharding.max_over_ranks(dist, [10.0 * (rank + 1), 3.0])
    q.put((rank, mine, inbox, ex.consumers, ex.producers, ms))
    dist.destroy_process_group()


def _run(world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=180) for _ in ps])
    for p in ps:
        p.join(timeout=60)
    return res


def test_two_rank_sharding_and_exchange():
    res = _run(2)
    assert res[0][1] == [0, 1, 4, 5, 8, 9] and res[1][1] == [2, 3, 6, 7, 10, 11]
    assert sorted(res[0][1] + res[1][1]) == list(range(12))  # every frame exactly once
    for rank, _, inbox, cons, prod, ms in res:
        other = 1 - rank
        assert cons == [other] and prod == [other]                    # 2 ranks: one consumer, one producer
        assert inbox == [[100.0 * other + 0], [100.0 * other + 1]]     # both pictures of the producer's batch arrived, slot by slot
        assert ms == [20.0, 3.0]                                      # max over ranks
    from svt_av1_psy_b200 import sharding
    assert sharding.whole_job_fps(20, 2, 100.0) == 400.0


def test_three_rank_owner_to_consumers_exchange():
    """owner -> two consumers, never all-to-all: with 3 ranks every rank receives exactly the pictures of the two ranks before it"""
    res = _run(3)
    for rank, _, inbox, cons, prod, _ in res:
        assert cons == [(rank + 1) % 3, (rank + 2) % 3] and prod == [(rank - 1) % 3, (rank - 2) % 3]
        for k in range(2):
            assert inbox[k] == [100.0 * p + k for p in prod]
    from svt_av1_psy_b200 import sharding
    assert sharding.reference_consumers(0, 1) == [] and sharding.reference_producers(0, 1) == []
    assert sharding.reference_consumers(5, 8) == [6, 7] and sharding.reference_producers(0, 8) == [7, 6]

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
    mine = sharding.frames_for_rank(10, rank, world, gop=2)
    local = torch.full((3, 5), float(rank + 1))
    g = sharding.exchange_recon(dist, local)
    ms = sharding.max_over_ranks(dist, [10.0 * (rank + 1), 3.0])
    q.put((rank, mine, g.tolist(), ms))
    dist.destroy_process_group()


def test_two_rank_sharding_and_exchange():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] == [0, 1, 4, 5, 8, 9] and res[1][1] == [2, 3, 6, 7]
    assert sorted(res[0][1] + res[1][1]) == list(range(10))  # every frame exactly once
    for r in res:
        assert r[2] == [[[1.0] * 5] * 3, [[2.0] * 5] * 3]   # both ranks hold both recon frames
        assert r[3] == [20.0, 3.0]                           # max over ranks
    from svt_av1_psy_b200 import sharding
    assert sharding.whole_job_fps(20, 2, 100.0) == 400.0

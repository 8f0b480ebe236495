"""Multi-process (world_size 2 and 3, gloo, CPU) test of the batch-shard + all-gather host logic
(d3roma_b200/distributed.py); the denoiser is replaced by a deterministic per-sample function of the per-sample
start noise, so the test also checks that results do not depend on the number of ranks."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_denoise(x, generator):
    noise = torch.cat([torch.randn((1, 1, 4, 6), generator=g) for g in generator])
    return x + noise


def _expected(n):
    from d3roma_b200 import distributed as D
    idx = list(range(n))
    x = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1).expand(n, 1, 4, 6)
    return _fake_denoise(x, D.sample_generators(idx, seed=5))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from d3roma_b200 import distributed as D
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)

    def make_batch(idx):
        return {"x": torch.tensor(idx, dtype=torch.float32).view(-1, 1, 1, 1).expand(len(idx), 1, 4, 6)}

    out = D.run_sharded(n, make_batch, _fake_denoise, batch_size=2, seed=5)
    q.put((rank, out))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 7), (3, 8), (2, 1)])
def test_sharded_eval_gathers_in_global_order(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _expected(n)
    for r in range(world):
        assert torch.equal(results[r], want), f"rank {r}"


def test_shard_range_partitions_exactly():
    from d3roma_b200.distributed import shard_range
    for n in (0, 1, 7, 16, 256):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1

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
    from tests._util import run_ranks
    results = dict(run_ranks(_worker, world, (n,)))
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


def _fake_eval(config, pipeline, disable_bar, generator=None, as_numpy=False, raw_disp=None, fxb=None, **kw):
    """Stand-in for evaluate.eval_batch on CPU: 11 fp64 'metrics' per sample that depend on the sample's own data and on
    its own start-noise generator only."""
    noise = torch.stack([torch.randn((), generator=g, dtype=torch.float64) for g in generator])
    base = raw_disp.double().flatten(1).mean(1) + fxb.double()
    rows = base[:, None] * torch.arange(1, 12, dtype=torch.float64)[None] + noise[:, None]
    return None, rows, None, None, None


def _eval_worker(rank, world, port, n, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from d3roma_b200 import distributed as D, synthetic
    if world > 1:
        D.init_from_env("gloo")
    mk = lambda idx: synthetic.make_batch(idx, h=8, w=12)  # noqa: E731
    mdict, rows = D.run_distributed_eval(None, None, n, 2, make_batch_fn=mk, seed=3, eval_fn=_fake_eval, device="cpu")
    q.put((rank, rows, sorted(mdict)))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def test_distributed_eval_metric_rows_independent_of_world_size():
    """distributed_evaluate.py:153-163: the gathered [N, 5] + [N, 6] fp64 metric rows come back in global sample order and
    are the same at world_size 1 and 2."""
    import numpy as np
    from tests._util import run_ranks
    got = {}
    for world in (1, 2):
        res = run_ranks(_eval_worker, world, (5,))
        for _, rows, keys in res:
            assert rows.shape == (5, 11) and rows.dtype == np.float64 and keys == ["depth", "disp"]
        got[world] = res
    for _, rows, _ in got[2]:
        assert np.array_equal(rows, got[1][0][1])


def test_synthetic_batch_keys_and_determinism():
    """stereo_datasets.py:520-544 batch keys; a sample depends only on (seed, global index), not on its batch."""
    from d3roma_b200 import synthetic
    a = synthetic.make_batch([0, 1, 2], h=8, w=12)
    b = synthetic.make_batch([2], h=8, w=12)
    for k in ("raw_disp", "normalized_disp", "sim_disp", "sim_mask", "normalized_rgb", "left_image", "right_image", "mask",
              "depth", "raw_depth", "fxb"):
        assert k in a
        assert torch.equal(a[k][2:3], b[k]), k
    assert a["normalized_rgb"].shape == (3, 3, 8, 12) and a["raw_disp"].shape == (3, 1, 8, 12) and a["fxb"].shape == (3,)
    assert a["normalized_rgb"].abs().max() <= 1 and a["normalized_disp"].abs().max() <= 1
    assert ((a["sim_disp"] == -1.0) == (a["sim_mask"] == 0)).all()      # holes normalise to -1 (disp 0)

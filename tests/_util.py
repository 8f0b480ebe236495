"""Shared helpers of the GPU parity tests."""
import torch


def quad_stats(t: torch.Tensor) -> torch.Tensor:
    """[B, H, W, C] -> fp64 [B, C/4, 2]: per (sample, 4-channel quad) sum and sum of squares."""
    B, C = t.shape[0], t.shape[-1]
    o = t.double().reshape(B, -1, C // 4, 4)
    return torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1)


def check_conv_stats(stats: torch.Tensor, out: torch.Tensor, ref: torch.Tensor) -> bool:
    """The GroupNorm quad statistics a conv kernel emits are those of its output tile: either of the stored (16-bit
    rounded) values - pixels-as-M / per-tap / split-K kernels - or of the fp32 values before rounding - the stmatrix
    epilogue of the halo-block kernel.  Both are checked tightly (a single missing pixel fails either)."""
    ok_rounded = torch.allclose(stats, quad_stats(out), rtol=1e-5, atol=1e-3)
    ok_exact = torch.allclose(stats, quad_stats(ref), rtol=2e-5, atol=5e-3)
    return bool(ok_rounded or ok_exact)


def run_ranks(target, world: int, args=(), attempts: int = 3, timeout: float = 180.0):
    """Spawn ``world`` processes running ``target(rank, world, port, *args, queue)`` on a fresh local port and collect one
    queue item per rank.  A rendezvous can fail for reasons unrelated to the code under test (the probed port taken by
    another process before the ranks bind it, a loaded host): a failed attempt is torn down and retried on a new port;
    only the last attempt's failure is raised."""
    import queue as _queue
    import socket
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    last = None
    for _ in range(attempts):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        q = ctx.Queue()
        procs = [ctx.Process(target=target, args=(r, world, port, *args, q)) for r in range(world)]
        for p in procs:
            p.start()
        items = []
        try:
            for _ in range(world):
                items.append(q.get(timeout=timeout))
            for p in procs:
                p.join(timeout=60)
            if all(p.exitcode == 0 for p in procs):
                return items
            last = RuntimeError(f"rank exit codes {[p.exitcode for p in procs]}")
        except _queue.Empty:
            last = TimeoutError(f"ranks did not report within {timeout} s (exit codes {[p.exitcode for p in procs]})")
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(timeout=10)
    raise last

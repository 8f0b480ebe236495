"""Developer check (torchrun, >= 2 GPUs of one NVLink domain): the peer-memory output gather (d3roma_b200.distributed.
PeerGather: Normalizer.denormalize fused with stores into every rank's buffer + flag handshake) must equal
denormalize + NCCL all_gather_into_tensor bit for bit over several epochs, keep the previous epoch's buffer intact, and
is timed against the NCCL form with CUDA events (max over ranks)."""
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import d3roma_b200 as d  # noqa: E402
from d3roma_b200 import distributed as D  # noqa: E402

rank, world, local = D.init_from_env()
dev = torch.device("cuda", local)
Bp, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 360, 640)
norm = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
pg = D.PeerGather({"disp": ((Bp, 1, H, W), torch.float32), "rows": ((Bp, 11), torch.float64)})
g = torch.Generator(device=dev).manual_seed(100 + rank)
ref_disp = torch.empty((world * Bp, 1, H, W), device=dev)
ref_rows = torch.empty((world * Bp, 11), dtype=torch.float64, device=dev)
prev = None
for epoch in range(5):
    y = torch.rand((Bp, 1, H, W), generator=g, device=dev) * 2 - 1
    rows = torch.rand((Bp, 11), generator=g, device=dev, dtype=torch.float64)
    pg.denormalize_into(y, norm)
    pg.push(rows, "rows")
    out = pg.finish()
    dist.all_gather_into_tensor(ref_disp, norm.denormalize(y))
    dist.all_gather_into_tensor(ref_rows, rows)
    got_disp, got_rows = out["disp"].reshape(world * Bp, 1, H, W), out["rows"].reshape(world * Bp, 11)
    assert torch.equal(got_disp, ref_disp), f"rank {rank} epoch {epoch}: maps differ"
    assert torch.equal(got_rows, ref_rows), f"rank {rank} epoch {epoch}: rows differ"
    if prev is not None:                      # the other parity still holds the previous epoch
        assert torch.equal(prev[0], prev[1]), f"rank {rank} epoch {epoch}: previous epoch's buffer was overwritten"
    prev = (out["disp"], ref_disp.reshape(world, Bp, 1, H, W).clone())
torch.cuda.synchronize()
dist.barrier()


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


y = torch.rand((Bp, 1, H, W), generator=g, device=dev) * 2 - 1
rows = torch.rand((Bp, 11), generator=g, device=dev, dtype=torch.float64)


def nccl_form():
    dist.all_gather_into_tensor(ref_disp, norm.denormalize(y))
    dist.all_gather_into_tensor(ref_rows, rows)


def peer_form():
    pg.denormalize_into(y, norm)
    pg.push(rows, "rows")
    pg.finish()


t_nccl, t_peer = timed(nccl_form), timed(peer_form)
if rank == 0:
    mb = Bp * H * W * 4 / 1e6
    print(f"peer gather OK on {world} ranks ({mb:.1f} MB per rank): denormalise + NCCL all-gather {t_nccl * 1e3:.0f} us, "
          f"fused peer-memory form {t_peer * 1e3:.0f} us per step (CUDA events, max over ranks)", flush=True)
pg.close()
dist.destroy_process_group()

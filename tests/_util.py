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

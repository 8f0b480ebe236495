cd $GRAFT_REPO_ROOT
python -m pytest tests/test_guided_gpu.py tests/test_sched_prepost_gpu.py tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -4
echo "== attention 8 warps"; python tools/bench_attn.py
echo "== attention 4 warps"; D3_ATTN_WARPS=4 python tools/bench_attn.py
python - <<'PY'
import torch
import d3roma_b200 as d
B,H,W=16,384,640
fg = d.FlowGuidance(flow_guidance_weight=1e-3, flow_guidance_mode="stereo")
n = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
x0 = torch.rand(B,1,H,W,device="cuda")*2-1
left, right = torch.rand(B,3,H,W,device="cuda")*2-1, torch.rand(B,3,H,W,device="cuda")*2-1
dirv = torch.randn(B,1,H,W,device="cuda")
def t(fn, n_=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n_*1e3
us = t(lambda: fg.grad(x0,left,right)); print(f"guidance grad (B=16): {us:.1f} us, {8*B*H*W*4/us/1e3:.0f} GB/s algorithmic")
us = t(lambda: fg.perturb_fused(x0, 1, None, n, 0.9, dirv, 1.0, left_image=left, right_image=right)); print(f"guidance fused step (B=16): {us:.1f} us, {11*B*H*W*4/us/1e3:.0f} GB/s algorithmic")
# smooth disparity (real-data-like: coalesced gathers)
x0s = torch.linspace(-0.6,-0.2,W,device="cuda").view(1,1,1,W).expand(B,1,H,W).contiguous()
us = t(lambda: fg.perturb_fused(x0s, 1, None, n, 0.9, dirv, 1.0, left_image=left, right_image=right)); print(f"guidance fused step, smooth disparity: {us:.1f} us")
PY

"""``UNet2DModel`` — drop-in for the reference's denoiser on the sampling path.

The reference builds ``diffusers.UNet2DModel(sample_size=[H,W], in_channels=Cin, out_channels=dc,
layers_per_block=2, block_out_channels=(128,128,256,256,512,512), down_block_types=("DownBlock2D",)*4 +
("AttnDownBlock2D","DownBlock2D"), up_block_types=("UpBlock2D","AttnUpBlock2D")+("UpBlock2D",)*4)`` at
``/root/reference/train.py:788-810`` and calls ``unet(model_input, t).sample`` at
``core/custom_pipelines.py:173``.  This class keeps that call surface (``__call__(sample, timestep,
return_dict=True)``, ``.sample_size``, ``.config``, ``load_state_dict`` with diffusers key names, ``.to``,
``.device``) and runs the forward entirely through ``libd3roma_b200.so``:

    conv_in      d3_conv_in_direct (split: per-loop conditioning term + per-step sample taps; fp32); generic
                 inputs: d3_im2col_input + d3_conv2d_igemm (K = 128)
    ResnetBlock  d3_groupnorm_silu -> d3_conv2d_igemm (+temb bias) -> d3_groupnorm_silu ->
                 d3_conv2d_igemm (+ folded 1x1 shortcut K-slabs over the *virtual* concat, or identity residual)
    Attention    d3_groupnorm_silu(no SiLU) -> igemm 1x1 (fused QKV) -> d3_attention_hd8 -> igemm 1x1 (+residual)
    Down/Up      igemm stride 2 / d3_upsample_nearest2x + igemm
    head         d3_conv_out_fused (GroupNorm + SiLU + conv 128 -> dc, fp32 NCHW out)
    time         d3_time_embed + d3_temb_project_all (all 32 ResNet projections in one launch)

Activations are 16-bit NHWC - fp16 by default (``compute_dtype="fp16"``: per-step max-abs error 1.2-1.6e-3 against
the fp32 oracle, inside the north-star bound of 1e-2), bf16 on request (same tensor-core rate, 1.0-1.25e-2) - or fp32
(``compute_dtype="fp32"``, CUDA-core reference-precision path); accumulation, GroupNorm statistics and the model
output are fp32/fp64.
There is no PyTorch fallback: without the CUDA library every call raises.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import ops

BLOCK_OUT = (128, 128, 256, 256, 512, 512)
DOWN_TYPES = ("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D")
UP_TYPES = ("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4


class UNet2DOutput(SimpleNamespace):
    """``.sample`` like diffusers' ``UNet2DOutput``."""


class _Act:
    """A bf16 NHWC activation together with the fp64 quad statistics of its producer."""

    __slots__ = ("t", "stats")

    def __init__(self, t: torch.Tensor, stats: Optional[torch.Tensor]):
        self.t, self.stats = t, stats


class _StatsArena:
    """One zero-initialised fp64 buffer per forward from which every producer takes its [B, C/4, 2] slice
    (a single memset per UNet forward instead of one per tensor)."""

    def __init__(self, device, nbytes_doubles: int):
        self.buf = torch.zeros(nbytes_doubles, dtype=torch.float64, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, B: int, C: int) -> torch.Tensor:
        n = B * (C // 4) * 2
        if self.off + n > self.buf.numel():
            raise RuntimeError("stats arena exhausted")
        v = self.buf[self.off:self.off + n].view(B, C // 4, 2)
        self.off += n
        return v


def _resnet_specs():
    """(prefix, cin_hidden, cin_skip, cout, has_attn) in execution order + structural markers."""
    boc = BLOCK_OUT
    specs = []
    out_c = boc[0]
    for i, t in enumerate(DOWN_TYPES):
        in_c, out_c = out_c, boc[i]
        for j in range(2):
            specs.append(("down", i, j, in_c if j == 0 else out_c, 0, out_c, t.startswith("Attn")))
    specs.append(("mid", 0, 0, boc[-1], 0, boc[-1], True))
    specs.append(("mid", 0, 1, boc[-1], 0, boc[-1], False))
    rev = list(reversed(boc))
    out_c = rev[0]
    for i, t in enumerate(UP_TYPES):
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        for j in range(3):
            skip = in_c if j == 2 else out_c
            rin = prev if j == 0 else out_c
            specs.append(("up", i, j, rin, skip, out_c, t.startswith("Attn")))
    return specs


class UNet2DModel:
    def __init__(self, sample_size: Sequence[int] = (384, 640), in_channels: int = 8, out_channels: int = 1,
                 layers_per_block: int = 2, block_out_channels: Sequence[int] = BLOCK_OUT,
                 down_block_types: Sequence[str] = DOWN_TYPES, up_block_types: Sequence[str] = UP_TYPES,
                 compute_dtype: str = "fp16", **unused):
        if (tuple(block_out_channels) != BLOCK_OUT or tuple(down_block_types) != DOWN_TYPES
                or tuple(up_block_types) != UP_TYPES or layers_per_block != 2):
            raise NotImplementedError("d3roma_b200.UNet2DModel implements the architecture of train.py:788-810 only")
        if 9 * in_channels > 128 or out_channels not in (1, 3):
            raise NotImplementedError("in_channels must be <= 14 and out_channels 1 or 3")
        self.sample_size = list(sample_size) if not isinstance(sample_size, int) else sample_size
        self.config = SimpleNamespace(sample_size=self.sample_size, in_channels=in_channels,
                                      out_channels=out_channels, layers_per_block=2,
                                      block_out_channels=tuple(block_out_channels),
                                      down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                                      attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5)
        self.in_channels, self.out_channels = in_channels, out_channels
        if compute_dtype not in ("bf16", "fp16", "fp32"):
            raise ValueError("compute_dtype must be 'bf16', 'fp16' (tcgen05 tensor-core paths) or 'fp32' "
                             "(CUDA-core reference-precision path)")
        self.compute_dtype = compute_dtype
        self._adt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[compute_dtype]
        # per-loop conv_in conditioning term: fp16 in both 16-bit modes (O(1) values, 11 significant bits - finer than
        # the 16-bit activation conv_in stores; 4 instead of 6 bytes per element and step, SURVEY 8d conv_in bytes)
        self._base_dtype = torch.float16
        self.device = torch.device("cpu")
        self.dtype = torch.float32  # dtype of the pipeline-facing tensors
        self._w: Dict[str, torch.Tensor] = {}
        self._specs = _resnet_specs()
        self._arena: Optional[_StatsArena] = None         # arena of the forward in flight
        self._eager_arena: Optional[_StatsArena] = None   # arena of the non-graph path (grown on demand)
        self._loaded = False
        self.identity_slab = os.environ.get("D3_UNET_IDENTITY_SLAB", "1") != "0"   # A/B switch, see load_state_dict
        # batch * H * W up to which the loop's kernels are launched with programmatic dependent launch (None: leave the
        # library setting alone, e.g. when D3_PDL is being A/B-tested)
        self.pdl_max_pixels = None if "D3_PDL" in os.environ else 2 * 384 * 640
        self.keep_state_dict = True   # keep a host copy of the loaded fp32 state_dict (for save_pretrained)
        self._sd_cpu = None
        self._graphs: Dict[tuple, dict] = {}
        self.graph_kernel_launches = 0  # kernels executed through CUDA-graph replays (for launch accounting)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """Pack a diffusers-keyed ``state_dict`` into the layouts the kernels read."""
        if not torch.cuda.is_available():
            raise RuntimeError("d3roma_b200.UNet2DModel needs a CUDA device (no CPU fallback)")
        dev = self.device if self.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        sd = {self._modern_key(k): v.detach() for k, v in sd.items()}
        used = set()

        def get(k):
            used.add(k)
            return sd[k].to(dev, torch.float32)

        W = self._w
        # conv_in as a 1x1 igemm over the 128-wide im2col row: [128][tap*Cin + c], zero padded
        w_in = get("conv_in.weight")
        cin = w_in.shape[1]
        assert cin == self.in_channels, "conv_in.weight does not match in_channels"
        w_in = w_in.permute(0, 2, 3, 1).reshape(128, 9 * cin)
        W["conv_in.w"] = torch.nn.functional.pad(w_in, (0, 128 - 9 * cin)).to(self._adt).contiguous()
        W["conv_in.b"] = get("conv_in.bias").contiguous()
        W["conv_in.oihw"] = get("conv_in.weight").contiguous()   # fp32, for the split (direct) conv_in
        self._cin_packs: Dict[int, tuple] = {}
        self._cond_base_cache = None
        # time MLP (transposed fp32)
        W["time.w1t"] = get("time_embedding.linear_1.weight").t().contiguous()
        W["time.b1"] = get("time_embedding.linear_1.bias").contiguous()
        W["time.w2t"] = get("time_embedding.linear_2.weight").t().contiguous()
        W["time.b2"] = get("time_embedding.linear_2.bias").contiguous()

        proj_w, proj_b, off = [], [], 0
        self._temb_off: Dict[str, int] = {}
        self._identity_slab = set()
        for kind, i, j, cin_h, cin_s, cout, attn in self._specs:
            p = self._rprefix(kind, i, j)
            W[p + ".n1.g"], W[p + ".n1.b"] = get(p + ".norm1.weight").contiguous(), get(p + ".norm1.bias").contiguous()
            W[p + ".n2.g"], W[p + ".n2.b"] = get(p + ".norm2.weight").contiguous(), get(p + ".norm2.bias").contiguous()
            W[p + ".w1"] = ops.pack_conv_weight(get(p + ".conv1.weight"), dtype=self._adt)
            b2 = get(p + ".conv2.bias")
            if cin_h + cin_s != cout:
                W[p + ".w2"] = ops.pack_conv_weight(get(p + ".conv2.weight"), get(p + ".conv_shortcut.weight"), dtype=self._adt)
                b2 = b2 + get(p + ".conv_shortcut.bias")
            elif cout == 128 and self._adt != torch.float32 and self.identity_slab:
                # identity residual of the 128-channel layers as two extra K-blocks with an identity "shortcut" weight:
                # exact (x * 1.0 accumulated in fp32) and 11 % more MMA work, where the residual read-back in the
                # pixels-as-N epilogue measured +34 % (profiles/r01_conv_experiments.txt)
                W[p + ".w2"] = ops.pack_conv_weight(get(p + ".conv2.weight"), torch.eye(128, device=dev),
                                                    dtype=self._adt)
                self._identity_slab.add(p)
            else:
                W[p + ".w2"] = ops.pack_conv_weight(get(p + ".conv2.weight"), dtype=self._adt)
            W[p + ".b2"] = b2.contiguous()
            proj_w.append(get(p + ".time_emb_proj.weight"))
            proj_b.append(get(p + ".time_emb_proj.bias") + get(p + ".conv1.bias"))
            self._temb_off[p] = off
            off += cout
            if attn:
                a = self._aprefix(kind, i, j)
                W[a + ".gn.g"], W[a + ".gn.b"] = get(a + ".group_norm.weight").contiguous(), get(a + ".group_norm.bias").contiguous()
                wq, wk, wv = get(a + ".to_q.weight"), get(a + ".to_k.weight"), get(a + ".to_v.weight")
                W[a + ".wqkv"] = torch.cat([wq, wk, wv], 0).to(self._adt).contiguous()
                W[a + ".bqkv"] = torch.cat([get(a + ".to_q.bias"), get(a + ".to_k.bias"), get(a + ".to_v.bias")]).contiguous()
                W[a + ".wo"] = get(a + ".to_out.0.weight").to(self._adt).contiguous()
                W[a + ".bo"] = get(a + ".to_out.0.bias").contiguous()
        W["temb.w"] = torch.cat(proj_w, 0).contiguous()
        W["temb.b"] = torch.cat(proj_b, 0).contiguous()
        self._temb_rows = off
        for i in range(5):
            p = f"down_blocks.{i}.downsamplers.0.conv"
            W[p + ".w"], W[p + ".b"] = ops.pack_conv_weight(get(p + ".weight"), dtype=self._adt), get(p + ".bias").contiguous()
            p = f"up_blocks.{i}.upsamplers.0.conv"
            for ph, wp in enumerate(ops.pack_upsample_weights(get(p + ".weight"), dtype=self._adt)):
                W[p + f".w{ph}"] = wp
            W[p + ".b"] = get(p + ".bias").contiguous()
        W["out.g"], W["out.b"] = get("conv_norm_out.weight").contiguous(), get("conv_norm_out.bias").contiguous()
        w_out = get("conv_out.weight")  # [dc,128,3,3] -> [dc][9][128]
        assert w_out.shape[0] == self.out_channels
        W["out.w"] = w_out.permute(0, 2, 3, 1).reshape(self.out_channels, 9, 128).contiguous()
        W["out.bias"] = get("conv_out.bias").contiguous()
        if strict:
            missing = set(sd) - used
            if missing:
                raise KeyError(f"unexpected keys in state_dict: {sorted(missing)[:5]} ...")
        self._loaded = True
        self._sd_cpu = {k: v.to("cpu") for k, v in sd.items()} if self.keep_state_dict else None
        self.reset_graphs()   # captured graphs hold raw pointers into the previous weight tensors
        return self

    @staticmethod
    def _modern_key(k: str) -> str:
        """Checkpoints saved by diffusers < 0.14 name the attention linears query / key / value / proj_attn
        (SURVEY Appendix A.6); map them onto to_q / to_k / to_v / to_out.0."""
        if ".attentions." in k:
            for old, new in ((".query.", ".to_q."), (".key.", ".to_k."), (".value.", ".to_v."),
                             (".proj_attn.", ".to_out.0.")):
                if old in k:
                    return k.replace(old, new)
        return k

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """The diffusers-keyed fp32 state_dict this model was loaded from (kept on the host)."""
        if self._sd_cpu is None:
            raise RuntimeError("state_dict() needs keep_state_dict=True at load time")
        return dict(self._sd_cpu)

    # ---- diffusers-layout persistence (train.py:682,701 `pipeline.save_pretrained`, evaluate.py:314 `from_pretrained`)
    def save_pretrained(self, path):
        import json
        import os
        from safetensors.torch import save_file
        os.makedirs(path, exist_ok=True)
        cfg = dict(vars(self.config), _class_name="UNet2DModel", compute_dtype=self.compute_dtype)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2)
        save_file({k: v.contiguous() for k, v in self.state_dict().items()},
                  os.path.join(path, "diffusion_pytorch_model.safetensors"))

    @classmethod
    def from_pretrained(cls, path, compute_dtype: Optional[str] = None, device="cuda"):
        import json
        import os
        with open(os.path.join(path, "config.json")) as f:
            cfg = json.load(f)
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}   # _class_name, _diffusers_version, _name_or_path
        if compute_dtype is not None:
            cfg["compute_dtype"] = compute_dtype
        net = cls(**cfg).to(device)
        st = os.path.join(path, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "diffusion_pytorch_model.bin"), map_location="cpu")
        return net.load_state_dict(sd)

    @staticmethod
    def _rprefix(kind, i, j):
        return {"down": f"down_blocks.{i}.resnets.{j}", "mid": f"mid_block.resnets.{j}",
                "up": f"up_blocks.{i}.resnets.{j}"}[kind]

    @staticmethod
    def _aprefix(kind, i, j):
        return {"down": f"down_blocks.{i}.attentions.{j}", "mid": "mid_block.attentions.0",
                "up": f"up_blocks.{i}.attentions.{j}"}[kind]

    def to(self, device=None, dtype=None):
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda":
                raise RuntimeError("d3roma_b200.UNet2DModel runs on CUDA only")
            if device.index is None:
                device = torch.device("cuda", torch.cuda.current_device())
            if self._loaded and device != self.device:
                self._w = {k: v.to(device) for k, v in self._w.items()}
                self._arena = self._eager_arena = None
                self._cin_packs, self._cond_base_cache = {}, None
                self.reset_graphs()
            self.device = device
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(self._w.values())

    # ------------------------------------------------------------------ forward pieces
    def _conv(self, x, w, b, ksize, B, stride=1, sc1=None, sc2=None, residual=None, want_stats=True):
        cout = w.shape[0]
        stats = self._arena.take(B, cout) if want_stats else None
        out = ops.conv2d_igemm(x, w, b, ksize=ksize, stride=stride, sc1=sc1, sc2=sc2, residual=residual, stats=stats)
        return _Act(out, stats)

    def _resnet(self, p: str, h: _Act, skip: Optional[_Act], cout: int, bias1: torch.Tensor, B: int,
                up_copy=None, pre: Optional[torch.Tensor] = None) -> _Act:
        """One ResnetBlock2D.  ``up_copy = (p_up, Cx, Cs)``: ``h`` is a skip tensor that the up-path resnet ``p_up``
        will consume as the second half of its concatenated norm1 input - norm1 here also writes that normalised copy
        (returned through ``self._pre_out``).  ``pre``: the concat buffer whose skip half was filled that way; only
        the hidden-state half is normalised now."""
        W = self._w
        if pre is not None:
            Cx = h.t.shape[3]
            cpg = pre.shape[3] // 32
            ops.groupnorm_silu(h.t, h.stats, W[p + ".n1.g"][:Cx], W[p + ".n1.b"][:Cx], groups=Cx // cpg,
                               out=pre[..., :Cx])
            a1 = pre
        elif up_copy is not None:
            p_up, Cx, Cs = up_copy
            Bn, Hh, Ww, _ = h.t.shape
            buf = torch.empty((Bn, Hh, Ww, Cx + Cs), dtype=h.t.dtype, device=h.t.device)
            cpg = (Cx + Cs) // 32
            a1 = ops.groupnorm_silu_dual(h.t, h.stats, W[p + ".n1.g"], W[p + ".n1.b"], W[p_up + ".n1.g"][Cx:],
                                         W[p_up + ".n1.b"][Cx:], buf[..., Cx:], groups_b=Cs // cpg)
            self._pre_out = buf
        else:
            a1 = ops.groupnorm_silu(h.t, h.stats, W[p + ".n1.g"], W[p + ".n1.b"],
                                    x2=None if skip is None else skip.t, stats2=None if skip is None else skip.stats)
        h1 = self._conv(a1, W[p + ".w1"], bias1, 3, B)
        del a1
        a2 = ops.groupnorm_silu(h1.t, h1.stats, W[p + ".n2.g"], W[p + ".n2.b"])
        cin = h.t.shape[3] + (0 if skip is None else skip.t.shape[3])
        if cin != cout:
            return self._conv(a2, W[p + ".w2"], W[p + ".b2"], 3, B, sc1=h.t, sc2=None if skip is None else skip.t)
        if p in self._identity_slab:
            return self._conv(a2, W[p + ".w2"], W[p + ".b2"], 3, B, sc1=h.t)
        return self._conv(a2, W[p + ".w2"], W[p + ".b2"], 3, B, residual=h.t)

    def _up_consumer(self, k: int):
        """(prefix, Cx, Cs) of the up-path resnet that pops skip ``k`` (of 18), or None when its norm1 groups would
        straddle the hidden/skip boundary (C = 384, 768) or the fp32 path is active."""
        if self._adt == torch.float32:
            return None
        kind, bi, bj, cin_h, cin_s, cout, attn = self._specs[14 + (17 - k)]   # 12 down + 2 mid resnets precede
        assert kind == "up"
        cpg = (cin_h + cin_s) // 32
        if cin_h % cpg or cin_s % cpg or (cin_h + cin_s) % 32:
            return None
        return (self._rprefix(kind, bi, bj), cin_h, cin_s)

    def _attention(self, a: str, h: _Act, B: int) -> _Act:
        W = self._w
        C = h.t.shape[3]
        n = ops.groupnorm_silu(h.t, h.stats, W[a + ".gn.g"], W[a + ".gn.b"], silu=False)
        qkv = self._conv(n, W[a + ".wqkv"], W[a + ".bqkv"], 1, B, want_stats=False)
        o = ops.attention_hd8(qkv.t, C)
        return self._conv(o, W[a + ".wo"], W[a + ".bo"], 1, B, residual=h.t)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def _cin_pack(self, Cs: int):
        """(sample-part, conditioning-part) fp32 weight packs of conv_in for ``Cs`` sample channels."""
        pk = self._cin_packs.get(Cs)
        if pk is None:
            w = self._w["conv_in.oihw"]
            pk = (ops.pack_conv_in_direct(w, 0, Cs), ops.pack_conv_in_direct(w, Cs, w.shape[1]))
            self._cin_packs[Cs] = pk
        return pk

    @torch.no_grad()
    def cond_base(self, cond: torch.Tensor, Cs: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``conv3x3(cond; W_in[:, Cs:]) + bias`` as fp32 NHWC [B,H,W,128]: the part of conv_in that does not change
        over the steps of a sampling loop.  Cached on (storage, version, shape) of ``cond`` so that a loop calling
        ``forward_split`` with the same conditioning tensor computes it once."""
        key = (cond.data_ptr(), cond._version, tuple(cond.shape), Cs)
        if out is None and self._cond_base_cache is not None and self._cond_base_cache[0] == key:
            return self._cond_base_cache[1]
        base = ops.conv_in_direct(cond, self._cin_pack(Cs)[1], bias=self._w["conv_in.b"], out=out,
                                  out_dtype=self._base_dtype)
        if out is None:
            self._cond_base_cache = (key, base, cond)  # keeps `cond` alive so the storage address cannot be reused
        return base

    @torch.no_grad()
    def time_bias(self, timestep, device=None, batch: int = 1) -> torch.Tensor:
        """Timestep -> the fused per-ResNet bias rows ``[nt, rows]`` (sinusoidal embedding, 2-layer MLP, all 32
        ``time_emb_proj`` + conv1 biases).  Depends on the scalar ``t`` only: a sampling loop can compute the rows of all
        its steps before the loop and hand them to ``forward_split(tb=...)``."""
        dev = self.device if device is None else device
        W = self._w
        if torch.is_tensor(timestep):
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
        else:
            t = torch.tensor([float(timestep)], dtype=torch.float32, device=dev)
        if t.numel() not in (1, batch):
            raise ValueError("timestep must be a scalar or have one entry per sample")
        semb = ops.time_embed(t, W["time.w1t"], W["time.b1"], W["time.w2t"], W["time.b2"])
        return ops.temb_project_all(semb, W["temb.w"], W["temb.b"])  # [nt, rows]

    @torch.no_grad()
    def forward_split(self, sample: torch.Tensor, cond: Optional[torch.Tensor], timestep,
                      cond_base: Optional[torch.Tensor] = None, arena: Optional[_StatsArena] = None,
                      tb: Optional[torch.Tensor] = None, ddim=None):
        """UNet forward on ``cat([sample, cond], 1)`` without materialising the concat.
        sample fp32 [B,dc,H,W], cond fp32 [B,Cc,H,W] (contiguous, CUDA); returns fp32 [B,out_channels,H,W].
        ``cond_base``: precomputed ``self.cond_base(cond, Cs)`` (the CUDA-graph path passes its static buffer).
        ``arena``: the GroupNorm statistics arena to use (a captured graph owns its own, so that a later call with a
        larger batch can never free memory a live graph still writes to).
        ``ddim``: a ``d3_ddim_coeffs`` struct (``MyDDIMScheduler.step_coeffs``) - the DDIM update of ``sample`` (eta = 0)
        is applied to the model output inside the last kernel and ``(prev_sample, pred_original_sample)`` is returned
        instead of the model output (bit-identical to ``scheduler.step`` on it)."""
        if not self._loaded:
            raise RuntimeError("UNet2DModel: load_state_dict() has not been called")
        B, Cs, H, Wd = sample.shape
        Cc = 0 if cond is None else cond.shape[1]
        if Cs + Cc != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {Cs}+{Cc}")
        if H % 32 or Wd % 32:
            raise ValueError(f"H and W must be multiples of 32 (got {H}x{Wd}); pad with InputPadder(divis_by=32)")
        dev = sample.device
        W = self._w
        # programmatic dependent launch pays in the latency regime only (c1: +2 %; batch 16: -2 %, profiles/r02_*)
        if self.pdl_max_pixels is not None:
            from . import _lib
            _lib.lib().d3_set_pdl(1 if B * H * Wd <= self.pdl_max_pixels else 0)
        need = B * 2 * 40000
        if arena is not None:
            if arena.buf.numel() < need or arena.buf.device != dev:
                raise RuntimeError("forward_split: the supplied statistics arena is too small for this batch")
            self._arena = arena
        elif (self._eager_arena is None or self._eager_arena.buf.numel() < need
              or self._eager_arena.buf.device != dev):
            self._arena = self._eager_arena = _StatsArena(dev, need)
        else:
            self._arena = self._eager_arena
        self._arena.reset()

        # timestep -> fused per-ResNet bias vectors (``tb``: precomputed by the caller with time_bias())
        if tb is None:
            tb = self.time_bias(timestep, dev, B)

        def bias1(p, cout):
            o = self._temb_off[p]
            return tb[:, o:o + cout] if tb.shape[0] > 1 else tb[0, o:o + cout]

        if cond is not None and Cs <= 3 and self._adt != torch.float32:
            # split conv_in: per-loop conditioning term (cached) + the 9*Cs taps of the noisy sample, all in fp32
            if cond_base is None:
                cond_base = self.cond_base(cond, Cs)
            st_in = self._arena.take(B, 128)
            h = _Act(ops.conv_in_direct(sample, self._cin_pack(Cs)[0], base=cond_base, out_dtype=self._adt,
                                        stats=st_in), st_in)
        else:
            col = ops.im2col_input(sample, cond, dtype=self._adt)
            h = self._conv(col, W["conv_in.w"], W["conv_in.b"], 1, B)
            del col
        skips: List[_Act] = [h]
        pre: Dict[int, torch.Tensor] = {}   # skip index -> concat buffer whose skip half is already normalised

        def resnet_on_skip(p, h, cout):
            """A resnet whose input is the most recent skip tensor: its norm1 pass also writes the up path's copy."""
            k = len(skips) - 1
            upc = self._up_consumer(k)
            self._pre_out = None
            out = self._resnet(p, h, None, cout, bias1(p, cout), B, up_copy=upc)
            if self._pre_out is not None:
                pre[k] = self._pre_out
                self._pre_out = None
            return out

        it = iter(self._specs)
        for i in range(6):
            for j in range(2):
                kind, bi, bj, cin_h, cin_s, cout, attn = next(it)
                p = self._rprefix(kind, bi, bj)
                h = resnet_on_skip(p, h, cout)     # h is skips[-1] here (conv_in / previous resnet / downsample)
                if attn:
                    h = self._attention(self._aprefix(kind, bi, bj), h, B)
                skips.append(h)
            if i < 5:
                p = f"down_blocks.{i}.downsamplers.0.conv"
                h = self._conv(h.t, W[p + ".w"], W[p + ".b"], 3, B, stride=2)
                skips.append(h)
        for j in range(2):
            kind, bi, bj, cin_h, cin_s, cout, attn = next(it)
            p = self._rprefix(kind, bi, bj)
            h = resnet_on_skip(p, h, cout) if j == 0 else self._resnet(p, h, None, cout, bias1(p, cout), B)
            if attn:
                h = self._attention(self._aprefix(kind, bi, bj), h, B)
        for i in range(6):
            for j in range(3):
                kind, bi, bj, cin_h, cin_s, cout, attn = next(it)
                p = self._rprefix(kind, bi, bj)
                k = len(skips) - 1
                h = self._resnet(p, h, skips.pop(), cout, bias1(p, cout), B, pre=pre.pop(k, None))
                if attn:
                    h = self._attention(self._aprefix(kind, bi, bj), h, B)
            if i < 5:
                # Upsample2D = nearest-2x + conv3x3, run as 4 sub-pixel phase convs on the low-res tensor
                p = f"up_blocks.{i}.upsamplers.0.conv"
                Bn, Hh, Ww, Cc = h.t.shape
                up = torch.empty((Bn, 2 * Hh, 2 * Ww, Cc), dtype=h.t.dtype, device=h.t.device)
                st = self._arena.take(B, Cc)
                for ph in range(4):
                    ops.conv2d_igemm(h.t, W[p + f".w{ph}"], W[p + ".b"], ksize=3, out=up, stats=st,
                                     up2x_phase=ph + 1)
                h = _Act(up, st)
        return ops.conv_out_fused(h.t, h.stats, W["out.g"], W["out.b"], W["out.w"], W["out.bias"], self.out_channels,
                                  ddim=None if ddim is None else (sample, ddim))

    @staticmethod
    def _cond_key(cond: Optional[torch.Tensor]):
        """Identity of a conditioning tensor's CONTENTS: storage address, in-place version counter and shape (the same
        key ``cond_base`` uses) - a caller that refills one preallocated buffer per frame bumps the version.  Whoever
        stores the key must also hold the tensor (an address is only unique among LIVE tensors)."""
        return None if cond is None else (cond.data_ptr(), cond._version, tuple(cond.shape))

    def _graph_for(self, sample: torch.Tensor, cond: Optional[torch.Tensor], tval: float) -> dict:
        """Capture (once per (B, H, W, Cs, Cc, device)) the UNet forward into a CUDA graph over static buffers.  The
        graph owns everything it touches: input / conditioning / timestep buffers, the conv_in conditioning term and
        its own GroupNorm statistics arena."""
        key = (tuple(sample.shape), None if cond is None else tuple(cond.shape), sample.device.index)
        g = self._graphs.get(key)
        if g is not None:
            return g
        B = sample.shape[0]
        g = {"x": torch.empty_like(sample), "c": None if cond is None else torch.empty_like(cond),
             "t": torch.zeros(1, dtype=torch.float32, device=sample.device),
             "arena": _StatsArena(sample.device, B * 2 * 40000)}
        g["x"].copy_(sample)
        if cond is not None:
            g["c"].copy_(cond)
        g["t"].fill_(tval)
        split = cond is not None and sample.shape[1] <= 3 and self._adt != torch.float32
        g["base"] = self.cond_base(g["c"], sample.shape[1], out=torch.empty(
            (B, sample.shape[2], sample.shape[3], 128), dtype=self._base_dtype, device=sample.device)) if split else None
        side = torch.cuda.Stream(device=sample.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            # warm-up: function attributes, allocator pools
            self.forward_split(g["x"], g["c"], g["t"], cond_base=g["base"], arena=g["arena"])
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        from . import _lib
        n0 = _lib.lib().d3_launch_count()
        with torch.cuda.graph(graph):
            g["y"] = self.forward_split(g["x"], g["c"], g["t"], cond_base=g["base"], arena=g["arena"])
        g["n_kernels"] = int(_lib.lib().d3_launch_count() - n0)
        g["graph"] = graph
        g["cond_key"], g["cond_ref"] = self._cond_key(cond), cond
        self._graphs[key] = g
        return g

    def _refresh_graph_cond(self, g: dict, cond: Optional[torch.Tensor], Cs: int) -> None:
        """Conditioning is constant across the steps of one sampling loop: copy it into the graph's static buffer (and
        recompute its conv_in term) only when its contents changed."""
        if cond is None:
            return
        key = self._cond_key(cond)
        if g["cond_key"] == key:
            return
        g["c"].copy_(cond)
        # the key names a storage address: hold the tensor so that the allocator cannot hand the same address (with a
        # fresh version counter) to a DIFFERENT conditioning tensor while the key is still trusted
        g["cond_key"], g["cond_ref"] = key, cond
        if g["base"] is not None:
            self.cond_base(g["c"], Cs, out=g["base"])

    @torch.no_grad()
    def forward_graphed(self, sample: torch.Tensor, cond: Optional[torch.Tensor], timestep) -> torch.Tensor:
        """Same as ``forward_split`` but replayed from a CUDA graph captured per (B, H, W, Cs, Cc): the ~180 kernel
        launches of one UNet forward become a single graph launch.  ``timestep`` must be one scalar for the batch.
        Returns a STATIC output buffer that the next call overwrites (the sampling loop consumes it immediately)."""
        tval = float(timestep)
        g = self._graph_for(sample, cond, tval)
        g["x"].copy_(sample)
        self._refresh_graph_cond(g, cond, sample.shape[1])
        g["t"].fill_(tval)
        g["graph"].replay()
        self.graph_kernel_launches += g["n_kernels"]
        return g["y"]

    @torch.no_grad()
    def prime_conditioning(self, sample_shape, cond: Optional[torch.Tensor], graphed: bool = True) -> None:
        """Launch the per-loop conditioning work (static-buffer copy + conv_in conditioning term) NOW, so that it runs
        on the GPU while the host is still busy (the pipeline calls this before drawing x_T on the CPU generators).
        The first forward of the loop then finds everything cached.  No-op when there is nothing to prepare."""
        if cond is None or self._adt == torch.float32 or sample_shape[1] > 3 or not self._loaded:
            return
        Cs = int(sample_shape[1])
        if graphed:
            g = self._graphs.get((tuple(sample_shape), tuple(cond.shape), cond.device.index))
            if g is not None:
                self._refresh_graph_cond(g, cond, Cs)
        else:
            self.cond_base(cond, Cs)

    def reset_graphs(self):
        self._graphs.clear()
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1   # invalidates pipeline-level whole-loop graphs

    @torch.no_grad()
    def __call__(self, sample: torch.Tensor, timestep, return_dict: bool = True):
        """``unet(model_input[B,Cin,H,W], t).sample`` (custom_pipelines.py:173; train.py:308 form with
        ``return_dict=False``)."""
        x = sample.to(device=self.device, dtype=torch.float32).contiguous()
        out = self.forward_split(x, None, timestep).to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNet2DOutput(sample=out)

    forward = __call__

"""Training step of the pixel-space denoiser on B200.

Reference: ``train_step`` in ``/root/reference/train.py:157-538`` - forward diffusion (``noise_scheduler.add_noise`` :278),
``model(inputs, timesteps)`` with one timestep per sample (:308), the masked / weighted MSE or L1 loss (:310-417),
``accelerator.backward(loss)`` (:455), ``clip_grad_norm_`` (:518), ``optimizer.step()`` (:524, ``torch.optim.AdamW``
:862) and the gradient all-reduce accelerate's DDP wrapper performs between ranks.

``UNetTrainer`` owns
  * one flat fp32 master buffer of all parameters, stored in the layouts the kernels read (conv weights as
    ``[Cout][tap*Cin + c]`` with the 1x1 shortcut appended as extra K columns), its 16-bit mirror for the tensor-core
    kernels, flat gradient / Adam moment buffers of the same layout, and the transposed ("data-gradient") weight packs;
  * a forward pass over the same sm_100a kernels as inference (``d3_conv2d_igemm`` with the quad statistics,
    ``d3_groupnorm_silu``, ``d3_attention_hd8``, ``d3_conv_out_fused``) that keeps every tensor the backward needs;
  * a backward pass that walks a tape of closures in reverse: data gradients of the convolutions are convolutions with
    flipped / transposed weights on ``d3_conv2d_igemm`` (stride-2 convs as four sub-pixel phases), weight gradients run on
    the split-K tcgen05 kernel ``d3_conv2d_wgrad``, everything else on the bandwidth kernels of ``csrc/train_ops.cu`` and
    ``csrc/attention_bwd.cu``;
  * gradient all-reduce over NCCL (``torch.distributed``; the only collective of data-parallel training), gradient-norm
    clipping and a fused AdamW step that also refreshes the 16-bit mirror.

There is no PyTorch autograd and no CPU fallback on this path; ``oracle/train_ref.py`` (autograd over the oracle UNet) is
the checker.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .unet import BLOCK_OUT, _Act, _resnet_specs, _StatsArena, UNet2DModel


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class UNetTrainer:
    def __init__(self, in_channels: int = 8, out_channels: int = 1, compute_dtype: str = "bf16", lr: float = 1e-4,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 1.0, loss_scale: float = 1.0, loss_type: str = "mse", device="cuda"):
        if compute_dtype not in ("bf16", "fp16"):
            raise ValueError("UNetTrainer: compute_dtype must be 'bf16' or 'fp16'")
        if 9 * in_channels > 128 or out_channels not in (1, 3):
            raise NotImplementedError("in_channels must be <= 14 and out_channels 1 or 3")
        if loss_type not in ("mse", "l1"):
            raise ValueError(f"loss type {loss_type} not supported")     # train.py:255-260
        self.in_channels, self.out_channels = in_channels, out_channels
        self.compute_dtype = compute_dtype
        self._adt = torch.bfloat16 if compute_dtype == "bf16" else torch.float16
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm, self.loss_scale, self.loss_type = max_grad_norm, float(loss_scale), loss_type
        self.overlap_allreduce = True   # data parallel: all-reduce the up path's gradients under the down path's backward
        self.device = torch.device(device)
        self.step_count = 0
        self._specs = _resnet_specs()
        self._build_layout()
        self._loaded = False
        self._tape: List = []
        self._grads: Dict[int, tuple] = {}
        self._consumed, self._bias_req = set(), {}
        self._tape_up_start = 0
        self._early_done = False
        self._early_reduce = None      # set by train_step for world_size > 1 (overlapped gradient all-reduce)
        self._reduce_stream = None
        self._early_work = None
        self.overlapped_reductions = 0   # steps whose up-path gradients were reduced under the backward

    # ------------------------------------------------------------------ parameter layout
    def _build_layout(self):
        L: Dict[str, Tuple[int, tuple]] = {}
        off = 0

        def add(name, *shape):
            nonlocal off
            n = 1
            for s in shape:
                n *= s
            L[name] = (off, tuple(shape))
            off = _align(off + n)

        add("conv_in.w", 128, 128)
        rows = 0
        self._up_range = [None, None]   # [start, end) of the up blocks' conv / attention weights in the flat buffers
        self._temb_off: Dict[str, int] = {}
        self._attn: List[Tuple[str, int]] = []
        for kind, i, j, cin_h, cin_s, cout, attn in self._specs:
            p = UNet2DModel._rprefix(kind, i, j)
            cin = cin_h + cin_s
            if kind == "up" and self._up_range[0] is None:
                self._up_range[0] = off
            add(p + ".w1", cout, 9 * cin)
            add(p + ".w2", cout, 9 * cout + (cin if cin != cout else 0))
            self._temb_off[p] = rows
            rows += cout
            if attn:
                a = UNet2DModel._aprefix(kind, i, j)
                add(a + ".wqkv", 3 * cout, cout)
                add(a + ".wo", cout, cout)
                self._attn.append((a, cout))
        self._up_range[1] = off
        for i in range(5):
            c = BLOCK_OUT[i]
            add(f"down_blocks.{i}.downsamplers.0.conv.w", c, 9 * c)
        rev = list(reversed(BLOCK_OUT))
        for i in range(5):
            c = rev[i]
            add(f"up_blocks.{i}.upsamplers.0.conv.w", c, 9 * c)
        self.n_tc = off                      # parameters mirrored in 16 bits for the tensor-core kernels
        self.rows_total = rows
        add("conv_in.b", 128)
        add("time.w1", 512, 128)
        add("time.b1", 512)
        add("time.w2", 512, 512)
        add("time.b2", 512)
        add("temb.w", rows, 512)             # all time_emb_proj weights, resnets in execution order
        add("temb.b", rows)                  # time_emb_proj.bias
        add("c1.b", rows)                    # conv1.bias
        add("c2.b", rows)                    # conv2.bias
        add("sc.b", rows)                    # conv_shortcut.bias (zero rows where a resnet has no shortcut)
        for kind, i, j, cin_h, cin_s, cout, attn in self._specs:
            p = UNet2DModel._rprefix(kind, i, j)
            add(p + ".n1.g", cin_h + cin_s)
            add(p + ".n1.b", cin_h + cin_s)
            add(p + ".n2.g", cout)
            add(p + ".n2.b", cout)
        for a, c in self._attn:
            add(a + ".gn.g", c)
            add(a + ".gn.b", c)
            add(a + ".bqkv", 3 * c)
            add(a + ".bo", c)
        for i in range(5):
            add(f"down_blocks.{i}.downsamplers.0.conv.b", BLOCK_OUT[i])
            add(f"up_blocks.{i}.upsamplers.0.conv.b", rev[i])
        add("out.g", 128)
        add("out.b", 128)
        add("out.w", self.out_channels, 9, 128)
        add("out.bias", self.out_channels)
        self.n_params = off
        self._layout = L

    def _view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self._layout[name]
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n].view(shape)

    def P(self, name):     # fp32 master
        return self._view(self._p, name)

    def G(self, name):     # fp32 gradient
        return self._view(self._g, name)

    def W16(self, name):   # 16-bit mirror (tensor-core weights only)
        return self._view(self._p16, name)

    # ------------------------------------------------------------------ weights in / out
    def _key_map(self):
        """(diffusers key, layout name, kind) triples; kind selects the (un)packing."""
        m = [("conv_in.weight", "conv_in.w", "conv_in"), ("conv_in.bias", "conv_in.b", "vec"),
             ("time_embedding.linear_1.weight", "time.w1", "vec"), ("time_embedding.linear_1.bias", "time.b1", "vec"),
             ("time_embedding.linear_2.weight", "time.w2", "vec"), ("time_embedding.linear_2.bias", "time.b2", "vec"),
             ("conv_norm_out.weight", "out.g", "vec"), ("conv_norm_out.bias", "out.b", "vec"),
             ("conv_out.weight", "out.w", "conv_out"), ("conv_out.bias", "out.bias", "vec")]
        for kind, i, j, cin_h, cin_s, cout, attn in self._specs:
            p = UNet2DModel._rprefix(kind, i, j)
            o = self._temb_off[p]
            m += [(p + ".norm1.weight", p + ".n1.g", "vec"), (p + ".norm1.bias", p + ".n1.b", "vec"),
                  (p + ".norm2.weight", p + ".n2.g", "vec"), (p + ".norm2.bias", p + ".n2.b", "vec"),
                  (p + ".conv1.weight", p + ".w1", "conv"), (p + ".conv2.weight", p + ".w2", "conv"),
                  (p + ".conv1.bias", ("c1.b", o, cout), "slice"), (p + ".conv2.bias", ("c2.b", o, cout), "slice"),
                  (p + ".time_emb_proj.weight", ("temb.w", o, cout), "slice"),
                  (p + ".time_emb_proj.bias", ("temb.b", o, cout), "slice")]
            if cin_h + cin_s != cout:
                m += [(p + ".conv_shortcut.weight", p + ".w2", "shortcut"),
                      (p + ".conv_shortcut.bias", ("sc.b", o, cout), "slice")]
            if attn:
                a = UNet2DModel._aprefix(kind, i, j)
                m += [(a + ".group_norm.weight", a + ".gn.g", "vec"), (a + ".group_norm.bias", a + ".gn.b", "vec"),
                      (a + ".to_out.0.weight", a + ".wo", "vec"), (a + ".to_out.0.bias", a + ".bo", "vec")]
                for r, n in enumerate(("to_q", "to_k", "to_v")):
                    m += [(a + f".{n}.weight", (a + ".wqkv", r * cout, cout), "slice"),
                          (a + f".{n}.bias", (a + ".bqkv", r * cout, cout), "slice")]
        for i in range(5):
            for side in ("down_blocks.%d.downsamplers.0.conv", "up_blocks.%d.upsamplers.0.conv"):
                p = side % i
                m += [(p + ".weight", p + ".w", "conv"), (p + ".bias", p + ".b", "vec")]
        return m

    def _slot(self, buf, name, kind, src_shape=None):
        """The view of ``buf`` a diffusers tensor maps to, shaped so that ``.copy_`` moves it (both directions)."""
        if kind == "slice":
            base, o, n = name
            return self._view(buf, base)[o:o + n]
        v = self._view(buf, name)
        if kind == "conv":       # [co, 9*ci (+ shortcut)] <- OIHW
            co = v.shape[0]
            ci = src_shape[1]
            return v[:, :9 * ci].view(co, 3, 3, ci).permute(0, 3, 1, 2)
        if kind == "shortcut":   # the trailing K columns of the conv2 pack
            co = v.shape[0]
            ci = src_shape[1]
            return v[:, 9 * co:9 * co + ci].view(co, ci, 1, 1)
        if kind == "conv_in":
            ci = self.in_channels
            return v[:, :9 * ci].view(128, 3, 3, ci).permute(0, 3, 1, 2)
        if kind == "conv_out":   # [dc, 9, 128] <- [dc, 128, 3, 3]
            return v.view(self.out_channels, 3, 3, 128).permute(0, 3, 1, 2)
        return v.view(src_shape) if src_shape is not None else v

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """diffusers-keyed fp32 ``state_dict`` -> flat master buffer in kernel layouts."""
        if not torch.cuda.is_available() and self.device.type == "cuda":
            raise RuntimeError("d3roma_b200.UNetTrainer needs a CUDA device (no CPU fallback)")
        dev = self.device
        sd = {UNet2DModel._modern_key(k): v for k, v in sd.items()}
        self._p = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        used = set()
        for key, name, kind in self._key_map():
            src = sd[key].detach().to(dev, torch.float32)
            self._slot(self._p, name, kind, tuple(src.shape)).copy_(src)
            used.add(key)
        extra = set(sd) - used
        if extra:
            raise KeyError(f"unexpected keys in state_dict: {sorted(extra)[:5]} ...")
        self._g = torch.zeros_like(self._p)
        self._m = torch.zeros_like(self._p)
        self._v = torch.zeros_like(self._p)
        self._p16 = torch.zeros(self.n_tc, dtype=self._adt, device=dev)
        self._derived = torch.zeros(2 * self.rows_total, dtype=torch.float32, device=dev)
        self._zero_bias = torch.zeros(2048, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._arena: Optional[_StatsArena] = None
        self._alloc_dgrad_packs()
        ops.cast_f32_to_16(self._p[:self.n_tc], self._p16)
        self._refresh_derived()
        self.step_count = 0
        self._loaded = True
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """The current parameters under diffusers key names (fp32, on the host)."""
        out = {}
        for key, name, kind in self._key_map():
            shape = None
            if kind == "conv":
                co, k = self._layout[name][1]
                ci = (k // 9) if not name.endswith(".w2") else co
                shape = (co, ci, 3, 3)
            elif kind == "shortcut":
                co, k = self._layout[name][1]
                shape = (co, k - 9 * co, 1, 1)
            out[key] = self._slot(self._p, name, kind, shape).detach().to("cpu").contiguous().clone()
        return out

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        """The gradients of the last backward under diffusers key names (fp32, host; loss scale removed)."""
        out = {}
        for key, name, kind in self._key_map():
            shape = None
            if kind == "conv":
                co, k = self._layout[name][1]
                ci = (k // 9) if not name.endswith(".w2") else co
                shape = (co, ci, 3, 3)
            elif kind == "shortcut":
                co, k = self._layout[name][1]
                shape = (co, k - 9 * co, 1, 1)
            out[key] = (self._slot(self._g, name, kind, shape).detach() / self.loss_scale).to("cpu").contiguous().clone()
        return out

    # ------------------------------------------------------------------ derived tensors
    def _refresh_derived(self):
        """Merged per-channel biases the forward kernels read: time_emb_proj.bias + conv1.bias, conv2.bias +
        conv_shortcut.bias (one contiguous row per kind, resnets in execution order)."""
        r = self.rows_total
        self._derived[:r] = ops.lincomb(1.0, self.P("temb.b"), 1.0, self.P("c1.b"))
        self._derived[r:] = ops.lincomb(1.0, self.P("c2.b"), 1.0, self.P("sc.b"))
        self._refresh_dgrad_packs()

    def _dgrad_entries(self):
        """(key, master name, col0, Cin, Cout, tap map) of every transposed weight pack the backward needs."""
        flip = [8 - j for j in range(9)]
        ent = []
        for kind, i, j, cin_h, cin_s, cout, attn in self._specs:
            p = UNet2DModel._rprefix(kind, i, j)
            cin = cin_h + cin_s
            ent.append((p + ".w1", p + ".w1", 0, cin, cout, flip))
            ent.append((p + ".w2", p + ".w2", 0, cout, cout, flip))
            if cin != cout:
                ent.append((p + ".w2.sc", p + ".w2", 9 * cout, cin, cout, [0]))
            if attn:
                a = UNet2DModel._aprefix(kind, i, j)
                ent.append((a + ".wqkv", a + ".wqkv", 0, cout, 3 * cout, [0]))
                ent.append((a + ".wo", a + ".wo", 0, cout, cout, [0]))
        for i in range(5):
            n = f"down_blocks.{i}.downsamplers.0.conv.w"
            c = BLOCK_OUT[i]
            # stride-2 conv: dx[2i+py, 2j+px] gathers dy rows {i+py-1, i+py} (a = 0, 1) through filter rows
            # py = 0: {-, 1};  py = 1: {2, 0}   (same for columns) - the sub-pixel phase form of d3_conv2d_igemm
            sel = {0: (-1, 1), 1: (2, 0)}
            for py in (0, 1):
                for px in (0, 1):
                    tm = []
                    for a_ in (0, 1):
                        for b_ in (0, 1):
                            kh, kw = sel[py][a_], sel[px][b_]
                            tm.append(-1 if kh < 0 or kw < 0 else kh * 3 + kw)
                    ent.append((n + f".ph{2 * py + px}", n, 0, c, c, tm))
            n = f"up_blocks.{i}.upsamplers.0.conv.w"
            c = list(reversed(BLOCK_OUT))[i]
            ent.append((n, n, 0, c, c, flip))
        return ent

    def _alloc_dgrad_packs(self):
        self._dg: Dict[str, torch.Tensor] = {}
        total = 0
        ent = self._dgrad_entries()
        for key, name, col0, cin, cout, tm in ent:
            total += _align(cin * len(tm) * cout)
        self._dg_buf = torch.zeros(total, dtype=self._adt, device=self.device)
        off = 0
        table = []
        for key, name, col0, cin, cout, tm in ent:
            n = cin * len(tm) * cout
            self._dg[key] = self._dg_buf[off:off + n].view(cin, len(tm) * cout)
            w_off, w_shape = self._layout[name]
            table.append((w_off, w_shape[1], off, col0, cin, cout, tm))
            off = _align(off + n)
        self._dg_entries = ent
        self._dg_table = ops.make_pack_table(table, self.device)

    def _refresh_dgrad_packs(self):
        """Every transposed weight pack from the fp32 master buffer, one launch (``d3_pack_taps16_batched``)."""
        ops.pack_taps16_batched(self._dg_table, self._p, self._dg_buf)

    # ------------------------------------------------------------------ gradient bookkeeping
    # Every forward tensor collects the gradient contributions of its consumers as a list of pending 16-bit tensors
    # (channel-slice views allowed).  They are summed as late as possible: a GroupNorm backward folds up to three of them
    # into its own pass, a data-gradient conv one (its `residual` input); only what exceeds that costs an add kernel.
    def _put(self, t, g):
        self._grads[id(t)] = (t, [g])

    def _acc(self, t, g):
        self._grads.setdefault(id(t), (t, []))[1].append(g)

    def _pending(self, t, keep: int, colsum=None):
        e = self._grads.pop(id(t), None)
        lst = [] if e is None else e[1]
        while len(lst) > keep:
            b, a = lst.pop(), lst.pop()
            last = len(lst) + 1 == keep          # the add that leaves `keep` tensors
            lst.append(ops.add_nhwc16(a, b, colsum=colsum if (last and keep == 1) else None))
        return lst

    def _take(self, t, colsum=None):
        """The summed gradient of ``t``.  ``colsum``: a dict ``{"bg": fp32 view, "done": bool}`` - when a final add is
        needed anyway, the channel sums of the result (a conv's bias gradient) are taken in that pass."""
        n = len(self._grads.get(id(t), (None, []))[1])
        if colsum is not None and n >= 2 and not colsum["done"] and colsum["bg"] is not None:
            lst = self._pending(t, 1, colsum=colsum["bg"])
            colsum["done"] = True
        else:
            lst = self._pending(t, 1)
        return lst[0] if lst else None

    def _consume(self, t) -> bool:
        """Mark a forward tensor as consumed; True for its FIRST consumer (whose backward therefore runs last)."""
        first = id(t) not in self._consumed
        self._consumed.add(id(t))
        return first

    # ------------------------------------------------------------------ forward pieces (each records its backward)
    def _gn(self, x1: _Act, gname: str, bname: str, x2: Optional[_Act] = None, silu: bool = True, colsum=None,
            holder: Optional[dict] = None):
        """GroupNorm(+SiLU) over ``x1`` (and ``x2``: the virtual concat of an up block).  ``holder``: a dict in which the
        same resnet's conv2 backward leaves ``"dsc"``, the gradient its 1x1 shortcut sends to the (concatenated) input -
        a dense tensor over exactly the channels of this GroupNorm's input, so it rides into dx as an addend instead of
        being split and summed per source."""
        g, b = self.P(gname), self.P(bname)
        a = ops.groupnorm_silu(x1.t, x1.stats, g, b, x2=None if x2 is None else x2.t,
                               stats2=None if x2 is None else x2.stats, silu=silu)
        first = self._consume(x1.t)
        if x2 is not None:
            self._consume(x2.t)
        # the first consumer's backward runs last: with a single source its dx (+ everything pending) IS the total
        # gradient of x1, so the bias gradient of the conv that produced x1 (channel sums) is taken in the same pass
        req = self._bias_req.get(id(x1.t)) if (first and colsum is None) else None

        def bwd():
            da = self._take(a)
            extra = holder.pop("dsc", None) if holder else None
            if x2 is None:
                addends = self._pending(x1.t, 3 if extra is None else 2)
            else:
                addends = []
            if extra is not None:
                addends = list(addends) + [extra]
            cs = colsum
            # single source: dx + the addends is x1's total gradient.  Concat: true for the x1 part when the only other
            # consumer of x1 was this resnet's shortcut (folded in through `extra`) - nothing else may be pending
            if req is not None and (x2 is None or (extra is not None and id(x1.t) not in self._grads)):
                cs, req["done"] = req["bg"], True
            dx = ops.groupnorm_silu_bwd(x1.t, x1.stats, g, b, da, x2=None if x2 is None else x2.t,
                                        stats2=None if x2 is None else x2.stats, silu=silu, addends=addends,
                                        dgamma=self.G(gname), dbeta=self.G(bname), colsum=cs)
            if x2 is None:
                self._put(x1.t, dx)
            else:
                c1 = x1.t.shape[3]
                self._acc(x1.t, dx[..., :c1])
                self._acc(x2.t, dx[..., c1:])
        self._tape.append(bwd)
        return a

    def _conv(self, x: torch.Tensor, wname: str, bias: torch.Tensor, ksize: int, B: int, *, stride: int = 1, sc=(),
              residual: Optional[torch.Tensor] = None, want_stats: bool = True, bias_grad: Optional[torch.Tensor] = None,
              need_dx: bool = True, sc_holder: Optional[dict] = None) -> _Act:
        w = self.W16(wname)
        cout = w.shape[0]
        stats = self._arena.take(B, cout) if want_stats else None
        sc = list(sc)
        out = ops.conv2d_igemm(x, w, bias, ksize=ksize, stride=stride, sc1=sc[0] if len(sc) > 0 else None,
                               sc2=sc[1] if len(sc) > 1 else None, residual=residual, stats=stats)
        for t_ in (x, residual, *sc):
            if t_ is not None:
                self._consume(t_)
        req = {"bg": bias_grad, "done": False}
        if bias_grad is not None:
            self._bias_req[id(out)] = req      # a GroupNorm that consumes `out` first may take the channel sums for us

        def bwd():
            dy = self._take(out, colsum=req)
            cin = x.shape[3]
            gw = self.G(wname)
            if bias_grad is not None and not req["done"]:
                ops.colsum_nhwc16(dy, bias_grad)
            ops.conv2d_wgrad(x, dy, gw, ksize=ksize, stride=stride, col0=0)
            col = ksize * ksize * cin
            for s in sc:
                ops.conv2d_wgrad(s, dy, gw, ksize=1, col0=col)
                col += s.shape[3]
            if need_dx:
                if stride == 1:
                    dx = ops.conv2d_igemm(dy, self._dg[wname], self._zero_bias[:cin], ksize=ksize, residual=self._take(x))
                    self._put(x, dx)
                else:
                    Bn, Ho, Wo, _ = dy.shape
                    dx = torch.empty((Bn, 2 * Ho, 2 * Wo, cin), dtype=dy.dtype, device=dy.device)
                    for ph in range(4):
                        ops.conv2d_igemm(dy, self._dg[wname + f".ph{ph}"], self._zero_bias[:cin], ksize=3, out=dx,
                                         up2x_phase=ph + 1)
                    self._acc(x, dx)
            if sc:
                csc = sum(s.shape[3] for s in sc)
                dsc = ops.conv2d_igemm(dy, self._dg[wname + ".sc"], self._zero_bias[:csc], ksize=1)
                if sc_holder is not None:
                    sc_holder["dsc"] = dsc       # picked up by the resnet's norm1 backward (same channels, same order)
                else:
                    o = 0
                    for s in sc:
                        self._acc(s, dsc[..., o:o + s.shape[3]])
                        o += s.shape[3]
            if residual is not None:
                self._acc(residual, dy)
        self._tape.append(bwd)
        return _Act(out, stats)

    def _resnet(self, p: str, h: _Act, skip: Optional[_Act], cout: int, rows: torch.Tensor, B: int) -> _Act:
        o = self._temb_off[p]
        r = self.rows_total
        cin = h.t.shape[3] + (0 if skip is None else skip.t.shape[3])
        holder = {} if cin != cout else None
        a1 = self._gn(h, p + ".n1.g", p + ".n1.b", x2=skip, holder=holder)
        h1 = self._conv(a1, p + ".w1", rows[:, o:o + cout], 3, B)
        # norm2's backward yields conv1's output gradient: its per-sample channel sums are the time-embedding gradient
        a2 = self._gn(h1, p + ".n2.g", p + ".n2.b", colsum=self._d_rows[:, o:o + cout])
        b2 = self._derived[r + o:r + o + cout]
        bg = self.G("c2.b")[o:o + cout]
        if cin != cout:
            sc = [h.t] + ([] if skip is None else [skip.t])
            return self._conv(a2, p + ".w2", b2, 3, B, sc=sc, bias_grad=bg, sc_holder=holder)
        return self._conv(a2, p + ".w2", b2, 3, B, residual=h.t, bias_grad=bg)

    def _attention(self, a: str, h: _Act, B: int) -> _Act:
        C = h.t.shape[3]
        n = self._gn(h, a + ".gn.g", a + ".gn.b", silu=False)
        qkv = self._conv(n, a + ".wqkv", self.P(a + ".bqkv"), 1, B, want_stats=False, bias_grad=self.G(a + ".bqkv"))
        o = ops.attention_hd8(qkv.t, C)
        self._consume(qkv.t)

        def bwd():
            d_o = self._take(o)
            self._put(qkv.t, ops.attention_hd8_bwd(qkv.t, o, d_o))
        self._tape.append(bwd)
        return self._conv(o, a + ".wo", self.P(a + ".bo"), 1, B, residual=h.t, bias_grad=self.G(a + ".bo"))

    # ------------------------------------------------------------------ forward
    def forward(self, sample: torch.Tensor, cond: Optional[torch.Tensor], timesteps: torch.Tensor) -> torch.Tensor:
        """``model(cat([sample, cond], 1), timesteps)`` (train.py:308): sample fp32 [B,dc,H,W], cond fp32 [B,Cc,H,W],
        timesteps [B]; returns the model output fp32 [B,out_channels,H,W] and records the backward tape."""
        if not self._loaded:
            raise RuntimeError("UNetTrainer: load_state_dict() has not been called")
        B, Cs, H, Wd = sample.shape
        Cc = 0 if cond is None else cond.shape[1]
        if Cs + Cc != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {Cs}+{Cc}")
        if H % 32 or Wd % 32:
            raise ValueError(f"H and W must be multiples of 32 (got {H}x{Wd}); pad with InputPadder(divis_by=32)")
        dev = sample.device
        self._tape, self._grads = [], {}
        self._consumed, self._bias_req = set(), {}
        need = B * 2 * 40000
        if self._arena is None or self._arena.buf.numel() < need or self._arena.buf.device != dev:
            self._arena = _StatsArena(dev, need)
        self._arena.reset()
        self._d_rows = torch.zeros((B, self.rows_total), dtype=torch.float32, device=dev)

        # time embedding (per-sample timesteps): sinusoid -> Linear -> SiLU -> Linear -> [SiLU -> all 32 projections]
        t = timesteps.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(B).contiguous()
        e0 = ops.timestep_sincos(t, 128)
        a1 = ops.linear_small(e0, self.P("time.w1"), self.P("time.b1"))
        emb = ops.linear_small(a1, self.P("time.w2"), self.P("time.b2"), act_in=1)
        rows = ops.linear_small(emb, self.P("temb.w"), self._derived[:self.rows_total], act_in=1)
        self._time_saved = (e0, a1, emb)

        col = ops.im2col_input(sample.contiguous(), None if cond is None else cond.contiguous(), dtype=self._adt)
        h = self._conv(col, "conv_in.w", self.P("conv_in.b"), 1, B, bias_grad=self.G("conv_in.b"), need_dx=False)
        skips: List[_Act] = [h]
        it = iter(self._specs)
        for i in range(6):
            for j in range(2):
                kind, bi, bj, cin_h, cin_s, cout, attn = next(it)
                h = self._resnet(UNet2DModel._rprefix(kind, bi, bj), h, None, cout, rows, B)
                if attn:
                    h = self._attention(UNet2DModel._aprefix(kind, bi, bj), h, B)
                skips.append(h)
            if i < 5:
                p = f"down_blocks.{i}.downsamplers.0.conv"
                h = self._conv(h.t, p + ".w", self.P(p + ".b"), 3, B, stride=2, bias_grad=self.G(p + ".b"))
                skips.append(h)
        for j in range(2):
            kind, bi, bj, cin_h, cin_s, cout, attn = next(it)
            h = self._resnet(UNet2DModel._rprefix(kind, bi, bj), h, None, cout, rows, B)
            if attn:
                h = self._attention(UNet2DModel._aprefix(kind, bi, bj), h, B)
        self._tape_up_start = len(self._tape)     # closures from here on belong to the up path (they run FIRST in backward)
        for i in range(6):
            for j in range(3):
                kind, bi, bj, cin_h, cin_s, cout, attn = next(it)
                h = self._resnet(UNet2DModel._rprefix(kind, bi, bj), h, skips.pop(), cout, rows, B)
                if attn:
                    h = self._attention(UNet2DModel._aprefix(kind, bi, bj), h, B)
            if i < 5:
                p = f"up_blocks.{i}.upsamplers.0.conv"
                src = h.t
                u = ops.upsample_nearest2x(src)
                self._consume(src)

                def bwd(src=src, u=u):
                    du = self._take(u)
                    self._put(src, ops.sumpool2x_nhwc16(du, addend=self._take(src)))
                self._tape.append(bwd)
                h = self._conv(u, p + ".w", self.P(p + ".b"), 3, B, bias_grad=self.G(p + ".b"))
        self._head = h
        self._consume(h.t)
        return ops.conv_out_fused(h.t, h.stats, self.P("out.g"), self.P("out.b"), self.P("out.w"), self.P("out.bias"),
                                  self.out_channels)

    __call__ = forward

    # ------------------------------------------------------------------ backward
    def zero_grad(self):
        self._g.zero_()

    def backward(self, d_out: torch.Tensor, bias_grad_done: bool = False) -> None:
        """Accumulate the parameter gradients of the last ``forward`` given d(loss)/d(model output) (fp32 NCHW)."""
        h = self._head
        if not bias_grad_done:
            self.G("out.bias").add_(d_out.sum(dim=(0, 2, 3)))
        da = ops.conv_out_bwd(h.t, h.stats, self.P("out.g"), self.P("out.b"), self.P("out.w"), d_out, self.G("out.w"))
        self._put(h.t, ops.groupnorm_silu_bwd(h.t, h.stats, self.P("out.g"), self.P("out.b"), da,
                                              dgamma=self.G("out.g"), dbeta=self.G("out.b")))
        for idx in range(len(self._tape) - 1, -1, -1):
            self._tape[idx]()
            if idx == self._tape_up_start and self._early_reduce is not None:
                self._early_reduce()              # the up path's weight gradients are final: start their all-reduce
        self._tape, self._grads, self._head = [], {}, None
        self._consumed, self._bias_req = set(), {}
        # time path: rows = Linear(SiLU(emb)); emb = Linear(SiLU(a1)); a1 = Linear(e0)
        e0, a1, emb = self._time_saved
        d_emb = ops.linear_small_bwd(emb, self.P("temb.w"), self._d_rows, act_in=1, dw=self.G("temb.w"), db=self.G("temb.b"))
        d_a1 = ops.linear_small_bwd(a1, self.P("time.w2"), d_emb, act_in=1, dw=self.G("time.w2"), db=self.G("time.b2"))
        ops.linear_small_bwd(e0, None, d_a1, act_in=0, dw=self.G("time.w1"), db=self.G("time.b1"), want_dx=False)
        # parameters that share a gradient: conv1.bias with time_emb_proj.bias, conv_shortcut.bias with conv2.bias
        self.G("c1.b").copy_(self.G("temb.b"))
        gsc, gc2 = self.G("sc.b"), self.G("c2.b")
        for kind, i, j, cin_h, cin_s, cout, attn in self._specs:
            if cin_h + cin_s != cout:
                o = self._temb_off[UNet2DModel._rprefix(kind, i, j)]
                gsc[o:o + cout].copy_(gc2[o:o + cout])
        self._time_saved = None

    # ------------------------------------------------------------------ optimizer
    def optimizer_step(self, world_size: int = 1) -> None:
        """``clip_grad_norm_`` + ``AdamW.step`` (train.py:518-524) on the flat buffers; with ``world_size > 1`` the
        gradients are first summed over the ranks (NCCL all-reduce) and averaged, as DDP does."""
        if world_size > 1:
            import torch.distributed as dist
            if self._early_done:
                # the up path's weights (59 % of the buffer) were reduced during the backward (side stream on the
                # device, an asynchronous work handle on the host / gloo)
                lo, hi = self._up_range
                dist.all_reduce(self._g[:lo])
                dist.all_reduce(self._g[hi:])
                if self._early_work is not None:
                    self._early_work.wait()
                    self._early_work = None
                else:
                    torch.cuda.current_stream().wait_stream(self._reduce_stream)
                self._early_done = False
            else:
                dist.all_reduce(self._g)
        self.step_count += 1
        self._sumsq.zero_()
        ops.sumsq(self._g, self._sumsq)
        ops.adamw_step(self._p, self._g, self._m, self._v, p16=self._p16, n16=self.n_tc, lr=self.lr, beta1=self.betas[0],
                       beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay, step=self.step_count,
                       sumsq_buf=self._sumsq, inv_loss_scale=1.0 / (self.loss_scale * world_size),
                       max_norm=self.max_grad_norm, norm_out=self._norm)
        self._refresh_derived()

    @property
    def grad_norm(self) -> float:
        return float(self._norm.item())

    # ------------------------------------------------------------------ the whole step
    def train_step(self, clean: torch.Tensor, cond: Optional[torch.Tensor], noise: torch.Tensor, timesteps: torch.Tensor,
                   noise_scheduler, mask: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None,
                   prediction_type: str = "sample", world_size: int = 1) -> torch.Tensor:
        """One optimisation step (train.py:157-538 for the pixel-space model): returns the loss (0-dim, device)."""
        noisy = noise_scheduler.add_noise(clean, noise, timesteps)          # :278
        if prediction_type == "sample":
            target = clean
        elif prediction_type == "epsilon":
            target = noise
        elif prediction_type == "v_prediction":
            target, mask = noise_scheduler.get_velocity(clean, noise, timesteps), None
        else:
            raise NotImplementedError(prediction_type)
        self.zero_grad()
        pred = self.forward(noisy, cond, timesteps)
        n = pred.numel()
        w = None if weights is None else weights.to(pred.device, torch.float32).contiguous()
        dy, loss_sum = ops.masked_loss_grad(pred, target, None if mask is None else mask.contiguous(), w,
                                            l1=self.loss_type == "l1", grad_scale=self.loss_scale / n,
                                            dbias=self.G("out.bias"))
        self._early_done = False
        if world_size > 1 and self.overlap_allreduce:
            import torch.distributed as dist
            if pred.is_cuda and self._reduce_stream is None:
                self._reduce_stream = torch.cuda.Stream(device=pred.device)

            def early():
                # every up-path closure has run: their weight gradients are complete while the whole down path is still
                # to come - reduce them under the rest of the backward
                up = self._g[self._up_range[0]:self._up_range[1]]
                if pred.is_cuda:
                    self._reduce_stream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(self._reduce_stream):
                        dist.all_reduce(up)
                else:
                    self._early_work = dist.all_reduce(up, async_op=True)
                self._early_done = True
                self.overlapped_reductions += 1
            self._early_reduce = early
        self.backward(dy, bias_grad_done=True)
        self._early_reduce = None
        self.optimizer_step(world_size)
        self.last_loss_sum = loss_sum          # fp64 [B]: sum of network_loss over each sample (train.py:422 `losses`)
        ls = loss_sum if w is None else loss_sum * w.double()
        return ls.sum() / n


def pyramid_noise_like(x: torch.Tensor, discount: float = 0.9) -> torch.Tensor:
    """``pyramid_noise_like`` of utils_d3roma/utils.py:122-133 (`noise_strategy: pyramid`, the setting of the pixel-space
    training configs): the same random draws in the same order (``torch.randn_like`` on x's device, one
    ``random.random()`` and one host ``torch.randn`` per level), the bilinear up-sampling + accumulation and the final
    division by the standard deviation on the device kernels."""
    import random
    b, c, w, h = x.shape            # (the reference's own naming: w = rows, h = columns)
    noise = torch.randn_like(x)
    for i in range(10):
        r = random.random() * 2 + 2
        w, h = max(1, int(w / (r ** i))), max(1, int(h / (r ** i)))
        noise_i = torch.randn(b, c, w, h).to(x)
        ops.bilinear_upsample_acc_(noise, noise_i, discount ** i)
        if w == 1 or h == 1:
            break
    return ops.div_by_std_(noise)


# ----------------------------------------------------------------------------------------------------------------------
# train_step with the reference's signature (train.py:157-165)
# ----------------------------------------------------------------------------------------------------------------------
COND_ORDER = {   # train.py:283-297: channels concatenated after the noisy disparity
    "rgb": ("rgb",), "rgb+raw": ("rgb", "raw"), "left+right": ("left", "right"), "left+right+raw": ("left", "right", "raw"),
    "rgb+left+right": ("rgb", "left", "right"), "rgb+left+right+raw": ("rgb", "left", "right", "raw"),
}


def train_step(accelerator, config, model: UNetTrainer, optimizer, lr_scheduler, noise_scheduler, t_sampler, step_counter,
               normalized_disp=None, normalized_rgb=None, mask=None, left_image=None, right_image=None, sim_disp=None,
               tokenizer=None, text_encoder=None, vae=None, empty_text_embed=None, **kwargs):
    """``train_step`` of the reference (train.py:157-538) for the pixel-space model, over a ``UNetTrainer``.

    Same positional / keyword arguments and the same return value ``(loss, {"t", "losses", "gradn",
    "abnormal_loss_detected"})``.  ``model`` is the ``UNetTrainer`` (it owns the parameters, the gradients and the AdamW
    state, so ``optimizer`` is unused and may be ``None``); ``accelerator`` only has to provide ``num_processes`` (``None``:
    taken from ``torch.distributed``); ``lr_scheduler`` (optional) is stepped after the update and its learning rate is
    used for the NEXT step, as ``lr_scheduler.step()`` does at :528.  ``config.ldm`` raises ``NotImplementedError``."""
    import torch.distributed as dist
    from .prepost import InputPadder
    if getattr(config, "ldm", False):
        raise NotImplementedError("d3roma_b200.training.train_step covers the pixel-space model (config.ldm = False)")
    if config.loss_type not in ("l1", "mse"):
        raise ValueError(f"loss type {config.loss_type} not supported")                    # :255-260
    if config.loss_type != model.loss_type:
        raise ValueError("config.loss_type differs from the UNetTrainer's loss_type")
    if config.noise_strategy not in ("randn", "pyramid"):
        raise NotImplementedError                                                           # :263-269
    if config.cond_channels not in COND_ORDER:
        raise NotImplementedError                                                           # :283-297
    if config.prediction_type not in ("epsilon", "v_prediction", "sample"):
        raise NotImplementedError                                                           # :310-415 ('v_pred_depth' is LDM-only)
    assert normalized_disp.shape[1] == config.depth_channels, "depth channel inconsistency"   # :181
    normalized_disp = normalized_disp.to(torch.float32).contiguous()
    mask = mask.to(torch.float32).contiguous()
    if config.clip_sample:                                                                   # :175-178
        lo_hi = ops.masked_minmax(normalized_disp.reshape(1, -1)).tolist()[0]
        if lo_hi[1] > config.clip_sample_range or lo_hi[0] < -config.clip_sample_range:
            normalized_disp, mask = ops.clamp_target_(normalized_disp.clone(), mask.clone(), float(config.clip_sample_range))
    padder = InputPadder(normalized_disp.shape, divis_by=config.divis_by)                    # :183-186
    normalized_disp, normalized_rgb, left_image, right_image, sim_disp = padder.pad(
        normalized_disp, normalized_rgb, left_image, right_image, sim_disp)
    mask = padder.pad_zero(mask)[0]
    if config.noise_strategy == "pyramid":                                                   # :263-264
        noise = pyramid_noise_like(normalized_disp)
    else:
        noise = torch.randn(normalized_disp.shape).to(normalized_disp.device)                # :266 (host generator, as there)
    bs = noise.shape[0]
    timesteps, weights = t_sampler.sample(bs, noise.device)                                  # :261
    final_rgb = normalized_rgb
    if "rgb" in config.cond_channels and getattr(config, "noise_rgb", False):               # :271-276
        noise_rgb = torch.randn(normalized_rgb.shape).to(normalized_rgb.device)
        noisy_rgb = noise_scheduler.add_noise(normalized_rgb, noise_rgb, timesteps)
        final_rgb = ops.lincomb(0.5, normalized_rgb, 0.5, noisy_rgb)
    parts = {"rgb": final_rgb, "left": left_image, "right": right_image, "raw": sim_disp}
    cond = torch.cat([parts[k] for k in COND_ORDER[config.cond_channels]], dim=1)
    if accelerator is not None:
        world = int(accelerator.num_processes)
    else:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    model.max_grad_norm = float(config.clip_grad_norm)                                       # :518
    loss = model.train_step(normalized_disp, cond, noise, timesteps, noise_scheduler, mask=mask, weights=weights,
                            prediction_type=config.prediction_type, world_size=world)
    per_sample = model.last_loss_sum / float(normalized_disp[0].numel())                     # :422 network_loss.mean((1,2,3))
    losses = {"loss": per_sample.to(torch.float32)}
    losses["mse"] = losses["loss"] ** 2 if config.loss_type == "l1" else losses["loss"]      # :423 (sic)
    gathered_t, gathered_losses = timesteps, losses
    if world > 1:                                                                            # :427-428 gather_for_metrics
        def gather(x):
            out = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(out, x.contiguous())
            return torch.cat(out)
        gathered_t = gather(timesteps)
        gathered_losses = {k: gather(v) for k, v in losses.items()}
    if lr_scheduler is not None:                                                             # :527-528
        lr_scheduler.step()
        model.lr = float(lr_scheduler.get_last_lr()[0])
    loss_value = float(loss.item())
    abnormal = bool(step_counter is not None and hasattr(step_counter, "avg_loss") and step_counter.avg_loss() > 0
                    and loss_value / step_counter.avg_loss() > 100)                          # :431 (reported, not acted on)
    return loss_value, {"t": gathered_t, "losses": gathered_losses, "gradn": model.grad_norm,
                        "abnormal_loss_detected": abnormal}

// Per-element arithmetic of the scheduler updates, shared by the stand-alone step kernels (scheduler.cu) and the
// conv_out gather that applies the DDIM update to the model output it has just produced (edge_convs.cu).  Every
// operation is an explicitly rounded fp32 intrinsic in the reference's order (core/scheduler_ddim.py:423-459,
// core/scheduler_ddpm.py:460-487): both users produce bit-identical results.
#pragma once

#include "common.cuh"
#include "internal.h"

namespace d3 {

D3_DEVINL float clip_x0(float x0, int mode, float range) {
  if (mode == 1) return fminf(fmaxf(x0, -range), range);
  if (mode == 2) return __fdiv_rn(fminf(fmaxf(x0, -1.0f), 1.0f), 1.0f);
  return x0;
}

D3_DEVINL void split_x0_eps(int ptype, float m, float s, float sa, float sb, float& x0, float& eps) {
  if (ptype == 0) {  // epsilon
    x0 = __fdiv_rn(__fsub_rn(s, __fmul_rn(sb, m)), sa);
    eps = m;
  } else if (ptype == 1) {  // sample
    x0 = m;
    eps = __fdiv_rn(__fsub_rn(s, __fmul_rn(sa, x0)), sb);
  } else {  // v_prediction
    x0 = __fsub_rn(__fmul_rn(sa, s), __fmul_rn(sb, m));
    eps = __fadd_rn(__fmul_rn(sa, m), __fmul_rn(sb, s));
  }
}

}  // namespace d3

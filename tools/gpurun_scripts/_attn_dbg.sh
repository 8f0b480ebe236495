for d in 0 64 95 31 127; do echo "== D3_ATTN_DBG=$d"; D3_ATTN_DBG=$d timeout 60 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from d3roma_b200 import ops
B,hw,heads=16,(45,80),5
C=64*heads
qkv=(torch.randn(B,hw[0],hw[1],3*C,device="cuda")*1.5).half()
q,k,v=qkv[...,:C],qkv[...,C:2*C],qkv[...,2*C:]
for _ in range(3): ops.attention_hd64_tc(q,k,v,heads)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): ops.attention_hd64_tc(q,k,v,heads)
b.record(); torch.cuda.synchronize()
print(f"{a.elapsed_time(b)/20:.3f} ms")
PY
done

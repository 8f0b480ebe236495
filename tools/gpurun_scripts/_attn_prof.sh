for d in 256 383; do echo "== D3_ATTN_DBG=$d"; D3_ATTN_DBG=$d timeout 60 python - <<'PY'
import sys, torch, ctypes
sys.path.insert(0, ".")
from d3roma_b200 import ops, _lib
B,hw,heads=16,(45,80),5
C=64*heads
qkv=(torch.randn(B,hw[0],hw[1],3*C,device="cuda")*1.5).half()
q,k,v=qkv[...,:C],qkv[...,C:2*C],qkv[...,2*C:]
for _ in range(3): ops.attention_hd64_tc(q,k,v,heads)
torch.cuda.synchronize()
lib=_lib.lib()
buf=(ctypes.c_ulonglong*16)()
lib.d3_debug_attn_profile(buf,1)
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
ops.attention_hd64_tc(q,k,v,heads)
b.record(); torch.cuda.synchronize()
lib.d3_debug_attn_profile(buf,0)
v=[int(x) for x in buf]
tiles=v[15]
print(f"{a.elapsed_time(b):.3f} ms; tiles(sum over CTAs)={tiles}")
names=["softmax total","wait s_full","S load+s_free","max+exchange","wait d_full","rescale","exps+pack","P store+arrive","MMA total","wait s_free","wait k_full","wait p_full","wait v_full"]
for i,n in enumerate(names):
    print(f"  {n:18s} {v[i]/max(tiles,1):9.1f} clk per tile")
PY
done

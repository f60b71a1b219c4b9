"""Runs the attention forward (and backward) of one shape a few times - the command profiled by ncu.
usage: python scripts/attn_one.py [tc|mma] [B Lt Lv heads d pair_mask] [packed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volta_b200 import _lib, ops

be = {"tc": _lib.ATTN_TC, "mma": _lib.ATTN_MMA}[sys.argv[1] if len(sys.argv) > 1 else "tc"]
B, Lt, Lv, heads, d, pm = (int(x) for x in (sys.argv[2:8] if len(sys.argv) >= 8 else (256, 36, 37, 12, 64, 9)))
packed = len(sys.argv) > 8 and sys.argv[8] == "packed"
A = heads * d
g = torch.Generator().manual_seed(0)
dev = "cuda"
qt = (torch.randn(B * Lt, 3 * A, generator=g) * 0.7).bfloat16().to(dev)
qv = (torch.randn(B * Lv, 3 * A, generator=g) * 0.7).bfloat16().to(dev)
lt = torch.randint(8, Lt + 1, (B,), generator=g).int()
lv = torch.randint(10, Lv + 1, (B,), generator=g).int()
mt = (torch.arange(Lt)[None] < lt[:, None]).int().to(dev)
mv = (torch.arange(Lv)[None] < lv[:, None]).int().to(dev)
ct = torch.zeros(B * Lt, A, dtype=torch.bfloat16, device=dev)
cv = torch.zeros(B * Lv, A, dtype=torch.bfloat16, device=dev)
lset, lsev = torch.zeros(B, heads, Lt, device=dev), torch.zeros(B, heads, Lv, device=dev)
kw = dict(len_t=lt.to(dev), len_v=lv.to(dev))
if packed:
    kw.update(cu_t=(torch.cumsum(lt, 0) - lt).int().to(dev), cu_v=(torch.cumsum(lv, 0) - lv).int().to(dev))
has_t, has_v = bool(pm & 7), bool(pm & 14)
tq, vq = bool(pm & 3), bool(pm & 12)


def run():
    ops.attention_fwd(qt if has_t else None, qv if has_v else None, mt if has_t else None, mv if has_v else None,
                      ct if tq else None, cv if vq else None, lset, lsev, B, Lt, Lv, heads, d, pm, 0.1, 0.1, 7, be, **kw)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 20
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
print("%s B=%d Lt=%d Lv=%d heads=%d d=%d pm=%d packed=%d: %.1f us per forward" % (sys.argv[1] if len(sys.argv) > 1 else "tc", B, Lt, Lv, heads, d, pm, packed, e0.elapsed_time(e1) / n * 1e3))

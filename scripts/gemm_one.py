"""Runs one tcgen05 GEMM configuration a few times (the command profiled by ncu for the roofline numbers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volta_b200 import _lib, ops

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
M, N, K = (int(x) for x in (sys.argv[2:5] if len(sys.argv) >= 5 else (9216, 3072, 768)))
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).bfloat16().cuda()
w = (torch.randn(N, K, generator=g) * 0.1).bfloat16().cuda()
bias = torch.zeros(N, device="cuda")
out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
pre = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
dW = torch.zeros(N, K, device="cuda")
dx = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
fns = {
    "fwd": lambda: ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS, bias=bias, backend=_lib.GEMM_TC),
    "gelu": lambda: ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS_GELU, bias=bias, aux_out=pre, backend=_lib.GEMM_TC),
    "wgrad": lambda: ops.gemm(out, a, dW, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ACCUM, backend=_lib.GEMM_TC),
    "dgrad": lambda: ops.gemm(out, w, dx, a_kmajor=True, b_kmajor=False, backend=_lib.GEMM_TC),
}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    fns[mode]()
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    fns[mode]()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("%s M=%d N=%d K=%d: %.4f ms  %.1f TFLOP/s" % (mode, M, N, K, ms, 2.0 * M * N * K / ms / 1e9))

"""First-contact diagnostics for the tcgen05 GEMM on a B200: prints error statistics per configuration
instead of asserting, so one GPU call tells us which operand layouts are right."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volta_b200 import _lib, ops

dev = "cuda"
print("device:", torch.cuda.get_device_name(0), "check:", _lib.load().vb_device_check(), flush=True)


def stats(name, got, ref):
    got, ref = got.double(), ref.double()
    err = (got - ref).norm() / ref.norm().clamp(min=1e-30)
    print("%-46s rel_l2=%.3e  max_abs=%.3e  ref_norm=%.3e got_norm=%.3e" % (name, err, (got - ref).abs().max(), ref.norm(), got.norm()), flush=True)
    return float(err)


def run(name, fn):
    try:
        t = time.time()
        r = fn()
        torch.cuda.synchronize()
        print("   [%s ok, %.2fs]" % (name, time.time() - t), flush=True)
        return r
    except Exception as e:  # noqa
        print("   [%s FAILED: %s]" % (name, str(e)[:300]), flush=True)


import subprocess
MODE = sys.argv[1] if len(sys.argv) > 1 else "main"
if MODE == "main":
    # every layout runs in its own process: a trapped kernel kills only that CUDA context
    for mode in ("kk", "kmn", "mnmn", "bench"):
        print("=== %s ===" % mode, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), mode], timeout=150, capture_output=True, text=True)
            print(r.stdout[-6000:], r.stderr[-3000:], "rc=%d" % r.returncode, flush=True)
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT", (e.stdout or b"")[-3000:], flush=True)
    sys.exit(0)

g = torch.Generator().manual_seed(0)
for (M, N, K) in [(128, 128, 64), (128, 256, 64), (256, 256, 128), (300, 768, 768), (9216, 768, 768)]:
    a = torch.randn(M, K, generator=g).bfloat16().to(dev)
    w = (torch.randn(N, K, generator=g) * 0.1).bfloat16().to(dev)
    ref = a.float() @ w.float().t()

    def kk():
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        ops.gemm(a, w, out, backend=_lib.GEMM_TC)
        torch.cuda.synchronize()
        return stats("KK  fwd   M%d N%d K%d" % (M, N, K), out.float(), ref)
    if MODE == "kk":
        run("kk", kk)

    def kmn():
        wt = w.t().contiguous()  # [K, N]: MN-major B
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        ops.gemm(a, wt, out, a_kmajor=True, b_kmajor=False, backend=_lib.GEMM_TC)
        torch.cuda.synchronize()
        return stats("K-MN dgrad M%d N%d K%d" % (M, N, K), out.float(), ref)
    if MODE == "kmn":
        run("kmn", kmn)

    def mnmn():
        at, wt = a.t().contiguous(), w.t().contiguous()  # [K, M], [K, N]
        out = torch.zeros(M, N, dtype=torch.float32, device=dev)
        ops.gemm(at, wt, out, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ACCUM, backend=_lib.GEMM_TC)
        torch.cuda.synchronize()
        return stats("MN-MN wgrad M%d N%d K%d" % (M, N, K), out, ref)
    if MODE == "mnmn":
        run("mnmn", mnmn)

if MODE != "bench":
    sys.exit(0)
# timing of the big forward GEMM
M, N, K = 9216, 3072, 768
a = torch.randn(M, K, generator=g).bfloat16().to(dev)
w = (torch.randn(N, K, generator=g) * 0.1).bfloat16().to(dev)
bias = torch.zeros(N, device=dev)
out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
pre = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)


def bench(label, fn, flops, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-40s %.3f ms  %.1f TFLOP/s" % (label, ms, flops / ms / 1e9), flush=True)


run("bench", lambda: bench("tc fwd 9216x3072x768 bias", lambda: ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS, bias=bias, backend=_lib.GEMM_TC), 2 * M * N * K))
run("bench", lambda: bench("tc fwd 9216x3072x768 bias+gelu", lambda: ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS_GELU, bias=bias, aux_out=pre, backend=_lib.GEMM_TC), 2 * M * N * K))
run("bench", lambda: bench("torch.matmul same shape", lambda: torch.matmul(a, w.t()), 2 * M * N * K))
w2 = (torch.randn(768, 768, generator=g) * 0.1).bfloat16().to(dev)
o2 = torch.zeros(M, 768, dtype=torch.bfloat16, device=dev)
run("bench", lambda: bench("tc fwd 9216x768x768", lambda: ops.gemm(a, w2, o2, backend=_lib.GEMM_TC), 2 * M * 768 * 768))
run("bench", lambda: bench("torch.matmul 9216x768x768", lambda: torch.matmul(a, w2.t()), 2 * M * 768 * 768))
dW = torch.zeros(3072, 768, device=dev)
run("bench", lambda: bench("tc wgrad 3072x768 over 9216", lambda: ops.gemm(out, a, dW, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ACCUM, backend=_lib.GEMM_TC), 2 * M * N * K))
dx = torch.zeros(M, 768, dtype=torch.bfloat16, device=dev)
run("bench", lambda: bench("tc dgrad 9216x768 over 3072", lambda: ops.gemm(out, w, dx, a_kmajor=True, b_kmajor=False, backend=_lib.GEMM_TC), 2 * M * N * K))

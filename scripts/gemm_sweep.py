"""Plain tcgen05 GEMM (no epilogue work) across tile configurations and shapes, next to torch.matmul (cuBLAS)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from volta_b200 import _lib, ops

shapes = [(8192, 8192, 8192), (9216, 3072, 768), (9216, 768, 3072), (9216, 2304, 768), (9216, 768, 768), (9472, 768, 768)]
tiles = ["192x1", "192x2", "256x2"]
if len(sys.argv) > 1:  # child: one tile configuration (the env var is read at launch time by the library)
    tile = sys.argv[1]
    for (M, N, K) in shapes:
        n = 3 if M * N >= 8192 * 8192 else 6
        sets = [(torch.randn(M, K, device="cuda").bfloat16(), (torch.randn(N, K, device="cuda") * 0.05).bfloat16(),
                 torch.empty(M, N, dtype=torch.bfloat16, device="cuda")) for _ in range(n)]
        def run(i):
            a, w, o = sets[i % n]
            if tile == "cublas":
                torch.matmul(a, w.t(), out=o)
            else:
                ops.gemm(a, w, o, backend=_lib.GEMM_TC)
        for i in range(3): run(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        iters = 20
        for i in range(iters): run(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        print("%-7s %5dx%5dx%5d %8.1f us %8.1f TF/s" % (tile, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)
        del sets
    sys.exit(0)
for t in tiles + ["cublas"]:
    env = dict(os.environ)
    if t != "cublas":
        env["VOLTA_B200_GEMM_TILE"] = t
    subprocess.run([sys.executable, os.path.abspath(__file__), t], env=env)

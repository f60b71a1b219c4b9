"""Per-kernel micro-benchmark at the headline workload's shapes (ctrl_vilbert_base, B=256, T=36, V=37).

    python scripts/opbench.py                 # every op, CUDA-event timing, us + achieved TFLOP/s or GB/s
    python scripts/opbench.py --only attn     # substring filter
    python scripts/opbench.py --once          # one launch per op after one warm-up (the command ncu profiles)

Buffers rotate through a pool larger than the 126 MB L2 so the timings are HBM-cold like the real step.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from volta_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--once", action="store_true")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()

dev = "cuda"
B, T, V, H, I, HEADS, D = args.batch, 36, 37, 768, 3072, 12, 64
Mt, Mv = B * T, B * V
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16


def rnd(*shape, dtype=bf, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


POOL = 4  # rotate 4 copies of every operand set: > L2 for all the large ops


def pool(fn):
    return [fn() for _ in range(POOL)]


OPS = []


def op(name, flops=0.0, bytes_=0.0):
    def deco(make):
        OPS.append((name, flops, bytes_, make))
        return make
    return deco


# ----------------------------------------------------------------------------------------- GEMMs
def gemm_case(name, M, N, K, epi, a_k=True, b_k=True, out_dtype=bf, drop=0.0):
    fl = 2.0 * M * N * K

    @op(name, flops=fl)
    def make():
        sets = []
        for _ in range(POOL):
            a = rnd(M, K) if a_k else rnd(K, M)
            b = rnd(N, K, scale=0.05) if b_k else rnd(K, N, scale=0.05)
            out = torch.zeros(M, N, dtype=out_dtype, device=dev)
            bias = torch.zeros(N, device=dev)
            aux_in = rnd(M, N) if epi in (_lib.EPI_BIAS_DROP_RES, _lib.EPI_MUL, _lib.EPI_ADD) else None
            aux_out = torch.zeros(M, N, dtype=bf, device=dev) if epi == _lib.EPI_BIAS_GELU else None
            sets.append((a, b, out, bias, aux_in, aux_out))

        def run(i):
            a, b, out, bias, aux_in, aux_out = sets[i % POOL]
            ops.gemm(a, b, out, a_kmajor=a_k, b_kmajor=b_k, epilogue=epi, bias=bias, aux_in=aux_in, aux_out=aux_out,
                     dropout_p=drop, seed=1234, backend=_lib.GEMM_TC)
        return run


gemm_case("gemm qkv      9216x2304x768 bias", Mt, 3 * H, H, _lib.EPI_BIAS)
gemm_case("gemm attn-out 9216x768x768 bias+drop+res", Mt, H, H, _lib.EPI_BIAS_DROP_RES, drop=0.1)
gemm_case("gemm ffn1     9216x3072x768 bias+gelu", Mt, I, H, _lib.EPI_BIAS_GELU)
gemm_case("gemm ffn2     9216x768x3072 bias+drop+res", Mt, H, I, _lib.EPI_BIAS_DROP_RES, drop=0.1)
gemm_case("dgrad ffn2    9216x3072x768 *gelu'", Mt, I, H, _lib.EPI_MUL, b_k=False)
gemm_case("dgrad ffn1    9216x768x3072 +res", Mt, H, I, _lib.EPI_ADD, b_k=False)
gemm_case("dgrad out     9216x768x768", Mt, H, H, _lib.EPI_NONE, b_k=False)
gemm_case("dgrad qkv     9216x768x2304 +res", Mt, H, 3 * H, _lib.EPI_ADD, b_k=False)
gemm_case("wgrad ffn1    3072x768 over 9216", I, H, Mt, _lib.EPI_ACCUM, a_k=False, b_k=False, out_dtype=torch.float32)
gemm_case("wgrad ffn2    768x3072 over 9216", H, I, Mt, _lib.EPI_ACCUM, a_k=False, b_k=False, out_dtype=torch.float32)
gemm_case("wgrad qkv     2304x768 over 9216", 3 * H, H, Mt, _lib.EPI_ACCUM, a_k=False, b_k=False, out_dtype=torch.float32)
gemm_case("wgrad out     768x768 over 9216", H, H, Mt, _lib.EPI_ACCUM, a_k=False, b_k=False, out_dtype=torch.float32)


# ----------------------------------------------------------------------------------------- attention
def attn_case(name, pm, bwd, p=0.1):
    has_t = bool(pm & 0b0111)  # bit = 2*query_stream + key_stream: tt, tv, vt touch text
    has_v = bool(pm & 0b1110)
    rows = (Mt if has_t else 0) + (Mv if has_v else 0)
    by = rows * 3 * H * 2 + rows * H * 2
    if bwd:
        by = rows * 3 * H * 2 * 2 + rows * H * 2 * 2

    @op(name, bytes_=by)
    def make():
        sets = []
        for _ in range(POOL):
            qt = rnd(Mt, 3 * H) if has_t else None
            qv = rnd(Mv, 3 * H) if has_v else None
            mt = torch.ones(B, T, dtype=torch.int32, device=dev)
            mv = torch.ones(B, V, dtype=torch.int32, device=dev)
            mt[:, 30:] = 0
            ct = torch.zeros(Mt, H, dtype=bf, device=dev) if has_t else None
            cv = torch.zeros(Mv, H, dtype=bf, device=dev) if has_v else None
            lt = torch.zeros(B, HEADS, T, device=dev) if has_t else None
            lv = torch.zeros(B, HEADS, V, device=dev) if has_v else None
            extra = {}
            if bwd:
                extra = dict(dctx_t=rnd(Mt, H) if has_t else None, dctx_v=rnd(Mv, H) if has_v else None,
                             dqkv_t=torch.zeros(Mt, 3 * H, dtype=bf, device=dev) if has_t else None,
                             dqkv_v=torch.zeros(Mv, 3 * H, dtype=bf, device=dev) if has_v else None)
            sets.append((qt, qv, mt if has_t else None, mv if has_v else None, ct, cv, lt, lv, extra))
        for s in sets:  # the backward needs a real forward's lse / ctx
            ops.attention_fwd(*s[:8], B, T, V, HEADS, D, pm, p, p, 77, _lib.ATTN_MMA)

        def run(i):
            s = sets[i % POOL]
            if bwd:
                ops.attention_bwd(*s[:8], B, T, V, HEADS, D, pm, p, p, 77, _lib.ATTN_MMA, **s[8])
            else:
                ops.attention_fwd(*s[:8], B, T, V, HEADS, D, pm, p, p, 77, _lib.ATTN_MMA)
        return run


attn_case("attn fwd tt", 1, False)
attn_case("attn fwd tt+vv", 9, False)
attn_case("attn fwd tv+vt", 6, False)
attn_case("attn bwd tt", 1, True)
attn_case("attn bwd tt+vv", 9, True)
attn_case("attn bwd tv+vt", 6, True)


# ----------------------------------------------------------------------------------------- LN / colsum
@op("ln fwd 9216x768", bytes_=Mt * H * 2 * 2)
def _ln_fwd():
    sets = pool(lambda: (rnd(Mt, H), torch.ones(H, device=dev), torch.zeros(H, device=dev), torch.zeros(Mt, H, dtype=bf, device=dev),
                         torch.zeros(Mt, device=dev), torch.zeros(Mt, device=dev)))
    return lambda i: ops.layernorm_fwd(*sets[i % POOL])


@op("ln bwd 9216x768 (+dropout copy, +dbias)", bytes_=Mt * H * 2 * 4)
def _ln_bwd():
    def mk():
        x = rnd(Mt, H)
        mean, rstd = x.float().mean(1), 1.0 / x.float().std(1)
        return (rnd(Mt, H), x, mean.contiguous(), rstd.contiguous(), torch.ones(H, device=dev), torch.zeros(Mt, H, dtype=bf, device=dev),
                torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(Mt, H, dtype=bf, device=dev), torch.zeros(H, device=dev))
    sets = pool(mk)

    def run(i):
        dy, x, mean, rstd, gam, dx, dg, db, dxd, dbias = sets[i % POOL]
        ops.layernorm_bwd(dy, x, mean, rstd, gam, dx, dg, db, 1.0, dxd, 0.1, 99, 0, dbias)
    return run


@op("colsum 9216x3072", bytes_=Mt * I * 2)
def _colsum():
    sets = pool(lambda: (rnd(Mt, I), torch.zeros(I, device=dev)))
    return lambda i: ops.colsum(*sets[i % POOL])


@op("colsum 9216x2304", bytes_=Mt * 3 * H * 2)
def _colsum2():
    sets = pool(lambda: (rnd(Mt, 3 * H), torch.zeros(3 * H, device=dev)))
    return lambda i: ops.colsum(*sets[i % POOL])


pk = {"tf": 1354.0, "gb": 6529.7}
print("%-48s %10s %12s %8s" % ("op", "us", "achieved", "frac"))
for name, fl, by, make in OPS:
    if args.only and args.only not in name:
        continue
    run = make()
    run(0)
    torch.cuda.synchronize()
    if args.once:
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only this launch is captured
        run(1)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    for i in range(3):
        run(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / args.iters * 1e3
    if fl:
        ach = fl / us / 1e6
        print("%-48s %10.1f %8.1f TF/s %7.1f%%" % (name, us, ach, 100 * ach / pk["tf"]))
    else:
        ach = by / us / 1e3
        print("%-48s %10.1f %8.1f GB/s %7.1f%%" % (name, us, ach, 100 * ach / pk["gb"]))
    del run
    torch.cuda.empty_cache()

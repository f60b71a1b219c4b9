"""Kernel-level parity on the B200: every C-ABI entry point against its PyTorch restatement (tests/emu_ops.py).
bf16 tolerances are relative L2 per tensor (stated per test); integer/index outputs must be exact."""
import math

import pytest
import torch

import emu_ops
from volta_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def rnd(*shape, dtype=torch.bfloat16, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


# ------------------------------------------------------------------ GEMM
GEMM_SHAPES = [(256, 256, 128), (300, 768, 768), (130, 1601, 768), (1000, 128, 3072), (72, 2304, 768), (520, 200, 72)]


@pytest.mark.parametrize("backend", ["tc", "simt"])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_forward_bias(backend, M, N, K):
    a, w, bias = rnd(M, K), rnd(N, K, scale=0.05), rnd(N, dtype=torch.float32)
    ldn = (N + 7) // 8 * 8
    out = torch.zeros(M, ldn, dtype=torch.bfloat16, device=DEV)[:, :N]
    ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS, bias=bias, backend=_lib.GEMM_TC if backend == "tc" else _lib.GEMM_SIMT)
    ref = a.float() @ w.float().t() + bias
    assert rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(256, 2, 1024), (500, 1, 1024), (37, 8, 200)])
def test_gemm_skinny_outputs(M, N, K, dtype):
    """N <= 8 (ITM / retrieval logit heads): warp-per-row kernel behind the SIMT back-end."""
    a, w, bias = rnd(M, K, dtype=dtype), rnd(N, K, dtype=dtype, scale=0.05), rnd(8, dtype=torch.float32)[:N]
    out = torch.zeros(M, 8, dtype=dtype, device=DEV)[:, :N]
    ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS, bias=bias, backend=_lib.GEMM_AUTO)
    assert rel(out.float(), a.float() @ w.float().t() + bias) < (6e-3 if dtype == torch.bfloat16 else 1e-5)
    out2 = torch.zeros(M, 8, dtype=dtype, device=DEV)[:, :N]
    ops.gemm(a, w, out2, backend=_lib.GEMM_SIMT)
    assert rel(out2.float(), a.float() @ w.float().t()) < (6e-3 if dtype == torch.bfloat16 else 1e-5)


@pytest.mark.parametrize("epi", ["none", "gelu", "relu", "res", "mul", "add", "drelu", "bias_add"])
@pytest.mark.parametrize("odt", [torch.bfloat16, torch.float32])
def test_gemm_tc_epilogues(epi, odt):
    M, N, K = 260, 392, 256
    a, w, bias = rnd(M, K), rnd(N, K, scale=0.1), rnd(N, dtype=torch.float32)
    aux = rnd(M, N, dtype=odt, seed=3)
    out, out_ref = torch.zeros(M, N, dtype=odt, device=DEV), torch.zeros(M, N, dtype=odt, device=DEV)
    pre, pre_ref = torch.zeros(M, N, dtype=odt, device=DEV), torch.zeros(M, N, dtype=odt, device=DEV)
    code = dict(none=_lib.EPI_NONE, gelu=_lib.EPI_BIAS_GELU, relu=_lib.EPI_BIAS_RELU, res=_lib.EPI_BIAS_DROP_RES,
                mul=_lib.EPI_MUL, add=_lib.EPI_ADD, drelu=_lib.EPI_DRELU, bias_add=_lib.EPI_BIAS_ADD)[epi]
    kw = dict(epilogue=code, bias=bias, aux_in=aux)
    ops.gemm(a, w, out, aux_out=pre, backend=_lib.GEMM_TC, **kw)
    emu_ops.gemm(a, w, out_ref, aux_out=pre_ref, **kw)
    assert rel(out.float(), out_ref.float()) < 8e-3
    if epi == "gelu":
        assert rel(pre.float(), pre_ref.float()) < 8e-3


@pytest.mark.parametrize("backend", ["tc", "simt"])
def test_gemm_dgrad_mn_major_b(backend):
    """dX[M,K] = dY[M,N] . W[N,K]: B operand is MN-major (row-major [reduction, out])."""
    M, N, K = 300, 520, 384
    dy, w = rnd(M, N), rnd(N, K, scale=0.1)
    out = torch.zeros(M, K, dtype=torch.bfloat16, device=DEV)
    ops.gemm(dy, w, out, a_kmajor=True, b_kmajor=False, backend=_lib.GEMM_TC if backend == "tc" else _lib.GEMM_SIMT)
    assert rel(out.float(), dy.float() @ w.float()) < 6e-3


@pytest.mark.parametrize("backend", ["tc", "simt"])
@pytest.mark.parametrize("M,N,K,split", [(1000, 256, 384, 0), (4096, 768, 768, 0), (333, 1601, 128, 1), (2000, 130, 264, 3)])
def test_gemm_wgrad_accumulate(backend, M, N, K, split):
    """dW[N,K] += dY[M,N]^T . X[M,K]: both operands MN-major, fp32 atomic accumulation (split-K)."""
    # operands keep a 16-byte row pitch (the engine pads logits to a multiple of 8 columns)
    dy = rnd(M, (N + 7) // 8 * 8, scale=0.5)[:, :N]
    x = rnd(M, (K + 7) // 8 * 8, scale=0.5)[:, :K]
    base = rnd(N, K, dtype=torch.float32, seed=9)
    out = base.clone()
    ops.gemm(dy, x, out, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ACCUM, split_k=split,
             backend=_lib.GEMM_TC if backend == "tc" else _lib.GEMM_SIMT)
    ref = base + dy.float().t() @ x.float()
    assert rel(out, ref) < 2e-3


@pytest.mark.parametrize("clc", ["0", "1"])
@pytest.mark.parametrize("tile", ["128x1", "192x1", "192x2", "256x2"])
def test_gemm_tc_every_tile_configuration(tile, clc, monkeypatch):
    """Each instantiation of the tcgen05 kernel (tile width x CTA-group size) on all three operand layouts, ragged
    edges, several persistent rounds and split-K; CTA pairs (x2) run the cta_group::2 protocol.  clc=1: work items
    handed out by cluster launch control (one cluster per item, resident clusters cancel and take over the rest)
    instead of the static persistent stride."""
    monkeypatch.setenv("VOLTA_B200_GEMM_TILE", tile)
    monkeypatch.setenv("VOLTA_B200_GEMM_CLC", clc)
    for (M, N, K) in [(300, 776, 328), (9472, 1536, 256), (129, 200, 64)]:
        a, w, bias = rnd(M, K), rnd(N, K, scale=0.05), rnd(N, dtype=torch.float32)
        res = rnd(M, N, seed=2)
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS_ADD, bias=bias, aux_in=res, backend=_lib.GEMM_TC)
        assert rel(out.float(), a.float() @ w.float().t() + bias + res.float()) < 6e-3, (tile, M, N, K, "fwd")
        if tile != "192x2":  # an MN-major B splits between the CTAs of a pair only in 64-column blocks
            wt = w.t().contiguous()  # [K, N]
            out2 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
            ops.gemm(a, wt, out2, a_kmajor=True, b_kmajor=False, backend=_lib.GEMM_TC)
            assert rel(out2.float(), a.float() @ w.float().t()) < 6e-3, (tile, M, N, K, "dgrad")
            dy = rnd(M, N, scale=0.5, seed=4)
            dW = torch.ones(N, K, device=DEV)
            ops.gemm(dy, a, dW, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ACCUM, backend=_lib.GEMM_TC)
            assert rel(dW, 1.0 + dy.float().t() @ a.float()) < 2e-3, (tile, M, N, K, "wgrad")


@pytest.mark.parametrize("tile", ["auto", "192x1", "192x2", "256x2"])
def test_gemm_row_split_launch(tile, monkeypatch):
    """One launch, two weight matrices: rows below split_M use (B, bias), the rest (B2, bias2) - the text and vision
    rows of a two-stream sub-layer; forward (K-major B) and dgrad (MN-major B) layouts."""
    if tile != "auto":
        monkeypatch.setenv("VOLTA_B200_GEMM_TILE", tile)
    M, split, N, K = 1300, 512, 776, 320
    a = rnd(M, K)
    w1, w2 = rnd(N, K, scale=0.05), rnd(N, K, scale=0.05, seed=1)
    b1, b2 = rnd(N, dtype=torch.float32), rnd(N, dtype=torch.float32, seed=2)
    res = rnd(M, N, seed=3)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, w1, out, epilogue=_lib.EPI_BIAS_ADD, bias=b1, aux_in=res, backend=_lib.GEMM_TC, b2=w2, bias2=b2, split_M=split)
    ref = torch.cat([a[:split].float() @ w1.float().t() + b1, a[split:].float() @ w2.float().t() + b2]) + res.float()
    assert rel(out.float(), ref) < 6e-3
    if tile != "192x2":
        w1t, w2t = w1.t().contiguous(), w2.t().contiguous()   # [K, N]: MN-major B
        out2 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(a, w1t, out2, a_kmajor=True, b_kmajor=False, backend=_lib.GEMM_TC, b2=w2t, split_M=split)
        ref2 = torch.cat([a[:split].float() @ w1.float().t(), a[split:].float() @ w2.float().t()])
        assert rel(out2.float(), ref2) < 6e-3


@pytest.mark.parametrize("backend", ["tc", "simt"])
@pytest.mark.parametrize("count", [0, 1, 300, 777, 1300])
def test_gemm_device_side_extents(backend, count):
    """m_limit / k_limit: the row count lives in device memory (compacted masked-LM rows); tiles / k-blocks beyond it are
    not scheduled, rows below it match the plain GEMM, rows beyond the last touched tile keep their old contents."""
    be = _lib.GEMM_TC if backend == "tc" else _lib.GEMM_SIMT
    M, N, K = 1300, 776, 320
    npad = min(M, (count + 255) // 256 * 256)
    lim = torch.tensor([count, npad], dtype=torch.int32, device=DEV)
    a, w, bias = rnd(M, K), rnd(N, K, scale=0.05), rnd(N, dtype=torch.float32)
    a[count:] = 0                                   # the caller keeps the operands zero up to the padded count
    out = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS, bias=bias, backend=be, m_limit=lim[0:1])
    if count:
        assert rel(out[:count].float(), a[:count].float() @ w.float().t() + bias) < 6e-3
    assert bool((out[npad:] == 7.0).all()), "rows beyond the last scheduled tile were written"
    # dgrad layout (MN-major B)
    wt = w.t().contiguous()
    out2 = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, wt, out2, a_kmajor=True, b_kmajor=False, backend=be, m_limit=lim[0:1])
    if count:
        assert rel(out2[:count].float(), a[:count].float() @ w.float().t()) < 6e-3
    assert bool((out2[npad:] == 7.0).all())
    # wgrad: the reduction runs over the rows
    dy = rnd(M, N, scale=0.5, seed=4)
    dy[count:] = 0
    dW = torch.ones(N, K, device=DEV)
    ops.gemm(dy, a, dW, a_kmajor=False, b_kmajor=False, epilogue=_lib.EPI_ACCUM, backend=be, k_limit=lim[1:2])
    ref = 1.0 + dy[:count].float().t() @ a[:count].float()
    assert rel(dW, ref) < 2e-3


@pytest.mark.parametrize("backend", ["tc", "simt"])
@pytest.mark.parametrize("na,nb,split", [(300, 700, True), (0, 5, False), (1024, 0, True), (513, 1500, False)])
def test_gemm_two_row_segments_with_device_counts(backend, na, nb, split):
    """m_limit / m_limit2 / limit_seg_M: packed text rows [0, na) and packed vision rows [seg, seg + nb) of a joint
    buffer, optionally with per-segment weights (row-split launch).  Rows below the counts match the plain GEMM; rows
    beyond the last scheduled tile of a segment are not written."""
    be = _lib.GEMM_TC if backend == "tc" else _lib.GEMM_SIMT
    seg, M, N, K = 1024, 1024 + 1536, 776, 320
    lim = torch.tensor([na, nb], dtype=torch.int32, device=DEV)
    a = rnd(M, K)
    a[na:seg] = 0
    a[seg + nb:] = 0
    w1, w2 = rnd(N, K, scale=0.05), rnd(N, K, scale=0.05, seed=1)
    b1, b2 = rnd(N, dtype=torch.float32), rnd(N, dtype=torch.float32, seed=2)
    out = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    kw = dict(b2=w2, bias2=b2, split_M=seg) if split else {}
    if split and backend == "simt":
        pytest.skip("row-split launches are a tcgen05 feature")
    ops.gemm(a, w1, out, epilogue=_lib.EPI_BIAS, bias=b1, backend=be, m_limit=lim[0:1], m_limit2=lim[1:2], limit_seg_M=seg, **kw)
    wb, bb = (w2, b2) if split else (w1, b1)
    if na:
        assert rel(out[:na].float(), a[:na].float() @ w1.float().t() + b1) < 6e-3
    if nb:
        assert rel(out[seg:seg + nb].float(), a[seg:seg + nb].float() @ wb.float().t() + bb) < 6e-3
    pa, pb = (na + 255) // 256 * 256, (nb + 255) // 256 * 256
    assert bool((out[pa:seg] == 7.0).all()) and bool((out[seg + pb:] == 7.0).all())
    # dgrad layout
    w1t = w1.t().contiguous()
    out2 = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    kw2 = dict(b2=w2.t().contiguous(), split_M=seg) if split else {}
    ops.gemm(a, w1t, out2, a_kmajor=True, b_kmajor=False, backend=be, m_limit=lim[0:1], m_limit2=lim[1:2], limit_seg_M=seg, **kw2)
    if na:
        assert rel(out2[:na].float(), a[:na].float() @ w1.float().t()) < 6e-3
    if nb:
        assert rel(out2[seg:seg + nb].float(), a[seg:seg + nb].float() @ wb.float().t()) < 6e-3
    assert bool((out2[pa:seg] == 7.0).all()) and bool((out2[seg + pb:] == 7.0).all())


def test_zero_rows_device_bounds():
    buf = torch.full((700, 776), 3.0, dtype=torch.bfloat16, device=DEV)
    lim = torch.tensor([301, 512], dtype=torch.int32, device=DEV)
    ops.zero_rows(buf[100:], lim[0:1], lim[1:2])
    ref = torch.full_like(buf, 3.0)
    ref[401:612] = 0
    assert torch.equal(buf, ref)
    lim2 = torch.tensor([5, 9], dtype=torch.int32, device=DEV)
    ops.zero_rows(buf, lim2[0:1], lim2[1:2], second=(lim2[0:1], lim2[1:2], 650))   # two ranges, one launch
    ref[5:9] = 0
    ref[655:659] = 0
    assert torch.equal(buf, ref)


def test_compact_plan_gather_scatter_and_limited_criterion():
    """vb_compact_*: order-preserving packing of the labelled rows, exact; cross-entropy / column sums with a device-side
    row count equal the full-size calls on the same rows."""
    R, H, C = 3000, 768, 1000
    g = torch.Generator().manual_seed(3)
    labels = torch.where(torch.rand(R, generator=g) < 0.12, torch.randint(0, C, (R,), generator=g), torch.full((R,), -1)).to(DEV)
    x = rnd(R, H)
    lim = torch.zeros(2, dtype=torch.int32, device=DEV)
    src, pos = torch.zeros(R, dtype=torch.int32, device=DEV), torch.zeros(R, dtype=torch.int32, device=DEV)
    lab_c = torch.zeros(R, dtype=torch.int64, device=DEV)
    ops.compact_plan(labels, 256, lim, src, pos, lab_c)
    keep = (labels >= 0).nonzero()[:, 0]
    n = keep.numel()
    npad = (n + 255) // 256 * 256
    assert lim.tolist() == [n, npad]
    assert torch.equal(src[:n].long(), keep) and bool((src[n:] == -1).all())
    assert torch.equal(lab_c[:n], labels[keep]) and bool((lab_c[n:] == -1).all())
    assert torch.equal(pos[keep].long(), torch.arange(n, device=DEV)) and int((pos >= 0).sum()) == n
    xc = torch.full((R, H), 5.0, dtype=torch.bfloat16, device=DEV)
    ops.compact_gather(x, src, lim, xc)
    assert torch.equal(xc[:n], x[keep]) and bool((xc[n:npad] == 0).all()) and bool((xc[npad:] == 5.0).all())
    back = torch.full((R, H), 5.0, dtype=torch.bfloat16, device=DEV)
    ops.compact_scatter(xc, pos, back)
    ref = torch.zeros_like(x)
    ref[keep] = x[keep]
    assert torch.equal(back, ref)
    # no labelled row at all
    none = torch.full((R,), -1, dtype=torch.int64, device=DEV)
    ops.compact_plan(none, 256, lim, src, pos, lab_c)
    assert lim.tolist() == [0, 0] and bool((pos == -1).all())
    # criterion + column sums on the packed rows
    ops.compact_plan(labels, 256, lim, src, pos, lab_c)
    logits = rnd(R, C, scale=2.0, seed=5)
    lg_c = torch.zeros(R, C, dtype=torch.bfloat16, device=DEV)
    ops.compact_gather(logits, src, lim, lg_c)
    out = {}
    for name, (lg, lb, rl_f, rl_b) in dict(full=(logits, labels, None, None), packed=(lg_c, lab_c, lim[0:1], lim[1:2])).items():
        lse = torch.zeros(R, device=DEV)
        sums = torch.zeros(2, device=DEV)
        ops.xent_fwd(lg, lb, C, lse, sums[0:1], sums[1:2], row_limit=rl_f)
        d = torch.full((R, C), 3.0, dtype=torch.bfloat16, device=DEV)
        ops.xent_bwd(lg, lb, C, lse, sums[1:2], torch.ones(1, device=DEV), d, row_limit=rl_b)
        cs = torch.zeros(C, device=DEV)
        ops.colsum(d, cs, row_limit=rl_b)
        out[name] = (sums.clone(), d, cs)
    assert torch.allclose(out["full"][0], out["packed"][0], rtol=1e-5)
    assert torch.equal(out["packed"][1][:n], out["full"][1][keep])
    assert bool((out["packed"][1][n:npad] == 0).all()) and bool((out["packed"][1][npad:] == 3.0).all())
    assert rel(out["packed"][2], out["full"][2]) < 1e-5


def test_gemm_tc_strided_rows_and_inplace_add():
    """The pooler reads row 0 of every sample through lda = T*H and adds its dgrad in place through ldd."""
    B, T, H, P = 64, 12, 256, 512
    x = rnd(B * T, H)
    w, bias = rnd(P, H, scale=0.1), rnd(P, dtype=torch.float32)
    rows0 = x.view(B, T * H)[:, :H]
    out = torch.zeros(B, P, dtype=torch.bfloat16, device=DEV)
    ops.gemm(rows0, w, out, epilogue=_lib.EPI_BIAS_RELU, bias=bias, backend=_lib.GEMM_TC)
    assert rel(out.float(), torch.relu(rows0.float() @ w.float().t() + bias)) < 6e-3
    dpre = rnd(B, P, seed=5)
    dx = rnd(B * T, H, seed=6)
    ref = dx.clone().float().view(B, T, H)
    ref[:, 0] += dpre.float() @ w.float()
    d0 = dx.view(B, T * H)[:, :H]
    ops.gemm(dpre, w, d0, a_kmajor=True, b_kmajor=False, epilogue=_lib.EPI_ADD, aux_in=d0, backend=_lib.GEMM_TC)
    assert rel(dx.float().view(B, T, H), ref) < 6e-3


def test_gemm_fp32_simt_matches_fp32():
    M, N, K = 200, 96, 40
    a, w = rnd(M, K, dtype=torch.float32), rnd(N, K, dtype=torch.float32)
    out = torch.zeros(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(a, w, out)
    assert rel(out, a @ w.t()) < 1e-5


def test_gemm_dropout_mask_is_shared_with_layernorm_bwd():
    """BIAS_DROP_RES and vb_layernorm_bwd(dx_drop) must regenerate the same Philox mask."""
    M, H, p, seed = 256, 768, 0.1, 1234567
    a = torch.zeros(M, 64, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(H, 64, dtype=torch.bfloat16, device=DEV)
    bias = torch.ones(H, dtype=torch.float32, device=DEV)
    res = torch.zeros(M, H, dtype=torch.bfloat16, device=DEV)
    for backend in (_lib.GEMM_TC, _lib.GEMM_SIMT):
        out = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
        ops.gemm(a, w, out, epilogue=_lib.EPI_BIAS_DROP_RES, bias=bias, aux_in=res, dropout_p=p, seed=seed, backend=backend)
        keep = out.float() > 0
        assert abs(float(keep.float().mean()) - (1 - p)) < 0.01
        assert torch.allclose(out.float()[keep], torch.tensor(1 / (1 - p)).bfloat16().float().to(DEV))
        x = rnd(M, H)
        dy = rnd(M, H, seed=2)
        gamma = torch.ones(H, device=DEV)
        mean, rstd = torch.zeros(M, device=DEV), torch.ones(M, device=DEV)
        dx, dxd = torch.empty_like(x), torch.empty_like(x)
        ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, None, None, dx_drop=dxd, dropout_p=p, seed=seed)
        assert torch.equal(dxd.float() != 0, keep & (dx.float() != 0))


# ------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("H", [128, 768, 1024, 1536])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_layernorm_fwd_bwd(H, dtype):
    rows = 1000
    x, dy = rnd(rows, H, dtype=dtype, scale=2.0), rnd(rows, H, dtype=dtype, seed=1)
    gamma, beta = 1 + rnd(H, dtype=torch.float32, scale=0.1), rnd(H, dtype=torch.float32, scale=0.1)
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    y, y_ref = torch.empty_like(x), torch.empty_like(x)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    mean_r, rstd_r = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd)
    emu_ops.layernorm_fwd(x, gamma, beta, y_ref, mean_r, rstd_r)
    assert rel(y.float(), y_ref.float()) < tol
    assert rel(mean, mean_r) < 1e-4 and rel(rstd, rstd_r) < 1e-4
    dx, dx_ref = torch.empty_like(x), torch.empty_like(x)
    grads = [torch.zeros(H, device=DEV) for _ in range(6)]
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, grads[0], grads[1], dbias=grads[2])
    emu_ops.layernorm_bwd(dy, x, mean_r, rstd_r, gamma, dx_ref, grads[3], grads[4], dbias=grads[5])
    assert rel(dx.float(), dx_ref.float()) < tol
    for got, ref in zip(grads[:3], grads[3:]):
        assert rel(got, ref) < max(tol, 1e-4)


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_layernorm_skips_trailing_padding_rows(dropout):
    """vb_row_seqs: rows at or beyond their sample's valid length are written as zeros and contribute nothing; every
    other row and every column sum is bit-identical to the call without the skip on inputs whose dy is zero there."""
    B, T, V, H = 9, 36, 37, 768
    rows = B * (T + V)
    g = torch.Generator().manual_seed(1)
    len_t = torch.randint(0, T + 1, (B,), generator=g, dtype=torch.int32)
    len_v = torch.randint(0, V + 1, (B,), generator=g, dtype=torch.int32)
    len_t[0], len_v[1] = 0, 0                      # length 0 = the sample is computed in full
    live_t = (torch.arange(T)[None] < len_t[:, None]) | (len_t[:, None] == 0)
    live_v = (torch.arange(V)[None] < len_v[:, None]) | (len_v[:, None] == 0)
    live = torch.cat([live_t.reshape(-1), live_v.reshape(-1)]).to(DEV)
    x = rnd(rows, H, scale=2.0)
    dy = rnd(rows, H, seed=1) * live[:, None]      # gradients of padded rows are exactly zero in the model
    gamma, beta = 1 + rnd(H, dtype=torch.float32, scale=0.1), rnd(H, dtype=torch.float32, scale=0.1)
    seqs = (len_t.to(DEV), T, B * T, len_v.to(DEV), V)
    out = {}
    for name, sq in dict(full=None, skip=seqs).items():
        y = torch.full_like(x, 5.0)
        mean, rstd = torch.zeros(rows, device=DEV), torch.zeros(rows, device=DEV)
        ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, seqs=sq)
        dx, dxd = torch.full_like(x, 5.0), torch.full_like(x, 5.0)
        dg, db, dbias = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
        ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, dx_drop=dxd if dropout else None, dropout_p=dropout, seed=3,
                          dbias=dbias, seqs=sq)
        out[name] = (y, dx, dxd, dg, db, dbias)
    for i in range(3 if dropout else 2):
        assert torch.equal(out["skip"][i][live], out["full"][i][live])
        assert bool((out["skip"][i][~live] == 0).all())
    for i in (3, 4, 5):
        assert rel(out["skip"][i], out["full"][i]) < 1e-5   # fp32 atomics: order only


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_layernorm_two_streams_in_one_launch(dropout):
    """gamma2 / split_row: the vision rows of a joint buffer are normalised by the same launch with their own module's
    parameters; against two separate launches: identical rows, separate gradient sums (fp32 atomics: order only)."""
    Mt, Mv, H = 512, 777, 768
    rows = Mt + Mv
    x, dy = rnd(rows, H, scale=2.0), rnd(rows, H, seed=1)
    g1, b1 = 1 + rnd(H, dtype=torch.float32, scale=0.1), rnd(H, dtype=torch.float32, scale=0.1)
    g2, b2 = 1 + rnd(H, dtype=torch.float32, scale=0.1, seed=2), rnd(H, dtype=torch.float32, scale=0.1, seed=3)

    def run(merged):
        y, dx, dxd = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        mean, rstd = torch.zeros(rows, device=DEV), torch.zeros(rows, device=DEV)
        sums = [torch.zeros(H, device=DEV) for _ in range(6)]   # dg1, db1, dbias1, dg2, db2, dbias2
        dd = dxd if dropout else None
        if merged:
            ops.layernorm_fwd(x, g1, b1, y, mean, rstd, second=(g2, b2), split_row=Mt)
            ops.layernorm_bwd(dy, x, mean, rstd, g1, dx, sums[0], sums[1], dx_drop=dd, dropout_p=dropout, seed=3, dbias=sums[2],
                              second=(g2, sums[3], sums[4], sums[5]), split_row=Mt)
        else:
            for lo, hi, g_, b_, k, off in ((0, Mt, g1, b1, 0, 0), (Mt, rows, g2, b2, 3, Mt * H)):
                ops.layernorm_fwd(x[lo:hi], g_, b_, y[lo:hi], mean[lo:hi], rstd[lo:hi])
                ops.layernorm_bwd(dy[lo:hi], x[lo:hi], mean[lo:hi], rstd[lo:hi], g_, dx[lo:hi], sums[k], sums[k + 1],
                                  dx_drop=None if dd is None else dd[lo:hi], dropout_p=dropout, seed=3, offset=off, dbias=sums[k + 2])
        return y, dx, (dxd if dropout else dx), sums
    ya, dxa, dda, sa = run(True)
    yb, dxb, ddb, sb = run(False)
    assert torch.equal(ya, yb) and torch.equal(dxa, dxb) and torch.equal(dda, ddb)
    for a, b in zip(sa, sb):
        assert rel(a, b) < 1e-5


def test_layernorm_scale_accumulate():
    rows, H = 300, 768
    a, b = rnd(rows, H), rnd(rows, H, seed=1)
    g1, b1, g2, b2 = (1 + rnd(H, dtype=torch.float32, scale=0.1, seed=s) for s in range(4))
    y = torch.empty_like(a)
    ops.layernorm_fwd(a, g1, b1, y, out_scale=0.5)
    ops.layernorm_fwd(b, g2, b2, y, out_scale=0.5, accum=True)
    ref = 0.5 * torch.nn.functional.layer_norm(a.float(), (H,), g1, b1, 1e-12) + \
        0.5 * torch.nn.functional.layer_norm(b.float(), (H,), g2, b2, 1e-12)
    assert rel(y.float(), ref) < 8e-3


# ------------------------------------------------------------------ attention
ATTN_CASES = [  # (B, Lt, Lv, heads, d, pair_mask)
    (3, 36, 37, 12, 64, 1), (3, 36, 37, 12, 64, 8), (3, 36, 37, 12, 64, 6), (3, 36, 37, 12, 64, 15),
    (2, 30, 101, 8, 128, 6), (2, 30, 101, 8, 128, 8), (2, 10, 7, 2, 64, 15), (2, 38, 37, 12, 64, 9),
    (2, 56, 80, 4, 64, 15),
]


def _attn_inputs(B, Lt, Lv, heads, d, pm, dtype):
    A = heads * d
    has_t, has_v = bool(pm & 7), bool(pm & 14)
    qt = rnd(B * Lt, 3 * A, dtype=dtype, scale=0.7) if has_t else None
    qv = rnd(B * Lv, 3 * A, dtype=dtype, scale=0.7, seed=1) if has_v else None
    lt = ([Lt, max(3, Lt // 2), Lt - 1] * ((B + 2) // 3))[:B]
    lv = ([Lv - 2, Lv, max(2, Lv // 3)] * ((B + 2) // 3))[:B]
    mt = (torch.arange(Lt)[None] < torch.tensor(lt)[:, None]).int().to(DEV)
    mv = (torch.arange(Lv)[None] < torch.tensor(lv)[:, None]).int().to(DEV)
    return qt, qv, mt, mv, A


@pytest.mark.parametrize("backend", ["tc", "mma", "simt"])
@pytest.mark.parametrize("B,Lt,Lv,heads,d,pm", ATTN_CASES + [(7, 36, 37, 12, 64, 6), (9, 20, 25, 4, 64, 15), (5, 30, 101, 8, 128, 15)])
def test_attention_fwd_bwd(backend, B, Lt, Lv, heads, d, pm):
    """backend "tc": forward on the tcgen05 / TMEM / TMA kernel (several samples per 128-row tile), backward on the
    warp-level kernel from the lse the tcgen05 forward saved."""
    dtype = torch.bfloat16
    qt, qv, mt, mv, A = _attn_inputs(B, Lt, Lv, heads, d, pm, dtype)
    tq, vq = bool(pm & 3), bool(pm & 12)
    be = {"mma": _lib.ATTN_MMA, "simt": _lib.ATTN_SIMT, "tc": _lib.ATTN_TC}[backend]

    def bufs():
        return (torch.zeros(B * Lt, A, dtype=dtype, device=DEV) if tq else None,
                torch.zeros(B * Lv, A, dtype=dtype, device=DEV) if vq else None,
                torch.zeros(B, heads, Lt, device=DEV), torch.zeros(B, heads, Lv, device=DEV))
    ct, cv, lt, lv = bufs()
    rt_, rv_, rlt, rlv = bufs()
    args = (qt, qv, mt if qt is not None else None, mv if qv is not None else None)
    ops.attention_fwd(*args, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, be)
    emu_ops.attention_fwd(*args, rt_, rv_, rlt, rlv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, 0)
    for got, ref, on in ((ct, rt_, tq), (cv, rv_, vq)):
        if on:
            assert rel(got.float(), ref.float()) < 1e-2
    if tq:
        assert torch.allclose(lt, rlt, atol=2e-2)
    if vq:
        assert torch.allclose(lv, rlv, atol=2e-2)
    dct = rnd(B * Lt, A, dtype=dtype, seed=7) if tq else None
    dcv = rnd(B * Lv, A, dtype=dtype, seed=8) if vq else None
    dqt = torch.full_like(qt, 7.0) if qt is not None else None
    dqv = torch.full_like(qv, 7.0) if qv is not None else None
    rqt = torch.zeros_like(qt) if qt is not None else None
    rqv = torch.zeros_like(qv) if qv is not None else None
    ops.attention_bwd(*args, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, be, dctx_t=dct, dctx_v=dcv, dqkv_t=dqt, dqkv_v=dqv)
    emu_ops.attention_bwd(*args, rt_, rv_, rlt, rlv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, 0, dctx_t=dct, dctx_v=dcv, dqkv_t=rqt, dqkv_v=rqv)
    for got, ref in ((dqt, rqt), (dqv, rqv)):
        if got is not None:
            assert rel(got.float(), ref.float()) < 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Lt,Lv,heads,d,pm", [(3, 36, 37, 12, 64, 1), (3, 36, 37, 12, 64, 8), (3, 36, 37, 12, 64, 6),
                                                (3, 36, 37, 12, 64, 15), (2, 30, 101, 8, 128, 15), (5, 10, 7, 2, 64, 9)])
def test_attention_probs(dtype, B, Lt, Lv, heads, d, pm):
    """vb_attention_probs (the data of config.visualization): rows sum to 1 over the enabled key streams, masked keys
    get 0, values match softmax(QK^T/sqrt(d) + mask) of the emulation."""
    qt, qv, mt, mv, A = _attn_inputs(B, Lt, Lv, heads, d, pm, dtype)
    wt, wv = (Lt if qt is not None else 0), (Lv if qv is not None else 0)
    tq, vq = bool(pm & 3), bool(pm & 12)

    def bufs():
        return (torch.full((B, heads, Lt, wt + wv), 7.0, device=DEV) if tq else None,
                torch.full((B, heads, Lv, wt + wv), 7.0, device=DEV) if vq else None)
    args = (qt, qv, mt if qt is not None else None, mv if qv is not None else None)
    pt, pv = bufs()
    rt_, rv_ = bufs()
    ops.attention_probs(*args, pt, pv, B, Lt, Lv, heads, d, pm)
    emu_ops.attention_probs(*args, rt_, rv_, B, Lt, Lv, heads, d, pm)
    for got, ref in ((pt, rt_), (pv, rv_)):
        if got is None:
            continue
        assert torch.allclose(got, ref, atol=2e-6 if dtype == torch.float32 else 2e-3), float((got - ref).abs().max())
        assert torch.allclose(got.sum(-1), torch.ones_like(got[..., 0]), atol=1e-5)
        for off, m, on in ((0, mt, wt), (wt, mv, wv)):
            if on:
                dead = (m == 0)[:, None, None, :].expand(B, heads, got.shape[2], m.shape[1])
                assert float(got[..., off:off + m.shape[1]][dead].abs().max() if dead.any() else 0.0) == 0.0


@pytest.mark.parametrize("B,Lt,Lv,heads,d,pm", [(4, 36, 37, 12, 64, 1), (4, 36, 37, 12, 64, 9), (4, 36, 37, 12, 64, 6),
                                                (4, 36, 37, 12, 64, 15), (4, 30, 101, 8, 128, 6), (4, 56, 80, 4, 64, 15)])
@pytest.mark.parametrize("fwd_backend", ["mma", "tc"])
def test_attention_skips_trailing_padding(fwd_backend, B, Lt, Lv, heads, d, pm):
    """vb_attn_args::len_*: tiles of trailing padding are skipped.  Valid rows must match the full computation (masked
    keys have weight exactly 0), padded query rows come out as zeros, gradients of valid rows are unchanged when the
    upstream gradient of the padded rows is zero (nothing downstream reads them); a hole in the mask below the length
    still goes through the additive mask."""
    dtype = torch.bfloat16
    qt, qv, _, _, A = _attn_inputs(B, Lt, Lv, heads, d, pm, dtype)
    lens_t = torch.tensor([Lt, 5, 17, 0][:B])
    lens_v = torch.tensor([3, Lv, 0, 20][:B])
    mt = (torch.arange(Lt)[None] < lens_t[:, None]).int()
    mv = (torch.arange(Lv)[None] < lens_v[:, None]).int()
    mt[0, 2] = 0                      # holes below the length
    mv[1, 1] = 0
    mt, mv = mt.to(DEV), mv.to(DEV)
    lt_dev = ops.mask_lengths(mt, torch.zeros(B, dtype=torch.int32, device=DEV))
    lv_dev = ops.mask_lengths(mv, torch.zeros(B, dtype=torch.int32, device=DEV))
    assert lt_dev.tolist() == lens_t.tolist() and lv_dev.tolist() == lens_v.tolist()
    tq, vq = bool(pm & 3), bool(pm & 12)
    has_t, has_v = qt is not None, qv is not None
    vt = (torch.arange(Lt)[None] < lens_t[:, None]).reshape(-1).to(DEV)   # rows below the length (holes included)
    vv = (torch.arange(Lv)[None] < lens_v[:, None]).reshape(-1).to(DEV)
    # samples with no valid position are computed in full: treat all their rows as valid
    vt = vt | (lens_t == 0).repeat_interleave(Lt).to(DEV)
    vv = vv | (lens_v == 0).repeat_interleave(Lv).to(DEV)
    dct = rnd(B * Lt, A, dtype=dtype, seed=7) * vt[:, None] if tq else None
    dcv = rnd(B * Lv, A, dtype=dtype, seed=8) * vv[:, None] if vq else None
    res = {}
    for name, (l_t, l_v) in dict(full=(None, None), skip=(lt_dev, lv_dev)).items():
        ct = torch.full((B * Lt, A), 3.0, dtype=dtype, device=DEV) if tq else None
        cv = torch.full((B * Lv, A), 3.0, dtype=dtype, device=DEV) if vq else None
        lt, lv = torch.zeros(B, heads, Lt, device=DEV), torch.zeros(B, heads, Lv, device=DEV)
        args = (qt, qv, mt if has_t else None, mv if has_v else None)
        ops.attention_fwd(*args, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0,
                          _lib.ATTN_TC if fwd_backend == "tc" else _lib.ATTN_MMA, len_t=l_t, len_v=l_v)
        dqt = torch.full_like(qt, 7.0) if has_t else None
        dqv = torch.full_like(qv, 7.0) if has_v else None
        ops.attention_bwd(*args, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, _lib.ATTN_MMA, dctx_t=dct, dctx_v=dcv,
                          dqkv_t=dqt, dqkv_v=dqv, len_t=l_t, len_v=l_v)
        res[name] = (ct, cv, dqt, dqv)
    for i, valid in ((0, vt), (1, vv)):
        full, skip = res["full"][i], res["skip"][i]
        if full is not None:
            assert rel(skip[valid].float(), full[valid].float()) < 2e-3
            assert bool((skip[~valid] == 0).all()), "padded query rows must be written as zeros"
        fullg, skipg = res["full"][2 + i], res["skip"][2 + i]
        if fullg is not None:
            assert torch.isfinite(skipg.float()).all()
            assert rel(skipg[valid].float(), fullg[valid].float()) < 5e-3
            if bool((~valid).any()):
                assert float(skipg[~valid].float().abs().max()) == 0.0


@pytest.mark.parametrize("B,Lt,Lv,heads,d,pm", [(4, 36, 37, 12, 64, 1), (4, 36, 37, 12, 64, 9), (4, 36, 37, 12, 64, 6),
                                                (4, 30, 101, 8, 128, 6), (4, 56, 80, 4, 64, 15)])
@pytest.mark.parametrize("fwd_backend", ["mma", "tc"])
def test_attention_on_packed_rows(fwd_backend, B, Lt, Lv, heads, d, pm):
    """vb_attn_args::cu_*: sample b owns rows [cu[b], cu[b] + len[b]) of the qkv / ctx / dctx / dqkv buffers.  Against the
    padded layout with the same lengths: identical context rows and gradients for every valid position; rows beyond
    the packed count are not touched."""
    dtype = torch.bfloat16
    qt, qv, _, _, A = _attn_inputs(B, Lt, Lv, heads, d, pm, dtype)
    lens = {"t": torch.tensor([Lt, 5, 17, 9][:B]), "v": torch.tensor([3, Lv, 11, 20][:B])}
    L = {"t": Lt, "v": Lv}
    mask, ldev, cu, keep, n = {}, {}, {}, {}, {}
    for k in "tv":
        m = (torch.arange(L[k])[None] < lens[k][:, None]).int()
        if k == "t":
            m[0, 2] = 0                       # a hole below the length stays a (masked) row of the packed sample
        mask[k] = m.to(DEV)
        ldev[k] = lens[k].to(torch.int32).to(DEV)
        cu[k] = (torch.cumsum(lens[k], 0) - lens[k]).to(torch.int32).to(DEV)
        keep[k] = (torch.arange(L[k])[None] < lens[k][:, None]).reshape(-1).to(DEV)
        n[k] = int(lens[k].sum())
    tq, vq = bool(pm & 3), bool(pm & 12)
    q = {"t": qt, "v": qv}
    dctx = {"t": rnd(B * Lt, A, dtype=dtype, seed=7) if tq else None, "v": rnd(B * Lv, A, dtype=dtype, seed=8) if vq else None}

    def pack(x, k):
        if x is None:
            return None
        y = torch.full_like(x, 9.0)
        y[:n[k]] = x[keep[k]]
        return y
    res = {}
    for name in ("padded", "packed"):
        pk = name == "packed"
        qq = {k: (pack(q[k], k) if pk else q[k]) for k in "tv"}
        dc = {k: (pack(dctx[k], k) if pk else dctx[k]) for k in "tv"}
        ct = torch.full((B * Lt, A), 3.0, dtype=dtype, device=DEV) if tq else None
        cv = torch.full((B * Lv, A), 3.0, dtype=dtype, device=DEV) if vq else None
        lt, lv = torch.zeros(B, heads, Lt, device=DEV), torch.zeros(B, heads, Lv, device=DEV)
        args = (qq["t"], qq["v"], mask["t"] if qt is not None else None, mask["v"] if qv is not None else None)
        kw = dict(len_t=ldev["t"], len_v=ldev["v"])
        if pk:
            kw.update(cu_t=cu["t"], cu_v=cu["v"])
        ops.attention_fwd(*args, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, _lib.ATTN_TC if fwd_backend == "tc" else _lib.ATTN_MMA, **kw)
        dqt = torch.full_like(qt, 7.0) if qt is not None else None
        dqv = torch.full_like(qv, 7.0) if qv is not None else None
        ops.attention_bwd(*args, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, _lib.ATTN_MMA, dctx_t=dc["t"], dctx_v=dc["v"],
                          dqkv_t=dqt, dqkv_v=dqv, **kw)
        res[name] = dict(t=(ct, dqt), v=(cv, dqv))
    for k in "tv":
        for i in (0, 1):
            a, b = res["padded"][k][i], res["packed"][k][i]
            if a is None:
                continue
            assert torch.equal(b[:n[k]], a[keep[k]]), (k, i)
            fill = 3.0 if i == 0 else 7.0
            assert bool((b[n[k]:] == fill).all()), "rows beyond the packed count were written"


def test_attention_simt_fp32_exact():
    B, Lt, Lv, heads, d, pm = 2, 12, 9, 2, 32, 15
    qt, qv, mt, mv, A = _attn_inputs(B, Lt, Lv, heads, d, pm, torch.float32)
    ct, cv = torch.zeros(B * Lt, A, device=DEV), torch.zeros(B * Lv, A, device=DEV)
    lt, lv = torch.zeros(B, heads, Lt, device=DEV), torch.zeros(B, heads, Lv, device=DEV)
    rt_, rv_, rlt, rlv = torch.zeros_like(ct), torch.zeros_like(cv), torch.zeros_like(lt), torch.zeros_like(lv)
    ops.attention_fwd(qt, qv, mt, mv, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, _lib.ATTN_SIMT)
    emu_ops.attention_fwd(qt, qv, mt, mv, rt_, rv_, rlt, rlv, B, Lt, Lv, heads, d, pm, 0.0, 0.0, 0, 0)
    assert rel(ct, rt_) < 1e-5 and rel(cv, rv_) < 1e-5


def test_attention_dropout_fwd_bwd_consistent():
    """With dropout on, the MMA and SIMT back-ends must drop the same probabilities (same Philox indexing),
    and the backward must regenerate the forward's mask."""
    B, Lt, Lv, heads, d, pm = 2, 36, 37, 4, 64, 15
    qt, qv, mt, mv, A = _attn_inputs(B, Lt, Lv, heads, d, pm, torch.bfloat16)
    outs = []
    for be in (_lib.ATTN_MMA, _lib.ATTN_SIMT):
        ct, cv = torch.zeros(B * Lt, A, dtype=torch.bfloat16, device=DEV), torch.zeros(B * Lv, A, dtype=torch.bfloat16, device=DEV)
        lt, lv = torch.zeros(B, heads, Lt, device=DEV), torch.zeros(B, heads, Lv, device=DEV)
        ops.attention_fwd(qt, qv, mt, mv, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.3, 0.3, 99, be)
        dct, dcv = rnd(B * Lt, A, seed=7), rnd(B * Lv, A, seed=8)
        dqt, dqv = torch.zeros_like(qt), torch.zeros_like(qv)
        ops.attention_bwd(qt, qv, mt, mv, ct, cv, lt, lv, B, Lt, Lv, heads, d, pm, 0.3, 0.3, 99, be,
                          dctx_t=dct, dctx_v=dcv, dqkv_t=dqt, dqkv_v=dqv)
        outs.append((ct, cv, dqt, dqv))
    for a, b in zip(*outs):
        assert rel(a.float(), b.float()) < 2e-2


# ------------------------------------------------------------------ embeddings / losses / glue
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_embed_gather_and_box(dtype):
    rows, H, V = 500, 768, 1000
    g = torch.Generator().manual_seed(0)
    tabs = [rnd(V, H, dtype=torch.float32), rnd(64, H, dtype=torch.float32, seed=1), rnd(2, H, dtype=torch.float32, seed=2)]
    idx = [torch.randint(0, V, (rows,), generator=g).to(DEV), torch.randint(0, 64, (rows,), generator=g).to(DEV), None]
    out, ref = torch.zeros(rows, H, dtype=dtype, device=DEV), torch.zeros(rows, H, dtype=dtype, device=DEV)
    ops.embed_gather_fwd(tabs, idx, [0, 0, 1], out)
    emu_ops.embed_gather_fwd(tabs, idx, [0, 0, 1], ref)
    assert rel(out.float(), ref.float()) < 1e-6
    dout = rnd(rows, H, dtype=dtype, seed=5)
    gt = [torch.zeros_like(t) for t in tabs]
    gr = [torch.zeros_like(t) for t in tabs]
    ops.embed_gather_bwd(gt, idx, [0, 0, 1], dout)
    emu_ops.embed_gather_bwd(gr, idx, [0, 0, 1], dout)
    for a, b in zip(gt, gr):
        assert rel(a, b) < 1e-4
    # padding row skipped (nn.Embedding(padding_idx)) + keyed column sums for the tiny token-type table
    gt2, gr2 = torch.zeros_like(tabs[0]), torch.zeros_like(tabs[0])
    skip_id = int(idx[0][0])
    ops.embed_gather_bwd([gt2], [idx[0]], [0], dout, skip=[skip_id])
    emu_ops.embed_gather_bwd([gr2], [idx[0]], [0], dout, skip=[skip_id])
    assert rel(gt2, gr2) < 1e-4 and float(gt2[skip_id].abs().sum()) == 0.0
    keys = torch.randint(0, 3, (rows,), generator=g).to(DEV)
    kt, kr = torch.zeros(3, H, device=DEV), torch.zeros(3, H, device=DEV)
    ops.colsum_by_key(dout, keys, 3, kt)
    emu_ops.colsum_by_key(dout, keys, 3, kr)
    assert rel(kt, kr) < 1e-4
    loc = torch.rand(rows, 5, generator=g).to(DEV)
    W, b = rnd(H, 5, dtype=torch.float32), rnd(H, dtype=torch.float32)
    o2, r2 = out.clone(), out.clone()
    ops.box_project_fwd(loc, W, b, o2, accum=True)
    emu_ops.box_project_fwd(loc, W, b, r2, accum=True)
    assert rel(o2.float(), r2.float()) < 1e-5
    dW, db, rW, rb = torch.zeros_like(W), torch.zeros_like(b), torch.zeros_like(W), torch.zeros_like(b)
    ops.box_project_bwd(loc, dout, dW, db)
    emu_ops.box_project_bwd(loc, dout, rW, rb)
    assert rel(dW, rW) < 1e-4 and rel(db, rb) < 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("C_", [30522, 2, 1601])
def test_xent(dtype, C_):
    rows = 96
    ld = (C_ + 7) // 8 * 8
    g = torch.Generator().manual_seed(1)
    logits = rnd(rows, ld, dtype=dtype, scale=2.0)
    labels = torch.randint(0, C_, (rows,), generator=g)
    labels[torch.rand(rows, generator=g) < 0.6] = -1
    labels = labels.to(DEV)
    acc, acc_r = torch.zeros(2, device=DEV), torch.zeros(2, device=DEV)
    lse, lse_r = torch.zeros(rows, device=DEV), torch.zeros(rows, device=DEV)
    ops.xent_fwd(logits, labels, C_, lse, acc[0:1], acc[1:2])
    emu_ops.xent_fwd(logits, labels, C_, lse_r, acc_r[0:1], acc_r[1:2])
    assert torch.allclose(acc, acc_r, rtol=1e-4) and torch.allclose(lse, lse_r, atol=1e-4)
    dl = torch.tensor([0.7], device=DEV)
    d, d_r = torch.full_like(logits, 3.0), torch.zeros_like(logits)
    ops.xent_bwd(logits, labels, C_, lse, acc[1:2], dl, d)
    emu_ops.xent_bwd(logits, labels, C_, lse_r, acc_r[1:2], dl, d_r)
    assert rel(d[:, :C_].float(), d_r[:, :C_].float()) < (1e-5 if dtype == torch.float32 else 6e-3)
    # in place
    ops.xent_bwd(logits, labels, C_, lse, acc[1:2], dl, logits)
    assert rel(logits[:, :C_].float(), d_r[:, :C_].float()) < (1e-5 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("skip", [0, 1])
def test_kl(dtype, skip):
    B, V, C_ = 6, 9, 1601
    per = V - 1
    ld = 1608
    g = torch.Generator().manual_seed(2)
    logits = rnd(B * V, ld, dtype=dtype, scale=2.0)
    label = torch.where(torch.rand(B * per, generator=g) < 0.3, 1, -1).to(DEV)
    target = torch.softmax(3 * torch.randn(B * per, C_, generator=g), -1).to(DEV)
    acc, acc_r = torch.zeros(2, device=DEV), torch.zeros(2, device=DEV)
    lse, lse_r = torch.zeros(B * per, device=DEV), torch.zeros(B * per, device=DEV)
    ops.kl_fwd(logits, per, V, skip, label, target, C_, 1.5, lse, acc[0:1], acc[1:2])
    emu_ops.kl_fwd(logits, per, V, skip, label, target, C_, 1.5, lse_r, acc_r[0:1], acc_r[1:2])
    assert abs(float(acc[0]) - float(acc_r[0])) < 1e-3 * max(1.0, abs(float(acc_r[0]))) and float(acc[1]) == float(acc_r[1])
    dl = torch.tensor([1.3], device=DEV)
    d, d_r = torch.full_like(logits, 5.0), torch.zeros_like(logits)
    ops.kl_bwd(logits, per, V, skip, label, target, C_, 1.5, lse, acc[1:2], dl, d)
    emu_ops.kl_bwd(logits, per, V, skip, label, target, C_, 1.5, lse_r, acc_r[1:2], dl, d_r)
    assert rel(d[:, :C_].float(), d_r[:, :C_].float()) < (1e-4 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("kind,C_,skip,with_conf", [("xent", 1600, 0, True), ("xent", 1601, 1, False), ("xent", 400, 0, True),
                                                     ("mse", 2048, 0, False), ("huber", 2048, 0, False)])
def test_region_criteria(dtype, kind, C_, skip, with_conf):
    """xent_1600 / xent_400 / xent_1601 / mse_2048 / huber_2048 (volta/losses.py) against their restatement."""
    B, per = 5, 9
    V = per + skip
    code = {"xent": _lib.REGION_XENT, "mse": _lib.REGION_MSE, "huber": _lib.REGION_HUBER}[kind]
    ld = (C_ + 7) // 8 * 8
    logits = rnd(B * V, ld, dtype=dtype, scale=2.0)
    g = torch.Generator().manual_seed(3)
    label = torch.where(torch.rand(B * per, generator=g) < 0.4, 1, -1).to(DEV)
    idx = torch.randint(0, C_, (B * per,), generator=g).to(DEV) if kind == "xent" else None
    conf = torch.rand(B * per, generator=g).to(DEV) if with_conf else None
    vec = rnd(B * per, C_, dtype=torch.float32, seed=5) if kind != "xent" else None
    w = 6.667 if kind == "xent" else 6.667 / C_
    outs = []
    for mod in (ops, emu_ops):
        acc = torch.zeros(2, device=DEV)
        lse = torch.zeros(B * per, device=DEV)
        mod.region_loss_fwd(code, logits, per, V, skip, label, idx, conf, vec, C_, w, lse, acc[0:1], acc[1:2], True)
        d = torch.full_like(logits, 9.0)
        mod.region_loss_bwd(code, logits, per, V, skip, label, idx, conf, vec, C_, w, lse, acc[1:2], torch.tensor([0.7], device=DEV), d)
        outs.append((acc.clone(), d[:, :C_].float().clone()))
    (a0, d0), (a1, d1) = outs
    assert float(a0[1]) == float(a1[1]) == float((label == 1).sum())
    assert abs(float(a0[0]) - float(a1[0])) < 2e-3 * max(1.0, abs(float(a1[0])))
    assert rel(d0, d1) < (1e-4 if dtype == torch.float32 else 8e-3)


def test_eltwise_colsum_gather():
    a, b = rnd(300, 1024), rnd(300, 1024, seed=1)
    for op in (_lib.ELT_ADD, _lib.ELT_MUL, _lib.ELT_RELU, _lib.ELT_DRELU, _lib.ELT_GELU, _lib.ELT_DGELU, _lib.ELT_SWISH, _lib.ELT_DSWISH):
        o, r = torch.empty_like(a), torch.empty_like(a)
        ops.eltwise(op, a, o, b=b)
        emu_ops.eltwise(op, a, r, b=b)
        assert rel(o.float(), r.float()) < 4e-3, op
    f32 = torch.empty(300, 1024, device=DEV)
    ops.eltwise(_lib.ELT_COPY, a, f32)
    assert torch.equal(f32, a.float())
    cs, cr = torch.zeros(1024, device=DEV), a.float().sum(0)
    ops.colsum(a, cs)
    assert rel(cs, cr) < 1e-5
    for rows, cols, dtype in ((9472, 2304, torch.bfloat16), (1000, 776, torch.float32), (77, 1024, torch.bfloat16), (333, 100, torch.bfloat16)):
        big = rnd(rows, cols, dtype=dtype, seed=3)      # vector path (and its fall-backs: odd widths, few rows)
        cs = torch.full((cols,), 1.0, device=DEV)       # accumulates into the existing content
        ops.colsum(big, cs)
        assert rel(cs, big.float().sum(0) + 1.0) < 1e-5, (rows, cols, dtype)
    sl = rnd(512, 2048, seed=4)[:, 1024:]                # strided column slice
    cs = torch.zeros(1024, device=DEV)
    ops.colsum(sl, cs)
    assert rel(cs, sl.float().sum(0)) < 1e-5
    idx = torch.randperm(300)[:100].to(DEV)
    o = torch.empty(100, 1024, dtype=torch.bfloat16, device=DEV)
    ops.gather_rows(a, idx, o)
    assert torch.equal(o, a[idx])
    w = rnd(100000, dtype=torch.float32)
    h = torch.empty(100000, dtype=torch.bfloat16, device=DEV)
    ops.cast_bf16(w, h)
    assert torch.equal(h, w.bfloat16())
    keep = torch.empty_like(a)
    ops.eltwise(_lib.ELT_DROPOUT, torch.ones_like(a), keep, p=0.25, seed=5)
    frac = float((keep.float() > 0).float().mean())
    assert abs(frac - 0.75) < 0.01

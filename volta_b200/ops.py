"""Tensor-level wrappers over the C ABI: they take torch tensors (used only as device buffers), pass raw
pointers, extents and the current CUDA stream to libvolta_b200.so, and return.  No arithmetic happens
in PyTorch here."""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("volta_b200 kernels take float32 or bfloat16 tensors, got %s" % t.dtype)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _rows2d(t):
    """(rows, cols, ld) of a tensor viewed as a row-major matrix over its last dim."""
    assert t.stride(-1) == 1 or t.shape[-1] == 1, "innermost dimension must be contiguous"
    cols = t.shape[-1]
    rows = t.numel() // cols
    if t.dim() >= 2:
        ld = t.stride(-2)
        for i in range(t.dim() - 2):
            assert t.stride(i) == t.stride(i + 1) * t.shape[i + 1], "leading dimensions must be collapsible"
    else:
        ld = cols
    return rows, cols, ld


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("volta_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device")


# --------------------------------------------------------------------------- GEMM
def gemm(a, b, out, *, a_kmajor=True, b_kmajor=True, epilogue=_lib.EPI_NONE, bias=None, aux_in=None, aux_out=None,
         dropout_p=0.0, seed=0, offset=0, backend=_lib.GEMM_AUTO, split_k=0, M=None, N=None, K=None, b2=None, bias2=None,
         split_M=0, m_limit=None, k_limit=None, m_limit2=None, limit_seg_M=0, out2=None, k_seg=0, k_limit2=None, m_hint=0):
    """out[M,N] = epilogue(opA . opB^T); see vb_gemm in include/volta_b200.h.  m_limit / k_limit: int32 device tensors
    (one element) holding the number of rows / reduction elements that carry data (see vb_compact_plan)."""
    require_cuda(a, b, out)
    ar, ac, lda = _rows2d(a)
    br, bc, ldb = _rows2d(b)
    m, k = (ar, ac) if a_kmajor else (ac, ar)
    n, kb = (br, bc) if b_kmajor else (bc, br)
    M = m if M is None else M
    N = n if N is None else N
    K = k if K is None else K
    assert kb >= K and k >= K, "reduction extents differ: A has %d, B has %d" % (k, kb)
    orow, ocol, ldd = _rows2d(out)
    assert orow >= M and ocol >= N, "output too small: %s for M=%d N=%d" % (tuple(out.shape), M, N)
    assert a.dtype == b.dtype
    g = _lib.GemmArgs()
    g.A, g.B, g.D = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = lda, ldb, ldd
    g.ab_dtype, g.d_dtype = dt(a), dt(out)
    g.a_kmajor, g.b_kmajor = int(a_kmajor), int(b_kmajor)
    g.epilogue, g.backend = epilogue, backend
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    g.bias = _ptr(bias)
    ld_aux = 0
    for x in (aux_in, aux_out):
        if x is not None:
            assert x.dtype == out.dtype
            _, _, ld_aux = _rows2d(x)
    g.aux_in, g.aux_out, g.ld_aux = _ptr(aux_in), _ptr(aux_out), ld_aux
    g.alpha, g.dropout_p, g.seed, g.offset, g.split_k = 1.0, float(dropout_p), int(seed), int(offset), int(split_k)
    if b2 is not None:   # row-split launch: rows >= split_M use b2 / bias2
        assert b2.shape == b.shape and b2.stride() == b.stride() and b2.dtype == b.dtype and 0 < split_M < M
        g.B2, g.bias2, g.split_M = b2.data_ptr(), _ptr(bias2), int(split_M)
    for name, lim in (("m_limit", m_limit), ("k_limit", k_limit), ("m_limit2", m_limit2)):
        if lim is not None:
            assert lim.dtype == torch.int32 and lim.is_cuda and lim.numel() >= 1
            setattr(g, name, lim.data_ptr())
    g.limit_seg_M = int(limit_seg_M)
    g.m_hint = int(m_hint)
    if out2 is not None:   # two reduction segments (weight gradients of two streams in one launch): [0, k_seg) -> out, [k_seg, K) -> out2
        assert epilogue == _lib.EPI_ACCUM and 0 < k_seg < K and k_seg % 64 == 0
        assert out2.dtype == out.dtype and out2.shape == out.shape and out2.stride() == out.stride()
        g.D2, g.k_seg = out2.data_ptr(), int(k_seg)
        if k_limit2 is not None:
            assert k_limit2.dtype == torch.int32 and k_limit2.is_cuda
            g.k_limit2 = k_limit2.data_ptr()
    _lib.call("vb_gemm", C.byref(g), _stream())
    if GEMM_BENCH is not None:  # bench.py: device time of this launch, repeated back to back (no host gaps)
        ms = C.c_float(0.0)
        _lib.call("vb_gemm_bench", C.byref(g), _stream(), 5, C.byref(ms))
        Me = M if m_limit is None else min(M, int(m_limit[0].item()))   # executed extents of a device-limited launch
        if limit_seg_M:
            Me = min(limit_seg_M, Me) + (M - limit_seg_M if m_limit2 is None else min(M - limit_seg_M, int(m_limit2[0].item())))
        if out2 is not None:
            Ke = (k_seg if k_limit is None else min(k_seg, int(k_limit[0].item()))) + \
                 (K - k_seg if k_limit2 is None else min(K - k_seg, int(k_limit2[0].item())))
        else:
            Ke = K if k_limit is None else min(K, int(k_limit[0].item()))
        GEMM_BENCH.append((2.0 * Me * N * Ke, ms.value, _lib.load().vb_gemm_backend(C.byref(g)) == _lib.GEMM_TC,
                           "M%d N%d K%d a%d b%d epi%d" % (M, N, K, a_kmajor, b_kmajor, epilogue)))
    return out


GEMM_BENCH = None  # set to a list by bench.py


def gemm_backend(a, b, out, **kw):
    """Which back-end vb_gemm would pick (for tests / bench accounting)."""
    ar, ac, lda = _rows2d(a)
    br, bc, ldb = _rows2d(b)
    a_k, b_k = kw.get("a_kmajor", True), kw.get("b_kmajor", True)
    g = _lib.GemmArgs()
    g.A, g.B, g.D = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.M, g.K = (ar, ac) if a_k else (ac, ar)
    g.N = br if b_k else bc
    g.lda, g.ldb, g.ldd = lda, ldb, _rows2d(out)[2]
    g.ab_dtype, g.d_dtype = dt(a), dt(out)
    g.a_kmajor, g.b_kmajor = int(a_k), int(b_k)
    g.epilogue, g.backend = kw.get("epilogue", 0), kw.get("backend", 0)
    return _lib.load().vb_gemm_backend(C.byref(g))


# --------------------------------------------------------------------------- LayerNorm
def _row_seqs(seqs, rows):
    """seqs = (len_a, L_a, rows_a, len_b, L_b) or None -> vb_row_seqs (trailing-padding rows are skipped)."""
    if seqs is None:
        return None
    len_a, L_a, rows_a, len_b, L_b = seqs
    assert len_a.dtype == torch.int32 and rows_a <= rows and rows_a % L_a == 0 and len_a.numel() >= rows_a // L_a
    q = _lib.RowSeqs()
    q.len_a, q.L_a, q.rows_a = len_a.data_ptr(), int(L_a), int(rows_a)
    if len_b is not None and rows > rows_a:
        assert len_b.dtype == torch.int32 and (rows - rows_a) % L_b == 0 and len_b.numel() >= (rows - rows_a) // L_b
        q.len_b, q.L_b = len_b.data_ptr(), int(L_b)
    else:
        assert rows == rows_a
    return C.byref(q)


def layernorm_fwd(x, gamma, beta, y, mean=None, rstd=None, eps=1e-12, out_scale=1.0, accum=False, seqs=None,
                  second=None, split_row=0):
    """second = (gamma2, beta2): rows [split_row, rows) are normalised with the second module's parameters."""
    require_cuda(x, y)
    rows, H, ld = _rows2d(x)
    assert ld == H and _rows2d(y)[2] == H, "LayerNorm rows must be dense"
    g2, b2 = (second[0].data_ptr(), second[1].data_ptr()) if second is not None else (None, None)
    _lib.call("vb_layernorm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(mean), _ptr(rstd),
              rows, H, eps, out_scale, int(accum), _row_seqs(seqs, rows), g2, b2, int(split_row), dt(x), _stream())
    return y


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, out_scale=1.0, dx_drop=None, dropout_p=0.0, seed=0,
                  offset=0, dbias=None, seqs=None, second=None, split_row=0):
    """second = (gamma2, dgamma2, dbeta2, dbias2 or None): the rows [split_row, rows) belong to a second LayerNorm."""
    require_cuda(dy, x, dx)
    rows, H, ld = _rows2d(x)
    assert ld == H
    g2, dg2, db2, dbi2 = (_ptr(t) for t in second) if second is not None else (None, None, None, None)
    _lib.call("vb_layernorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
              dx.data_ptr(), _ptr(dgamma), _ptr(dbeta), rows, H, out_scale, _ptr(dx_drop), float(dropout_p), int(seed),
              int(offset), _ptr(dbias), _row_seqs(seqs, rows), g2, dg2, db2, dbi2, int(split_row), dt(x), _stream())
    return dx


# --------------------------------------------------------------------------- attention
def _attn_args(qkv_t, qkv_v, mask_t, mask_v, ctx_t, ctx_v, lse_t, lse_v, B, Lt, Lv, heads, d, pair_mask, p_t, p_v,
               seed, backend, dctx_t=None, dctx_v=None, dqkv_t=None, dqkv_v=None, len_t=None, len_v=None, cu_t=None, cu_v=None):
    a = _lib.AttnArgs()
    ref = qkv_t if qkv_t is not None else qkv_v
    require_cuda(ref)
    a.qkv_t, a.qkv_v = _ptr(qkv_t), _ptr(qkv_v)
    a.ld_t = qkv_t.stride(0) if qkv_t is not None else 0
    a.ld_v = qkv_v.stride(0) if qkv_v is not None else 0
    for m in (mask_t, mask_v):
        assert m is None or m.dtype == torch.int32
    a.mask_t, a.mask_v = _ptr(mask_t), _ptr(mask_v)
    a.ctx_t, a.ctx_v = _ptr(ctx_t), _ptr(ctx_v)
    a.ld_ctx_t = ctx_t.stride(0) if ctx_t is not None else 0
    a.ld_ctx_v = ctx_v.stride(0) if ctx_v is not None else 0
    a.lse_t, a.lse_v = _ptr(lse_t), _ptr(lse_v)
    a.dctx_t, a.dctx_v, a.dqkv_t, a.dqkv_v = _ptr(dctx_t), _ptr(dctx_v), _ptr(dqkv_t), _ptr(dqkv_v)
    a.B, a.Lt, a.Lv, a.heads, a.d, a.pair_mask = B, Lt, Lv, heads, d, pair_mask
    a.dtype = dt(ref)
    a.scale = 1.0 / (d ** 0.5)
    a.dropout_p_t, a.dropout_p_v, a.seed, a.offset, a.backend = float(p_t), float(p_v), int(seed), 0, backend
    for ln in (len_t, len_v):
        assert ln is None or (ln.dtype == torch.int32 and ln.numel() >= B)
    a.len_t, a.len_v = _ptr(len_t if qkv_t is not None else None), _ptr(len_v if qkv_v is not None else None)
    for cu in (cu_t, cu_v):
        assert cu is None or (cu.dtype == torch.int32 and cu.numel() >= B)
    a.cu_t, a.cu_v = _ptr(cu_t if qkv_t is not None else None), _ptr(cu_v if qkv_v is not None else None)
    return a


def zero_rows(buf, row_begin, row_end, row_offset=0, max_rows=256, second=None):
    """buf[row_offset + *row_begin : row_offset + *row_end] = 0 with device-side bounds (int32 tensors);
    second = (row_begin2, row_end2, row_offset2): another range of the same buffer in the same launch."""
    _, cols, ld = _rows2d(buf)
    b2, e2, o2 = (second[0].data_ptr(), second[1].data_ptr(), int(second[2])) if second is not None else (None, None, 0)
    _lib.call("vb_zero_rows", buf.data_ptr(), ld, cols, row_begin.data_ptr(), row_end.data_ptr(), int(row_offset), b2, e2, o2,
              int(max_rows), dt(buf), _stream())
    return buf


def mask_lengths(mask, out):
    """out[b] = 1 + index of the last set position of mask[b, :] (int32 [B, L]); 0 for an all-zero row."""
    assert mask.dtype == torch.int32 and mask.dim() == 2 and mask.is_contiguous() and out.dtype == torch.int32
    _lib.call("vb_mask_lengths", mask.data_ptr(), mask.shape[0], mask.shape[1], out.data_ptr(), _stream())
    return out


def attention_fwd(*args, **kw):
    a = _attn_args(*args, **kw)
    _lib.call("vb_attention_fwd", C.byref(a), _stream())


def attention_probs(qkv_t, qkv_v, mask_t, mask_v, probs_t, probs_v, B, Lt, Lv, heads, d, pair_mask):
    """vb_attention_probs: the probabilities the fused kernels never materialise, for inspection (visualization).
    probs_t fp32 [B, heads, Lt, Lt+Lv] / probs_v fp32 [B, heads, Lv, Lt+Lv] (None for a stream without queries)."""
    a = _attn_args(qkv_t, qkv_v, mask_t, mask_v, None, None, None, None, B, Lt, Lv, heads, d, pair_mask, 0.0, 0.0, 0, _lib.ATTN_SIMT)
    for t in (probs_t, probs_v):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda)
    _lib.call("vb_attention_probs", C.byref(a), _ptr(probs_t), _ptr(probs_v), _stream())


def attention_bwd(*args, **kw):
    a = _attn_args(*args, **kw)
    _lib.call("vb_attention_bwd", C.byref(a), _stream())


# --------------------------------------------------------------------------- embeddings
def _table_arrays(tables, idx, const_rows):
    n = len(tables)
    tp = (C.c_void_p * n)(*[_ptr(t) for t in tables])
    ip = (C.c_void_p * n)(*[_ptr(i) for i in idx])
    cr = (C.c_int64 * n)(*[int(c) for c in const_rows])
    return n, tp, ip, cr


def embed_gather_fwd(tables, idx, const_rows, out, accum=False):
    rows, H, ld = _rows2d(out)
    for i in idx:
        assert i is None or (i.dtype == torch.int64 and i.is_contiguous() and i.numel() == rows)
    n, tp, ip, cr = _table_arrays(tables, idx, const_rows)
    _lib.call("vb_embed_gather_fwd", tp, ip, cr, n, out.data_ptr(), ld, rows, H, int(accum), dt(out), _stream())
    return out


def embed_gather_bwd(dtables, idx, const_rows, dout, skip=None):
    rows, H, ld = _rows2d(dout)
    n, tp, ip, cr = _table_arrays(dtables, idx, const_rows)
    sk = (C.c_int64 * n)(*[int(-1 if skip is None or s is None else s) for s in (skip or [None] * n)])
    _lib.call("vb_embed_gather_bwd", tp, ip, cr, sk, n, dout.data_ptr(), ld, rows, H, dt(dout), _stream())


def colsum_by_key(a, keys, nkeys, out):
    """out[keys[r], :] += a[r, :] for nkeys <= 4 distinct keys (fp32 out [nkeys, cols])."""
    rows, cols, lda = _rows2d(a)
    assert keys.dtype == torch.int64 and keys.numel() == rows and out.dtype == torch.float32
    _lib.call("vb_colsum_by_key", a.data_ptr(), dt(a), lda, rows, cols, keys.data_ptr(), nkeys, out.data_ptr(), _stream())
    return out


def box_project_fwd(loc, W, b, out, accum=False):
    rows, H, ld = _rows2d(out)
    nl = loc.shape[-1]
    assert loc.dtype == torch.float32 and loc.is_contiguous() and W.shape == (H, nl)
    _lib.call("vb_box_project_fwd", loc.data_ptr(), nl, W.data_ptr(), _ptr(b), out.data_ptr(), ld, rows, H, int(accum),
              dt(out), _stream())
    return out


def box_project_bwd(loc, dout, dW, db):
    rows, H, ld = _rows2d(dout)
    _lib.call("vb_box_project_bwd", loc.data_ptr(), loc.shape[-1], dout.data_ptr(), ld, _ptr(dW), _ptr(db), rows, H,
              dt(dout), _stream())


# --------------------------------------------------------------------------- losses
def _limit_ptr(lim):
    if lim is None:
        return None
    assert lim.dtype == torch.int32 and lim.is_cuda and lim.numel() >= 1
    return lim.data_ptr()


def xent_fwd(logits, labels, C_, lse, loss_sum, count, row_limit=None):
    rows, _, ld = _rows2d(logits)
    assert labels.dtype == torch.int64 and labels.numel() == rows
    _lib.call("vb_xent_fwd", logits.data_ptr(), ld, labels.data_ptr(), rows, C_, lse.data_ptr(), loss_sum.data_ptr(),
              count.data_ptr(), _limit_ptr(row_limit), dt(logits), _stream())


def xent_bwd(logits, labels, C_, lse, count, dloss, dlogits, row_limit=None):
    rows, _, ld = _rows2d(logits)
    assert _rows2d(dlogits)[2] == ld
    _lib.call("vb_xent_bwd", logits.data_ptr(), ld, labels.data_ptr(), rows, C_, lse.data_ptr(), count.data_ptr(),
              dloss.data_ptr(), dlogits.data_ptr(), _limit_ptr(row_limit), dt(logits), _stream())


def compact_plan(labels, pad, limits, src_index, pos, labels_c):
    """Order-preserving compaction plan of the rows with label >= 0 (vb_compact_plan)."""
    rows = labels.numel()
    assert labels.dtype == torch.int64 and labels_c.dtype == torch.int64 and labels_c.numel() >= rows
    for t_ in (limits, src_index, pos):
        assert t_.dtype == torch.int32
    assert limits.numel() >= 2 and src_index.numel() >= rows and pos.numel() >= rows
    _lib.call("vb_compact_plan", labels.data_ptr(), rows, int(pad), limits.data_ptr(), src_index.data_ptr(), pos.data_ptr(),
              labels_c.data_ptr(), _stream())


def compact_gather(src, src_index, limits, dst):
    rows, cols, ld_dst = _rows2d(dst)
    _lib.call("vb_compact_gather", src.data_ptr(), _rows2d(src)[2], src_index.data_ptr(), limits.data_ptr(), dst.data_ptr(),
              ld_dst, rows, cols, dt(src), _stream())
    return dst


def compact_scatter(src_c, pos, dst):
    rows, cols, ld_dst = _rows2d(dst)
    assert pos.numel() >= rows
    _lib.call("vb_compact_scatter", src_c.data_ptr(), _rows2d(src_c)[2], pos.data_ptr(), dst.data_ptr(), ld_dst, rows, cols,
              dt(src_c), _stream())
    return dst


def kl_fwd(logits, per, per_ld, skip, label, target, C_, weight, lse, loss_sum, count):
    _, _, ld = _rows2d(logits)
    rows = label.numel()
    assert label.dtype == torch.int64 and target.dtype == torch.float32 and target.is_contiguous()
    _lib.call("vb_kl_fwd", logits.data_ptr(), ld, per, per_ld, skip, label.data_ptr(), target.data_ptr(), rows, C_,
              float(weight), lse.data_ptr(), loss_sum.data_ptr(), count.data_ptr(), dt(logits), _stream())


def kl_bwd(logits, per, per_ld, skip, label, target, C_, weight, lse, count, dloss, dlogits):
    _, _, ld = _rows2d(logits)
    rows = label.numel()
    _lib.call("vb_kl_bwd", logits.data_ptr(), ld, per, per_ld, skip, label.data_ptr(), target.data_ptr(), rows, C_,
              float(weight), lse.data_ptr(), count.data_ptr(), dloss.data_ptr(), dlogits.data_ptr(), dt(logits),
              _stream())


def region_loss_fwd(kind, logits, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, C_, weight, lse, loss_sum, count, add_count):
    _, _, ld = _rows2d(logits)
    rows = label.numel()
    assert label.dtype == torch.int64 and (tgt_idx is None or tgt_idx.dtype == torch.int64)
    assert conf is None or (conf.dtype == torch.float32 and conf.is_contiguous())
    ld_t = 0
    if tgt_vec is not None:
        assert tgt_vec.dtype == torch.float32
        _, _, ld_t = _rows2d(tgt_vec)
    _lib.call("vb_region_loss_fwd", kind, logits.data_ptr(), ld, per, per_ld, skip, label.data_ptr(), _ptr(tgt_idx), _ptr(conf),
              _ptr(tgt_vec), ld_t, rows, C_, float(weight), _ptr(lse), loss_sum.data_ptr(), count.data_ptr(), int(add_count),
              dt(logits), _stream())


def region_loss_bwd(kind, logits, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, C_, weight, lse, count, dloss, dlogits):
    _, _, ld = _rows2d(logits)
    rows = label.numel()
    ld_t = _rows2d(tgt_vec)[2] if tgt_vec is not None else 0
    _lib.call("vb_region_loss_bwd", kind, logits.data_ptr(), ld, per, per_ld, skip, label.data_ptr(), _ptr(tgt_idx), _ptr(conf),
              _ptr(tgt_vec), ld_t, rows, C_, float(weight), _ptr(lse), count.data_ptr(), dloss.data_ptr(), dlogits.data_ptr(),
              dt(logits), _stream())


def bce_logits_fwd(logits, target, row_loss, scale, loss_out):
    """vb_bce_logits_fwd: logits [rows, C] (bf16 / fp32), target fp32 [rows, C]; loss_out[0] = scale * sum."""
    rows, C_, ld = _rows2d(logits)
    require_cuda(logits, target)
    assert target.dtype == torch.float32 and _rows2d(target)[:2] == (rows, C_) and row_loss.numel() >= rows
    _lib.call("vb_bce_logits_fwd", logits.data_ptr(), ld, target.data_ptr(), _rows2d(target)[2], rows, C_, row_loss.data_ptr(),
              float(scale), loss_out.data_ptr(), dt(logits), _stream())
    return loss_out


def bce_logits_bwd(logits, target, dloss, scale, dlogits):
    rows, C_, ld = _rows2d(logits)
    require_cuda(logits, target, dlogits)
    assert target.dtype == torch.float32 and dloss.dtype == torch.float32 and dlogits.dtype == logits.dtype
    _lib.call("vb_bce_logits_bwd", logits.data_ptr(), ld, target.data_ptr(), _rows2d(target)[2], rows, C_, dloss.data_ptr(),
              float(scale), dlogits.data_ptr(), _rows2d(dlogits)[2], dt(logits), _stream())
    return dlogits


def argmax_pick(logits, labels=None, arg=None, pick=None):
    """vb_argmax_pick over the rows of logits [rows, C]: arg int64 [rows], pick fp32 [rows] (label at the argmax, or the max)."""
    rows, C_, ld = _rows2d(logits)
    require_cuda(logits)
    assert labels is None or (labels.dtype == torch.float32 and _rows2d(labels)[:2] == (rows, C_))
    assert (arg is None or arg.dtype == torch.int64) and (pick is None or pick.dtype == torch.float32)
    _lib.call("vb_argmax_pick", logits.data_ptr(), ld, _ptr(labels), 0 if labels is None else _rows2d(labels)[2], rows, C_,
              _ptr(arg), _ptr(pick), dt(logits), _stream())
    return arg, pick


def loss_finalize(loss_sum, count, out):
    _lib.call("vb_loss_finalize", loss_sum.data_ptr(), count.data_ptr(), out.data_ptr(), out.numel(), _stream())
    return out


# --------------------------------------------------------------------------- element-wise glue
def eltwise(op, a, out, b=None, alpha=1.0, p=0.0, seed=0, offset=0):
    require_cuda(a, out)
    rows, cols, lda = _rows2d(a)
    ldb = _rows2d(b)[2] if b is not None else 0
    ldo = _rows2d(out)[2]
    _lib.call("vb_eltwise", op, a.data_ptr(), dt(a), lda, _ptr(b), dt(b) if b is not None else 0, ldb, out.data_ptr(),
              dt(out), ldo, rows, cols, float(alpha), float(p), int(seed), int(offset), _stream())
    return out


def colsum(a, out, row_limit=None):
    rows, cols, lda = _rows2d(a)
    assert out.dtype == torch.float32 and out.numel() >= cols
    _lib.call("vb_colsum", a.data_ptr(), dt(a), lda, rows, cols, out.data_ptr(), _limit_ptr(row_limit), _stream())
    return out


def gather_rows(a, idx, out):
    rows, cols, ldo = _rows2d(out)
    assert idx.dtype == torch.int64 and idx.numel() == rows
    _lib.call("vb_gather_rows", a.data_ptr(), _rows2d(a)[2], idx.data_ptr(), out.data_ptr(), ldo, rows, cols, dt(a), _stream())
    return out


def scatter_rows_add(a, idx, out):
    rows, cols, lda = _rows2d(a)
    _lib.call("vb_scatter_rows_add", a.data_ptr(), lda, idx.data_ptr(), out.data_ptr(), _rows2d(out)[2], rows, cols,
              dt(a), _stream())
    return out


def cast_bf16(src, dst):
    """fp32 -> bf16 copy of one flat buffer (weight refresh after an optimizer step)."""
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    item = torch.tensor([src.data_ptr(), dst.data_ptr(), src.numel()], dtype=torch.int64).to(src.device)
    _lib.call("vb_multi_cast_bf16", item.data_ptr(), 1, src.numel(), _stream())
    return dst


def box_sincos(boxes, dim, out):
    rows, _, ldo = _rows2d(out)
    _lib.call("vb_box_sincos", boxes.data_ptr(), boxes.shape[-1], dim, out.data_ptr(), ldo, rows, dt(out), _stream())
    return out


def sumsq(x, out):
    assert x.dtype == torch.float32 and out.dtype == torch.float32 and x.is_contiguous()
    _lib.call("vb_sumsq", x.data_ptr(), x.numel(), out.data_ptr(), _stream())
    return out


def adamw_flat(p, g, m, v, p16, chunk_group, group_hp, beta1, beta2, eps, sumsq_buf=None, max_norm=0.0, rule="adamw"):
    """rule "adamw": vb_adamw_flat, group_hp rows {lr, wd, step_size, active};  "radam": vb_radam_flat, rows
    {wd * lr, rectified, step_size, active}."""
    assert p.dtype == g.dtype == m.dtype == v.dtype == torch.float32 and chunk_group.dtype == torch.int16
    assert group_hp.dtype == torch.float32 and group_hp.shape[-1] == 4 and chunk_group.numel() * 64 == p.numel()
    _lib.call({"adamw": "vb_adamw_flat", "radam": "vb_radam_flat"}[rule], p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(p16), p.numel(),
              chunk_group.data_ptr(), group_hp.data_ptr(), float(beta1), float(beta2), float(eps), _ptr(sumsq_buf),
              float(max_norm), _stream())


def fill(buf, value=0):
    _lib.call("vb_fill", buf.data_ptr(), buf.numel() * buf.element_size(), int(value), _stream())


# --------------------------------------------------------------------------- VL-BERT input preparation
def vlbert_prep(feat, mask_emb, out, masked, kind, K, mvrc, end_kind):
    rows, F = feat.shape
    assert feat.dtype == torch.float32 and feat.is_contiguous() and mask_emb.dtype == torch.float32
    _lib.call("vb_vlbert_prep", feat.data_ptr(), mask_emb.data_ptr(), out.data_ptr(), out.stride(0), masked.data_ptr(),
              kind.data_ptr(), rows, F, K, int(mvrc), int(end_kind), dt(out), _stream())


def vlbert_index(ids, K, tpos, opos, text_end, last_row):
    B, T = ids.shape
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    _lib.call("vb_vlbert_index", ids.data_ptr(), B, T, K, tpos.data_ptr(), opos.data_ptr(), text_end.data_ptr(),
              last_row.data_ptr(), _stream())


def batched_colsum(a, batch, rows, cols, out):
    assert a.is_contiguous() and out.dtype == torch.float32 and out.numel() == batch * cols
    _lib.call("vb_batched_colsum", a.data_ptr(), dt(a), batch, rows, cols, out.data_ptr(), _stream())
    return out


def region_batch(raw_feat, raw_box, image_hw, num_boxes, sel, global_pos, feat, loc, mask, label, count):
    """Region batch assembly (vb_region_batch): detector output -> image_feat / image_loc / image_mask / image_label."""
    require_cuda(raw_feat, raw_box, image_hw, num_boxes, sel, feat, loc, mask, label, count)
    B, R, F = raw_feat.shape
    for t in (raw_feat, raw_box, image_hw, num_boxes, sel, feat, loc, mask, label, count):
        assert t.is_contiguous()
    assert raw_box.shape == (B, R, 4) and raw_box.dtype == torch.float32 and image_hw.shape == (B, 2)
    assert image_hw.dtype == torch.float32 and num_boxes.dtype == torch.int32 and sel.dtype == torch.uint8
    V = R + (1 if global_pos else 0)
    assert feat.shape == (B, V, F) and loc.shape[:2] == (B, V) and mask.shape == (B, V) and label.shape == (B, R)
    assert loc.dtype == torch.float32 and mask.dtype == torch.int64 and label.dtype == torch.int64 and count.dtype == torch.int32
    _lib.call("vb_region_batch", raw_feat.data_ptr(), dt(raw_feat), raw_box.data_ptr(), image_hw.data_ptr(),
              num_boxes.data_ptr(), sel.data_ptr(), B, R, F, loc.shape[2], global_pos, feat.data_ptr(), dt(feat),
              loc.data_ptr(), mask.data_ptr(), label.data_ptr(), count.data_ptr(), _stream())


def retrieval_ranks(scores, targets, axis, rank, topk=None):
    """vb_retrieval_ranks: places of the matching pairs in the descending order of every row (axis 0) / column (axis 1)."""
    require_cuda(scores, targets, rank, topk)
    C, I = scores.shape
    assert scores.dtype == torch.float32 and targets.dtype == torch.uint8 and targets.shape == (C, I)
    assert scores.stride(1) == 1 and targets.stride(1) == 1
    assert rank.dtype == torch.int32 and rank.is_contiguous() and rank.numel() == (C if axis == 0 else I)
    k = 0
    if topk is not None:
        assert topk.dtype == torch.int32 and topk.is_contiguous() and topk.shape[0] == C
        k = topk.shape[1]
    _lib.call("vb_retrieval_ranks", scores.data_ptr(), scores.stride(0), targets.data_ptr(), targets.stride(0), C, I, axis,
              rank.data_ptr(), _ptr(topk), k, _stream())

"""CPU emulation of the C ABI (include/volta_b200.h) in plain PyTorch - TEST INFRASTRUCTURE ONLY.

`install()` monkey-patches `volta_b200.ops` so that the engine's host logic (the launch plan, the
hand-written backward, the flat parameter store) can be exercised on a machine without a GPU and compared
with the oracle.  Each function restates the documented semantics of one entry point; the GPU tests check
the real kernels against these same functions.  Nothing in the product imports this file.
"""
import math

import torch
import torch.nn.functional as F

from volta_b200 import _lib, engine, ops


def _gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _dgelu(x):
    cdf = 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    pdf = torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    return cdf + x * pdf


def gemm(a, b, out, *, a_kmajor=True, b_kmajor=True, epilogue=_lib.EPI_NONE, bias=None, aux_in=None, aux_out=None,
         dropout_p=0.0, seed=0, offset=0, backend=0, split_k=0, M=None, N=None, K=None, b2=None, bias2=None, split_M=0,
         m_limit=None, k_limit=None, m_limit2=None, limit_seg_M=0, out2=None, k_seg=0, k_limit2=None, m_hint=0):
    if out2 is not None:   # two reduction segments: rows [0, k_seg) of the (MN-major) operands -> out, rows [k_seg, K) -> out2
        assert epilogue == _lib.EPI_ACCUM and not a_kmajor and not b_kmajor and k_seg % 64 == 0
        gemm(a[:k_seg], b[:k_seg], out, a_kmajor=False, b_kmajor=False, epilogue=epilogue, k_limit=k_limit)
        gemm(a[k_seg:], b[k_seg:], out2, a_kmajor=False, b_kmajor=False, epilogue=epilogue, k_limit=k_limit2)
        return out
    if limit_seg_M:      # two row segments with their own device-side counts (and, in a row-split launch, weights)
        assert a_kmajor and epilogue != _lib.EPI_ACCUM and (b2 is None or split_M == limit_seg_M)
        sg = limit_seg_M
        kw = dict(a_kmajor=a_kmajor, b_kmajor=b_kmajor, epilogue=epilogue, dropout_p=dropout_p, N=N, K=K)
        for rows, bb_, bias_, lim in ((slice(0, sg), b, bias, m_limit),
                                     (slice(sg, None), b if b2 is None else b2, bias if (b2 is None or bias2 is None) else bias2, m_limit2)):
            gemm(a[rows], bb_, out[rows], bias=bias_, aux_in=None if aux_in is None else aux_in[rows],
                 aux_out=None if aux_out is None else aux_out[rows], m_limit=lim, **kw)
        return out
    if m_limit is not None or k_limit is not None:   # device-side extents: only that many rows / reduction elements exist
        A0 = a if a_kmajor else a.t()
        B0 = b if b_kmajor else b.t()
        M = A0.shape[0] if M is None else M
        K = A0.shape[1] if K is None else K
        if m_limit is not None:
            M = min(M, (int(m_limit[0]) + 255) // 256 * 256)   # whole 256-row tiles, like the kernels
        if k_limit is not None:
            K = min(K, int(k_limit[0]))
        if M == 0 or K == 0:
            return out
        return gemm(a, b, out, a_kmajor=a_kmajor, b_kmajor=b_kmajor, epilogue=epilogue, bias=bias, aux_in=aux_in, aux_out=aux_out,
                    dropout_p=dropout_p, M=M, N=N, K=K)
    if b2 is not None:   # row-split launch = two ordinary GEMMs over the two row ranges
        assert a_kmajor and epilogue != _lib.EPI_ACCUM
        kw = dict(a_kmajor=a_kmajor, b_kmajor=b_kmajor, epilogue=epilogue, dropout_p=dropout_p, N=N, K=K)
        for rows, bb_, bias_ in ((slice(0, split_M), b, bias), (slice(split_M, None), b2, bias2 if bias2 is not None else bias)):
            gemm(a[rows], bb_, out[rows], bias=bias_, aux_in=None if aux_in is None else aux_in[rows],
                 aux_out=None if aux_out is None else aux_out[rows], **kw)
        return out
    A = a.float() if a_kmajor else a.float().t()
    B = b.float() if b_kmajor else b.float().t()
    M = A.shape[0] if M is None else M
    N = B.shape[0] if N is None else N
    K = A.shape[1] if K is None else K
    acc = A[:M, :K] @ B[:N, :K].t()
    o = out[:M, :N]
    ai = aux_in[:M, :N].float() if aux_in is not None else None
    bb = bias[:N].float() if bias is not None else None
    E = _lib
    assert dropout_p == 0.0, "the CPU emulation covers the deterministic (eval / p=0) path"
    if epilogue == E.EPI_NONE:
        r = acc
    elif epilogue == E.EPI_BIAS:
        r = acc + bb
    elif epilogue == E.EPI_BIAS_GELU:
        pre = acc + bb
        aux_out[:M, :N] = _dgelu(pre).to(aux_out.dtype)
        r = _gelu(pre)
    elif epilogue == E.EPI_BIAS_RELU:
        r = torch.relu(acc + bb)
    elif epilogue == E.EPI_BIAS_DROP_RES:
        r = acc + bb + ai
    elif epilogue == E.EPI_MUL:
        r = acc * ai
    elif epilogue == E.EPI_ADD:
        r = acc + ai
    elif epilogue == E.EPI_DRELU:
        r = acc * (ai > 0)
    elif epilogue == E.EPI_BIAS_ADD:
        r = acc + bb + ai
    elif epilogue == E.EPI_ACCUM:
        r = o.float() + acc
    else:
        raise ValueError(epilogue)
    o.copy_(r.to(out.dtype))
    return out


def _dead_rows(seqs, rows):
    """Rows that are trailing padding of their sample (vb_row_seqs); None = no row is skipped."""
    if seqs is None:
        return None
    len_a, L_a, rows_a, len_b, L_b = seqs
    parts = []
    for ln, L, n in ((len_a, L_a, rows_a), (len_b, L_b, rows - rows_a)):
        if n == 0:
            continue
        if ln is None:
            parts.append(torch.zeros(n, dtype=torch.bool))
            continue
        j = torch.arange(n) % L
        v = ln[:n // L].repeat_interleave(L)
        parts.append((v > 0) & (j >= v))
    return torch.cat(parts)


def _seq_slice(seqs, lo, hi, rows):
    """vb_row_seqs restricted to the row range [lo, hi) that starts a segment (emulation of two-stream launches)."""
    if seqs is None:
        return None
    len_a, L_a, rows_a, len_b, L_b = seqs
    if lo == 0:
        assert hi == rows_a
        return (len_a, L_a, rows_a, None, 0)
    assert lo == rows_a
    return None if len_b is None else (len_b, L_b, hi - lo, None, 0)


def layernorm_fwd(x, gamma, beta, y, mean=None, rstd=None, eps=1e-12, out_scale=1.0, accum=False, seqs=None, second=None,
                  split_row=0):
    if second is not None and 0 < split_row < x.shape[0]:
        rows = x.shape[0]
        for lo, hi, g_, b_ in ((0, split_row, gamma, beta), (split_row, rows, second[0], second[1])):
            layernorm_fwd(x[lo:hi], g_, b_, y[lo:hi], None if mean is None else mean[lo:hi], None if rstd is None else rstd[lo:hi],
                          eps, out_scale, accum, _seq_slice(seqs, lo, hi, rows))
        return y
    dead = None if accum else _dead_rows(seqs, x.shape[0])
    if dead is not None:
        y2, m2, r2 = torch.empty_like(y), torch.empty(x.shape[0]), torch.empty(x.shape[0])
        layernorm_fwd(x, gamma, beta, y2, m2, r2, eps, out_scale, accum)
        y.copy_(torch.where(dead[:, None], torch.zeros_like(y2), y2))
        if mean is not None:
            mean.copy_(torch.where(dead, torch.zeros_like(m2), m2))
        if rstd is not None:
            rstd.copy_(torch.where(dead, torch.zeros_like(r2), r2))
        return y
    xf = x.float()
    mu = xf.mean(-1, keepdim=True)
    var = (xf - mu).pow(2).mean(-1, keepdim=True)
    rs = torch.rsqrt(var + eps)
    r = (gamma * ((xf - mu) * rs) + beta) * out_scale
    if accum:
        r = r + y.float()
    y.copy_(r.to(y.dtype))
    if mean is not None:
        mean.copy_(mu.reshape(-1))
    if rstd is not None:
        rstd.copy_(rs.reshape(-1))
    return y


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, out_scale=1.0, dx_drop=None, dropout_p=0.0, seed=0,
                  offset=0, dbias=None, seqs=None, second=None, split_row=0):
    assert dropout_p == 0.0
    if second is not None and 0 < split_row < x.shape[0]:
        rows = x.shape[0]
        g2, dg2, db2, dbi2 = second
        for lo, hi, g_, dg_, db_, dbi_ in ((0, split_row, gamma, dgamma, dbeta, dbias), (split_row, rows, g2, dg2, db2, dbi2)):
            layernorm_bwd(dy[lo:hi], x[lo:hi], mean[lo:hi], rstd[lo:hi], g_, dx[lo:hi], dg_, db_, out_scale, None, 0.0, 0, 0, dbi_,
                          _seq_slice(seqs, lo, hi, rows))
        return dx
    dead = _dead_rows(seqs, x.shape[0])
    if dead is not None:   # skipped rows: no contribution, zero dx
        dy = torch.where(dead[:, None], torch.zeros_like(dy), dy)
        x = torch.where(dead[:, None], torch.zeros_like(x), x)
    dyo = dy.float() * out_scale
    xh = (x.float() - mean[:, None]) * rstd[:, None]
    g = dyo * gamma
    c1 = g.mean(-1, keepdim=True)
    c2 = (g * xh).mean(-1, keepdim=True)
    r = rstd[:, None] * (g - c1 - xh * c2)
    dx.copy_(r.to(dx.dtype))
    if dgamma is not None:
        dgamma += (dyo * xh).sum(0)
    if dbeta is not None:
        dbeta += dyo.sum(0)
    if dbias is not None:
        dbias += dx.float().sum(0)
    return dx


def _attn_core(qkv_t, qkv_v, mask_t, mask_v, B, Lt, Lv, heads, d, pair_mask):
    A = heads * d
    streams = {}
    for s, (qkv, L) in enumerate(((qkv_t, Lt), (qkv_v, Lv))):
        if qkv is None:
            continue
        x = qkv.float().view(B, L, -1)
        q, k, v = (x[..., i * A:(i + 1) * A].reshape(B, L, heads, d).transpose(1, 2) for i in range(3))
        streams[s] = (q, k, v)
    madd = {0: None if mask_t is None else (1.0 - mask_t.float()) * -10000.0,
            1: None if mask_v is None else (1.0 - mask_v.float()) * -10000.0}
    return A, streams, madd


def _attn_forward_math(qkv_t, qkv_v, mask_t, mask_v, B, Lt, Lv, heads, d, pair_mask):
    A, streams, madd = _attn_core(qkv_t, qkv_v, mask_t, mask_v, B, Lt, Lv, heads, d, pair_mask)
    out = {}
    for qs in streams:
        segs = [ks for ks in streams if (pair_mask >> (qs * 2 + ks)) & 1]
        if not segs:
            continue
        q = streams[qs][0]
        sc = torch.cat([q @ streams[ks][1].transpose(-1, -2) / math.sqrt(d) + madd[ks][:, None, None, :] for ks in segs], -1)
        lse = torch.logsumexp(sc, -1)
        p = torch.exp(sc - lse[..., None])
        vv = torch.cat([streams[ks][2] for ks in segs], 2)
        ctx = (p @ vv).transpose(1, 2).reshape(B * q.shape[2], A)
        out[qs] = (ctx, lse)
    return out


def _unpack_rows(x, cu, lens, B, L):
    """Packed rows (vb_attn_args::cu_*: sample b at [cu[b], cu[b] + len[b])) -> the padded [B*L, C] layout, zeros elsewhere."""
    if x is None or cu is None:
        return x
    out = torch.zeros(B * L, x.shape[1], dtype=x.dtype)
    for b in range(B):
        n, c0 = int(lens[b]), int(cu[b])
        out[b * L:b * L + n] = x[c0:c0 + n]
    return out


def _pack_rows_into(src, dst, cu, lens, B, L):
    for b in range(B):
        n, c0 = int(lens[b]), int(cu[b])
        dst[c0:c0 + n] = src[b * L:b * L + n].to(dst.dtype)


def attention_fwd(qkv_t, qkv_v, mask_t, mask_v, ctx_t, ctx_v, lse_t, lse_v, B, Lt, Lv, heads, d, pair_mask, p_t, p_v,
                  seed, backend, len_t=None, len_v=None, cu_t=None, cu_v=None, **kw):
    assert p_t == 0.0 and p_v == 0.0
    out = _attn_forward_math(_unpack_rows(qkv_t, cu_t, len_t, B, Lt), _unpack_rows(qkv_v, cu_v, len_v, B, Lv), mask_t, mask_v,
                             B, Lt, Lv, heads, d, pair_mask)
    for qs, (ctx, lse) in out.items():
        dst, ldst, cu, ln, L = (ctx_t, lse_t, cu_t, len_t, Lt) if qs == 0 else (ctx_v, lse_v, cu_v, len_v, Lv)
        if cu is None:
            dst.copy_(ctx.to(dst.dtype))
        else:
            _pack_rows_into(ctx, dst, cu, ln, B, L)
        ldst.copy_(lse)


def attention_probs(qkv_t, qkv_v, mask_t, mask_v, probs_t, probs_v, B, Lt, Lv, heads, d, pair_mask):
    A, streams, madd = _attn_core(qkv_t, qkv_v, mask_t, mask_v, B, Lt, Lv, heads, d, pair_mask)
    off = {0: 0, 1: Lt if qkv_t is not None else 0}
    for qs, out in ((0, probs_t), (1, probs_v)):
        if out is None or qs not in streams:
            continue
        segs = [ks for ks in streams if (pair_mask >> (qs * 2 + ks)) & 1]
        out.zero_()
        if not segs:
            continue
        q = streams[qs][0]
        sc = torch.cat([q @ streams[ks][1].transpose(-1, -2) / math.sqrt(d) + madd[ks][:, None, None, :] for ks in segs], -1)
        p = torch.softmax(sc, -1)
        o = 0
        for ks in segs:
            n = streams[ks][1].shape[2]
            out[..., off[ks]:off[ks] + n] = p[..., o:o + n]
            o += n


def mask_lengths(mask, out):
    idx = torch.arange(1, mask.shape[1] + 1, dtype=torch.int32)
    out.copy_(((mask != 0).to(torch.int32) * idx).max(1).values)
    return out


def attention_bwd(qkv_t, qkv_v, mask_t, mask_v, ctx_t, ctx_v, lse_t, lse_v, B, Lt, Lv, heads, d, pair_mask, p_t, p_v,
                  seed, backend, dctx_t=None, dctx_v=None, dqkv_t=None, dqkv_v=None, len_t=None, len_v=None, cu_t=None,
                  cu_v=None, **kw):
    qkv_t, dctx_t = _unpack_rows(qkv_t, cu_t, len_t, B, Lt), _unpack_rows(dctx_t, cu_t, len_t, B, Lt)
    qkv_v, dctx_v = _unpack_rows(qkv_v, cu_v, len_v, B, Lv), _unpack_rows(dctx_v, cu_v, len_v, B, Lv)
    with torch.enable_grad():
        a = qkv_t.detach().float().requires_grad_(True) if qkv_t is not None else None
        b = qkv_v.detach().float().requires_grad_(True) if qkv_v is not None else None
        out = _attn_forward_math(a, b, mask_t, mask_v, B, Lt, Lv, heads, d, pair_mask)
        tot = 0
        for qs, (ctx, _) in out.items():
            g = dctx_t if qs == 0 else dctx_v
            tot = tot + (ctx * g.float()).sum()
        tot.backward()
    for leaf, dst, cu, ln, L in ((a, dqkv_t, cu_t, len_t, Lt), (b, dqkv_v, cu_v, len_v, Lv)):
        if leaf is None:
            continue
        g = leaf.grad if leaf.grad is not None else torch.zeros_like(leaf)
        if cu is None:
            dst.copy_(g.to(dst.dtype))
        else:
            _pack_rows_into(g, dst, cu, ln, B, L)


def embed_gather_fwd(tables, idx, const_rows, out, accum=False):
    r = out.float() if accum else torch.zeros_like(out, dtype=torch.float32)
    for t, i, c in zip(tables, idx, const_rows):
        r = r + (t[i] if i is not None else t[c][None, :])
    out.copy_(r.to(out.dtype))
    return out


def colsum_by_key(a, keys, nkeys, out):
    out.index_add_(0, keys, a.float())
    return out


def embed_gather_bwd(dtables, idx, const_rows, dout, skip=None):
    d = dout.float()
    skip = skip or [None] * len(dtables)
    for t, i, c, sk in zip(dtables, idx, const_rows, skip):
        if t is None:
            continue
        if i is not None:
            t.index_add_(0, i, d if sk is None else d * (i != sk)[:, None])
        else:
            t[c] += d.sum(0)


def box_project_fwd(loc, W, b, out, accum=False):
    r = loc @ W.t() + (b if b is not None else 0)
    if accum:
        r = r + out.float()
    out.copy_(r.to(out.dtype))
    return out


def box_project_bwd(loc, dout, dW, db):
    d = dout.float()
    dW += d.t() @ loc
    if db is not None:
        db += d.sum(0)


def xent_fwd(logits, labels, C_, lse, loss_sum, count, row_limit=None):
    if row_limit is not None:
        n = int(row_limit[0])
        logits, labels, lse = logits[:n], labels[:n], lse[:n]
    x = logits[:, :C_].float()
    live = labels >= 0
    l = torch.logsumexp(x, -1)
    lse.copy_(torch.where(live, l, torch.zeros_like(l)))
    picked = x.gather(1, labels.clamp(min=0)[:, None])[:, 0]
    loss_sum += ((l - picked) * live).sum()
    count += live.sum()


def xent_bwd(logits, labels, C_, lse, count, dloss, dlogits, row_limit=None):
    if row_limit is not None:
        n = int(row_limit[0])
        logits, labels, lse, dlogits = logits[:n], labels[:n], lse[:n], dlogits[:n]
    x = logits[:, :C_].float()
    live = (labels >= 0)[:, None]
    g = dloss / count.clamp(min=1.0)
    p = torch.exp(x - lse[:, None])
    onehot = F.one_hot(labels.clamp(min=0), C_).float()
    dlogits[:, :C_] = torch.where(live, (p - onehot) * g, torch.zeros_like(p)).to(dlogits.dtype)


def _kl_rows(per, per_ld, skip, rows, device=None):
    r = torch.arange(rows, device=device)
    return (r // per) * per_ld + (r % per) + skip


def kl_fwd(logits, per, per_ld, skip, label, target, C_, weight, lse, loss_sum, count):
    rows = label.numel()
    x = logits[_kl_rows(per, per_ld, skip, rows, logits.device), :C_].float()
    live = label == 1
    l = torch.logsumexp(x, -1)
    lse.copy_(torch.where(live, l, torch.zeros_like(l)))
    kl = torch.where(target > 0, target * (torch.log(target.clamp(min=1e-45)) - (x - l[:, None])), torch.zeros_like(x)).sum(-1)
    loss_sum += weight * (kl * live).sum()
    count += live.sum()


def kl_bwd(logits, per, per_ld, skip, label, target, C_, weight, lse, count, dloss, dlogits):
    rows = label.numel()
    idx = _kl_rows(per, per_ld, skip, rows, logits.device)
    x = logits[idx, :C_].float()
    live = (label == 1)[:, None]
    g = weight * dloss / count.clamp(min=1.0)
    p = torch.exp(x - lse[:, None])
    d = g * (p * target.sum(-1, keepdim=True) - target) * live
    full = torch.zeros(dlogits.shape[0], C_, device=logits.device)
    full[idx] = d
    dlogits[:, :C_] = full.to(dlogits.dtype)


def region_loss_fwd(kind, logits, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, C_, weight, lse, loss_sum, count, add_count):
    rows = label.numel()
    x = logits[_kl_rows(per, per_ld, skip, rows, logits.device), :C_].float()
    live = label == 1
    if kind == _lib.REGION_XENT:
        l = torch.logsumexp(x, -1)
        lse.copy_(torch.where(live, l, torch.zeros_like(l)))
        per_row = (l - x.gather(1, tgt_idx[:, None]).squeeze(1)) * (conf if conf is not None else 1.0)
    else:
        d = x - tgt_vec[:, :C_]
        per_row = ((d * d) if kind == _lib.REGION_MSE else torch.where(d.abs() < 1, 0.5 * d * d, d.abs() - 0.5)).sum(-1)
    loss_sum += weight * (per_row * live).sum()
    if add_count:
        count += live.sum()


def region_loss_bwd(kind, logits, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, C_, weight, lse, count, dloss, dlogits):
    rows = label.numel()
    idx = _kl_rows(per, per_ld, skip, rows, logits.device)
    x = logits[idx, :C_].float()
    live = (label == 1)[:, None]
    g = weight * dloss / count.clamp(min=1.0)
    if kind == _lib.REGION_XENT:
        p = torch.exp(x - lse[:, None])
        onehot = torch.zeros_like(x).scatter_(1, tgt_idx[:, None], 1.0)
        d = g * (p - onehot) * (conf[:, None] if conf is not None else 1.0) * live
    else:
        e = x - tgt_vec[:, :C_]
        d = g * (2 * e if kind == _lib.REGION_MSE else e.clamp(-1, 1)) * live
    full = torch.zeros(dlogits.shape[0], C_, device=logits.device)
    full[idx] = d
    dlogits[:, :C_] = full.to(dlogits.dtype)


def bce_logits_fwd(logits, target, row_loss, scale, loss_out):
    """vb_bce_logits_fwd as documented in include/volta_b200.h."""
    x = logits.float()
    l = x.clamp(min=0) - x * target + torch.log1p(torch.exp(-x.abs()))
    row_loss[:x.shape[0]] = l.sum(1)
    loss_out[0] = row_loss[:x.shape[0]].sum() * scale
    return loss_out


def bce_logits_bwd(logits, target, dloss, scale, dlogits):
    dlogits.copy_(((torch.sigmoid(logits.float()) - target) * dloss * scale).to(dlogits.dtype))
    return dlogits


def argmax_pick(logits, labels=None, arg=None, pick=None):
    x = logits.float()
    a = x.argmax(1)          # first of equal maxima
    if arg is not None:
        arg.copy_(a)
    if pick is not None:
        pick.copy_((labels if labels is not None else x).gather(1, a[:, None])[:, 0])
    return arg, pick


def loss_finalize(loss_sum, count, out):
    out.copy_(loss_sum / count.clamp(min=1.0))
    return out


def eltwise(op, a, out, b=None, alpha=1.0, p=0.0, seed=0, offset=0):
    E = _lib
    af = a.float()
    bf = b.float() if b is not None else None
    if op == E.ELT_COPY:
        r = af
    elif op == E.ELT_ADD:
        r = af + bf
    elif op == E.ELT_MUL:
        r = af * bf
    elif op == E.ELT_RELU:
        r = torch.relu(af)
    elif op == E.ELT_DRELU:
        r = af * (bf > 0)
    elif op == E.ELT_GELU:
        r = _gelu(af)
    elif op == E.ELT_DGELU:
        r = af * _dgelu(bf)
    elif op == E.ELT_SWISH:
        r = af * torch.sigmoid(af)
    elif op == E.ELT_DSWISH:
        sg = torch.sigmoid(af)
        r = sg * (1.0 + af * (1.0 - sg))
    elif op == E.ELT_SCALE:
        r = af * alpha
    elif op == E.ELT_DROPOUT:
        assert p == 0.0
        r = af
    elif op == E.ELT_AXPY:
        r = out.float() + alpha * af
    else:
        raise ValueError(op)
    out.copy_(r.reshape(out.shape).to(out.dtype))
    return out


def colsum(a, out, row_limit=None):
    a2 = a.reshape(-1, a.shape[-1])
    if row_limit is not None:
        a2 = a2[:int(row_limit[0])]
    out[:a.shape[-1]] += a2.float().sum(0)
    return out


def compact_plan(labels, pad, limits, src_index, pos, labels_c):
    rows = labels.numel()
    keep = (labels >= 0).nonzero()[:, 0]
    n = keep.numel()
    limits[0] = n
    limits[1] = min(rows, (n + pad - 1) // pad * pad)
    src_index[:rows] = -1
    src_index[:n] = keep.to(torch.int32)
    pos[:rows] = -1
    pos[keep] = torch.arange(n, dtype=torch.int32)
    labels_c[:rows] = -1
    labels_c[:n] = labels[keep]


def compact_gather(src, src_index, limits, dst):
    n, npad = int(limits[0]), int(limits[1])
    dst[:n] = src[src_index[:n].long()]
    dst[n:npad] = 0
    return dst


def zero_rows(buf, row_begin, row_end, row_offset=0, max_rows=256, second=None):
    buf[row_offset + int(row_begin[0]):row_offset + int(row_end[0])] = 0
    if second is not None:
        buf[second[2] + int(second[0][0]):second[2] + int(second[1][0])] = 0
    return buf


def compact_scatter(src_c, pos, dst):
    rows = dst.shape[0]
    p = pos[:rows].long()
    dst.zero_()
    live = p >= 0
    dst[live] = src_c[p[live]]
    return dst


def gather_rows(a, idx, out):
    out.copy_(a[idx])
    return out


def scatter_rows_add(a, idx, out):
    out.index_add_(0, idx, a)
    return out


def cast_bf16(src, dst):
    dst.copy_(src.to(torch.bfloat16))
    return dst


def box_sincos(boxes, dim, out):
    cx = (boxes[:, 0] + boxes[:, 2]) / 2 * 100
    cy = (boxes[:, 1] + boxes[:, 3]) / 2 * 100
    w = (boxes[:, 2] - boxes[:, 0]) * 100
    h = (boxes[:, 3] - boxes[:, 1]) * 100
    pos = torch.stack([cx, cy, w, h], -1)
    freq = 1000 ** (torch.arange(dim, dtype=torch.float32, device=boxes.device) / float(dim))
    ang = pos[..., None] / freq
    out.copy_(torch.cat([ang.sin(), ang.cos()], -1).reshape(out.shape).to(out.dtype))
    return out


def vlbert_prep(feat, mask_emb, out, masked, kind, K, mvrc, end_kind):
    m = (feat == 0).all(-1)
    out.copy_(torch.where(m[:, None], mask_emb[None, :].expand_as(feat), feat).to(out.dtype))
    masked.copy_(m.long())
    r = torch.arange(feat.shape[0], device=feat.device)
    k = torch.where(m & bool(mvrc), torch.ones_like(r), torch.zeros_like(r))
    kind.copy_(torch.where(r % K == K - 1, torch.full_like(r, end_kind), k))


def vlbert_index(ids, K, tpos, opos, text_end, last_row):
    B, T = ids.shape
    te = (ids != 0).sum(1)
    text_end.copy_(te)
    t = torch.arange(T, device=ids.device)
    tpos.copy_(torch.where(t >= te.min(), t + K, t).repeat(B))
    o = te[:, None].expand(B, K).clone()
    o[:, -1] += 1
    opos.copy_(o.reshape(-1))
    last_row.copy_(torch.arange(B, device=ids.device) * K + K - 1)


def batched_colsum(a, batch, rows, cols, out):
    out.copy_(a.float().reshape(batch, rows, cols).sum(1).reshape(out.shape))
    return out


def fill(buf, value=0):
    assert value == 0
    buf.zero_()


def sumsq(x, out):
    out += (x.double() ** 2).sum().float()
    return out


def adamw_flat(p, g, m, v, p16, chunk_group, group_hp, beta1, beta2, eps, sumsq_buf=None, max_norm=0.0, rule="adamw"):
    """vb_adamw_flat / vb_radam_flat as documented in include/volta_b200.h: chunk_group[i] = hyper-parameter group of
    elements [64 i, 64 i + 64) (-1 = not optimised), group_hp[g] = (lr, weight_decay, step_size, active) for "adamw",
    (wd * lr, rectified, step_size, active) for "radam"."""
    coef = 1.0
    if sumsq_buf is not None and max_norm > 0:
        coef = min(1.0, max_norm / (float(sumsq_buf.sqrt()) + 1e-6))
    grp = chunk_group.to(torch.int64).repeat_interleave(64)
    live = grp >= 0
    hp = group_hp[grp.clamp(min=0)]
    live = live & (hp[:, 3] != 0)
    gg = g * coef
    m_new = beta1 * m + (1 - beta1) * gg
    v_new = beta2 * v + (1 - beta2) * gg * gg
    if rule == "adamw":
        x = p - hp[:, 2] * m_new / (v_new.sqrt() + eps)
        x = torch.where(hp[:, 1] > 0, x - hp[:, 0] * hp[:, 1] * x, x)
    else:
        x = p - hp[:, 0] * p
        x = torch.where(hp[:, 1] != 0, x - hp[:, 2] * m_new / (v_new.sqrt() + eps), x - hp[:, 2] * m_new)
    m.copy_(torch.where(live, m_new, m))
    v.copy_(torch.where(live, v_new, v))
    p.copy_(torch.where(live, x, p))
    if p16 is not None:
        p16.copy_(torch.where(live, x, p16.float()).to(p16.dtype))


_NAMES = ["gemm", "layernorm_fwd", "layernorm_bwd", "attention_fwd", "attention_bwd", "embed_gather_fwd",
          "embed_gather_bwd", "box_project_fwd", "box_project_bwd", "xent_fwd", "xent_bwd", "kl_fwd", "kl_bwd", "region_loss_fwd", "region_loss_bwd",
          "loss_finalize", "bce_logits_fwd", "bce_logits_bwd", "argmax_pick", "attention_probs", "eltwise", "colsum", "colsum_by_key", "vlbert_prep", "vlbert_index", "batched_colsum", "gather_rows", "scatter_rows_add", "cast_bf16", "box_sincos", "fill",
          "compact_plan", "compact_gather", "compact_scatter", "mask_lengths", "zero_rows", "sumsq", "adamw_flat"]


def install(monkeypatch):
    """Route volta_b200.ops through the emulation for the duration of one test."""
    g = globals()
    for n in _NAMES:
        monkeypatch.setattr(ops, n, g[n])
    monkeypatch.setattr(ops, "require_cuda", lambda *a: None)
    monkeypatch.setattr(engine, "_check_inputs", lambda *a: None)

"""VL-BERT joint embeddings on the engine (reference: `VLBertEmbeddings`, volta/embeddings.py:184-301).

Forward, per region row: sin/cos box embedding || appearance feature (all-zero rows replaced by the learned
mask feature) -> dropout -> Linear(4096 -> H) -> ReLU = `final`; object token = LN_obj(final) + one of three
1-row "linguistic" tables (object / masked-object / end) + position + type 2.  Per text token: word + LN_text(final
of the sample's LAST region) + position + type, where the position ids reproduce the reference's quirk (text
positions move by K for every sample as soon as p >= min over the batch of text_end, embeddings.py:285-287).
One LayerNorm over [text; regions], dropout.  The backward is written by hand below.
"""
import torch

from . import ops
from ._lib import (ELT_AXPY, ELT_COPY, ELT_DRELU, EPI_BIAS, EPI_BIAS_RELU, EPI_NONE)


def _bidx(rt):
    from .engine import _POS_CACHE
    key = ("bidx", rt.device, rt.B, rt.T)
    if key not in _POS_CACHE:
        _POS_CACHE[key] = torch.arange(rt.B, dtype=torch.int64).repeat_interleave(rt.T).to(rt.device)
    return _POS_CACHE[key]


def _ling_tables(e):
    """The three 1-row tables, adjacent in the flat store (see BertPreTrainedModel._fused_groups)."""
    ts = [e.object_linguistic_embeddings.weight]
    if e.with_mvrc_loss:
        ts.append(e.object_mask_word_embedding.weight)
    ts.append(e.end_embedding.weight)
    return ts


def vlbert_embeddings_fwd(rt, bert, ids, feat, loc2, seg, x, s):
    from .engine import Pair, _count, dropout_, linear_fwd, ln_fwd
    e, cfg = bert.embeddings, bert.config
    B, T, K = rt.B, rt.T, rt.V
    F = feat.shape[-1]
    H, Hv = cfg.hidden_size, cfg.v_hidden_size
    dim = e.v_coordinate_embeddings_dim
    assert 8 * dim == F, "VL-BERT: 4 * 2 * v_coordinate_embeddings_dim must equal v_feature_size"
    assert x.all is not None, "VL-BERT needs hidden_size == v_hidden_size for its joint LayerNorm"
    dev, i64 = rt.device, torch.int64
    feat2 = feat.reshape(B * K, F).float().contiguous()
    # --- index preparation (device, no host sync)
    tpos, opos = torch.empty(B * T, dtype=i64, device=dev), torch.empty(B * K, dtype=i64, device=dev)
    text_end, last_row = torch.empty(B, dtype=i64, device=dev), torch.empty(B, dtype=i64, device=dev)
    ops.vlbert_index(ids, K, tpos, opos, text_end, last_row)
    masked, kind = torch.empty(B * K, dtype=i64, device=dev), torch.empty(B * K, dtype=i64, device=dev)
    ling = _ling_tables(e)
    # --- region features
    z = rt.empty(B * K, 2 * F)
    ops.box_sincos(loc2, dim, z[:, :F])
    ops.vlbert_prep(feat2, rt.w32(e.object_mask_visual_embedding.weight).view(-1), z[:, F:], masked, kind, K,
                    e.with_mvrc_loss, len(ling) - 1)
    _count(3)
    s["z_drop"] = (rt.p(e.obj_downsample[0].p), rt.seed())
    dropout_(rt, z, *s["z_drop"])
    final = rt.empty(B * K, Hv)
    linear_fwd(rt, z, e.obj_downsample[1], epilogue=EPI_BIAS_RELU, out=final)
    obj_in = final
    if e.visual_1x1_object is not None:
        obj_in = linear_fwd(rt, final, e.visual_1x1_object)
    u = Pair.alloc(rt, rt.Mt, H, rt.Mv, Hv)
    _, s["obj_stats"] = ln_fwd(rt, obj_in, e.visual_ln_object, y=u.v)
    ops.embed_gather_fwd([rt.w32(*ling)], [kind], [0], u.v, accum=True)
    ops.embed_gather_fwd([rt.w32(e.position_embeddings.weight), rt.w32(e.token_type_embeddings.weight)],
                         [opos, None], [0, 2], u.v, accum=True)
    # --- text tokens
    last = rt.empty(B, Hv)
    ops.gather_rows(final, last_row, last)
    t_in = last
    if e.visual_1x1_text is not None:
        t_in = linear_fwd(rt, last, e.visual_1x1_text)
    tv, s["txt_stats"] = ln_fwd(rt, t_in, e.visual_ln_text)
    tv32 = rt.empty(B, H, dtype=torch.float32)
    ops.eltwise(ELT_COPY, tv, tv32)
    tables = [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight]
    idx = [ids.reshape(-1), tpos, seg.reshape(-1), _bidx(rt)]
    ops.embed_gather_fwd([rt.w32(t) for t in tables] + [tv32], idx, [0, 0, 0, 0], u.t)
    _count(5)
    # --- joint LayerNorm + dropout
    _, s["stats"] = ln_fwd(rt, u.all, e.LayerNorm, y=x.all)
    s["drop"] = (rt.p(e.dropout.p), rt.seed())
    dropout_(rt, x.all, *s["drop"])
    s.update(u=u, z=z, final=final, obj_in=obj_in, last=last, t_in=t_in, tpos=tpos, opos=opos, masked=masked, okind=kind,
             last_row=last_row, ids=ids.reshape(-1), seg=seg.reshape(-1), F=F)


def vlbert_embeddings_bwd(rt, bert, dx, s):
    from .engine import _count, dropout_, linear_bwd, ln_bwd
    e, cfg = bert.embeddings, bert.config
    B, T, K, F = rt.B, rt.T, rt.V, s["F"]
    H = cfg.hidden_size
    dropout_(rt, dx.all, *s["drop"])
    du = ln_bwd(rt, dx.all, s["u"].all, s["stats"], e.LayerNorm)
    dut, duv = du[:rt.Mt], du[rt.Mt:]
    # ---- text side
    ops.embed_gather_bwd([rt.g32(e.word_embeddings.weight)], [s["ids"]], [0], dut, skip=[0])
    gpos = rt.g32(e.position_embeddings.weight)
    ptmp = rt.zeros(T, H, dtype=torch.float32)
    ops.colsum(dut.view(B, T * H), ptmp.view(-1))              # every sample uses the same T position rows
    ops.scatter_rows_add(ptmp, s["tpos"][:T].contiguous(), gpos)
    ops.colsum_by_key(dut, s["seg"], e.token_type_embeddings.weight.shape[0], rt.g32(e.token_type_embeddings.weight))
    dtv32 = rt.empty(B, H, dtype=torch.float32)
    ops.batched_colsum(dut, B, T, H, dtv32)                      # the per-sample visual vector is added to all T tokens
    dtv = rt.empty(B, H)
    ops.eltwise(ELT_COPY, dtv32, dtv)
    _count(6)
    dt_in = ln_bwd(rt, dtv, s["t_in"], s["txt_stats"], e.visual_ln_text)
    dlast = dt_in
    if e.visual_1x1_text is not None:
        dlast = linear_bwd(rt, s["last"], dt_in, e.visual_1x1_text)
    # ---- region side
    ops.embed_gather_bwd([gpos], [s["opos"]], [0], duv)
    ops.colsum(duv, rt.g32(e.token_type_embeddings.weight)[2])
    ling = _ling_tables(e)
    ops.colsum_by_key(duv, s["okind"], len(ling), rt.g32(*ling))
    _count(3)
    dobj_in = ln_bwd(rt, duv, s["obj_in"], s["obj_stats"], e.visual_ln_object)
    dfinal = dobj_in
    if e.visual_1x1_object is not None:
        dfinal = linear_bwd(rt, s["final"], dobj_in, e.visual_1x1_object)
    ops.scatter_rows_add(dlast, s["last_row"], dfinal)           # unique rows: last region of each sample
    dpre = rt.empty(*dfinal.shape)
    ops.eltwise(ELT_DRELU, dfinal, dpre, b=s["final"])
    _count(2)
    dz = linear_bwd(rt, s["z"], dpre, e.obj_downsample[1])       # dz is needed only for the learned mask feature
    dropout_(rt, dz, *s["z_drop"])
    tmp = rt.zeros(2, F, dtype=torch.float32)
    ops.colsum_by_key(dz[:, F:], s["masked"], 2, tmp)
    gm = rt.g32(e.object_mask_visual_embedding.weight)
    ops.eltwise(ELT_AXPY, tmp[1:2], gm.view(1, -1), alpha=1.0)
    _count(2)

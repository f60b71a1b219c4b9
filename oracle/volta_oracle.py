"""CPU oracle for the VOLTA encoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 restatement of the algorithm in the
reference's `volta/encoders.py` / `volta/embeddings.py` / `volta/losses.py`.
It exists so the CUDA path can be checked; it is not shipped and nothing in
the `volta_b200` package imports it.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py` (the `cpu_baseline` leg and `--impl reference`) may use it.

Parity status: PINNED.  The reference has no golden vectors of its own for
this path, so the oracle is pinned against outputs of the
reference itself: `tests/golden/make_golden.py` imports the unmodified
reference from /root/reference in the build container, runs it on seeded
weights + seeded synthetic batches for every model family, and commits the
outputs under `tests/golden/*.pt`; `tests/test_oracle_golden.py` checks this
file against them (fp32, atol 2e-5).

It is written functionally over a flat dict of parameters keyed by the
reference's `state_dict()` names, which is deliberately not how the reference
is organised (nn.Module tree) - what is restated is the arithmetic.
Citations are `file:line` under /root/reference.
"""
import math

import torch
import torch.nn.functional as F

NEG = -10000.0  # additive key mask, volta/encoders.py:986,991


# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-12):
    """TF-style LN, biased variance, eps inside the sqrt (volta/encoders.py:57-61)."""
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return w * ((x - mu) / torch.sqrt(var + eps)) + b


def linear(x, P, name):
    """y = x W^T + b with W stored [out, in] (nn.Linear)."""
    return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))


def gelu(x):
    """Exact erf GELU (volta/encoders.py:136)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


_ACT = {"gelu": gelu, "relu": F.relu, "swish": lambda x: x * torch.sigmoid(x)}  # volta/encoders.py:157


def _drop(x, p, train):
    return F.dropout(x, p, training=True) if (train and p > 0) else x


def _sets(cfg):
    g = lambda k: set(getattr(cfg, k))
    return dict(tt=g("tt_attn_sublayers"), tv=g("tv_attn_sublayers"), vt=g("vt_attn_sublayers"),
                vv=g("vv_attn_sublayers"), fft=g("t_ff_sublayers"), ffv=g("v_ff_sublayers"),
                shared=g("shared_sublayers"), single_ln=g("single_ln_sublayers"))


def depth(cfg):
    s = _sets(cfg)
    return len(s["tt"] | s["tv"] | s["vt"] | s["vv"]) + len(s["fft"] | s["ffv"])


# ----------------------------------------------------------------------------
# embeddings
# ----------------------------------------------------------------------------
def _text_sum(P, pre, ids, seg):
    """word + position + token-type (volta/embeddings.py:55-65)."""
    T = ids.shape[1]
    pos = torch.arange(T, device=ids.device)
    # nn.Embedding(padding_idx=0): the padding row receives no gradient from the lookup (volta/embeddings.py:46)
    return (F.embedding(ids, P[pre + "word_embeddings.weight"], padding_idx=0) + P[pre + "position_embeddings.weight"][pos][None]
            + P[pre + "token_type_embeddings.weight"][seg])


def embed_text(P, cfg, ids, seg, train=False):
    """BertEmbeddings (volta/embeddings.py:39-70)."""
    pre = "bert.embeddings."
    x = layer_norm(_text_sum(P, pre, ids, seg), P[pre + "LayerNorm.weight"], P[pre + "LayerNorm.bias"])
    return _drop(x, cfg.hidden_dropout_prob, train)


def embed_image_vilbert(P, cfg, feat, loc, train=False):
    """ViLBertImageEmbeddings (volta/embeddings.py:127-146)."""
    pre = "bert.v_embeddings."
    s = linear(feat, P, pre + "image_embeddings") + linear(loc, P, pre + "image_location_embeddings")
    return _drop(layer_norm(s, P[pre + "LayerNorm.weight"], P[pre + "LayerNorm.bias"]),
                 cfg.v_hidden_dropout_prob, train)


def embed_image_lxmert(P, cfg, feat, loc, train=False):
    """LxmertImageEmbeddings (volta/embeddings.py:149-172)."""
    pre = "bert.v_embeddings."
    a = layer_norm(linear(feat, P, pre + "image_embeddings"), P[pre + "ImgLayerNorm.weight"], P[pre + "ImgLayerNorm.bias"])
    b = layer_norm(linear(loc, P, pre + "image_location_embeddings"), P[pre + "LocLayerNorm.weight"], P[pre + "LocLayerNorm.bias"])
    return _drop((a + b) / 2, cfg.v_hidden_dropout_prob, train)


def embed_visualbert(P, cfg, ids, feat, loc, seg, train=False):
    """VisualBertEmbeddings (volta/embeddings.py:304-398); image_loc is ignored there."""
    pre = "bert.embeddings."
    xt = _text_sum(P, pre, ids, seg)
    xv = (linear(feat, P, pre + "projection") + P[pre + "position_embeddings_visual.weight"][0]
          + P[pre + "token_type_embeddings_visual.weight"][1])
    x = layer_norm(torch.cat([xt, xv], 1), P[pre + "LayerNorm.weight"], P[pre + "LayerNorm.bias"])
    x = _drop(x, cfg.hidden_dropout_prob, train)
    return x[:, :xt.shape[1]], x[:, xt.shape[1]:]


def embed_uniter(P, cfg, ids, feat, loc, seg, train=False):
    """UniterEmbeddings (volta/embeddings.py:401-457)."""
    pre = "bert.embeddings."
    xt = layer_norm(_text_sum(P, pre, ids, seg), P[pre + "LayerNorm.weight"], P[pre + "LayerNorm.bias"])
    xt = _drop(xt, cfg.hidden_dropout_prob, train)
    a = layer_norm(linear(feat, P, pre + "image_embeddings"), P[pre + "image_layer_norm.weight"], P[pre + "image_layer_norm.bias"])
    b = layer_norm(linear(loc, P, pre + "image_location_embeddings"), P[pre + "image_location_layer_norm.weight"],
                   P[pre + "image_location_layer_norm.bias"])
    xv = a + b + P[pre + "token_type_embeddings.weight"][1]
    xv = layer_norm(xv, P[pre + "v_LayerNorm.weight"], P[pre + "v_LayerNorm.bias"])
    return xt, _drop(xv, cfg.hidden_dropout_prob, train)


def box_sincos(boxes, dim):
    """coordinate_embeddings (volta/embeddings.py:102-124): boxes [B,K,>=4] -> [B,K,4,2*dim]."""
    cx = (boxes[..., 0] + boxes[..., 2]) / 2 * 100
    cy = (boxes[..., 1] + boxes[..., 3]) / 2 * 100
    w = (boxes[..., 2] - boxes[..., 0]) * 100
    h = (boxes[..., 3] - boxes[..., 1]) * 100
    pos = torch.stack([cx, cy, w, h], -1)
    freq = 1000 ** (torch.arange(dim, dtype=boxes.dtype, device=boxes.device) / float(dim))
    ang = pos[..., None] / freq
    return torch.cat([ang.sin(), ang.cos()], -1)


def embed_vlbert(P, cfg, ids, feat, loc, seg, train=False):
    """VLBertEmbeddings (volta/embeddings.py:184-301).  Does not mutate `feat`."""
    pre = "bert.embeddings."
    B, K, D = feat.shape
    T = ids.shape[1]
    masked = (feat == 0).all(-1)                                                   # :243
    feat = torch.where(masked[..., None], P[pre + "object_mask_visual_embedding.weight"][0].expand_as(feat), feat)
    coord = box_sincos(loc, cfg.v_coordinate_embeddings_dim).reshape(B * K, -1)
    z = torch.cat([coord, feat.reshape(B * K, -1)], -1)
    z = _drop(z, cfg.v_attention_probs_dropout_prob, train)                        # obj_downsample[0]
    final = F.relu(linear(z, P, pre + "obj_downsample.1")).view(B, K, -1)
    obj_vis = final
    if (pre + "visual_1x1_object.weight") in P:
        obj_vis = linear(final, P, pre + "visual_1x1_object")
    obj_vis = layer_norm(obj_vis, P[pre + "visual_ln_object.weight"], P[pre + "visual_ln_object.bias"])
    obj_ling = P[pre + "object_linguistic_embeddings.weight"][0].expand(B, K, -1)
    if cfg.visual_target_weights.get("6", 0) > 0:
        obj_ling = torch.where(masked[..., None], P[pre + "object_mask_word_embedding.weight"][0].expand_as(obj_ling), obj_ling)
    last = torch.zeros(K, dtype=torch.bool, device=feat.device)
    last[-1] = True
    obj_ling = torch.where(last[None, :, None], P[pre + "end_embedding.weight"][0].expand_as(obj_ling), obj_ling)  # :265
    obj = obj_ling + obj_vis
    txt_vis = final[:, -1:].expand(B, T, -1)                                       # :271
    if (pre + "visual_1x1_text.weight") in P:
        txt_vis = linear(txt_vis, P, pre + "visual_1x1_text")
    txt_vis = layer_norm(txt_vis, P[pre + "visual_ln_text.weight"], P[pre + "visual_ln_text.bias"])
    txt = F.embedding(ids, P[pre + "word_embeddings.weight"], padding_idx=0) + txt_vis
    text_end = (ids != 0).sum(1, keepdim=True)                                     # :278-279
    # :285-287 - the reference adds K in place to an EXPANDED [1,T]->[B,T] view, so the shift lands in the one
    # shared row: position p moves by K for every sample as soon as p >= text_end of ANY sample (min over batch).
    tpos = torch.arange(T, device=ids.device)
    tpos = torch.where(tpos >= text_end.min(), tpos + K, tpos)[None].expand(B, T)
    opos = text_end.expand(B, K).clone()
    opos[:, -1] += 1                                                               # :288-290
    tok = P[pre + "token_type_embeddings.weight"]
    posw = P[pre + "position_embeddings.weight"]
    xt = txt + posw[tpos] + tok[seg]
    xv = obj + posw[opos] + tok[2]
    x = layer_norm(torch.cat([xt, xv], 1), P[pre + "LayerNorm.weight"], P[pre + "LayerNorm.bias"])
    x = _drop(x, cfg.hidden_dropout_prob, train)
    return x[:, :T], x[:, T:]


def embed(P, cfg, ids, feat, loc, seg, train=False):
    """Embedding dispatch of BertModel.forward (volta/encoders.py:963-967)."""
    kind = cfg.image_embeddings
    if kind == "vilbert":
        return embed_text(P, cfg, ids, seg, train), embed_image_vilbert(P, cfg, feat, loc, train)
    if kind == "lxmert":
        return embed_text(P, cfg, ids, seg, train), embed_image_lxmert(P, cfg, feat, loc, train)
    if kind == "visualbert":
        return embed_visualbert(P, cfg, ids, feat, loc, seg, train)
    if kind == "uniter":
        return embed_uniter(P, cfg, ids, feat, loc, seg, train)
    if kind == "vl-bert":
        return embed_vlbert(P, cfg, ids, feat, loc, seg, train)
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# encoder sub-layers
# ----------------------------------------------------------------------------
def _heads(x, n):
    B, L, H = x.shape
    return x.view(B, L, n, H // n).transpose(1, 2)


def attention_sublayer(P, cfg, n, xt, xv, mt, mv, train=False, taps=None):
    """BertGatedAttention n (volta/encoders.py:163-449).  mt/mv: additive [B,1,1,L]."""
    s = _sets(cfg)
    pre = "bert.encoder.layer.%d." % n
    ht = cfg.sublayer2num_attention_heads.get(str(n), cfg.num_attention_heads)
    hv = cfg.sublayer2v_num_attention_heads.get(str(n), cfg.v_num_attention_heads)
    has = {k: n in s[k] for k in ("tt", "tv", "vt", "vv")}
    text, vision = has["tt"] or has["tv"], has["vv"] or has["vt"]

    if text:
        qt, kt, vt_ = (_heads(linear(xt, P, pre + "attention_self." + w), ht) for w in ("query", "key", "value"))
    if vision:
        qv, kv, vv_ = (_heads(linear(xv, P, pre + "attention_self.v_" + w), hv) for w in ("query", "key", "value"))

    def scores(q, k, m):  # scale after QK^T, then additive key mask (:260-263)
        return q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]) + m

    def attend(q, segs, p):
        # joint softmax over the concatenated key axis, text keys first (:285-314)
        sc = torch.cat([scores(q, k, m) for k, _, m in segs], -1)
        pr = _drop(torch.softmax(sc, -1), p, train)
        out, o = 0, 0
        for k, v, _ in segs:
            out = out + pr[..., o:o + k.shape[2]] @ v
            o += k.shape[2]
        B, h, L, d = out.shape
        return out.transpose(1, 2).reshape(B, L, h * d), pr

    ot, ov = xt, xv
    if taps is not None:     # a stream the sub-layer does not touch: every slot None (encoders.py:343-354)
        taps["attn_t.%d" % n], taps["attn_v.%d" % n] = (dict(intra_attn=None, inter_attn=None, queries=None, keys=None) for _ in range(2))
    if text:
        segs = ([(kt, vt_, mt)] if has["tt"] else []) + ([(kv, vv_, mv)] if has["tv"] else [])
        ctx, pr = attend(qt, segs, cfg.attention_probs_dropout_prob)
        if taps is not None:
            taps["ctx_t.%d" % n] = ctx
            # what BertGatedSelfAttention returns with config.visualization (encoders.py:342-348): intra = own-modality block
            nt = kt.shape[2] if has["tt"] else 0
            taps["attn_t.%d" % n] = {"intra_attn": pr[..., :nt] if has["tt"] else None, "inter_attn": pr[..., nt:] if has["tv"] else None,
                                     "queries": qt, "keys": kt}
        d = _drop(linear(ctx, P, pre + "attention_output.dense"), cfg.hidden_dropout_prob, train)
        ot = layer_norm(d + xt, P[pre + "attention_output.LayerNorm.weight"], P[pre + "attention_output.LayerNorm.bias"])
    if vision:
        segs = ([(kt, vt_, mt)] if has["vt"] else []) + ([(kv, vv_, mv)] if has["vv"] else [])
        ctx, pr = attend(qv, segs, cfg.v_attention_probs_dropout_prob)
        if taps is not None:
            taps["ctx_v.%d" % n] = ctx
            nt = kt.shape[2] if has["vt"] else 0       # vision rows: [vt | vv] (encoders.py:300-306, 349-354)
            taps["attn_v.%d" % n] = {"intra_attn": pr[..., nt:] if has["vv"] else None, "inter_attn": pr[..., :nt] if has["vt"] else None,
                                     "queries": qv, "keys": kv}
        d = _drop(linear(ctx, P, pre + "attention_output.v_dense"), cfg.v_hidden_dropout_prob, train)
        ov = layer_norm(d + xv, P[pre + "attention_output.v_LayerNorm.weight"], P[pre + "attention_output.v_LayerNorm.bias"])
    return ot, ov


def ffn_sublayer(P, cfg, n, xt, xv, train=False):
    """BertGatedFeedForward n (volta/encoders.py:452-590)."""
    s = _sets(cfg)
    pre = "bert.encoder.layer.%d." % n
    ot, ov = xt, xv
    if n in s["fft"]:
        h = _ACT[cfg.hidden_act](linear(xt, P, pre + "intermediate.dense"))
        d = _drop(linear(h, P, pre + "output.dense"), cfg.hidden_dropout_prob, train)
        ot = layer_norm(d + xt, P[pre + "output.LayerNorm.weight"], P[pre + "output.LayerNorm.bias"])
    if n in s["ffv"]:
        h = _ACT[cfg.v_hidden_act](linear(xv, P, pre + "intermediate.v_dense"))
        d = _drop(linear(h, P, pre + "output.v_dense"), cfg.v_hidden_dropout_prob, train)
        ov = layer_norm(d + xv, P[pre + "output.v_LayerNorm.weight"], P[pre + "output.v_LayerNorm.bias"])
    return ot, ov


def additive_mask(mask):
    """(1 - m) * -10000 broadcast over heads and queries (volta/encoders.py:974-991)."""
    return ((1.0 - mask.float()) * NEG)[:, None, None, :]


def bert_model(P, cfg, ids, feat, loc, seg=None, mask_t=None, mask_v=None, train=False, taps=None):
    """BertModel.forward (volta/encoders.py:954-1017).

    Returns (layers_t, layers_v, pooled_t, pooled_v): per-sub-layer hidden
    states of both streams (lists of length depth) and the pooled outputs."""
    if mask_t is None:
        mask_t = torch.ones_like(ids)
    if seg is None:
        seg = torch.zeros_like(ids)
    if mask_v is None:
        mask_v = torch.ones(feat.shape[:2], dtype=ids.dtype, device=ids.device)
    xt, xv = embed(P, cfg, ids, feat, loc, seg, train)
    if taps is not None:
        taps["emb_t"], taps["emb_v"] = xt, xv
    mt, mv = additive_mask(mask_t), additive_mask(mask_v)
    s = _sets(cfg)
    attn = s["tt"] | s["tv"] | s["vt"] | s["vv"]
    lt, lv = [], []
    for n in range(depth(cfg)):
        if n in attn:
            xt, xv = attention_sublayer(P, cfg, n, xt, xv, mt, mv, train, taps)
        else:
            xt, xv = ffn_sublayer(P, cfg, n, xt, xv, train)
        lt.append(xt)
        lv.append(xv)
    pooled_t = pooled_v = None
    fm = cfg.fusion_method
    if fm == "vl-bert_vqa":  # VLBertTextPooler (:610-623)
        end = (ids != 0).sum(1) - 2
        pooled_t = F.relu(linear(xt[torch.arange(xt.shape[0]), end], P, "bert.t_pooler.dense"))
    elif fm != "none":
        pooled_t = F.relu(linear(xt[:, 0], P, "bert.t_pooler.dense"))          # :596-607
    if fm not in ("none", "text", "vl-bert_vqa"):
        pooled_v = F.relu(linear(xv[:, 0], P, "bert.v_pooler.dense"))          # :626-637
    return lt, lv, pooled_t, pooled_v


# ----------------------------------------------------------------------------
# heads and losses
# ----------------------------------------------------------------------------
def fuse_pooled(cfg, pooled_t, pooled_v, p_drop=0.1, train=False):
    """fusion + dropout(0.1) (volta/encoders.py:766-778, 1184-1195)."""
    fm = cfg.fusion_method
    if fm == "sum":
        z = pooled_t + pooled_v
    elif fm == "mul":
        z = pooled_t * pooled_v
    elif fm in ("text", "vl-bert_vqa"):
        z = pooled_t
    elif fm == "none":
        return None
    else:
        raise ValueError("Invalid fusion method: %s" % fm)
    return _drop(z, p_drop, train)


def pretraining_heads(P, cfg, xt, xv, pooled_t, pooled_v, train=False):
    """BertPreTrainingHeads.forward (volta/encoders.py:643-784)."""
    pre = "cls.predictions."
    h = layer_norm(_ACT[cfg.hidden_act](linear(xt, P, pre + "transform.dense")),
                   P[pre + "transform.LayerNorm.weight"], P[pre + "transform.LayerNorm.bias"])
    mlm = F.linear(h, P[pre + "decoder.weight"]) + P[pre + "bias"]
    pooled = fuse_pooled(cfg, pooled_t, pooled_v, 0.1, train)
    itm = linear(pooled, P, "cls.bi_seq_relationship") if cfg.fusion_method not in ("none", "vl-bert_vqa") else None
    pre = "cls.imagePredictions."
    hv = _ACT[cfg.hidden_act](linear(xv, P, pre + "transform.dense"))
    if getattr(cfg, "image_head_ln", True):
        hv = layer_norm(hv, P[pre + "transform.LayerNorm.weight"], P[pre + "transform.LayerNorm.bias"])
    img = {ix: linear(hv, P, pre + "decoder_dict.%s" % ix)
           for ix, w in cfg.visual_target_weights.items() if w > 0}
    return mlm, img, itm, pooled


def kl_1601(pred, weight, label, target):
    """volta/losses.py:16-22."""
    loss = F.kl_div(F.log_softmax(pred, dim=2), target, reduction="none")
    picked = (label == 1)
    return weight * torch.sum(loss * picked.unsqueeze(2).float()) / max(int(picked.sum()), 1)


def xent_masked(pred, weight, label, target, conf=None):
    """xent_1600 / xent_400 / xent_1601 (volta/losses.py:87-133)."""
    loss = F.cross_entropy(pred.reshape(-1, pred.shape[-1]), target.reshape(-1), reduction="none")
    if conf is not None:
        loss = loss * conf.reshape(-1)
    picked = (label == 1)
    return weight * torch.sum(loss * picked.reshape(-1)) / max(int(picked.sum()), 1)


def regress_masked(pred, weight, label, target, kind):
    """mse_2048 / huber_2048 (volta/losses.py:25-33, 109-118)."""
    fn = F.mse_loss if kind == "mse" else F.smooth_l1_loss
    loss = fn(pred, target, reduction="none")
    m = (label == 1).unsqueeze(2)
    return weight * torch.sum(loss * m.float()) / max(int(m.expand_as(loss).sum()), 1)


def pretraining_losses(P, cfg, batch, train=False, taps=None):
    """BertForVLPreTraining.forward (volta/encoders.py:1044-1112): (mlm, img, itm) losses."""
    lt, lv, pt, pv = bert_model(P, cfg, batch["input_ids"], batch["image_feat"], batch["image_loc"],
                                batch["segment_ids"], batch["input_mask"], batch["image_mask"], train, taps)
    mlm, img, itm, pooled = pretraining_heads(P, cfg, lt[-1], lv[-1], pt, pv, train)
    if taps is not None:
        taps.update(seq_t=lt[-1], seq_v=lv[-1], pooled_t=pt, pooled_v=pv, mlm_logits=mlm, itm_logits=itm,
                    layers_t=lt, layers_v=lv)
        for ix, v in img.items():
            taps["img_logits." + ix] = v
    img_loss = 0
    for ix, w in cfg.visual_target_weights.items():
        if w <= 0:
            continue
        pred = img[ix][:, :-1] if cfg.add_global_imgfeat == "last" else img[ix][:, int(cfg.add_global_imgfeat is not None):]
        lab = batch["image_label"]
        if ix == "0":
            img_loss = img_loss + kl_1601(pred, w, lab, batch["image_cls"])
        elif ix == "1":
            img_loss = img_loss + regress_masked(pred, w, lab, batch.get("image_feat_target", batch["image_feat"]), "mse")
        elif ix == "5":
            img_loss = img_loss + regress_masked(pred, w, lab, batch.get("image_feat_target", batch["image_feat"]), "huber")
        elif ix == "3":
            img_loss = img_loss + xent_masked(pred, w, lab, batch["obj_labels"], batch["obj_confs"])
        elif ix == "4":
            img_loss = img_loss + xent_masked(pred, w, lab, batch["attr_labels"], batch["attr_confs"])
        elif ix == "6":
            img_loss = img_loss + xent_masked(pred, w, lab, batch["obj_labels"])
        else:
            raise NotImplementedError("visual target %s (nce_2048 draws random negatives)" % ix)
    mlm_loss = F.cross_entropy(mlm.view(-1, cfg.vocab_size), batch["lm_label_ids"].view(-1), ignore_index=-1)
    itm_loss = (F.cross_entropy(itm.view(-1, 2), batch["is_match"].view(-1), ignore_index=-1)
                if itm is not None else torch.zeros(()))
    return mlm_loss, img_loss, itm_loss


def retrieval_scores(P, cfg, batch, clf="clfs_dict.TASK8"):
    """BertForVLTasks.forward for a `VL-logit` task (volta/encoders.py:1158-1206), eval mode."""
    lt, lv, pt, pv = bert_model(P, cfg, batch["input_ids"], batch["image_feat"], batch["image_loc"],
                                batch["segment_ids"], batch["input_mask"], batch["image_mask"])
    return linear(fuse_pooled(cfg, pt, pv), P, clf)


def task_prediction(P, cfg, batch, task_id, task_cfg):
    """BertForVLTasks.forward for every head type (volta/encoders.py:1128-1206; SimpleClassifier :787-814), eval mode:
    the dropouts (:1185-1190, :1197) are identities."""
    lt, lv, pt, pv = bert_model(P, cfg, batch["input_ids"], batch["image_feat"], batch["image_loc"],
                                batch["segment_ids"], batch["input_mask"], batch["image_mask"])
    pooled = fuse_pooled(cfg, pt, pv)
    ttype = task_cfg[task_id]["type"]
    pre = "clfs_dict.%s" % task_id

    def simple_classifier(x):  # Linear -> GeLU -> BertLayerNorm -> Linear
        h = gelu(linear(x, P, pre + ".logit_fc.0"))
        h = layer_norm(h, P[pre + ".logit_fc.2.weight"], P[pre + ".logit_fc.2.bias"])
        return linear(h, P, pre + ".logit_fc.3")

    if ttype.startswith("V-logit"):  # :1197-1199: per-region logit + additive mask of padded regions
        xv = lv[-1]
        if task_cfg[task_id].get("num_clf_layers", 1) == 2:
            y = linear(gelu(linear(xv, P, pre + ".0")), P, pre + ".3")
        else:
            y = linear(xv, P, pre)
        return y + ((1.0 - batch["image_mask"].float()) * -10000.0).unsqueeze(2)
    if ttype == "VL-binary-classifier":  # NLVR2: two consecutive samples form one example (:1200-1202)
        return simple_classifier(pooled.view(-1, pooled.shape[1] * 2))
    if ttype in ("VL-classifier", "VL-classifier-GQA"):
        return simple_classifier(pooled)
    return linear(pooled, P, pre)  # VL-tri-classifier, VL-logit


# ----------------------------------------------------------------------------
# helpers for tests / baseline
# ----------------------------------------------------------------------------
def params_from_state_dict(state, requires_grad=False, dtype=torch.float32):
    """Flat param dict; entries that alias in `state` stay aliased (shared sub-layers, tied decoder)."""
    by_ptr, P = {}, {}
    for k, v in state.items():
        key = (v.data_ptr(), tuple(v.shape))
        if key not in by_ptr:
            t = v.detach().to(dtype).clone()
            t.requires_grad_(requires_grad)
            by_ptr[key] = t
        P[k] = by_ptr[key]
    return P


def fwd_bwd(P, cfg, batch, train=False):
    """One oracle training step (sum of the three losses, as train_concap.py:291-303)."""
    losses = pretraining_losses(P, cfg, batch, train)
    total = losses[0] + losses[1] + losses[2]
    total.backward()
    return [float(l.detach()) for l in losses]

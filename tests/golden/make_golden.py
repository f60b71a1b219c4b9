"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

The reference (`/root/reference/volta`) is imported as-is; the only shims are
`sys.modules` stubs for boto3/botocore (imported at the top of
`volta/utils.py:20-22`, never called on this path) and a no-op `Tensor.cuda`
(`VisualBertEmbeddings` hard-codes `.cuda()`, `volta/embeddings.py:373,383`).
Weights and inputs are NOT stored: both are regenerated from seeds by
`volta_b200.synthetic` (a checksum of the weights is stored to detect RNG
drift).  Stored are the reference's outputs: per-sub-layer hidden states,
logits, the three losses, and per-parameter gradient norm + projection.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from volta_b200.config import named_config_dict  # noqa: E402
from volta_b200.synthetic import pretraining_batch, seeded_state_dict  # noqa: E402

MINI = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, v_hidden_size=128,
            v_num_attention_heads=2, v_intermediate_size=256, pooler_size=128, v_pooler_size=128,
            vocab_size=1024, max_position_embeddings=128, v_feature_size=64)

# name -> (base config, overrides, B, T, V)
CASES = {
    "mini_ctrl_vilbert": ("ctrl_vilbert_base", MINI, 3, 10, 7),
    "mini_vilbert": ("vilbert_base", dict(MINI, v_hidden_size=256, v_intermediate_size=128,
                                          sublayer2attn_hidden_size={str(n): 256 for n in (12, 16, 20, 24, 28, 32)},
                                          sublayer2num_attention_heads={str(n): 2 for n in (12, 16, 20, 24, 28, 32)}), 2, 9, 12),
    "mini_ctrl_lxmert": ("ctrl_lxmert", MINI, 3, 10, 7),
    "mini_ctrl_visualbert": ("ctrl_visualbert_base", MINI, 3, 10, 7),
    "mini_ctrl_uniter": ("ctrl_uniter_base", MINI, 3, 10, 7),
    "mini_ctrl_vlbert": ("ctrl_vl-bert_base", dict(MINI, v_coordinate_embeddings_dim=8), 3, 10, 7),
    # the non-ctrl configs exercise the other masked-region criteria (volta/losses.py): lxmert = xent_1600 + xent_400 +
    # huber_2048 (no global feature, 4-d boxes, fusion "text"); vl-bert_base = xent_1601 (global feature last, no poolers)
    "mini_lxmert": ("lxmert", dict(MINI, v_feature_size=2048), 3, 10, 7),
    "mini_vlbert_base": ("vl-bert_base", dict(MINI, v_coordinate_embeddings_dim=8), 3, 10, 7),
    "full_ctrl_vilbert": ("ctrl_vilbert_base", {}, 2, 12, 9),
    "full_ctrl_visualbert": ("ctrl_visualbert_base", {}, 2, 12, 9),
    "full_ctrl_lxmert": ("ctrl_lxmert", {}, 2, 12, 9),
    "full_vilbert": ("vilbert_base", {}, 2, 10, 13),
}


def case_config_dict(name):
    base, over, B, T, V = CASES[name]
    d = named_config_dict(base)
    d.update(over)
    for k in ("hidden_dropout_prob", "attention_probs_dropout_prob", "v_hidden_dropout_prob",
              "v_attention_probs_dropout_prob"):
        d[k] = 0.0
    return d, (B, T, V)


def case_batch(name):
    d, (B, T, V) = case_config_dict(name)
    glob = d.get("add_global_imgfeat")
    return pretraining_batch(B, T, V, vocab_size=d["vocab_size"], feat_dim=d["v_feature_size"],
                             num_locs=d["num_locs"], seed=77, global_first=glob == "first", global_last=glob == "last",
                             detector_labels=any(ix in ("3", "4", "6") for ix, w in d["visual_target_weights"].items() if w > 0))


def grad_projection(g):
    """Deterministic signed checksum of a gradient tensor."""
    n = g.numel()
    w = torch.cos(torch.arange(n, dtype=torch.float64) * 0.37 + 0.1)
    return float((g.double().reshape(-1) * w).sum())


def weights_checksum(state):
    return float(sum(v.double().sum() for k, v in sorted(state.items())))


def import_reference():
    for name in ("boto3", "botocore", "botocore.exceptions"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["botocore.exceptions"].ClientError = type("ClientError", (Exception,), {})
    sys.path.insert(0, "/root/reference")
    torch.Tensor.cuda = lambda self, *a, **k: self
    from volta.config import BertConfig
    from volta.encoders import BertForVLPreTraining
    return BertConfig, BertForVLPreTraining


def run_case(name, BertConfig, BertForVLPreTraining):
    d, (B, T, V) = case_config_dict(name)
    cfg = BertConfig.from_dict(d)
    model = BertForVLPreTraining(cfg)
    state = seeded_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(state)
    model.eval()
    b = case_batch(name)
    out = {"config": d, "shape": (B, T, V), "weights_checksum": weights_checksum(state),
           "state_keys": sorted(state.keys())}
    with torch.no_grad():
        lt, lv, pt, pv, _ = model.bert(b["input_ids"], b["image_feat"].clone(), b["image_loc"], b["segment_ids"],
                                       b["input_mask"], b["image_mask"], output_all_encoded_layers=True)
        mlm, img, itm, pooled = model.cls(lt[-1], lv[-1], pt, pv)
    full = name.startswith("full")
    keep = (lambda xs: [xs[i] for i in (0, len(xs) // 2, len(xs) - 1)]) if full else (lambda xs: list(xs))
    out["layers_t"] = [x.clone() for x in keep(lt)]
    out["layers_v"] = [x.clone() for x in keep(lv)]
    out["layer_ids"] = [0, len(lt) // 2, len(lt) - 1] if full else list(range(len(lt)))
    out["pooled_t"], out["pooled_v"] = pt, pv
    out["mlm_logits"] = mlm[:, :, :64].clone() if full else mlm
    ix0 = sorted(img)[0]
    out["img_target"] = ix0
    out["img_logits"] = img[ix0][:, :, :64].clone() if full else img[ix0]
    out["itm_logits"] = itm
    model.zero_grad()
    losses = model(b["input_ids"], b["image_feat"].clone(), b["image_loc"], b["segment_ids"], b["input_mask"],
                   b["image_mask"], b["lm_label_ids"], b["image_label"], b["image_cls"], b.get("obj_labels"), b.get("obj_confs"),
                   b.get("attr_labels"), b.get("attr_confs"), None, b["is_match"])
    out["losses"] = [float(l) for l in losses]
    (losses[0] + losses[1] + losses[2]).sum().backward()
    gn, gp = {}, {}
    for k, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        gn[k] = float(g.double().norm())
        gp[k] = grad_projection(g)
    out["grad_norm"], out["grad_proj"] = gn, gp
    return out


def main():
    BertConfig, BertForVLPreTraining = import_reference()
    only = sys.argv[1:]
    for name in CASES:
        if only and name not in only:
            continue
        torch.manual_seed(0)
        out = run_case(name, BertConfig, BertForVLPreTraining)
        path = os.path.join(HERE, name + ".pt")
        torch.save(out, path)
        print("%-24s losses=%s  %.1f KB" % (name, ["%.6f" % l for l in out["losses"]], os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

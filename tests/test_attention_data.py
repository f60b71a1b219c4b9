"""config.visualization + output_all_attention_masks (volta/encoders.py:342-356, 873-875): the per-sub-layer attention
data of the reference - intra- / inter-modal probability blocks after the joint softmax, queries, keys - against golden
outputs of the UNMODIFIED reference (tests/golden/make_golden_attn.py).  CPU: the oracle's taps and the engine's host
logic with the C ABI emulated; GPU: vb_attention_probs through the module API."""
import os

import pytest
import torch

import emu_ops
import volta_b200
from golden_util import ROOT
from oracle import volta_oracle as O
from volta_b200.synthetic import seeded_state_dict

import make_golden as G  # noqa: E402  (case table; does not import the reference)

CASES = ["mini_ctrl_vilbert", "mini_ctrl_visualbert", "mini_vilbert"]


def _load(name):
    g = torch.load(os.path.join(ROOT, "tests", "golden", "attn_" + name + ".pt"), weights_only=False)
    cfg = volta_b200.BertConfig.from_dict(g["config"])
    g["name"] = name
    return g, cfg, G.case_batch(name)


def _check(got_t, got_v, g, tol, b=None, oracle=False):
    """Rows of padded queries are compared for the oracle only: the reference computes them from the hidden states of
    padding positions, which the engine does not define past the embeddings (as for the layer outputs,
    tests/test_engine_emulated.py); columns of padded keys are compared everywhere (probability 0)."""
    assert len(got_t) == len(got_v) == g["num_attention_sublayers"]
    b = b if b is not None else G.case_batch(g["name"])
    rows = {"t": b["input_mask"].bool().cpu(), "v": b["image_mask"].bool().cpu()}
    for side, got in (("t", got_t), ("v", got_v)):
        for i, ref in enumerate(g[side]):
            if ref is None:
                assert got[i] is None, (side, i)
                continue
            for k, r in ref.items():
                if r is None:
                    assert got[i][k] is None, (side, i, k)
                    continue
                a = got[i][k].float().cpu()
                assert a.shape == r.shape, (side, i, k, a.shape, r.shape)
                if not oracle:
                    keep = rows[side][:, None, :, None].expand_as(r)
                    a, r = a[keep], r[keep]
                err = float((a - r).norm() / r.norm().clamp(min=1e-12))
                assert err < tol, (side, i, k, err)


@pytest.mark.parametrize("name", CASES)
def test_oracle_attention_taps_match_reference(name):
    g, cfg, b = _load(name)
    model = volta_b200.BertForVLPreTraining(cfg)
    P = O.params_from_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    taps = {}
    with torch.no_grad():
        O.bert_model(P, cfg, b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"], taps=taps)
    layers = sorted(int(k.split(".")[1]) for k in taps if k.startswith("attn_t.") or k.startswith("attn_v."))
    layers = sorted(set(layers))
    got_t = [taps.get("attn_t.%d" % n) for n in layers]
    got_v = [taps.get("attn_v.%d" % n) for n in layers]
    _check(got_t, got_v, g, 2e-5, oracle=True)


@pytest.mark.parametrize("name", CASES)
def test_engine_attention_data_match_reference_emulated(name, monkeypatch):
    emu_ops.install(monkeypatch)
    g, cfg, b = _load(name)
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision("fp32").eval()
    with torch.no_grad():
        _, _, _, _, (at, av) = model.bert(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"],
                                          b["image_mask"], output_all_attention_masks=True)
        _check(at, av, g, 2e-4)
        # not requested: empty lists, as in the reference; requested without config.visualization: one None per sub-layer
        assert model.bert(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"])[4] == ([], [])
    cfg2 = volta_b200.BertConfig.from_dict(dict(g["config"], visualization=False))
    m2 = volta_b200.BertForVLPreTraining(cfg2)
    m2.set_precision("fp32").eval()
    with torch.no_grad():
        _, _, _, _, (at, av) = m2.bert(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"],
                                       b["image_mask"], output_all_attention_masks=True)
    assert len(at) == g["num_attention_sublayers"] and all(x is None for x in at) and all(x is None for x in av)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
@pytest.mark.parametrize("name", CASES)
def test_attention_data_match_reference_gpu(name, precision, tol):
    g, cfg, b = _load(name)
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision(precision).eval().cuda()
    d = {k: v.cuda() for k, v in b.items()}
    with torch.no_grad():
        _, _, _, _, (at, av) = model.bert(d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"],
                                          d["image_mask"], output_all_attention_masks=True)
        _check(at, av, g, tol)
        mlm, img, itm, attn, pooled = model(d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"],
                                            d["image_mask"], output_all_attention_masks=True)
        _check(attn[0], attn[1], g, tol)

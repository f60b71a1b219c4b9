"""from_pretrained / save_pretrained (reference: volta/utils.py:341-560): local checkpoints, legacy gamma/beta names,
the HuggingFace-BERT -> sub-layer key remap, bare-BERT <-> model-with-heads prefixes, tied decoder."""
import os
import sys
import types

import pytest
import torch

import volta_b200
from golden_util import load_case
from volta_b200.synthetic import seeded_state_dict


def _mini():
    _, cfg, _ = load_case("mini_ctrl_vilbert")
    return cfg


def test_save_and_reload_round_trip(tmp_path):
    cfg = _mini()
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=3))
    model.save_pretrained(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["config.json", "pytorch_model.bin"]
    again, info = volta_b200.BertForVLPreTraining.from_pretrained(str(tmp_path), config=cfg, output_loading_info=True)
    assert info["missing_keys"] == [] and info["unexpected_keys"] == []
    assert not again.training                                         # eval() by default, as the reference
    a, b = model.state_dict(), again.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert again.cls.predictions.decoder.weight is again.bert.embeddings.word_embeddings.weight   # still tied
    # the weights file alone (train_concap.py passes a file path) and the bare BertModel (prefix stripped) also load
    bare = volta_b200.BertModel.from_pretrained(os.path.join(str(tmp_path), "pytorch_model.bin"), config=cfg)
    assert all(torch.equal(v, a["bert." + k]) for k, v in bare.state_dict().items())
    assert volta_b200.BertModel.from_pretrained(str(tmp_path / "nope"), config=cfg) is None


def test_hf_bert_checkpoint_is_remapped_onto_the_schedule():
    cfg = _mini()
    src = volta_b200.BertForVLPreTraining(cfg)
    state = seeded_state_dict(src.state_dict(), seed=5)
    inv_attn = {v: int(k) for k, v in cfg.bert_layer2attn_sublayer.items()}
    inv_ff = {v: int(k) for k, v in cfg.bert_layer2ff_sublayer.items()}
    hf = {}
    for k, v in state.items():
        if ".v_" in k or "v_embeddings" in k or "v_pooler" in k or k.startswith("cls.") or "t_pooler" in k:
            continue
        if ".layer." in k:
            n = int(k.split(".layer.")[1].split(".")[0])
            if ".attention_" in k:
                if n not in inv_attn:
                    continue
                k = k.replace(".layer.%d.attention_" % n, ".layer.%d.attention." % inv_attn[n])
            else:
                if n not in inv_ff:
                    continue
                k = k.replace(".layer.%d." % n, ".layer.%d." % inv_ff[n])
        k = k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta")  # legacy names
        hf[k] = v
    model, info = volta_b200.BertForVLPreTraining.from_pretrained("unused", config=cfg, state_dict=hf, from_hf=True,
                                                                  output_loading_info=True)
    got = model.state_dict()
    for n_hf in range(12):
        a, f = cfg.bert_layer2attn_sublayer[str(n_hf)], cfg.bert_layer2ff_sublayer[str(n_hf)]
        for key in ("bert.encoder.layer.%d.attention_self.query.weight" % a, "bert.encoder.layer.%d.attention_output.LayerNorm.bias" % a,
                    "bert.encoder.layer.%d.intermediate.dense.weight" % f, "bert.encoder.layer.%d.output.dense.bias" % f):
            assert torch.equal(got[key], state[key]), key
    assert torch.equal(got["bert.embeddings.word_embeddings.weight"], state["bert.embeddings.word_embeddings.weight"])
    assert info["unexpected_keys"] == []
    assert any("v_query" in k for k in info["missing_keys"]) and any(k.startswith("cls.") for k in info["missing_keys"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/volta"), reason="the reference is only mounted in the build container")
def test_same_result_as_the_reference_loader(tmp_path):
    cfg = _mini()
    model = volta_b200.BertForVLPreTraining(cfg)
    state = seeded_state_dict(model.state_dict(), seed=7)
    legacy = {k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta"): v for k, v in state.items()}
    path = str(tmp_path / "pytorch_model.bin")
    torch.save(legacy, path)
    mine = volta_b200.BertForVLPreTraining.from_pretrained(path, config=cfg)
    for name in ("boto3", "botocore", "botocore.exceptions"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["botocore.exceptions"].ClientError = type("ClientError", (Exception,), {})
    sys.path.insert(0, "/root/reference")
    try:
        from volta.config import BertConfig as RefConfig
        from volta.encoders import BertForVLPreTraining as RefModel
        ref = RefModel.from_pretrained(path, config=RefConfig.from_dict(cfg.to_dict()), default_gpu=False)
    finally:
        sys.path.remove("/root/reference")
    a, b = mine.state_dict(), ref.state_dict()
    assert sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a)

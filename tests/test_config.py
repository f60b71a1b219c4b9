"""BertConfig / config/*.json - the drop-in contract of the sub-layer schedules (reference: volta/config.py:11-169 and
the eight JSON files under config/)."""
import json
import os

import pytest

import volta_b200
from volta_b200.config import NAMED_CONFIGS, BertConfig, named_config_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/config"


@pytest.mark.parametrize("name", NAMED_CONFIGS)
def test_shipped_json_round_trips_and_builds(name):
    path = os.path.join(ROOT, "config", name + ".json")
    d = json.load(open(path))
    assert d == named_config_dict(name), "config/%s.json is stale: run scripts/write_configs.py" % name
    cfg = BertConfig.from_json_file(path)
    for k, v in d.items():
        assert getattr(cfg, k) == v, (name, k)
    # the schedule is contiguous and every sub-layer is an attention or a feed-forward sub-layer, as BertEncoder asserts
    attn = set(cfg.tt_attn_sublayers) | set(cfg.tv_attn_sublayers) | set(cfg.vt_attn_sublayers) | set(cfg.vv_attn_sublayers)
    ff = set(cfg.t_ff_sublayers) | set(cfg.v_ff_sublayers)
    assert not (attn & ff) and sorted(attn | ff) == list(range(len(attn | ff)))
    model = volta_b200.BertForVLPreTraining(BertConfig.from_dict(dict(d, hidden_size=128, num_attention_heads=2, intermediate_size=256,
                                                                     v_hidden_size=128, v_num_attention_heads=2, v_intermediate_size=256,
                                                                     pooler_size=128, v_pooler_size=128, vocab_size=512,
                                                                     v_feature_size=64, v_coordinate_embeddings_dim=8 if "vl-bert" in name else None,
                                                                     sublayer2attn_hidden_size={}, sublayer2num_attention_heads={})))
    assert len(model.bert.encoder.layer) == len(attn | ff)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")
@pytest.mark.parametrize("name", NAMED_CONFIGS)
def test_shipped_json_equals_the_reference_file(name):
    assert json.load(open(os.path.join(ROOT, "config", name + ".json"))) == json.load(open(os.path.join(REF, name + ".json")))

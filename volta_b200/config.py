"""BertConfig: the JSON -> attribute-bag contract of the VOLTA encoders.

Mirrors the behaviour of the reference `volta/config.py:11-169` (constructor
defaults, `from_dict`, `from_json_file`, `to_dict`, `to_json_string`): every
JSON key is copied over the constructor defaults, so any of the reference's
`config/*.json` files loads unchanged.  The sub-layer schedule keys
(`tt/tv/vt/vv_attn_sublayers`, `t/v_ff_sublayers`, `shared_sublayers`,
`single_ln_sublayers`, `sublayer2*`) are the drop-in contract that
`volta_b200.plan` compiles into a static launch plan.

`named_config()` rebuilds the reference's shipped schedules programmatically
(they are data, generated here from the published layer recipes, see
SURVEY.md Appendix A) so the GPU box does not need the reference tree.
"""
import copy
import json

_DEFAULTS = dict(
    hidden_size=768,
    num_attention_heads=12,
    intermediate_size=3072,
    pooler_size=768,
    hidden_act="gelu",
    hidden_dropout_prob=0.1,
    attention_probs_dropout_prob=0.1,
    max_position_embeddings=512,
    type_vocab_size=2,
    num_locs=5,
    v_coordinate_embeddings_dim=None,
    add_global_imgfeat=None,
    image_embeddings="vilbert",
    initializer_range=0.02,
    v_feature_size=2048,
    v_hidden_size=768,
    v_num_attention_heads=12,
    v_intermediate_size=3072,
    v_pooler_size=1024,
    v_attention_probs_dropout_prob=0.1,
    v_hidden_act="gelu",
    v_hidden_dropout_prob=0.1,
    v_initializer_range=0.2,
    visual_target_weights={"0": 1},
    fixed_layers=[],
    fusion_method="mul",
    objective=0,
    clf_hidden_size=1536,
    image_head_ln=True,
    model="bert",
    visualization=False,
    tt_attn_sublayers=[],
    tv_attn_sublayers=[],
    vt_attn_sublayers=[],
    vv_attn_sublayers=[],
    t_ff_sublayers=[],
    v_ff_sublayers=[],
    shared_sublayers=[],
    single_ln_sublayers=[],
    sublayer2attn_hidden_size={},
    sublayer2num_attention_heads={},
    sublayer2intermediate_size={},
    sublayer2v_attn_hidden_size={},
    sublayer2v_num_attention_heads={},
    sublayer2v_intermediate_size={},
    bert_layer2attn_sublayer={},
    bert_layer2ff_sublayer={},
)


class BertConfig(object):
    """Configuration of a `BertModel` (reference: volta/config.py:11-153)."""

    def __init__(self, vocab_size_or_config_json_file, **kwargs):
        unknown = set(kwargs) - set(_DEFAULTS)
        if unknown:
            raise TypeError("unexpected BertConfig arguments: %s" % sorted(unknown))
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r", encoding="utf-8") as reader:
                for key, value in json.loads(reader.read()).items():
                    self.__dict__[key] = value
        elif isinstance(vocab_size_or_config_json_file, int):
            self.vocab_size = vocab_size_or_config_json_file
            for key, value in _DEFAULTS.items():
                self.__dict__[key] = copy.deepcopy(kwargs.get(key, value))
        else:
            raise ValueError(
                "First argument must be either a vocabulary size (int)"
                "or the path to a pretrained model config file (str)"
            )

    @classmethod
    def from_dict(cls, json_object):
        config = BertConfig(vocab_size_or_config_json_file=-1)
        for key, value in json_object.items():
            config.__dict__[key] = value
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls.from_dict(json.loads(reader.read()))

    def __repr__(self):
        return str(self.to_json_string())

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"


# --------------------------------------------------------------------------- #
# Shipped schedules, rebuilt from their layer recipes. #
# --------------------------------------------------------------------------- #
def _base(**over):
    d = dict(
        attention_probs_dropout_prob=0.1, hidden_act="gelu", hidden_dropout_prob=0.1,
        hidden_size=768, initializer_range=0.02, intermediate_size=3072,
        max_position_embeddings=512, num_attention_heads=12, pooler_size=1024,
        type_vocab_size=2, vocab_size=30522, num_locs=5, add_global_imgfeat="first",
        image_embeddings="vilbert", v_attention_probs_dropout_prob=0.1, v_hidden_act="gelu",
        v_hidden_dropout_prob=0.1, v_feature_size=2048, visual_target_weights={"0": 1.0},
        v_hidden_size=768, v_initializer_range=0.02, v_pooler_size=1024,
        v_num_attention_heads=12, v_intermediate_size=3072, fusion_method="mul",
        clf_hidden_size=1536, shared_sublayers=[], single_ln_sublayers=[],
        sublayer2attn_hidden_size={}, sublayer2num_attention_heads={},
        sublayer2intermediate_size={}, sublayer2v_attn_hidden_size={},
        sublayer2v_num_attention_heads={}, sublayer2v_intermediate_size={},
    )
    d.update(over)
    return d


def _schedule(blocks):
    """blocks: list of (kind, flags) in execution order.

    kind "attn": flags is a subset string of {"tt","tv","vt","vv"}; kind "ff":
    flags subset of {"t","v"}.  Returns the six schedule lists plus the
    text-BERT layer -> sub-layer maps (every tt sub-layer / following text FFN
    in order is one HuggingFace BERT layer, reference volta/utils.py:475-498).
    """
    out = {k: [] for k in ("tt", "tv", "vt", "vv", "t", "v")}
    for n, (kind, flags) in enumerate(blocks):
        for f in flags:
            out[f].append(n)
    sched = dict(
        tt_attn_sublayers=out["tt"], tv_attn_sublayers=out["tv"],
        vt_attn_sublayers=out["vt"], vv_attn_sublayers=out["vv"],
        t_ff_sublayers=out["t"], v_ff_sublayers=out["v"],
    )
    attn_map, ff_map, li = {}, {}, 0
    pending = None
    for n, (kind, flags) in enumerate(blocks):
        if kind == "attn" and "tt" in flags and li < 12:
            pending = n
        elif kind == "ff" and "t" in flags and pending is not None:
            attn_map[str(li)] = pending
            ff_map[str(li)] = n
            li += 1
            pending = None
    sched["bert_layer2attn_sublayer"] = attn_map
    sched["bert_layer2ff_sublayer"] = ff_map
    return sched


def _vilbert_blocks():
    b = [("attn", ("tt",)), ("ff", ("t",))] * 6
    b += [("attn", ("tv", "vt")), ("ff", ("t", "v")), ("attn", ("tt", "vv")), ("ff", ("t", "v"))] * 6
    return b


def _lxmert_blocks():
    b = [("attn", ("tt", "vv")), ("ff", ("t", "v"))] * 5
    b += [("attn", ("tt",)), ("ff", ("t",))] * 4
    b += [("attn", ("tv", "vt")), ("attn", ("tt", "vv")), ("ff", ("t", "v"))] * 5
    return b


def _single_stream_blocks():
    return [("attn", ("tt", "tv", "vt", "vv")), ("ff", ("t", "v"))] * 12


def named_config_dict(name):
    """The JSON dictionary of one of the reference's shipped model configs."""
    if name == "ctrl_vilbert_base":
        d = _base()
        d.update(_schedule(_vilbert_blocks()))
        # In ViLBERT the text-BERT layers 6..11 are the tt+vv sub-layers and the FFN after them.
        d["bert_layer2attn_sublayer"] = {str(i): n for i, n in enumerate([0, 2, 4, 6, 8, 10, 14, 18, 22, 26, 30, 34])}
        d["bert_layer2ff_sublayer"] = {str(i): n for i, n in enumerate([1, 3, 5, 7, 9, 11, 15, 19, 23, 27, 31, 35])}
    elif name == "vilbert_base":
        d = named_config_dict("ctrl_vilbert_base")
        co = [12, 16, 20, 24, 28, 32]
        d.update(v_hidden_size=1024, v_num_attention_heads=8, v_intermediate_size=1024,
                 sublayer2attn_hidden_size={str(n): 1024 for n in co},
                 sublayer2num_attention_heads={str(n): 8 for n in co})
    elif name in ("ctrl_lxmert", "lxmert"):
        d = _base(image_embeddings="lxmert")
        blocks = _lxmert_blocks()
        d.update(_schedule(blocks))
        d["shared_sublayers"] = [n for n, (k, f) in enumerate(blocks) if k == "attn" and "tv" in f]
        if name == "lxmert":
            d.update(pooler_size=768, num_locs=4, fusion_method="text",
                     visual_target_weights={"3": 6.667, "4": 6.667, "5": 6.667})
            for k in ("add_global_imgfeat", "v_pooler_size"):
                d.pop(k)
    elif name in ("ctrl_visualbert_base", "ctrl_uniter_base", "ctrl_vl-bert_base", "vl-bert_base"):
        kind = {"ctrl_visualbert_base": "visualbert", "ctrl_uniter_base": "uniter"}.get(name, "vl-bert")
        d = _base(image_embeddings=kind)
        d.update(_schedule(_single_stream_blocks()))
        d["shared_sublayers"] = list(range(24))
        d["single_ln_sublayers"] = list(range(24))
        if kind == "vl-bert":
            d.update(type_vocab_size=3, v_coordinate_embeddings_dim=256, image_head_ln=False)
        if name == "vl-bert_base":
            d.update(num_locs=4, add_global_imgfeat="last", visual_target_weights={"6": 1.0},
                     fusion_method="none", clf_hidden_size=768)
            for k in ("pooler_size", "v_pooler_size"):
                d.pop(k)
    else:
        raise ValueError("unknown model config name: %s" % name)
    return d


NAMED_CONFIGS = ("ctrl_vilbert_base", "vilbert_base", "ctrl_lxmert", "lxmert", "ctrl_visualbert_base",
                 "ctrl_uniter_base", "ctrl_vl-bert_base", "vl-bert_base")


def named_config(name, **overrides):
    d = named_config_dict(name)
    d.update(overrides)
    return BertConfig.from_dict(d)

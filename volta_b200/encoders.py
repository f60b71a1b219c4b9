"""VOLTA encoder module API on top of the B200 engine.

Same public classes, constructor arguments, forward signatures, return tuples and `state_dict()` keys as
the reference's `volta/encoders.py` (BertGated* sub-layers :163-590, poolers :596-637, heads :643-814,
BertEncoder :820-888, BertModel :918-1017, BertForVLPreTraining :1020-1114, BertForVLTasks :1117-1206), so
a caller written against the reference (train_concap.py:187-197,286-289; eval_retrieval.py:138,191) can
switch imports.  The modules here only OWN parameters and describe the sub-layer schedule; all arithmetic
is done by `volta_b200.engine`, which compiles the schedule into a static plan of fused sm_100a kernels
(libvolta_b200.so).  There is no PyTorch fallback: calling a model on a CPU tensor raises.
"""
import logging
import math
import weakref

import torch
from torch import nn

from .config import BertConfig
from .embeddings import (BertEmbeddings, BertLayerNorm, RobertaEmbeddings, _Declared, _D, _L, _N,
                         dual_embeddings, shared_embeddings)
from .losses import pre_vis_targets

logger = logging.getLogger(__name__)


def _get(table, n, default):
    return table.get(str(n), default)


# ----------------------------------------------------------------------------------------------------
# gated sub-layers (parameter containers + schedule flags)
# ----------------------------------------------------------------------------------------------------
class BertGatedSelfAttention(_Declared):
    def __init__(self, config, layer_num):
        super().__init__()
        c, n = config, layer_num
        hidden = _get(c.sublayer2attn_hidden_size, n, c.hidden_size)
        heads = _get(c.sublayer2num_attention_heads, n, c.num_attention_heads)
        v_hidden = _get(c.sublayer2v_attn_hidden_size, n, c.v_hidden_size)
        v_heads = _get(c.sublayer2v_num_attention_heads, n, c.v_num_attention_heads)
        if hidden % heads != 0:
            raise ValueError("The text hidden size (%d) is not a multiple of the number of attention heads (%d)" % (hidden, heads))
        if v_hidden % v_heads != 0:
            raise ValueError("The vision hidden size (%d) is not a multiple of the number of attention heads (%d)" % (v_hidden, heads))
        self.num_attention_heads, self.attention_head_size, self.all_head_size = heads, hidden // heads, hidden
        self.v_num_attention_heads, self.v_attention_head_size, self.v_all_head_size = v_heads, v_hidden // v_heads, v_hidden
        self.visualization = c.visualization
        self.has_tt, self.has_tv = n in c.tt_attn_sublayers, n in c.tv_attn_sublayers
        self.has_vt, self.has_vv = n in c.vt_attn_sublayers, n in c.vv_attn_sublayers
        self.has_text, self.has_vision = self.has_tt or self.has_tv, self.has_vv or self.has_vt
        self.share_layer = n in c.shared_sublayers
        if self.has_tv or self.has_vt:
            assert hidden == v_hidden, "hidden_size != v_hidden_size"
            assert heads == v_heads, "num_attention_heads != v_num_attention_heads"
        if self.has_text:
            self._declare([(w, _L(c.hidden_size, hidden)) for w in ("query", "key", "value")]
                          + [("dropout", _D(c.attention_probs_dropout_prob))])
        if self.has_text and self.has_vision and self.share_layer:
            assert hidden == v_hidden, "hidden_size != v_hidden_size"
            self.v_query, self.v_key, self.v_value, self.v_dropout = self.query, self.key, self.value, self.dropout
        elif self.has_vision:
            self._declare([("v_" + w, _L(c.v_hidden_size, v_hidden)) for w in ("query", "key", "value")]
                          + [("v_dropout", _D(c.v_attention_probs_dropout_prob))])

    @property
    def pair_mask(self):
        return int(self.has_tt) | int(self.has_tv) << 1 | int(self.has_vt) << 2 | int(self.has_vv) << 3


class _GatedOut(_Declared):
    """dense -> dropout -> +residual -> LayerNorm for the streams a sub-layer touches."""

    def _build(self, c, n, has_t, has_v, in_t, in_v, ln_v):
        self.has_language, self.has_vision = has_t, has_v
        self.share_layer = n in c.shared_sublayers
        self.single_ln = n in c.single_ln_sublayers
        if self.single_ln:
            assert has_t and has_v and self.share_layer, "Missing language, vision or sharing"
        if has_t:
            self._declare([("dense", _L(in_t, c.hidden_size)), ("dropout", _D(c.hidden_dropout_prob)),
                           ("LayerNorm", _N(c.hidden_size))])
        if has_t and has_v and self.share_layer:
            assert in_t == in_v and c.hidden_size == c.v_hidden_size, "hidden_size != v_hidden_size"
            self.v_dense, self.v_dropout, self.v_LayerNorm = self.dense, self.dropout, self.LayerNorm
        elif has_v:
            self._declare([("v_dense", _L(in_v, c.v_hidden_size)), ("v_dropout", _D(c.v_hidden_dropout_prob)),
                           ("v_LayerNorm", _N(ln_v))])


class BertGatedSelfOutput(_GatedOut):
    def __init__(self, config, layer_num):
        super().__init__()
        c, n = config, layer_num
        hidden = _get(c.sublayer2attn_hidden_size, n, c.hidden_size)
        v_hidden = _get(c.sublayer2v_attn_hidden_size, n, c.v_hidden_size)
        has_t = (n in c.tt_attn_sublayers) or (n in c.tv_attn_sublayers)
        has_v = (n in c.vv_attn_sublayers) or (n in c.vt_attn_sublayers)
        # the reference sizes v_LayerNorm by the sub-layer's v_hidden (volta/encoders.py:392)
        self._build(c, n, has_t, has_v, hidden, v_hidden, v_hidden)


class BertGatedAttention(nn.Module):
    def __init__(self, config, layer_num):
        super().__init__()
        self.attention_self = BertGatedSelfAttention(config, layer_num)
        self.attention_output = BertGatedSelfOutput(config, layer_num)

    def forward(self, t_input_tensor, v_input_tensor, t_attention_mask, v_attention_mask):
        """Inputs as the reference (additive masks [bs,1,1,L]); returns (t_out, v_out, None, None)."""
        from . import engine
        t, v = engine.standalone_sublayer(self, "attn", t_input_tensor, v_input_tensor, t_attention_mask, v_attention_mask)
        return t, v, None, None


class BertGatedIntermediate(_Declared):
    def __init__(self, config, layer_num):
        super().__init__()
        c, n = config, layer_num
        self.has_language, self.has_vision = n in c.t_ff_sublayers, n in c.v_ff_sublayers
        self.share_layer = n in c.shared_sublayers
        inter = _get(c.sublayer2intermediate_size, n, c.intermediate_size)
        v_inter = _get(c.sublayer2v_intermediate_size, n, c.v_intermediate_size)
        self.hidden_act, self.v_hidden_act = c.hidden_act, c.v_hidden_act
        if self.has_language:
            self._declare([("dense", _L(c.hidden_size, inter))])
        if self.has_language and self.has_vision and self.share_layer:
            assert c.hidden_size == c.v_hidden_size, "hidden_size != v_hidden_size"
            assert inter == v_inter, "intermediate_size != v_intermediate_size"
            self.v_dense = self.dense
        elif self.has_vision:
            self._declare([("v_dense", _L(c.v_hidden_size, v_inter))])


class BertGatedOutput(_GatedOut):
    def __init__(self, config, layer_num):
        super().__init__()
        c, n = config, layer_num
        inter = _get(c.sublayer2intermediate_size, n, c.intermediate_size)
        v_inter = _get(c.sublayer2v_intermediate_size, n, c.v_intermediate_size)
        has_t, has_v = n in c.t_ff_sublayers, n in c.v_ff_sublayers
        if has_t and has_v and n in c.shared_sublayers:
            assert inter == v_inter, "intermediate_size != v_intermediate_size"
        self._build(c, n, has_t, has_v, inter, v_inter, c.v_hidden_size)


class BertGatedFeedForward(nn.Module):
    def __init__(self, config, layer_num):
        super().__init__()
        self.intermediate = BertGatedIntermediate(config, layer_num)
        self.output = BertGatedOutput(config, layer_num)

    def forward(self, t_input_tensor, v_input_tensor):
        from . import engine
        return engine.standalone_sublayer(self, "ff", t_input_tensor, v_input_tensor, None, None)


# ----------------------------------------------------------------------------------------------------
# poolers and heads
# ----------------------------------------------------------------------------------------------------
class BertTextPooler(_Declared):
    def __init__(self, config):
        super().__init__()
        self._declare([("dense", _L(config.hidden_size, config.pooler_size)), ("activation", nn.ReLU)])


class VLBertTextPooler(BertTextPooler):
    """Pools the token at position text_end - 2 instead of [CLS] (reference: volta/encoders.py:610-623)."""


class BertImagePooler(_Declared):
    def __init__(self, config):
        super().__init__()
        self._declare([("dense", _L(config.v_hidden_size, config.v_pooler_size)), ("activation", nn.ReLU)])


class BertPredictionHeadTransform(_Declared):
    def __init__(self, config):
        super().__init__()
        self.hidden_act = config.hidden_act
        self._declare([("dense", _L(config.hidden_size, config.hidden_size)), ("LayerNorm", _N(config.hidden_size))])


class BertImgPredictionHeadTransform(_Declared):
    def __init__(self, config):
        super().__init__()
        self.hidden_act = config.hidden_act  # the reference uses the TEXT activation here (volta/encoders.py:664-665)
        self._declare([("dense", _L(config.v_hidden_size, config.v_hidden_size))])
        self.has_ln = bool(config.image_head_ln)
        if self.has_ln:
            self._declare([("LayerNorm", _N(config.v_hidden_size))])


class BertLMPredictionHead(nn.Module):
    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        vocab, hidden = bert_model_embedding_weights.shape
        self.decoder = nn.Linear(hidden, vocab, bias=False)
        self.decoder.weight = bert_model_embedding_weights  # tied to the input embedding
        self.bias = nn.Parameter(torch.zeros(vocab))


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.predictions = BertLMPredictionHead(config, bert_model_embedding_weights)


class BertOnlyNSPHead(_Declared):
    def __init__(self, config):
        super().__init__()
        self._declare([("seq_relationship", _L(config.hidden_size, 2))])


class BertImagePredictionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.transform = BertImgPredictionHeadTransform(config)
        self.decoder_dict = nn.ModuleDict({
            ix: nn.Linear(config.v_hidden_size, num) for ix, num in pre_vis_targets.items()
            if config.visual_target_weights.get(ix, 0) > 0})


def _xavier_init(module):
    """Head initialisation of the reference (volta/encoders.py:753-764, 800-811)."""
    if isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=0.02)
    elif isinstance(module, nn.Linear):
        nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, BertLayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)


class BertPreTrainingHeads(nn.Module):
    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.predictions = BertLMPredictionHead(config, bert_model_embedding_weights)
        self.has_itm = config.fusion_method not in {"none", "vl-bert_vqa"}
        if self.has_itm:
            self.bi_seq_relationship = nn.Linear(config.pooler_size, 2)
        self.imagePredictions = BertImagePredictionHead(config)
        self.fusion_method = config.fusion_method
        self.dropout = nn.Dropout(0.1)
        self.apply(_xavier_init)

    init_weights = staticmethod(_xavier_init)


class GeLU(nn.Module):
    def forward(self, x):
        from .engine import GeluFn
        return GeluFn.apply(x)


class SimpleClassifier(nn.Module):
    def __init__(self, in_dim, hid_dim, out_dim, dropout_prob=0.0):
        super().__init__()
        self.logit_fc = nn.Sequential(nn.Linear(in_dim, hid_dim), GeLU(), BertLayerNorm(hid_dim, eps=1e-12),
                                      nn.Linear(hid_dim, out_dim))
        self.apply(_xavier_init)

    init_weights = staticmethod(_xavier_init)


# ----------------------------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------------------------
class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        c = config
        attn = set(c.tt_attn_sublayers + c.tv_attn_sublayers + c.vt_attn_sublayers + c.vv_attn_sublayers)
        ff = set(c.t_ff_sublayers + c.v_ff_sublayers)
        depth = len(attn) + len(ff)
        self.num2type = {n: "attn" for n in attn}
        self.num2type.update({n: "ff" for n in ff})
        assert len(self.num2type) == depth, "Overlapping attn-ff sublayer numbers"
        assert min(self.num2type) == 0 and max(self.num2type) == depth - 1, "Non contiguous sublayer numbers"
        self.layer = nn.ModuleList([BertGatedAttention(c, n) if self.num2type[n] == "attn" else BertGatedFeedForward(c, n)
                                    for n in range(depth)])


class BertPreTrainedModel(nn.Module):
    """Weight initialisation + `from_pretrained` (reference: volta/utils.py:361-560, volta/encoders.py:891-915)."""
    config_class = BertConfig
    base_model_prefix = "bert"

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        if not isinstance(config, BertConfig):
            raise ValueError("Parameter config in `%s(config)` should be an instance of class `BertConfig`." % self.__class__.__name__)
        self.config = config
        self._engine_state = None

    def init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, BertLayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def _tie_or_clone_weights(self, first_module, second_module):
        first_module.weight = second_module.weight

    def tie_weights(self):
        pass

    # ---- engine hooks ----
    @property
    def precision(self):
        return getattr(self, "_precision", "bf16")

    def set_precision(self, precision):
        """'bf16' (tensor-core path, default) or 'fp32' (validation mode: every kernel in fp32, SIMT GEMMs)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        for m in self.modules():
            if isinstance(m, BertPreTrainedModel):
                m._precision = precision
        return self

    def _fused_groups(self):
        """Parameters that must be adjacent in the flat store (Q/K/V of one attention module)."""
        groups, seen = [], set()
        for m in self.modules():
            if isinstance(m, BertGatedSelfAttention):
                for pre in ("", "v_"):
                    q = getattr(m, pre + "query", None)
                    if q is None or id(q) in seen:
                        continue
                    seen.add(id(q))
                    k, v = getattr(m, pre + "key"), getattr(m, pre + "value")
                    groups.append([q.weight, k.weight, v.weight])
                    groups.append([q.bias, k.bias, v.bias])
            elif type(m).__name__ == "VLBertEmbeddings":
                from .vlbert import _ling_tables
                groups.append(_ling_tables(m))  # object / masked-object / end rows form one 2-3 row table
        return groups

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        from .pretrained import from_pretrained
        return from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs)

    def save_pretrained(self, save_directory):
        from .pretrained import save_pretrained
        return save_pretrained(self, save_directory)


class BertModel(BertPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.shared_embeddings = False
        model_kind = getattr(config, "model", "bert")
        if model_kind == "bert":
            self.embeddings = BertEmbeddings(config)
        elif model_kind == "roberta":
            self.embeddings = RobertaEmbeddings(config)
        if config.image_embeddings in dual_embeddings:
            self.v_embeddings = dual_embeddings[config.image_embeddings](config)
        self.encoder = BertEncoder(config)
        self.fusion_method = config.fusion_method
        if config.fusion_method == "vl-bert_vqa":
            self.t_pooler = VLBertTextPooler(config)
        elif config.fusion_method != "none":
            self.t_pooler = BertTextPooler(config)
        if config.fusion_method not in {"none", "text", "vl-bert_vqa"}:
            assert config.pooler_size == config.v_pooler_size, "pooler_size != v_pooler_size"
            self.v_pooler = BertImagePooler(config)
        self.apply(self.init_weights)
        if config.image_embeddings in shared_embeddings:
            # created after init_weights, exactly as the reference does (volta/encoders.py:949-952)
            self.embeddings = shared_embeddings[config.image_embeddings](config)
            self.shared_embeddings = True

    def forward(self, input_txt, input_imgs, image_loc, token_type_ids=None, attention_mask=None,
                image_attention_mask=None, output_all_encoded_layers=False, output_all_attention_masks=False):
        from . import engine
        return engine.bert_model_forward(self, input_txt, input_imgs, image_loc, token_type_ids, attention_mask,
                                         image_attention_mask, output_all_encoded_layers, output_all_attention_masks)


class BertForVLPreTraining(BertPreTrainedModel):
    """BERT model with multimodal pre-training heads (MLM + MRM + ITM)."""

    def __init__(self, config):
        super().__init__(config)
        self.bert = BertModel(config)
        self.bert.__dict__["_engine_parent"] = weakref.ref(self)  # one flat parameter store for the whole model
        self.cls = BertPreTrainingHeads(config, self.bert.embeddings.word_embeddings.weight)
        self.loss_fct = nn.CrossEntropyLoss(ignore_index=-1)
        self.visual_target_weights = config.visual_target_weights
        print("model's visual targets are ", [ix for ix, w in config.visual_target_weights.items() if w > 0])
        self.add_global_imgfeat = int(config.add_global_imgfeat is not None)
        self.tie_weights()

    def tie_weights(self):
        self._tie_or_clone_weights(self.cls.predictions.decoder, self.bert.embeddings.word_embeddings)

    def forward(self, input_ids, image_feat, image_loc, token_type_ids=None, attention_mask=None,
                image_attention_mask=None, masked_lm_labels=None, image_label=None, image_cls=None, obj_labels=None,
                obj_confs=None, attr_labels=None, attr_confs=None, image_attrs=None, next_sentence_label=None,
                output_all_attention_masks=False):
        from . import engine
        return engine.pretraining_forward(self, input_ids, image_feat, image_loc, token_type_ids, attention_mask,
                                          image_attention_mask, masked_lm_labels, image_label, image_cls, obj_labels,
                                          obj_confs, attr_labels, attr_confs, image_attrs, next_sentence_label,
                                          output_all_attention_masks)


class BertForVLTasks(BertPreTrainedModel):
    def __init__(self, config, task_cfg, task_ids, dropout_prob=0.1):
        super().__init__(config)
        self.bert = BertModel(config)
        # one flat parameter store for encoder + task heads: the data-parallel hooks and FusedAdamW are keyed by the store's
        # model (this object); head gradients (written by autograd) enter the flat gradient buffer in prepare_grads
        self.bert.__dict__["_engine_parent"] = weakref.ref(self)
        self.dropout = nn.Dropout(dropout_prob)
        self.config, self.task_cfg = config, task_cfg
        task2clf = {}
        for task_id in task_ids:
            task_type = task_cfg[task_id]["type"]
            if task_type in {"VL-classifier", "VL-classifier-GQA"}:
                task2clf[task_id] = SimpleClassifier(config.pooler_size, config.clf_hidden_size, task_cfg[task_id]["num_labels"])
            elif task_type == "VL-binary-classifier":
                task2clf[task_id] = SimpleClassifier(config.pooler_size * 2, config.clf_hidden_size, 2)
            elif task_type == "VL-tri-classifier":
                task2clf[task_id] = nn.Linear(config.pooler_size, 3)
            elif task_type == "VL-logit":
                task2clf[task_id] = nn.Linear(config.pooler_size, 1)
            elif task_type.startswith("V-logit"):
                if task_cfg[task_id].get("num_clf_layers", 1) == 2:
                    task2clf[task_id] = nn.Sequential(nn.Linear(config.v_hidden_size, config.v_hidden_size), GeLU(),
                                                      nn.Dropout(config.v_attention_probs_dropout_prob, inplace=False),
                                                      nn.Linear(config.v_hidden_size, 1))
                else:
                    task2clf[task_id] = nn.Linear(config.v_hidden_size, 1)
            else:
                raise ValueError("Undefined task type: %s" % task_type)
        self.clfs_dict = nn.ModuleDict(task2clf)
        self.fusion_method = config.fusion_method
        self.apply(self.init_weights)

    def _engine_written(self):
        """Parameters whose gradients the engine's hand-written backward writes (the rest - the task heads - get theirs
        from autograd and stay `None` when a step does not use them, as in the reference)."""
        return list(self.bert.parameters())

    def forward(self, input_txt, input_imgs, image_loc, task_id, token_type_ids=None, attention_mask=None,
                image_attention_mask=None, output_all_encoded_layers=False, output_all_attention_masks=False):
        from . import engine
        return engine.tasks_forward(self, input_txt, input_imgs, image_loc, task_id, token_type_ids, attention_mask,
                                    image_attention_mask, output_all_encoded_layers, output_all_attention_masks)

"""Embedding modules of the VOLTA encoders: parameter containers carrying the reference's names.

The class / attribute names and parameter shapes are the drop-in contract (`state_dict()` keys of
`volta/embeddings.py` in the reference: text :39-70, ViLBERT :127-146, LXMERT :149-172, VL-BERT :184-238,
VisualBERT :304-334, UNITER :401-431).  No arithmetic lives here: `volta_b200.engine` runs the fused
CUDA kernels (gather-sum, region-projection GEMM, box FMAs, LayerNorm, dropout) over these parameters.
Each class is declared as a table {attribute: constructor}; `_Declared` instantiates it in order, which
keeps registration order (and therefore `state_dict()` / `parameters()` order) equal to the reference's.
"""
import torch
from torch import nn


class BertLayerNorm(nn.Module):
    """LayerNorm parameters, TF style: biased variance, eps inside the sqrt (reference: volta/encoders.py:48-61;
    apex FusedLayerNorm when installed).  Calling it runs vb_layernorm_fwd/bwd on a CUDA tensor."""

    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        from .engine import LayerNormFn
        return LayerNormFn.apply(x, self.weight, self.bias, self.variance_epsilon)


class _Declared(nn.Module):
    def _declare(self, table):
        for name, make in table:
            setattr(self, name, make())


def _E(n, h, pad=None):
    return lambda: nn.Embedding(n, h, padding_idx=pad)


def _L(i, o):
    return lambda: nn.Linear(i, o)


def _N(h):
    return lambda: BertLayerNorm(h, eps=1e-12)


def _D(p):
    return lambda: nn.Dropout(p)


def _text_tables(c):
    return [("word_embeddings", _E(c.vocab_size, c.hidden_size, 0)),
            ("position_embeddings", _E(c.max_position_embeddings, c.hidden_size)),
            ("token_type_embeddings", _E(c.type_vocab_size, c.hidden_size))]


class BertEmbeddings(_Declared):
    def __init__(self, config):
        super().__init__()
        self._declare(_text_tables(config) + [("LayerNorm", _N(config.hidden_size)), ("dropout", _D(config.hidden_dropout_prob))])


class RobertaEmbeddings(BertEmbeddings):
    padding_idx = 1  # position ids start at padding_idx + 1


class ViLBertImageEmbeddings(_Declared):
    def __init__(self, c):
        super().__init__()
        self._declare([("image_embeddings", _L(c.v_feature_size, c.v_hidden_size)),
                       ("image_location_embeddings", _L(c.num_locs, c.v_hidden_size)),
                       ("LayerNorm", _N(c.v_hidden_size)), ("dropout", _D(c.v_hidden_dropout_prob))])


class LxmertImageEmbeddings(_Declared):
    def __init__(self, c):
        super().__init__()
        self._declare([("image_embeddings", _L(c.v_feature_size, c.v_hidden_size)),
                       ("image_location_embeddings", _L(c.num_locs, c.v_hidden_size)),
                       ("ImgLayerNorm", _N(c.v_hidden_size)), ("LocLayerNorm", _N(c.v_hidden_size)),
                       ("dropout", _D(c.v_hidden_dropout_prob))])


dual_embeddings = {"vilbert": ViLBertImageEmbeddings, "lxmert": LxmertImageEmbeddings}


class VLBertEmbeddings(_Declared):
    def __init__(self, c):
        super().__init__()
        H = c.hidden_size
        self.hidden_size, self.initializer_range = H, c.initializer_range
        self.with_mvrc_loss = c.visual_target_weights.get("6", 0) > 0
        self.v_coordinate_embeddings_dim = c.v_coordinate_embeddings_dim
        t = [("obj_downsample", lambda: nn.Sequential(nn.Dropout(c.v_attention_probs_dropout_prob),
                                                      nn.Linear(2 * c.v_feature_size, c.v_hidden_size), nn.ReLU(inplace=True))),
             ("object_linguistic_embeddings", _E(1, H))]
        if self.with_mvrc_loss:
            t.append(("object_mask_word_embedding", _E(1, H)))
        t += [("object_mask_visual_embedding", _E(1, c.v_feature_size)), ("end_embedding", _E(1, H))] + _text_tables(c)
        self.visual_1x1_text = self.visual_1x1_object = None
        if c.v_hidden_size != H:
            t += [("visual_1x1_text", _L(c.v_hidden_size, H)), ("visual_1x1_object", _L(c.v_hidden_size, H))]
        t += [("visual_ln_text", _N(H)), ("visual_ln_object", _N(H)), ("LayerNorm", _N(H)), ("dropout", _D(c.hidden_dropout_prob))]
        self._declare(t)
        self.init_weights()

    def init_weights(self):
        nn.init.xavier_uniform_(self.obj_downsample[1].weight)
        self.object_mask_visual_embedding.weight.data.zero_()
        for emb in [self.object_linguistic_embeddings] + ([self.object_mask_word_embedding] if self.with_mvrc_loss else []):
            emb.weight.data.normal_(mean=0.0, std=self.initializer_range)
        for ln in (self.visual_ln_text, self.visual_ln_object):
            ln.weight.data.zero_()
        self.LayerNorm.weight.data.fill_(1.0)
        self.LayerNorm.bias.data.zero_()


class VisualBertEmbeddings(_Declared):
    def __init__(self, c):
        super().__init__()
        H = c.hidden_size
        self.hidden_size, self.initializer_range = H, c.initializer_range
        self._declare(_text_tables(c) + [
            ("LayerNorm", _N(H)), ("dropout", _D(c.hidden_dropout_prob)), ("projection", _L(c.v_feature_size, H)),
            ("token_type_embeddings_visual", _E(c.type_vocab_size, H, 0)),
            ("position_embeddings_visual", _E(c.max_position_embeddings, H))])
        self.special_initialize()

    def special_initialize(self):
        """Visual type / position tables start as copies of the text ones."""
        for dst, src in ((self.token_type_embeddings_visual, self.token_type_embeddings),
                         (self.position_embeddings_visual, self.position_embeddings)):
            dst.weight = nn.Parameter(src.weight.data.clone(), requires_grad=True)


class UniterEmbeddings(_Declared):
    def __init__(self, c):
        super().__init__()
        H = c.hidden_size
        self.hidden_size, self.initializer_range = H, c.initializer_range
        self._declare(_text_tables(c) + [
            ("LayerNorm", _N(H)), ("dropout", _D(c.hidden_dropout_prob)),
            ("image_embeddings", _L(c.v_feature_size, c.v_hidden_size)),
            ("image_location_embeddings", _L(c.num_locs, c.v_hidden_size)),
            ("image_layer_norm", _N(H)), ("image_location_layer_norm", _N(H)),
            ("v_LayerNorm", _N(H)), ("v_dropout", _D(c.hidden_dropout_prob))])
        self.special_initialize()

    def special_initialize(self):
        """v_LayerNorm starts as a copy of LayerNorm."""
        self.v_LayerNorm.weight = nn.Parameter(self.LayerNorm.weight.data.clone(), requires_grad=True)
        self.v_LayerNorm.bias = nn.Parameter(self.LayerNorm.bias.data.clone(), requires_grad=True)


shared_embeddings = {"vl-bert": VLBertEmbeddings, "visualbert": VisualBertEmbeddings, "uniter": UniterEmbeddings}

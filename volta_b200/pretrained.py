"""Checkpoint loading / saving with the reference's conventions (reference: volta/utils.py:341-560).

`Model.from_pretrained(path, config=config, from_hf=..., default_gpu=..., output_loading_info=..., state_dict=...)`
builds `Model(config, ...)` and loads a LOCAL checkpoint into it:
  * `path` is a directory holding `pytorch_model.bin` or the file itself (train_concap.py:187-197, train_task.py:181);
    there is no network in this deployment, so the reference's S3 shortcut names are not resolved;
  * legacy LayerNorm names `gamma` / `beta` become `weight` / `bias` (utils.py:461-472);
  * with `from_hf=True`, HuggingFace BERT layer keys are moved onto the sub-layer schedule:
    `.layer.N.attention.` -> `.layer.<bert_layer2attn_sublayer[N]>.attention_`, and `.intermediate.` / `.output.` ->
    `.layer.<bert_layer2ff_sublayer[N]>.` (utils.py:475-498);
  * a checkpoint of the bare `BertModel` loads into a model with heads and vice versa (`bert.` prefix, utils.py:521-531);
  * the decoder stays tied to the word embeddings, the model comes back in `eval()` mode.
Missing / unexpected keys are reported like the reference does (logged on the default GPU, returned with
`output_loading_info=True`); shape mismatches raise.
"""
import logging
import os
import re

import torch

logger = logging.getLogger(__name__)

WEIGHTS_NAME = "pytorch_model.bin"
CONFIG_NAME = "config.json"


def _archive_file(path):
    if os.path.isdir(path):
        return os.path.join(path, WEIGHTS_NAME)
    return path


def _legacy_layernorm_names(state):
    out = {}
    for k, v in state.items():
        out[k.replace("gamma", "weight").replace("beta", "bias")] = v
    return out


_LAYER = re.compile(r"\.layer\.(\d+)\.(attention|intermediate|output)\.")


def _remap_hf_bert_layers(state, config):
    """One HF BERT layer = one attention sub-layer + one feed-forward sub-layer of the schedule."""
    attn_map = getattr(config, "bert_layer2attn_sublayer", {}) or {}
    ff_map = getattr(config, "bert_layer2ff_sublayer", {}) or {}
    out = {}
    for k, v in state.items():
        m = _LAYER.search(k)
        if m is None:
            out[k] = v
            continue
        n, part = int(m.group(1)), m.group(2)
        if part == "attention":
            new = ".layer.%d.attention_" % attn_map.get(str(n), n)
        else:
            new = ".layer.%d.%s." % (ff_map.get(str(n), n), part)
        out[k[:m.start()] + new + k[m.end():]] = v
    return out


def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
    config = kwargs.pop("config", None)
    state_dict = kwargs.pop("state_dict", None)
    kwargs.pop("cache_dir", None)
    from_tf = kwargs.pop("from_tf", False)
    from_hf = kwargs.pop("from_hf", False)
    output_loading_info = kwargs.pop("output_loading_info", False)
    default_gpu = kwargs.pop("default_gpu", True)
    assert config is not None, "from_pretrained needs config= (as in the reference, volta/utils.py:413)"
    if from_tf:
        raise NotImplementedError("TensorFlow checkpoints are not supported; convert to a PyTorch state_dict first")

    archive = _archive_file(pretrained_model_name_or_path)
    if state_dict is None:
        if not os.path.isfile(archive):
            logger.error("'%s' is not a local checkpoint (a directory with %s or a state_dict file); shortcut names of the "
                         "reference's S3 archive map cannot be downloaded here", pretrained_model_name_or_path, WEIGHTS_NAME)
            return None  # the reference also returns None after logging (utils.py:432-444)
        if default_gpu:
            logger.info("loading weights file %s", archive)
        state_dict = torch.load(archive, map_location="cpu")

    model = cls(config, *model_args, **kwargs)

    state = _legacy_layernorm_names(dict(state_dict))
    if from_hf:
        state = _remap_hf_bert_layers(state, config)

    prefix = cls.base_model_prefix
    has_base = hasattr(model, prefix)
    ckpt_has_base = any(k.startswith(prefix) for k in state)
    target = model
    if not has_base and ckpt_has_base:       # checkpoint with heads -> bare BertModel
        state = {k[len(prefix) + 1:]: v for k, v in state.items() if k.startswith(prefix + ".")}
    elif has_base and not ckpt_has_base:     # bare-BERT checkpoint -> model with heads
        target = getattr(model, prefix)

    own = target.state_dict()
    errors = ["size mismatch for %s: checkpoint %s, model %s" % (k, tuple(v.shape), tuple(own[k].shape))
              for k, v in state.items() if k in own and tuple(v.shape) != tuple(own[k].shape)]
    if errors:
        raise RuntimeError("Error(s) in loading state_dict for %s:\n\t%s" % (model.__class__.__name__, "\n\t".join(errors)))
    result = target.load_state_dict(state, strict=False)
    missing, unexpected = list(result.missing_keys), list(result.unexpected_keys)
    if default_gpu:
        if missing:
            logger.info("Weights of %s not initialized from pretrained model: %s", model.__class__.__name__, missing)
        if unexpected:
            logger.info("Weights from pretrained model not used in %s: %s", model.__class__.__name__, unexpected)

    if hasattr(model, "tie_weights"):
        model.tie_weights()
    model.eval()
    if output_loading_info:
        return model, {"missing_keys": missing, "unexpected_keys": unexpected, "error_msgs": []}
    return model


def save_pretrained(model, save_directory):
    """`pytorch_model.bin` + `config.json` in `save_directory` (reference: volta/utils.py:341-359); a data-parallel
    wrapper is unwrapped first."""
    assert os.path.isdir(save_directory), "Saving path should be a directory where the model and configuration can be saved"
    to_save = model.module if hasattr(model, "module") else model
    with open(os.path.join(save_directory, CONFIG_NAME), "w", encoding="utf-8") as f:
        f.write(to_save.config.to_json_string())
    torch.save(to_save.state_dict(), os.path.join(save_directory, WEIGHTS_NAME))

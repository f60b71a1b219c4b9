"""The BENCHMARKED execution plan against the oracle, directly, at full width (H=768, 12 heads, 36 sub-layers).

The golden cases (tests/golden, B = 2-3) never reach the shapes at which the default plan switches on: rows packed to
the front of each stream (device-side counts), row-split GEMM launches for the two streams, the masked-LM head on the
labelled rows only, and CUDA-graph replay of the whole step.  Here the full-size models run that plan - asserted through
engine.last_plan / the graph step's replay counter - on seeded weights and a seeded ragged batch, and the three losses
and EVERY parameter gradient are compared with oracle.volta_oracle.fwd_bwd (the CPU restatement of
volta/encoders.py:1044-1114, pinned to the reference by tests/golden).

Tolerances: bf16 losses 2e-2 relative, the whole gradient 4e-2 and every parameter's gradient 6e-2 (relative L2, floored
at 2e-3 of the whole gradient's norm; observed on a B200: whole 0.027-0.035, worst tensor 0.046-0.052 - key / projection
weights whose gradients are small differences; ReLU-gated pooler weights 3e-1: units within bf16 resolution of zero flip - 0.23 seen at batch 8); fp32 validation mode 2e-4 / 2e-3."""
import pytest
import torch

import volta_b200
from oracle import volta_oracle as O
from volta_b200 import engine
from volta_b200.config import named_config
from volta_b200.synthetic import pretraining_batch, seeded_state_dict

pytestmark = pytest.mark.gpu
NO_DROPOUT = dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, v_hidden_dropout_prob=0.0,
                  v_attention_probs_dropout_prob=0.0)


def _losses(model, d):
    return model(d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"], d["image_mask"],
                 d["lm_label_ids"], d["image_label"], d["image_cls"], None, None, None, None, None, d["is_match"])


def _oracle(cfg, state, b):
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    P = O.params_from_state_dict(state, requires_grad=True)
    ref_losses = O.fwd_bwd(P, cfg, b)      # eval arithmetic: the CUDA side trains with every dropout rate set to 0
    return P, [float(x) for x in ref_losses]


def _compare(model, P, got_losses, ref_losses, loss_tol, grad_tol, pooler_tol, tag):
    for a, r in zip(got_losses, ref_losses):
        assert abs(a - r) < loss_tol * max(1.0, abs(r)), (tag, got_losses, ref_losses)
    uniq = {id(v): v for v in P.values()}
    total_ref = sum(float(v.grad.norm()) ** 2 for v in uniq.values() if v.grad is not None) ** 0.5
    bad, worst, num, den = [], (0.0, None), 0.0, 0.0
    for key, p in model.named_parameters():
        ref = P[key].grad if P[key].grad is not None else torch.zeros_like(P[key])
        got = p.grad.float().cpu() if p.grad is not None else torch.zeros_like(ref)
        err = float((got - ref).norm())
        num, den = num + err ** 2, den + float(ref.norm()) ** 2
        scale = max(float(ref.norm()), 2e-3 * total_ref)
        tol = pooler_tol if "pooler" in key else grad_tol
        if err / scale > worst[0] and "pooler" not in key:
            worst = (err / scale, key)
        if err > tol * scale:
            bad.append((key, err / scale))
    print("%s: losses %s (oracle %s); gradient rel-L2 whole %.4f, worst tensor %.4f (%s)" %
          (tag, ["%.4f" % x for x in got_losses], ["%.4f" % x for x in ref_losses], (num / den) ** 0.5, worst[0], worst[1]))
    assert not bad, (tag, bad[:8], len(bad))
    assert (num / den) ** 0.5 < (4e-2 if grad_tol > 1e-2 else grad_tol), (tag, "whole-gradient rel-L2", (num / den) ** 0.5)


def _build(name, precision, B, T, V, seed):
    cfg = named_config(name, **NO_DROPOUT)
    model = volta_b200.BertForVLPreTraining(cfg)
    state = seeded_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(state)
    model.cls.dropout.p = 0.0            # pooled-output dropout is hard-coded to 0.1 in the reference (encoders.py:750)
    model.set_precision(precision).train().cuda()
    b = pretraining_batch(B, T, V, seed=seed)
    d = {k: v.cuda() for k, v in b.items()}
    return cfg, model, state, b, d


@pytest.mark.parametrize("name", ["ctrl_vilbert_base", "ctrl_visualbert_base", "ctrl_lxmert"])
def test_default_plan_full_width_matches_oracle_bf16(name):
    """B=64, T=36, V=37: 2304 text rows = 9 tiles.  Packed rows + row-split launches (two-stream models) / joint launches
    (shared sub-layers) + masked-LM rows only + CUDA-graph replay: the third call is a replay, and it is the one compared."""
    cfg, model, state, b, d = _build(name, "bf16", 64, 36, 37, seed=77)
    got = None
    for it in range(3):
        model.zero_grad(set_to_none=True)
        l = _losses(model, d)
        (l[0] + l[1] + l[2]).sum().backward()
        got = [float(x) for x in l]
    steps = [s for s in model._graph_steps.values() if not isinstance(s, str)]
    assert len(steps) == 1 and steps[0].replays >= 1, "the compared step must be a CUDA-graph replay"
    plan = steps[0].plan
    assert plan["packed"] and plan["mlm_compact"], plan
    if name == "ctrl_vilbert_base":
        # 24 two-stream sub-layers (6 x [co-attention, ffn, self-attention, ffn]); the first one follows a text-only
        # sub-layer, whose output is not a joint buffer, and runs per-stream launches
        assert plan["row_split"] == 23 and plan["joint"] == 0, plan
    elif name == "ctrl_visualbert_base":
        assert plan["joint"] == 24, plan                                   # every sub-layer shares its weights between the streams
    else:
        # 5 shared cross layers (the first follows the text-only block: per-stream launches); 10 + 10 two-stream (self, ffn)
        assert plan["joint"] == 4 and plan["row_split"] == 20, plan
    torch.cuda.synchronize()
    P, ref = _oracle(cfg, state, b)
    _compare(model, P, got, ref, 2e-2, 6e-2, 3e-1, name + " bf16 default plan B=64")


def test_c1_shape_matches_oracle_bf16():
    """BASELINE.json configs[0]: ctrl_vilbert_base, batch 8, 36 text tokens, 36 regions + global.  288 text rows do not end
    on a 256-row tile edge: the packed buffers pad their text segment."""
    cfg, model, state, b, d = _build("ctrl_vilbert_base", "bf16", 8, 36, 37, seed=78)
    got = None
    for it in range(3):
        model.zero_grad(set_to_none=True)
        l = _losses(model, d)
        (l[0] + l[1] + l[2]).sum().backward()
        got = [float(x) for x in l]
    steps = [s for s in model._graph_steps.values() if not isinstance(s, str)]
    assert len(steps) == 1 and steps[0].replays >= 1
    assert steps[0].plan["packed"] == bool(engine.PACK_ANY_ROWS), steps[0].plan
    P, ref = _oracle(cfg, state, b)
    _compare(model, P, got, ref, 2e-2, 6e-2, 3e-1, "ctrl_vilbert_base bf16 C1 B=8")


def test_c1_shape_matches_oracle_fp32_validation_mode():
    """The fp32 validation mode (every kernel in fp32) on C1's shape: losses to 2e-4, every parameter gradient to 2e-3."""
    cfg, model, state, b, d = _build("ctrl_vilbert_base", "fp32", 8, 36, 37, seed=78)
    model.zero_grad(set_to_none=True)
    l = _losses(model, d)
    (l[0] + l[1] + l[2]).sum().backward()
    got = [float(x) for x in l]
    P, ref = _oracle(cfg, state, b)
    _compare(model, P, got, ref, 2e-4, 2e-3, 2e-3, "ctrl_vilbert_base fp32 C1 B=8")

"""Pins the CPU oracle (oracle/volta_oracle.py) to outputs of the unmodified reference.

The golden files were produced by tests/golden/make_golden.py from the live
reference; weights and inputs are regenerated from seeds."""
import pytest
import torch

from golden_util import MINI_CASES, FULL_CASES, load_case, rel_l2
from make_golden import grad_projection, weights_checksum
from oracle import volta_oracle as O
from volta_b200.synthetic import seeded_state_dict
import volta_b200

ATOL = 2e-5


def _seeded_params(cfg, g, requires_grad):
    model = volta_b200.BertForVLPreTraining(cfg)
    sd = model.state_dict()
    assert sorted(sd.keys()) == g["state_keys"], "state_dict keys differ from the reference"
    state = seeded_state_dict(sd, seed=0)
    assert abs(weights_checksum(state) - g["weights_checksum"]) < 1e-6 * max(1.0, abs(g["weights_checksum"]))
    return O.params_from_state_dict(state, requires_grad=requires_grad), model


@pytest.mark.parametrize("name", MINI_CASES + FULL_CASES)
def test_oracle_forward_matches_reference(name):
    g, cfg, b = load_case(name)
    P, _ = _seeded_params(cfg, g, False)
    taps = {}
    with torch.no_grad():
        losses = O.pretraining_losses(P, cfg, b, taps=taps)
    for i, n in enumerate(g["layer_ids"]):
        assert torch.allclose(taps["layers_t"][n], g["layers_t"][i], atol=ATOL), (name, "t", n)
        assert torch.allclose(taps["layers_v"][n], g["layers_v"][i], atol=ATOL), (name, "v", n)
    for key in ("pooled_t", "pooled_v"):
        if g[key] is not None:
            assert torch.allclose(taps[key], g[key], atol=ATOL)
    k = g["mlm_logits"].shape[-1]
    assert torch.allclose(taps["mlm_logits"][..., :k], g["mlm_logits"], atol=1e-4)
    k = g["img_logits"].shape[-1]
    assert torch.allclose(taps["img_logits." + g.get("img_target", "0")][..., :k], g["img_logits"], atol=1e-4)
    if g["itm_logits"] is not None:
        assert torch.allclose(taps["itm_logits"], g["itm_logits"], atol=ATOL)
    for mine, ref in zip(losses, g["losses"]):
        assert abs(float(mine) - ref) < 1e-4 * max(1.0, abs(ref)), (name, float(mine), ref)


@pytest.mark.parametrize("name", MINI_CASES + ["full_ctrl_vilbert"])
def test_oracle_gradients_match_reference(name):
    g, cfg, b = load_case(name)
    P, model = _seeded_params(cfg, g, True)
    O.fwd_bwd(P, cfg, b)
    checked = 0
    for key, _ in model.named_parameters():
        gr = P[key].grad
        gr = torch.zeros_like(P[key]) if gr is None else gr
        n_ref, p_ref = g["grad_norm"][key], g["grad_proj"][key]
        assert abs(float(gr.double().norm()) - n_ref) <= 2e-4 * max(n_ref, 1e-3), (name, key)
        assert abs(grad_projection(gr) - p_ref) <= 2e-4 * max(n_ref, 1e-3) * 4, (name, key)
        checked += 1
    assert checked == len(g["grad_norm"])


# ---- BertForVLTasks heads (volta/encoders.py:1117-1206) ---------------------------------------------------------
import make_golden_tasks  # noqa: E402


@pytest.mark.parametrize("name", list(make_golden_tasks.CASES))
def test_oracle_task_heads_match_reference(name):
    g = torch.load(__import__("os").path.join(__import__("golden_util").ROOT, "tests", "golden", name + ".pt"), weights_only=False)
    cfg = volta_b200.BertConfig.from_dict(g["config"])
    model = volta_b200.BertForVLTasks(cfg, g["task_cfg"], g["tasks"])
    sd = model.state_dict()
    assert sorted(sd.keys()) == g["state_keys"], "state_dict keys differ from the reference"
    state = seeded_state_dict(sd, seed=0)
    assert abs(weights_checksum(state) - g["weights_checksum"]) < 1e-6 * max(1.0, abs(g["weights_checksum"]))
    P = O.params_from_state_dict(state)
    b = make_golden_tasks.case_batch(name)
    with torch.no_grad():
        for t in g["tasks"]:
            mine = O.task_prediction(P, cfg, b, t, g["task_cfg"])
            assert mine.shape == g["pred"][t].shape, (name, t)
            assert torch.allclose(mine, g["pred"][t], atol=5e-5, rtol=1e-5), (name, t, float((mine - g["pred"][t]).abs().max()))

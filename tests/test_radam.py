"""FusedRAdam (vb_radam_flat) against trajectories of the UNMODIFIED reference RAdam (volta/optimization.py:9-93;
tests/golden/make_golden_radam.py): 12 steps, one group per parameter with its own lr / weight decay, through the switch
of the variance rectification (step 6) and the wrap of the reference's ten-slot step-size buffer.  CPU: host logic with
the kernel emulated from its documented semantics; GPU: the kernel.  fp32 arithmetic, tolerance 2e-6 relative per tensor
(the kernel contracts b*m + (1-b)*g into FMAs, the reference rounds twice)."""
import os

import pytest
import torch

import emu_ops
from golden_util import ROOT  # (puts tests/golden on sys.path)
import make_golden_radam as R  # noqa: E402  (setup shared with the generator; does not import the reference)
from volta_b200.optim import FusedRAdam


def _run(device, checks_at):
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "radam.pt"), weights_only=False)
    model = R.radam_setup()
    model.set_precision("fp32" if device == "cpu" else "bf16")
    model.to(device)
    named = list(model.named_parameters())
    groups = [{"params": [p], "lr": R.lr_of(n), "weight_decay": R.wd_of(n)} for n, p in named]
    opt = FusedRAdam(groups, lr=2e-3, model=model)
    assert opt.defaults["eps"] == 1e-8 and opt.defaults["betas"] == (0.9, 0.999) and "correct_bias" not in opt.param_groups[0]
    for step in range(1, gold["steps"] + 1):
        for (_, p), g in zip(named, R.grads_of(named, step)):
            p.grad = g.to(device)
        opt.step()
        if step in gold["checks"] and step in checks_at:
            for n, p in named:
                ref = gold["checks"][step][n]
                got = (R.summary(p), R.summary(opt.state[p]["exp_avg"]), R.summary(opt.state[p]["exp_avg_sq"]))
                for (rs, ra), (gs, ga), what in zip(ref, got, ("p", "exp_avg", "exp_avg_sq")):
                    assert abs(ga - ra) <= 2e-6 * max(ra, 1e-12), (step, n, what, ga, ra)
                    assert abs(gs - rs) <= 2e-6 * max(ra, 1e-12), (step, n, what, gs, rs)
    for n, p in named:
        if n in gold["full"]:
            for ref, got in zip(gold["full"][n], (p, opt.state[p]["exp_avg"], opt.state[p]["exp_avg_sq"])):
                got = got.detach().float().cpu()
                assert float((got - ref).abs().max()) <= 2e-6 * float(ref.abs().max()), n
    assert all(int(opt.state[p]["step"]) == gold["steps"] for _, p in named)
    return model, opt


def test_fused_radam_host_logic_matches_reference_trajectory(monkeypatch):
    emu_ops.install(monkeypatch)
    _run("cpu", (5, 6, 12))


def test_fused_radam_skips_parameters_without_gradients(monkeypatch):
    """`if p.grad is None: continue` (volta/optimization.py:27-28): untouched weights, moments and step count."""
    emu_ops.install(monkeypatch)
    model = R.radam_setup()
    model.set_precision("fp32")
    named = list(model.named_parameters())
    opt = FusedRAdam([{"params": [p], "lr": 1e-3, "weight_decay": 0.01} for _, p in named], model=model)
    before = [p.detach().clone() for _, p in named]
    for i, ((_, p), g) in enumerate(zip(named, R.grads_of(named, 1))):
        p.grad = g if i % 2 == 0 else None
    opt.step()
    for i, ((n, p), b) in enumerate(zip(named, before)):
        assert (not torch.equal(p.detach(), b)) == (i % 2 == 0), n
        assert int(opt.state[p]["step"]) == (1 if i % 2 == 0 else 0)


@pytest.mark.gpu
def test_fused_radam_matches_reference_trajectory_gpu():
    model, opt = _run("cuda", (5, 6, 12))
    # the bf16 operand copy the GEMMs read was rewritten by the same launch
    from volta_b200 import engine
    store = engine.model_store(model)
    assert torch.equal(store.flat16.float(), store.flat32.to(torch.bfloat16).float())

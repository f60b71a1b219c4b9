"""FusedAdamW's host logic on CPU (flat-store layout, per-parameter groups -> 64-element chunks, per-group step sizes,
clip folding, LR schedulers, state_dict round trip) with vb_adamw_flat / vb_sumsq emulated from their documented
semantics (tests/emu_ops.py), against the plain-PyTorch restatement of the reference's step (tests/adamw_ref.py:
train_concap.py:199-234, 305-313).  The real kernels are checked against the same restatement in test_optim_gpu.py."""
import pytest
import torch

import emu_ops
import volta_b200
from adamw_ref import NO_DECAY, adamw_step
from golden_util import load_case
from volta_b200.optim import FusedAdamW, clip_grad_norm_
from volta_b200.synthetic import pretraining_batch, seeded_state_dict


def _losses(model, d):
    return model(d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"], d["image_mask"],
                 d["lm_label_ids"], d["image_label"], d["image_cls"], None, None, None, None, None, d["is_match"])


def _model(name="mini_ctrl_vilbert"):
    _, cfg, _ = load_case(name)
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision("fp32").eval()
    return cfg, model


@pytest.mark.parametrize("per_param_groups,clip", [(False, 0.0), (True, 0.5)])
def test_fused_adamw_host_logic_matches_restated_adamw(per_param_groups, clip, monkeypatch):
    emu_ops.install(monkeypatch)
    cfg, model = _model()
    named = list(model.named_parameters())
    lr_of = (lambda n: 3e-4 * (0.1 if "encoder.layer.1." in n or "embeddings" in n else 1.0)) if per_param_groups else (lambda n: 3e-4)
    wd_of = lambda n: 0.0 if any(nd in n for nd in NO_DECAY) else 0.01
    if per_param_groups:   # train_concap.py:212-224
        groups = [{"params": [p], "lr": lr_of(n), "weight_decay": wd_of(n)} for n, p in named]
    else:                  # train_concap.py:203-210
        groups = [{"params": [p for n, p in named if wd_of(n) > 0], "weight_decay": 0.01},
                  {"params": [p for n, p in named if wd_of(n) == 0], "weight_decay": 0.0}]
    opt = FusedAdamW(groups, lr=3e-4, betas=(0.9, 0.999), eps=1e-6, model=model, max_grad_norm=clip)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 1.0 / (1 + it))     # any scheduler that edits group["lr"]
    ref_named = [(n, torch.nn.Parameter(p.detach().clone())) for n, p in named]
    state = {}
    for it in range(1, 4):
        b = pretraining_batch(4, 8, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=50 + it)
        opt.zero_grad()
        l = _losses(model, b)
        (l[0] + l[1] + l[2]).sum().backward()
        for (n, p), (_, q) in zip(named, ref_named):
            q.grad = p.grad.detach().clone()
        factor = 1.0 / it                                                          # LambdaLR after it-1 scheduler steps
        adamw_step(ref_named, state, 3e-4, it, clip=clip, lr_of=lambda n: lr_of(n) * factor, wd_of=wd_of)
        opt.step()
        sched.step()
        worst = max(float((p.detach() - q.detach()).abs().max() / q.detach().abs().max().clamp(min=1e-8))
                    for (_, p), (_, q) in zip(named, ref_named))
        assert worst < 2e-6, (it, worst)
    sd = opt.state_dict()
    assert all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in sd["state"].values()) and len(sd["state"]) == len(named)
    assert all(int(v["step"]) == 3 for v in sd["state"].values())
    # resume: a fresh optimizer loaded from the state dict continues exactly like the original
    opt2 = FusedAdamW([{k: v for k, v in g.items()} for g in groups], lr=3e-4, betas=(0.9, 0.999), eps=1e-6, model=model,
                      max_grad_norm=clip)
    opt2.load_state_dict(sd)
    k3 = named[3][1]
    assert torch.equal(opt2.state[k3]["exp_avg"], opt.state[k3]["exp_avg"])
    before = {n: p.detach().clone() for n, p in named}
    opt.step()
    after1 = {n: p.detach().clone() for n, p in named}
    with torch.no_grad():
        for n, p in named:
            p.copy_(before[n])
    opt2.step()
    for n, p in named:
        assert torch.equal(p.detach(), after1[n]), n


def test_groups_that_split_a_fused_chunk_are_rejected(monkeypatch):
    """Parameters of one fused group share 64-element chunks only with themselves; a bias of 2 elements given its own
    hyper-parameters next to a neighbour in the same chunk would be a silent error - it must raise."""
    emu_ops.install(monkeypatch)
    cfg, model = _model()
    with pytest.raises(ValueError):
        FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.999), model=None)
    with pytest.raises(ValueError):
        FusedAdamW([{"params": [p], "betas": (0.9, 0.99 if i else 0.999)} for i, p in enumerate(model.parameters())],
                   lr=1e-3, model=model)
    other = torch.nn.Parameter(torch.zeros(3))
    opt = FusedAdamW([{"params": list(model.parameters()) + [other]}], lr=1e-3, model=model)
    b = pretraining_batch(2, 8, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=1)
    l = _losses(model, b)
    (l[0] + l[1] + l[2]).sum().backward()
    with pytest.raises(ValueError):
        opt.step()


def test_clip_grad_norm_on_the_flat_buffer_emulated(monkeypatch):
    emu_ops.install(monkeypatch)
    cfg, model = _model("mini_ctrl_visualbert")
    b = pretraining_batch(4, 8, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=3)
    l = _losses(model, b)
    (l[0] + l[1] + l[2]).sum().backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    norm = clip_grad_norm_(model, 0.25 * float(total))
    assert abs(float(norm) - float(total)) < 1e-4 * float(total)
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, grads[n] * 0.25, rtol=1e-4, atol=1e-9), n

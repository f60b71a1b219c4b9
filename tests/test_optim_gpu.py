"""FusedAdamW / clip_grad_norm_ (volta_b200/optim.py, csrc/optim.cu) against the plain-PyTorch restatement of the
reference's optimizer step (tests/adamw_ref.py)."""
import pytest
import torch

import volta_b200
from adamw_ref import NO_DECAY, adamw_step
from golden_util import load_case
from volta_b200 import engine
from volta_b200.optim import FusedAdamW, clip_grad_norm_
from volta_b200.synthetic import pretraining_batch, seeded_state_dict

pytestmark = pytest.mark.gpu


def _losses(model, d):
    return model(d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"], d["image_mask"],
                 d["lm_label_ids"], d["image_label"], d["image_cls"], None, None, None, None, None, d["is_match"])


@pytest.mark.parametrize("per_param_groups,clip", [(False, 0.0), (True, 0.5)])
def test_fused_adamw_matches_restated_adamw(per_param_groups, clip):
    _, cfg, _ = load_case("mini_ctrl_vilbert")
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.train().cuda()
    named = list(model.named_parameters())
    lr_of = (lambda n: 3e-4 * (0.1 if "encoder.layer.1." in n or "embeddings" in n else 1.0)) if per_param_groups else (lambda n: 3e-4)
    wd_of = lambda n: 0.0 if any(nd in n for nd in NO_DECAY) else 0.01
    if per_param_groups:   # train_concap.py:212-224: one group per parameter with its own lr / weight decay
        groups = [{"params": [p], "lr": lr_of(n), "weight_decay": wd_of(n)} for n, p in named]
    else:                  # train_concap.py:203-210
        groups = [{"params": [p for n, p in named if wd_of(n) > 0], "weight_decay": 0.01},
                  {"params": [p for n, p in named if wd_of(n) == 0], "weight_decay": 0.0}]
    opt = FusedAdamW(groups, lr=3e-4, betas=(0.9, 0.999), eps=1e-6, model=model, max_grad_norm=clip)
    ref = {n: p.detach().clone() for n, p in named}          # fp32 shadow driven by the restated optimizer
    ref_named = [(n, torch.nn.Parameter(t)) for n, t in ref.items()]
    state = {}
    for it in range(1, 6):
        b = pretraining_batch(8, 12, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=50 + it)
        d = {k: v.cuda() for k, v in b.items()}
        opt.zero_grad()
        l = _losses(model, d)
        (l[0] + l[1] + l[2]).sum().backward()
        for (n, p), (_, q) in zip(named, ref_named):
            q.grad = p.grad.detach().clone()
        adamw_step(ref_named, state, 3e-4, it, clip=clip, lr_of=lr_of, wd_of=wd_of)
        opt.step()
        worst = max(float((p.detach() - q.detach()).abs().max() / q.detach().abs().max().clamp(min=1e-8)) for (_, p), (_, q) in zip(named, ref_named))
        assert worst < 2e-6, (it, worst)
        store = engine.model_store(model)
        assert torch.equal(store.flat16, store.flat32.bfloat16())     # the bf16 operand copy was refreshed by the same kernel
    sd = opt.state_dict()
    assert all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in sd["state"].values()) and len(sd["state"]) == len(named)
    k = named[3][0]
    assert torch.allclose(opt.state[named[3][1]]["exp_avg"], state[k][0], atol=1e-7)


def test_clip_grad_norm_on_the_flat_buffer():
    _, cfg, _ = load_case("mini_ctrl_visualbert")
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.train().cuda()
    b = pretraining_batch(4, 12, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=3)
    d = {k: v.cuda() for k, v in b.items()}
    l = _losses(model, d)
    (l[0] + l[1] + l[2]).sum().backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    norm = clip_grad_norm_(model, 0.25 * float(total))
    assert abs(float(norm) - float(total)) < 1e-4 * float(total)
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, grads[n] * 0.25, rtol=1e-4, atol=1e-9), n

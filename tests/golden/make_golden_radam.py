"""Golden optimizer trajectories from the UNMODIFIED reference `RAdam` (volta/optimization.py:9-93; build container only).

    python tests/golden/make_golden_radam.py       # writes tests/golden/radam.pt

Setup shared with tests/test_radam.py (`radam_setup`): the parameters of the mini_ctrl_vilbert pre-training model
(seeded weights), one group per parameter as train_task.py:206-217 builds them (own lr, weight decay 0 for biases /
LayerNorm), 12 steps of seeded gradients - the variance rectification switches on at step 6 (N_sma >= 5), the ten-slot
step-size buffer wraps at step 10, and the groups' different learning rates exercise the buffer's first-group-wins reuse.
Stored: float64 (sum, sum of |.|) of every parameter / exp_avg / exp_avg_sq after steps 5, 6 and 12, and the full
tensors of a few parameters after step 12.
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
STEPS, CHECK = 12, (5, 6, 12)
FULL = ("bert.embeddings.LayerNorm.weight", "bert.encoder.layer.0.attention_self.query.bias",
        "bert.encoder.layer.2.attention_output.dense.weight", "bert.encoder.layer.3.intermediate.dense.bias", "bert.v_pooler.dense.weight", "cls.bi_seq_relationship.weight")
NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")


def radam_setup():
    """(named parameters as fresh leaf tensors, group dicts, gradient generator) - no reference import."""
    import make_golden as G
    import volta_b200
    from volta_b200.synthetic import seeded_state_dict
    d, _ = G.case_config_dict("mini_ctrl_vilbert")
    model = volta_b200.BertForVLPreTraining(volta_b200.BertConfig.from_dict(d))
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    return model


def lr_of(name):
    return 1e-3 if ".v_" in name or "v_embeddings" in name else 2e-3


def wd_of(name):
    return 0.0 if any(nd in name for nd in NO_DECAY) else 0.01


def grads_of(named, step):
    g = torch.Generator().manual_seed(1000 + step)
    return [torch.randn(p.shape, generator=g) * (0.01 * (1 + i % 3)) for i, (_, p) in enumerate(named)]


def summary(t):
    t = t.detach().double()
    return (float(t.sum()), float(t.abs().sum()))


if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("ref_optimization", "/root/reference/volta/optimization.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    model = radam_setup()
    named = [(n, torch.nn.Parameter(p.detach().clone())) for n, p in model.named_parameters()]
    groups = [{"params": [p], "lr": lr_of(n), "weight_decay": wd_of(n)} for n, p in named]
    opt = ref.RAdam(groups, lr=2e-3)
    out = {"steps": STEPS, "checks": {}, "full": {}}
    for step in range(1, STEPS + 1):
        for (_, p), g in zip(named, grads_of(named, step)):
            p.grad = g
        opt.step()
        if step in CHECK:
            out["checks"][step] = {n: (summary(p), summary(opt.state[p]["exp_avg"]), summary(opt.state[p]["exp_avg_sq"]))
                                   for n, p in named}
    for n, p in named:
        if n in FULL:
            out["full"][n] = (p.detach().clone(), opt.state[p]["exp_avg"].clone(), opt.state[p]["exp_avg_sq"].clone())
    assert len(out["full"]) == len(FULL), sorted(set(FULL) - set(out["full"]))
    path = os.path.join(HERE, "radam.pt")
    torch.save(out, path)
    print(len(named), "parameters,", STEPS, "steps ->", path, os.path.getsize(path) // 1024, "KB")

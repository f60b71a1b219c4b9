"""Host-logic regressions found by the round-1 review, run on CPU with the C ABI emulated (tests/emu_ops.py):

  * BertForVLTasks owns ONE flat store (encoder + task heads): the data-parallel hooks fire, FusedAdamW updates it, task
    heads a step does not use are left alone (reference: train_task.py:222-253 wraps exactly this model);
  * the bf16 operand copy is rebuilt after a `p.data`-style optimizer step (pytorch_transformers.AdamW / RAdam write
    through `p.data`, which does not bump `p._version`);
  * frozen layers (volta/train_utils.py:250-255 freeze_layers, config.fixed_layers) contribute nothing to the flat
    gradient norm - it equals torch.nn.utils.clip_grad_norm_(model.parameters())."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emu_ops
import volta_b200
from golden_util import ROOT, load_case
from volta_b200.optim import FusedAdamW, clip_grad_norm_
from volta_b200.synthetic import pretraining_batch, seeded_state_dict

import make_golden_tasks  # noqa: E402


def _tasks_model(precision="fp32"):
    g = torch.load(os.path.join(ROOT, "tests", "golden", "tasks_mini_ctrl_vilbert.pt"), weights_only=False)
    cfg = volta_b200.BertConfig.from_dict(g["config"])
    model = volta_b200.BertForVLTasks(cfg, g["task_cfg"], g["tasks"])
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision(precision).eval()
    return g, cfg, model


def _tasks_step(model, g, task, seed_shift=0.0):
    b = make_golden_tasks.case_batch("tasks_mini_ctrl_vilbert")
    feat = b["image_feat"] + seed_shift
    vil, _, _, _ = model(b["input_ids"], feat, b["image_loc"], task, b["segment_ids"], b["input_mask"], b["image_mask"])
    keep = vil > -5000.0
    (vil * keep).float().pow(2).sum().backward()


def test_tasks_model_has_one_store_and_fused_adamw_updates_it(monkeypatch):
    emu_ops.install(monkeypatch)
    from volta_b200 import engine
    g, cfg, model = _tasks_model()
    assert engine.model_store(model.bert) is engine.model_store(model)
    tasks = list(g["tasks"])
    used, unused = tasks[0], tasks[1:]
    assert unused, "the golden task case holds several heads"
    groups = [{"params": [p], "weight_decay": 0.01} for p in model.parameters()]
    opt = FusedAdamW(groups, lr=1e-2, betas=(0.9, 0.999), eps=1e-6, model=model)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt.zero_grad()
    _tasks_step(model, g, used)
    store = engine.model_store(model)
    base = store.grad32.data_ptr()
    for n, p in model.named_parameters():
        if n.startswith("bert.") or n.startswith("clfs_dict.%s." % used):
            assert p.grad is not None and p.grad.data_ptr() == base + 4 * store.offsets[id(p)][0], n   # lives in grad32
        elif n.startswith("clfs_dict."):
            assert p.grad is None, n                                                                    # unused head
    opt.step()
    moved = {n: float((p.detach() - before[n]).abs().max()) for n, p in model.named_parameters()}
    assert moved["bert.encoder.layer.0.attention_self.query.weight"] > 0
    assert any(v > 0 for n, v in moved.items() if n.startswith("clfs_dict.%s." % used))
    for n, v in moved.items():
        if any(n.startswith("clfs_dict.%s." % t) for t in unused):
            assert v == 0.0, ("a head without a gradient was updated (weight decay / moments)", n)
    sd = opt.state_dict()["state"]
    steps = sorted({int(v["step"]) for v in sd.values()})
    assert steps == [0, 1], steps        # unused heads did not step
    # an optimizer step with no backward since zero_grad is a no-op, as in torch
    opt.zero_grad()
    snap = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt.step()
    assert all(torch.equal(p.detach(), snap[n]) for n, p in model.named_parameters())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddp_tasks_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emu_ops.install(pytest.MonkeyPatch())
        from volta_b200.parallel import DistributedDataParallel
        g, cfg, ref = _tasks_model()
        task = list(g["tasks"])[0]
        singles = []
        for r in range(world):
            ref.zero_grad(set_to_none=True)
            _tasks_step(ref, g, task, seed_shift=0.05 * r)
            singles.append({k: (None if p.grad is None else p.grad.clone()) for k, p in ref.named_parameters()})
        g, cfg, model = _tasks_model()
        ddp = DistributedDataParallel(model, message_size=100_000, delay_allreduce=(rank < 0))
        model.zero_grad(set_to_none=True)
        _tasks_step(ddp, g, task, seed_shift=0.05 * rank)
        worst, n = 0.0, 0
        for k, p in model.named_parameters():
            if singles[0][k] is None:
                continue
            expect = sum(s[k] for s in singles) / world
            worst = max(worst, float((p.grad - expect).norm()) / max(float(expect.norm()), 1e-9))
            n += 1
        out.put((rank, (worst, n, ddp.last_num_buckets)))
    except Exception:  # noqa
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_tasks_model_under_ddp_averages_gradients_two_ranks_gloo():
    """The data-parallel wrapper around BertForVLTasks (train_task.py:253): every rank ends the backward with the MEAN
    of the ranks' gradients - encoder (written by the engine) and head (written by autograd) alike."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_tasks_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, rep in results:
        assert isinstance(rep, tuple), rep
        worst, n, buckets = rep
        assert worst < 1e-6 and n > 100 and buckets >= 1, (rank, rep)


class _DataStyleSGD(torch.optim.Optimizer):
    """Writes through `p.data` like pytorch_transformers.AdamW (`p.data.addcdiv_`) and volta/optimization.py RAdam
    (`p.data.copy_`): such updates leave `p._version` untouched."""

    def __init__(self, params, lr):
        super().__init__(params, dict(lr=lr))

    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    p.data.add_(p.grad.data, alpha=-group["lr"])


def _pretrain(precision):
    _, cfg, _ = load_case("mini_ctrl_vilbert")
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision(precision).eval()
    return cfg, model


def _losses(model, d):
    return model(d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"], d["image_mask"],
                 d["lm_label_ids"], d["image_label"], d["image_cls"], None, None, None, None, None, d["is_match"])


def test_bf16_operands_follow_a_p_data_style_optimizer(monkeypatch):
    emu_ops.install(monkeypatch)
    from volta_b200 import engine
    cfg, model = _pretrain("bf16")
    b = pretraining_batch(4, 8, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=3)
    opt = _DataStyleSGD(model.parameters(), lr=0.05)
    versions = sum(p._version for p in model.parameters())
    for _ in range(2):
        opt.zero_grad()
        l = _losses(model, b)
        (l[0] + l[1] + l[2]).sum().backward()
        opt.step()
    assert sum(p._version for p in model.parameters()) == versions, "this optimizer is meant to be invisible to _version"
    with torch.no_grad():
        got = [float(x) for x in _losses(model, b)]
    store = engine.model_store(model)
    assert torch.equal(store.flat16.float(), store.flat32.bfloat16().float()), "the bf16 operand copy is stale"
    store.refresh()
    with torch.no_grad():
        want = [float(x) for x in _losses(model, b)]
    assert got == want, (got, want)


def test_flat_gradient_norm_ignores_frozen_layers(monkeypatch):
    emu_ops.install(monkeypatch)
    from volta_b200 import engine
    cfg, model = _pretrain("fp32")
    for n, p in model.named_parameters():          # freeze_layers(model) with fixed_layers = [0, 1]
        if ".encoder.layer.0." in n or ".encoder.layer.1." in n:
            p.requires_grad_(False)
    b = pretraining_batch(4, 8, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=3)
    l = _losses(model, b)
    (l[0] + l[1] + l[2]).sum().backward()
    for n, p in model.named_parameters():
        assert (p.grad is None) == (not p.requires_grad), n
    torch_norm = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None))
    store = engine.model_store(model)
    flat_norm = store.grad32.double().norm()
    assert abs(float(flat_norm) - float(torch_norm)) < 1e-6 * float(torch_norm), (float(flat_norm), float(torch_norm))
    norm = clip_grad_norm_(model, 1e9)
    assert abs(float(norm) - float(torch_norm)) < 1e-4 * float(torch_norm)
    # the unfrozen gradients are what an unfrozen model computes
    cfg2, full = _pretrain("fp32")
    l = _losses(full, b)
    (l[0] + l[1] + l[2]).sum().backward()
    for (n, p), (_, q) in zip(model.named_parameters(), full.named_parameters()):
        if p.requires_grad:
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-8), n

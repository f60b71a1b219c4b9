"""Host-logic tests (no GPU): the engine's launch plan, hand-written backward and flat parameter store are run
with the C ABI emulated in PyTorch (tests/emu_ops.py) in the fp32 validation mode and compared with the
oracle on the golden cases' seeded weights and inputs."""
import pytest
import torch

import emu_ops
import volta_b200
from volta_b200 import ops as ops_mod
from golden_util import MINI_CASES, load_case, rel_l2
from oracle import volta_oracle as O
from volta_b200.synthetic import seeded_state_dict

CASES = list(MINI_CASES)


def _build(name):
    g, cfg, b = load_case(name)
    model = volta_b200.BertForVLPreTraining(cfg)
    state = seeded_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(state)
    model.set_precision("fp32").eval()
    return g, cfg, b, model, state


@pytest.mark.parametrize("name", CASES + ["mini_ctrl_vilbert+img_compact"])
def test_engine_forward_backward_matches_oracle(name, monkeypatch):
    emu_ops.install(monkeypatch)
    if name.endswith("+img_compact"):     # the optional compaction of the masked-region head (engine.IMG_COMPACT)
        from volta_b200 import engine
        monkeypatch.setattr(engine, "IMG_COMPACT", True)
        name = name.split("+")[0]
    g, cfg, b, model, state = _build(name)
    losses = model(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"],
                   b["lm_label_ids"], b["image_label"], b["image_cls"], b.get("obj_labels"), b.get("obj_confs"), b.get("attr_labels"),
                   b.get("attr_confs"), None, b["is_match"])
    for mine, ref in zip(losses, g["losses"]):
        assert abs(float(mine) - ref) < 2e-4 * max(1.0, abs(ref)), (name, float(mine), ref)
    (losses[0] + losses[1] + losses[2]).sum().backward()
    P = O.params_from_state_dict(state, requires_grad=True)
    O.fwd_bwd(P, cfg, b)
    worst = 0.0
    for key, p in model.named_parameters():
        ref = P[key].grad if P[key].grad is not None else torch.zeros_like(P[key])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        scale = max(float(ref.norm()), 1e-6)
        err = float((got - ref).norm()) / scale
        worst = max(worst, err)
        assert err < 2e-3, (name, key, err, float(ref.norm()), float(got.norm()))
    assert worst < 2e-3


@pytest.mark.parametrize("name", ["mini_ctrl_vilbert", "mini_ctrl_visualbert"])
def test_engine_layer_taps_match_reference(name, monkeypatch):
    emu_ops.install(monkeypatch)
    g, cfg, b, model, _ = _build(name)
    with torch.no_grad():
        lt, lv, pt, pv, _ = model.bert(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"],
                                       b["input_mask"], b["image_mask"], output_all_encoded_layers=True)
    B = b["input_ids"].shape[0]
    # padded positions are not computed (they are zeros; the reference leaves finite garbage nobody reads - SURVEY.md
    # 8a' item 3): compare the valid rows, and check that the skipped ones really are zeros in the two-stream sub-layers
    vt, vv = b["input_mask"].bool(), b["image_mask"].bool()
    for i, n in enumerate(g["layer_ids"]):
        mt, mv = lt[n].view(B, -1, lt[n].shape[-1]).float(), lv[n].view(B, -1, lv[n].shape[-1]).float()
        assert torch.allclose(mt[vt], g["layers_t"][i][vt], atol=1e-4), (name, n)
        assert torch.allclose(mv[vv], g["layers_v"][i][vv], atol=1e-4), (name, n)
    assert torch.allclose(pt.float(), g["pooled_t"], atol=1e-4)
    assert torch.allclose(pv.float(), g["pooled_v"], atol=1e-4)


@pytest.mark.parametrize("name", ["mini_ctrl_vilbert", "mini_ctrl_lxmert", "mini_vilbert"])
def test_row_split_launch_plan_matches_oracle(name, monkeypatch):
    """Two-stream sub-layers run their text and vision rows through ONE row-split GEMM per Linear when the rows share a
    buffer (engine.split_ok; needs 256-row alignment on the device).  Here the plan is forced on under emulation:
    same losses and gradients as the oracle, and the split path was really taken."""
    from volta_b200 import engine
    emu_ops.install(monkeypatch)
    taken = []
    real = engine.split_ok

    def forced(rt, x, p_t, p_v, *pairs):
        ok = x.all is not None and p_t == p_v and all(a.weight.shape == b.weight.shape for a, b in pairs)
        taken.append(ok)
        return ok
    monkeypatch.setattr(engine, "split_ok", forced)
    g, cfg, b, model, state = _build(name)
    losses = model(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"],
                   b["lm_label_ids"], b["image_label"], b["image_cls"], None, None, None, None, None, b["is_match"])
    for mine, ref in zip(losses, g["losses"]):
        assert abs(float(mine) - ref) < 2e-4 * max(1.0, abs(ref)), (name, float(mine), ref)
    (losses[0] + losses[1] + losses[2]).sum().backward()
    P = O.params_from_state_dict(state, requires_grad=True)
    O.fwd_bwd(P, cfg, b)
    for key, p in model.named_parameters():
        ref = P[key].grad if P[key].grad is not None else torch.zeros_like(P[key])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert float((got - ref).norm()) < 2e-3 * max(float(ref.norm()), 1e-6), (name, key)
    assert any(taken) == (name != "mini_vilbert"), taken   # vilbert_base: the streams differ in width - never split
    assert real is not forced


@pytest.mark.parametrize("name,shape", [("mini_ctrl_vilbert", (32, 8, 9)), ("mini_ctrl_lxmert", (32, 8, 9)),
                                        ("mini_ctrl_visualbert", (32, 8, 9)), ("mini_ctrl_uniter", (32, 8, 9)),
                                        ("mini_ctrl_vlbert", (32, 8, 9)), ("mini_ctrl_vilbert", (30, 8, 9)),
                                        ("mini_ctrl_lxmert", (37, 7, 8)), ("mini_ctrl_visualbert", (30, 8, 9))])
def test_packed_rows_plan_matches_oracle(name, shape, monkeypatch):
    """engine.bert_fwd packs the valid rows of both streams between the embeddings and the heads (device-side counts for
    the GEMMs / LayerNorms, per-sample offsets for the attention).  The plan needs the tensor-core kernels on the device;
    here it is forced on under the emulation - the host logic (packing plan, limits, zero-row invariant, un-packing) must
    reproduce the oracle's losses and every parameter gradient on a batch whose text rows end on a 256-row boundary.
    ctrl_lxmert / the single-stream families have SHARED sub-layers (one GEMM over the joint buffer: two runs of valid
    rows, weight and bias gradients reduced per run) - the plan behind VOLTA_B200_PACK_SHARED=1.  Shapes whose B*T is not a
    multiple of 256 (30 x 8, 37 x 7) pad the text segment of the packed buffers to the next tile edge
    (VOLTA_B200_PACK_ANY_ROWS=1)."""
    from volta_b200 import engine
    from volta_b200.synthetic import pretraining_batch
    emu_ops.install(monkeypatch)
    taken = []

    def forced(rt, bert, keep_layers):
        ok = not keep_layers and engine._pad256(rt.Mt) != rt.Mv
        taken.append(ok)
        return ok
    monkeypatch.setattr(engine, "_pack_supported", forced)
    monkeypatch.setattr(engine, "PLAN_EMULATED", True)   # row-split GEMMs and two-segment weight-gradient launches (vb_gemm_args::k_seg / D2)
    grouped = []
    real_gemm = ops_mod.gemm

    def spy(*a, **k):
        if k.get("out2") is not None:
            grouped.append(k["k_seg"])
        return real_gemm(*a, **k)
    monkeypatch.setattr(ops_mod, "gemm", spy)
    g, cfg, _, model, state = _build(name)
    b = pretraining_batch(*shape, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=11)
    losses = model(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"],
                   b["lm_label_ids"], b["image_label"], b["image_cls"], None, None, None, None, None, b["is_match"])
    assert taken and all(taken)
    (losses[0] + losses[1] + losses[2]).sum().backward()
    assert len(grouped) >= 10 and all(k % 256 == 0 for k in grouped), grouped[:4]   # the grouped plan was taken
    P = O.params_from_state_dict(state, requires_grad=True)
    ref = O.pretraining_losses(P, cfg, b)
    for mine, r in zip(losses, ref):
        assert abs(float(mine) - float(r)) < 2e-4 * max(1.0, abs(float(r))), (float(mine), float(r))
    sum(ref).backward()
    for key, p in model.named_parameters():
        rg = P[key].grad if P[key].grad is not None else torch.zeros_like(P[key])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        err = float((got - rg).norm()) / max(float(rg.norm()), 1e-6)
        assert err < 2e-3, (name, key, err)


def _tasks_cases():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_tasks
    return make_golden_tasks


@pytest.mark.parametrize("name", ["tasks_mini_ctrl_vilbert", "tasks_mini_ctrl_visualbert"])
def test_task_heads_host_logic_matches_reference(name, monkeypatch):
    """engine.tasks_forward (fusion of the pooled outputs, every head type, V-logit masking) with the C ABI emulated:
    predictions equal the unmodified reference's (tests/golden/tasks_*.pt), and a task loss back-propagates into every
    parameter the prediction depends on (fine-tuning goes through the same hand-written backward)."""
    import os
    from golden_util import ROOT
    mk = _tasks_cases()
    emu_ops.install(monkeypatch)
    g = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
    cfg = volta_b200.BertConfig.from_dict(g["config"])
    model = volta_b200.BertForVLTasks(cfg, g["task_cfg"], g["tasks"])
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision("fp32").eval()
    b = mk.case_batch(name)
    for t in g["tasks"]:
        model.zero_grad(set_to_none=True)
        vil, _, _, _ = model(b["input_ids"], b["image_feat"], b["image_loc"], t, b["segment_ids"], b["input_mask"], b["image_mask"])
        ref = g["pred"][t]
        got = vil.detach().float().reshape(ref.shape)
        keep = ref > -5000.0
        assert torch.equal(keep, got > -5000.0), (name, t)
        assert float((got[keep] - ref[keep]).norm() / ref[keep].norm().clamp(min=1e-6)) < 2e-4, (name, t)
        (vil.float().reshape(ref.shape) * keep).sum().backward()
        touched = [k for k, p in model.named_parameters() if p.grad is not None and float(p.grad.abs().sum()) > 0]
        assert any(k.startswith("bert.encoder.layer.0.") for k in touched), (name, t)
        assert any(k.startswith("clfs_dict.%s." % t) for k in touched), (name, t, [k for k in touched if "clfs" in k])


@pytest.mark.parametrize("act", ["swish", "relu"])
def test_other_activations_match_oracle(act, monkeypatch):
    """hidden_act / v_hidden_act in {relu, swish} (ACT2FN, volta/encoders.py:157; no shipped config uses them): losses and
    every parameter gradient against the oracle, which applies the same table."""
    emu_ops.install(monkeypatch)
    g, cfg, b = load_case("mini_ctrl_vilbert")
    cfg.hidden_act = cfg.v_hidden_act = act
    model = volta_b200.BertForVLPreTraining(cfg)
    state = seeded_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(state)
    model.set_precision("fp32").eval()
    losses = model(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"],
                   b["lm_label_ids"], b["image_label"], b["image_cls"], None, None, None, None, None, b["is_match"])
    (losses[0] + losses[1] + losses[2]).sum().backward()
    P = O.params_from_state_dict(state, requires_grad=True)
    ref = O.fwd_bwd(P, cfg, b)
    for mine, r in zip(losses, ref):
        assert abs(float(mine) - r) < 2e-4 * max(1.0, abs(r)), (act, float(mine), r)
    for key, p in model.named_parameters():
        rg = P[key].grad if P[key].grad is not None else torch.zeros_like(P[key])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert float((got - rg).norm()) / max(float(rg.norm()), 1e-6) < 2e-3, (act, key)


def _make_ragged(b, seed):
    """Shortens the samples of a (full-length) synthetic batch in place: text lengths 3 ... T (first sample full, second
    the minimum), 1 ... V-1 detected regions behind the global row; labels / features beyond the new lengths are cleared
    the way the loader leaves padding (ids 0, lm label -1, region label -1, zero features)."""
    g = torch.Generator().manual_seed(seed)
    B, T = b["input_ids"].shape
    V = b["image_mask"].shape[1]
    lens = torch.randint(3, T + 1, (B,), generator=g)
    nbox = torch.randint(1, V, (B,), generator=g)
    lens[0], lens[1] = T, 3
    nbox[0], nbox[1] = V - 1, 1
    tpos, vpos = torch.arange(T)[None], torch.arange(V)[None]
    tm = (tpos < lens[:, None]).long()
    vm = (vpos < (nbox + 1)[:, None]).long()            # global row + nbox regions
    b["input_mask"] = tm
    b["input_ids"] = b["input_ids"] * tm
    b["input_ids"][torch.arange(B), lens - 1] = 102 % int(b["input_ids"].max() + 1)
    b["lm_label_ids"] = torch.where(tm.bool(), b["lm_label_ids"], torch.full_like(b["lm_label_ids"], -1))
    b["lm_label_ids"][:, 1] = b["input_ids"][:, 1].clamp(min=1)      # one labelled token per sample survives
    b["image_mask"] = vm
    b["image_feat"] = b["image_feat"] * vm[..., None].float()
    b["image_label"] = torch.where(vm[:, 1:].bool(), b["image_label"], torch.full_like(b["image_label"], -1))
    b["image_label"][:, 0] = 1                                       # one masked region per sample survives
    return b, lens, nbox


@pytest.mark.parametrize("name,shape", [("mini_ctrl_vilbert", (32, 8, 9)), ("mini_ctrl_lxmert", (32, 8, 9)),
                                        ("mini_ctrl_visualbert", (32, 8, 9)), ("mini_ctrl_vilbert", (19, 11, 7))])
def test_packed_rows_plan_on_ragged_batches_matches_oracle(name, shape, monkeypatch):
    """The packed plan where it matters: samples of very different lengths (3 ... T tokens, 1 ... V-1 regions, one sample
    at each extreme), so that the packed buffers hold ~60 % of the rows, whole 256-row groups stay empty and the
    per-sample offsets differ everywhere.  Losses and every parameter gradient against the oracle; the un-packed
    sequence outputs equal the oracle's on the valid rows and are zero on padding."""
    from volta_b200 import engine
    from volta_b200.synthetic import pretraining_batch
    emu_ops.install(monkeypatch)
    taken = []

    def forced(rt, bert, keep_layers):
        ok = not keep_layers
        taken.append(ok)
        return ok
    monkeypatch.setattr(engine, "_pack_supported", forced)
    monkeypatch.setattr(engine, "PLAN_EMULATED", True)
    g, cfg, _, model, state = _build(name)
    b = pretraining_batch(*shape, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=23, full_length=True)
    b, lens, nbox = _make_ragged(b, seed=5)
    losses = model(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"],
                   b["lm_label_ids"], b["image_label"], b["image_cls"], None, None, None, None, None, b["is_match"])
    assert taken and all(taken) and engine.last_plan["packed"]
    (losses[0] + losses[1] + losses[2]).sum().backward()
    P = O.params_from_state_dict(state, requires_grad=True)
    ref = O.pretraining_losses(P, cfg, b)
    for mine, r in zip(losses, ref):
        assert abs(float(mine) - float(r)) < 2e-4 * max(1.0, abs(float(r))), (name, float(mine), float(r))
    sum(ref).backward()
    for key, p in model.named_parameters():
        rg = P[key].grad if P[key].grad is not None else torch.zeros_like(P[key])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        err = float((got - rg).norm()) / max(float(rg.norm()), 1e-6)
        assert err < 2e-3, (name, key, err)
    with torch.no_grad():
        st, sv, pt, pv, _ = model.bert(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"])
        lt_, lv_, rpt, rpv = O.bert_model(O.params_from_state_dict(state), cfg, b["input_ids"], b["image_feat"], b["image_loc"],
                                          b["segment_ids"], b["input_mask"], b["image_mask"])[:4]
        rt_, rv_ = lt_[-1], lv_[-1]
    tm, vm = b["input_mask"].bool(), b["image_mask"].bool()
    assert torch.allclose(st.float()[tm], rt_[tm], atol=2e-4) and torch.allclose(sv.float()[vm], rv_[vm], atol=2e-4)
    assert float(st.float()[~tm].abs().max()) == 0.0 and float(sv.float()[~vm].abs().max()) == 0.0
    assert torch.allclose(pt.float(), rpt, atol=2e-4) and torch.allclose(pv.float(), rpv, atol=2e-4)

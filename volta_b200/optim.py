"""Optimizer step for volta_b200 models: AdamW + global-norm clipping in one pass over the flat buffers.

The reference's step after the hot path (train_concap.py:199-234, 305-313) is
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip); optimizer.step(); scheduler.step(); optimizer.zero_grad()
with `pytorch_transformers.AdamW` over one parameter group PER PARAMETER (own lr: 0.1x for BERT-initialised weights;
weight decay 0 for biases / LayerNorm).  With ~500 parameters that is ~3000 small launches per step - as long as the whole
forward + backward here.  `FusedAdamW` keeps the same constructor contract (a `torch.optim.Optimizer` over the same list
of group dicts, so `WarmupLinearSchedule` / any LR scheduler that edits `group["lr"]` works, and `state_dict()` carries
`step`, `exp_avg`, `exp_avg_sq` per parameter) but executes one kernel, `vb_adamw_flat`, over store.flat32 / grad32
(+ `vb_sumsq` when clipping): the moments live in two flat fp32 buffers of the same layout, and the bf16 tensor-core
copy of the weights is written by the same kernel, so the next forward does not need the fp32 -> bf16 refresh pass.

pytorch_transformers 1.1.0 is not installed (requirements.txt:39): its update rule is restated in
csrc/optim.cu and pinned by tests against a plain PyTorch restatement of the same published algorithm.  `FusedRAdam`
is the optimizer the reference ships itself (volta/optimization.py:9-93) on the same buffers, pinned by trajectories of
the unmodified class (tests/golden/make_golden_radam.py).
"""
import torch

from . import engine, ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 model=None, max_grad_norm=0.0):
        """`params`: parameters or group dicts of ONE volta_b200 model (pass `model=` or let it be found through the
        parameters' flat store).  `max_grad_norm` > 0 folds clip_grad_norm_(model.parameters(), max_grad_norm) into the
        step (the separate torch call would walk 500 gradient views)."""
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {} - should be in [0.0, 1.0[".format(betas))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias))
        if model is None:
            raise ValueError("FusedAdamW needs model= (the volta_b200 model whose flat store holds these parameters)")
        self.model = model
        self._writes_bf16_operands = True     # store.py: this optimizer's step leaves the bf16 operand copy fresh
        self.max_grad_norm = float(max_grad_norm)
        b = {tuple(g["betas"]) for g in self.param_groups}
        e = {float(g["eps"]) for g in self.param_groups}
        if len(b) != 1 or len(e) != 1:
            raise ValueError("FusedAdamW: betas and eps must be the same for every group (lr and weight_decay may differ)")
        self._layout = None

    # ------------------------------------------------------------------------------------------ flat layout
    def _ensure(self):
        store = engine.model_store(self.model)
        dev = next(self.model.parameters()).device
        store.ensure(dev, self.model.precision)
        key = (store.flat32.data_ptr(), store.total, len(self.param_groups))
        if self._layout is not None and self._layout["key"] == key:
            return store
        if store.total % 64:
            raise RuntimeError("flat store is not 64-element aligned")
        chunk_group = torch.full((store.total // 64,), -1, dtype=torch.int16)   # 0xFFFF = not optimised (frozen / absent)
        if len(self.param_groups) >= 0xFFFF:
            raise ValueError("too many parameter groups")
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if id(p) not in store.offsets:
                    raise ValueError("FusedAdamW: a parameter does not belong to the model's flat store")
                off, n = store.offsets[id(p)]
                c0, c1 = off // 64, (off + n + 63) // 64
                if (chunk_group[c0:c1] != -1).any() and not (chunk_group[c0:c1] == gi).all():
                    raise ValueError("FusedAdamW: two hyper-parameter groups share a 64-element chunk of the flat store "
                                     "(parameters of one fused group - e.g. q/k/v weights - must share lr and weight decay)")
                chunk_group[c0:c1] = gi
        m = torch.zeros(store.total, dtype=torch.float32, device=dev)
        v = torch.zeros(store.total, dtype=torch.float32, device=dev)
        old = self._layout
        self._layout = dict(key=key, chunk_group=chunk_group.to(dev), chunk_group_host=chunk_group, none_sig=None, active=None,
                            m=m, v=v, sumsq=torch.zeros(1, dtype=torch.float32, device=dev), steps=[0] * len(self.param_groups))
        for group in self.param_groups:          # per-parameter state as views, for state_dict() compatibility
            for p in group["params"]:
                off, n = store.offsets[id(p)]
                st = self.state[p]
                prev_m, prev_v = st.get("exp_avg"), st.get("exp_avg_sq")
                st["exp_avg"] = m[off:off + n].view(p.shape)
                st["exp_avg_sq"] = v[off:off + n].view(p.shape)
                if prev_m is not None:            # re-layout or load_state_dict: keep the moments
                    st["exp_avg"].copy_(prev_m)
                    st["exp_avg_sq"].copy_(prev_v)
                st.setdefault("step", 0)
        if old is not None:
            self._layout["steps"] = old["steps"] + [0] * (len(self.param_groups) - len(old["steps"]))
        return store

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._layout = None          # moments were replaced by loaded tensors: re-attach them to the flat buffers
        store = self._ensure()
        for gi, group in enumerate(self.param_groups):
            steps = [int(self.state[p].get("step", 0)) for p in group["params"]]
            self._layout["steps"][gi] = max(steps) if steps else 0
        return store is not None

    # ------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        store = self._ensure()
        lay = self._layout
        # Parameters without a gradient are skipped, as torch optimizers do (`if p.grad is None: continue`: task heads a
        # step did not use keep their weights, moments and step count).  Gradients that autograd produced outside the flat
        # buffer (a head whose encoder backward did not run) are adopted first.
        base = None if store.grad32 is None else store.grad32.data_ptr()
        none_sig, foreign = [], False
        for group in self.param_groups:
            for p in group["params"]:
                g = p.grad
                none_sig.append(g is None)
                if g is not None and (base is None or g.data_ptr() != base + 4 * store.offsets[id(p)][0]):
                    foreign = True
        if all(none_sig):
            return loss  # no backward has run since the gradients were cleared: nothing to do, as in torch
        if foreign:
            store.prepare_grads()
        none_sig = tuple(none_sig)
        if none_sig != lay["none_sig"]:
            lay["none_sig"] = none_sig
            if any(none_sig):
                active = lay["chunk_group_host"].clone()
                it = iter(none_sig)
                for group in self.param_groups:
                    for p in group["params"]:
                        if next(it):
                            off, n = store.offsets[id(p)]
                            active[off // 64:(off + n + 63) // 64] = -1
                lay["active"] = active.to(store.flat32.device)
            else:
                lay["active"] = lay["chunk_group"]
        beta1, beta2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        rows = []
        it = iter(none_sig)
        for gi, group in enumerate(self.param_groups):
            live = [not next(it) for _ in group["params"]]
            if not any(live):
                rows.append((0.0, 0.0, 0.0, 1.0))   # no parameter of this group has a gradient: its chunks are masked out
                continue
            lay["steps"][gi] += 1
            t = lay["steps"][gi]
            rows.append(self._group_row(group, t, beta1, beta2))
            for p, on in zip(group["params"], live):
                if on:
                    self.state[p]["step"] = t
        hp = torch.tensor(rows, dtype=torch.float32)
        hp_dev = hp.to(store.flat32.device, non_blocking=True)
        sumsq = None
        if self.max_grad_norm > 0.0:
            sumsq = lay["sumsq"]
            sumsq.zero_()
            ops.sumsq(store.grad32, sumsq)
        p16 = store.flat16 if store.precision == "bf16" else None
        ops.adamw_flat(store.flat32, store.grad32, lay["m"], lay["v"], p16, lay["active"], hp_dev, beta1, beta2, eps,
                       sumsq, self.max_grad_norm, rule=self._rule)
        engine._count(2 if sumsq is not None else 1)
        return loss

    _rule = "adamw"

    def _group_row(self, group, t, beta1, beta2):
        """(lr, weight_decay, step_size, active) of vb_adamw_flat for a group at its step t."""
        lr = float(group["lr"])
        step_size = lr
        if group.get("correct_bias", True):
            step_size = lr * (1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t)
        return (lr, float(group["weight_decay"]), step_size, 1.0)

    def grad_norm(self):
        """Global gradient norm of the last clipped step (device tensor; reading it synchronises)."""
        return None if self._layout is None else self._layout["sumsq"].sqrt()


class FusedRAdam(FusedAdamW):
    """The reference's own `RAdam` (volta/optimization.py:9-93; `train_task.py --optim RAdam`, :227-228) over the flat
    buffers: same constructor contract (`RAdam(optimizer_grouped_parameters, lr=base_lr)`, defaults betas (0.9, 0.999),
    eps 1e-8, weight_decay 0), same `state_dict()` entries (`step`, `exp_avg`, `exp_avg_sq`), one `vb_radam_flat` launch.

    The host forms what the reference forms per parameter: the length of the approximated SMA N_sma(t) and the step size
    (:52-81), weight decay `p += -wd * lr * p` BEFORE the update (:83-84).  It also keeps the reference's ten-slot
    `buffer` (:12, 52-81): a step size is computed once per step count `t` - with the learning rate of the FIRST group
    that reaches `t` - and reused by every later group at the same `t`, whatever its own learning rate (groups with
    different learning rates therefore share the first one's step size; their weight decay uses their own).  That is
    the reference's behaviour and the golden vectors (tests/golden/make_golden_radam.py, made by running the unmodified
    class) contain it."""

    _rule = "radam"

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, model=None, max_grad_norm=0.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, model=model,
                         max_grad_norm=max_grad_norm)
        for g in self.param_groups:
            g.pop("correct_bias", None)         # the reference's defaults: lr, betas, eps, weight_decay
        self.defaults.pop("correct_bias", None)
        self.buffer = [[None, None, None] for _ in range(10)]

    def _group_row(self, group, t, beta1, beta2):
        buffered = self.buffer[int(t % 10)]
        if t == buffered[0]:
            n_sma, step_size = buffered[1], buffered[2]
        else:
            buffered[0] = t
            beta2_t = beta2 ** t
            n_sma_max = 2 / (1 - beta2) - 1
            n_sma = n_sma_max - 2 * t * beta2_t / (1 - beta2_t)
            buffered[1] = n_sma
            if n_sma >= 5:
                step_size = (group["lr"] * ((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max /
                                            (n_sma_max - 2)) ** 0.5 / (1 - beta1 ** t))
            else:
                step_size = group["lr"] / (1 - beta1 ** t)
            buffered[2] = step_size
        return (float(group["weight_decay"]) * float(group["lr"]), 1.0 if n_sma >= 5 else 0.0, float(step_size), 1.0)


def clip_grad_norm_(model, max_norm):
    """torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) on the flat gradient buffer: one reduction kernel +
    one scaling kernel instead of a walk over ~500 gradient views.  Returns the total norm (device tensor)."""
    store = engine.model_store(model)
    if store.grad32 is None:
        return torch.zeros((), device=next(model.parameters()).device)
    acc = torch.zeros(1, dtype=torch.float32, device=store.grad32.device)
    ops.sumsq(store.grad32, acc)
    norm = acc.sqrt()
    coef = (max_norm / (norm + 1e-6)).clamp(max=1.0)
    store.grad32.mul_(coef)   # torch op on purpose: a scalar-tensor multiply needs no host sync
    return norm.reshape(())

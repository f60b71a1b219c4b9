"""Flat parameter / gradient storage in HBM.

All parameters of a model live in ONE fp32 buffer (`flat32`), their gradients in one fp32 buffer of the
same layout (`grad32`), and - in bf16 mode - the tensor-core operand copies in one bf16 buffer
(`flat16`).  The `nn.Parameter`s the reference API exposes (state_dict names, optimizer param groups)
are views into `flat32`; `p.grad` are views into `grad32`.

Why: (1) Q/K/V weights of one attention module are laid out back to back, so the three Linears of the
reference (volta/encoders.py:204-206) run as one N = 3*A GEMM and one wgrad without any packing copy;
(2) the fp32 -> bf16 operand refresh after an optimizer step is one kernel over one buffer; (3) the
data-parallel all-reduce (reference: apex/apex/parallel/distributed.py:425-475, flatten -> all_reduce
-> unflatten) needs no flatten/unflatten passes - buckets are plain slices of `grad32`.
"""
import torch

from . import ops

# ---- staleness of the bf16 operand copy ------------------------------------------------------------------------------
# `p._version` sees in-place updates made through the parameter itself (`with no_grad(): p.add_()`), NOT updates made
# through `p.data` - which is how the reference's optimizers write (pytorch_transformers.AdamW: `p.data.addcdiv_`;
# volta/optimization.py RAdam: `p.data.copy_`, call sites train_task.py:222-234).  Every `torch.optim.Optimizer.step`
# therefore bumps a global epoch through torch's optimizer post-step hook, and a store re-casts its operands when the
# epoch moved since its last cast.  FusedAdamW writes the bf16 copy itself and does not bump it.  Edits made neither
# through the parameter nor by an optimizer (`p.data.mul_()` by hand) need `store.refresh()`.
_OPT_EPOCH = [0]


def _on_optimizer_step(optimizer, args, kwargs):
    if not getattr(optimizer, "_writes_bf16_operands", False):
        _OPT_EPOCH[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(_on_optimizer_step)
except ImportError:  # pragma: no cover - torch < 2.0: fall back to re-casting after every backward
    register_optimizer_step_post_hook = None

_ZERO_STREAMS = {}


def _zero_stream(device):
    key = str(device)
    if key not in _ZERO_STREAMS:
        _ZERO_STREAMS[key] = torch.cuda.Stream(device=device)
    return _ZERO_STREAMS[key]


ALIGN = 64  # elements: every group starts on a 256-byte boundary (TMA wants 16 B, 128 B keeps sectors whole)


class ParamStore:
    def __init__(self, model):
        self.model = model
        self.flat32 = None
        self.flat16 = None
        self.grad32 = None
        self.offsets = {}      # id(param) -> (offset, numel)
        self.params = []       # unique parameters in layout order
        self.total = 0
        self._sig = None
        self.precision = None
        self._backwards = 0    # backwards run on this store (the fallback staleness signal without the optimizer hook)
        self._trash = None     # where the gradients of frozen parameters are written (never read)
        self._prezero = None   # event: grad32 was zeroed on the side stream during the forward (prezero_grads)

    # ---- layout -------------------------------------------------------------------------
    def _layout(self):
        groups = self.model._fused_groups() if hasattr(self.model, "_fused_groups") else []
        group_of = {}
        for g in groups:
            for p in g:
                group_of[id(p)] = g
        seen, order = set(), []
        for p in self.model.parameters():
            if id(p) in seen:
                continue
            members = group_of.get(id(p), [p])
            if any(id(m) in seen for m in members):
                raise RuntimeError("fused parameter group overlaps another group")
            order.append(members)
            seen.update(id(m) for m in members)
        off = 0
        self.offsets, self.params = {}, []
        self.groups = order
        for members in order:
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            for p in members:
                self.offsets[id(p)] = (off, p.numel())
                self.params.append(p)
                off += p.numel()
        self.total = (off + ALIGN - 1) // ALIGN * ALIGN

    def _is_flat(self, device):
        if self.flat32 is None or self.flat32.device != device:
            return False
        base = self.flat32.data_ptr()
        for p in self.params:
            off, _ = self.offsets[id(p)]
            if p.data_ptr() != base + 4 * off:
                return False
        return True

    def ensure(self, device, precision):
        """Make sure parameters are flat on `device` and the compute copy is fresh."""
        if not self._is_flat(device) or len(self.params) != len({id(p) for p in self.model.parameters()}):
            self._layout()
            flat = torch.zeros(self.total, dtype=torch.float32, device=device)
            for p in self.params:
                off, n = self.offsets[id(p)]
                flat[off:off + n].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            for p in self.params:
                off, n = self.offsets[id(p)]
                p.data = flat[off:off + n].view(p.shape)
                # gradients that exist already (an optimizer step before the first forward, accumulation across a
                # re-layout) stay: they are foreign tensors now and prepare_grads() adopts them into the new grad32
                if p.grad is not None and p.grad.device != flat.device:
                    p.grad = None
            self.flat32 = flat
            self.grad32 = None
            self.flat16 = None
            self._sig = None
        if precision != self.precision:
            self.precision, self._sig = precision, None
        if precision == "bf16":
            sig = (sum(p._version for p in self.params), _OPT_EPOCH[0],
                   self._backwards if register_optimizer_step_post_hook is None else 0)
            if self.flat16 is None:
                self.flat16 = torch.empty(self.total, dtype=torch.bfloat16, device=device)
                self._sig = None
            if sig != self._sig:
                ops.cast_bf16(self.flat32, self.flat16)
                self._sig = sig

    def refresh(self):
        """Force the bf16 operand copy to be rebuilt at the next forward (after out-of-band weight edits)."""
        self._sig = None

    # ---- views --------------------------------------------------------------------------
    def _view(self, buf, ps):
        off0, n0 = self.offsets[id(ps[0])]
        end = off0
        for p in ps:
            off, n = self.offsets[id(p)]
            if off != end:
                raise RuntimeError("parameters are not contiguous in the flat buffer")
            end = off + n
        lead = sum(p.shape[0] for p in ps)
        return buf[off0:end].view((lead,) + tuple(ps[0].shape[1:]))

    def cw(self, *ps):
        """Compute-dtype view (bf16 copy in bf16 mode, the fp32 master otherwise) of one parameter or of
        several consecutive ones stacked along dim 0."""
        return self._view(self.flat16 if self.precision == "bf16" else self.flat32, ps)

    def w32(self, *ps):
        return self._view(self.flat32, ps)

    def g32(self, *ps):
        """Gradient view of one parameter / several consecutive ones.  Frozen parameters (requires_grad False:
        config.fixed_layers, volta/train_utils.py:250-255 freeze_layers) get a scratch view instead: the kernels still
        write, nothing reads, and their slice of grad32 stays zero - so the flat gradient norm (optim.py) and the
        all-reduced buckets hold exactly what `clip_grad_norm_(model.parameters())` / apex DDP see in the reference."""
        if not any(p.requires_grad for p in ps):
            n = sum(p.numel() for p in ps)
            if self._trash is None or self._trash.numel() < n:   # sized in prepare_grads; only reached by out-of-band freezes
                self._trash = torch.empty(n, dtype=torch.float32, device=self.flat32.device)
            lead = sum(p.shape[0] for p in ps)
            return self._trash[:n].view((lead,) + tuple(ps[0].shape[1:]))
        return self._view(self.grad32, ps)

    # ---- gradients ----------------------------------------------------------------------
    def prezero_grads(self):
        """Forward-time hint of a training step: when no parameter holds a gradient (zero_grad(set_to_none=True), the
        reference's loop - train_concap.py:313), the backward would begin with the memset of the whole flat gradient
        buffer (968 MB for ctrl_vilbert: ~0.15 ms at HBM speed, exposed).  Issue it now on a side stream, ordered after
        everything already queued (the optimizer step that read the previous gradients), so it runs under the forward;
        prepare_grads waits on its event instead of filling."""
        if self._prezero is not None or self.flat32 is None or not self.flat32.is_cuda:
            return
        for p in self.params:
            if p.grad is not None:
                return             # gradients are being accumulated (or were produced outside): prepare_grads decides
        if self.grad32 is None:
            self.grad32 = torch.empty(self.total, dtype=torch.float32, device=self.flat32.device)
        main = torch.cuda.current_stream(self.flat32.device)
        side = _zero_stream(self.flat32.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.fill(self.grad32, 0)
            self._prezero = torch.cuda.Event()
            self._prezero.record(side)

    def prepare_grads(self):
        """Attach `p.grad` views into grad32; zero what is not being accumulated into."""
        if self.grad32 is None:
            self.grad32 = torch.empty(self.total, dtype=torch.float32, device=self.flat32.device)
        self._backwards += 1
        frozen = [sum(p.numel() for p in g) for g in self.groups if not any(p.requires_grad for p in g)]
        if frozen and (self._trash is None or self._trash.numel() < max(frozen)):
            self._trash = torch.empty(max(frozen), dtype=torch.float32, device=self.flat32.device)
        written = getattr(self.model, "_engine_written", None)
        written = None if written is None else {id(p) for p in written()}
        base = self.grad32.data_ptr()
        ours, missing, foreign = [], [], []
        for p in self.params:
            off, n = self.offsets[id(p)]
            g = p.grad
            if g is None:
                missing.append(p)
            elif g.data_ptr() == base + 4 * off:
                ours.append(p)
            else:
                foreign.append(p)
        prezero, self._prezero = self._prezero, None
        if not ours:
            if prezero is not None:   # already zeroed under the forward (prezero_grads)
                torch.cuda.current_stream(self.grad32.device).wait_event(prezero)
            else:
                ops.fill(self.grad32, 0)  # start of a step: one memset for every gradient
            zero_each = False
        else:
            zero_each = True          # accumulating into existing grads: only touch what is missing
        for p in missing:
            off, n = self.offsets[id(p)]
            view = self.grad32[off:off + n].view(p.shape)
            if zero_each:
                view.zero_()
            # parameters outside the engine's backward (task heads a step does not use) keep grad None, as in the reference
            if p.requires_grad and (written is None or id(p) in written):
                p.grad = view
        for p in foreign:
            off, n = self.offsets[id(p)]
            view = self.grad32[off:off + n].view(p.shape)
            view.copy_(p.grad)
            p.grad = view

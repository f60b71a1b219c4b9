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
                p.grad = None
            self.flat32 = flat
            self.grad32 = None
            self.flat16 = None
            self._sig = None
        if precision != self.precision:
            self.precision, self._sig = precision, None
        if precision == "bf16":
            sig = sum(p._version for p in self.params)
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
        return self._view(self.grad32, ps)

    # ---- gradients ----------------------------------------------------------------------
    def prepare_grads(self):
        """Attach `p.grad` views into grad32; zero what is not being accumulated into."""
        if self.grad32 is None:
            self.grad32 = torch.empty(self.total, dtype=torch.float32, device=self.flat32.device)
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
        if not ours:
            ops.fill(self.grad32, 0)  # start of a step: one memset for every gradient
            zero_each = False
        else:
            zero_each = True          # accumulating into existing grads: only touch what is missing
        for p in missing:
            off, n = self.offsets[id(p)]
            view = self.grad32[off:off + n].view(p.shape)
            if zero_each:
                view.zero_()
            if p.requires_grad:
                p.grad = view
        for p in foreign:
            off, n = self.offsets[id(p)]
            view = self.grad32[off:off + n].view(p.shape)
            view.copy_(p.grad)
            p.grad = view

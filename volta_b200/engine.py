"""Execution engine: runs the VOLTA encoder forward and backward as an explicit plan of fused sm_100a kernels.

The reference executes ~3000 ATen ops per forward through nn.Module.__call__ and autograd
(SURVEY.md 2.2).  Here one `torch.autograd.Function` node covers the whole model: its forward walks the
config's sub-layer schedule and launches the fused kernels of libvolta_b200.so (include/volta_b200.h),
recording what each block needs for its hand-written backward; its backward walks the record in reverse
and writes parameter gradients straight into the flat fp32 gradient buffer (store.py).  PyTorch supplies
device memory, the current stream and the autograd hook - no arithmetic.

Data layout in HBM: hidden states are row-major [rows, H] in the compute dtype (bf16, or fp32 in the
validation mode); when both streams have the same width they share ONE buffer [B*T + B*V, H] (text rows
first) so that sub-layers whose weights are shared between the streams (`shared_sublayers`) run as a
single GEMM over all rows - a single-stream model (VisualBERT / UNITER / VL-BERT) is then an ordinary BERT
layer over T+V tokens, which is what the reference's four-block gated attention with joint softmax
computes (volta/encoders.py:285-306).
"""
import math
import os

import torch

from . import _lib, ops
from ._lib import (EPI_ACCUM, EPI_ADD, EPI_BIAS, EPI_BIAS_ADD, EPI_BIAS_DROP_RES, EPI_BIAS_GELU, EPI_BIAS_RELU,
                   EPI_DRELU, EPI_MUL, EPI_NONE, ELT_ADD, ELT_COPY, ELT_DGELU, ELT_DRELU, ELT_DROPOUT, ELT_GELU,
                   ELT_MUL, ELT_RELU)
from .losses import pre_vis_targets
from .store import ParamStore

LN_EPS = 1e-12

# counters the benchmark reads: kernels launched through the C ABI in this process
launch_count = 0


def _count(n=1):
    global launch_count
    launch_count += n


# What the last forward's launch plan did (tests assert that the plan they mean to check was the one taken):
# packed = encoder ran on packed rows; row_split = sub-layers whose two streams shared row-split GEMM launches;
# joint = shared sub-layers run as one GEMM over both streams; mlm_compact = masked-LM head on the labelled rows only.
last_plan = {}


def _pad8(n):
    return (n + 7) // 8 * 8


class Pair:
    """Hidden states of both streams.  `all` is the joint [Mt+Mv, H] buffer when the widths match."""

    def __init__(self, t, v, joint=None):
        self.t, self.v, self.all = t, v, joint

    @staticmethod
    def alloc(rt, Mt, Ht, Mv, Hv, dtype=None):
        if Ht == Hv:
            joint = rt.empty(Mt + Mv, Ht, dtype=dtype)
            return Pair(joint[:Mt], joint[Mt:], joint)
        return Pair(rt.empty(Mt, Ht, dtype=dtype), rt.empty(Mv, Hv, dtype=dtype), None)


class Runtime:
    """Per-call context: shapes, compute dtype, dropout seeds, parameter views."""

    def __init__(self, model, device, B, T, V, training, grad=False, seed_base=None):
        self.model = model
        self.store = model_store(model)
        self.precision = self.store.model.precision
        self.cdtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        self.device = device
        self.B, self.T, self.V = B, T, V
        self.Mt, self.Mv = B * T, B * V
        self.training = training
        self.grad_enabled = grad  # autograd.Function.forward runs with grad mode off: the caller decides
        # CPU generator: no device sync.  Under graph capture the per-step seed lives on the device (graphs.py).
        if seed_base is not None:
            self._seed = seed_base
        else:
            self._seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
        self.quiet = False               # graphs.py: no progress notifications while the forward is captured
        self.progress_override = None    # graphs.py: receives the notifications while the backward is captured
        self._site = 0
        self.gemm_backend = {"auto": _lib.GEMM_AUTO, "tc": _lib.GEMM_TC, "simt": _lib.GEMM_SIMT}[
            os.environ.get("VOLTA_B200_GEMM", "auto")]
        self.attn_backend = {"auto": _lib.ATTN_AUTO, "mma": _lib.ATTN_MMA, "simt": _lib.ATTN_SIMT, "tc": _lib.ATTN_TC}[
            os.environ.get("VOLTA_B200_ATTN", "auto")]
        if self.precision == "fp32":
            self.gemm_backend, self.attn_backend = _lib.GEMM_SIMT, _lib.ATTN_SIMT
        self.store.ensure(device, self.precision)
        # (valid text rows, valid vision rows) of a recent batch of this shape - see graphs.py; steers GEMM tile choices only
        h = getattr(self.store, "row_hint", None)
        self.row_hint = tuple(h[3:]) if (h is not None and tuple(h[:3]) == (B, T, V)) else None

    def empty(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.cdtype, device=self.device)

    def zeros(self, *shape, dtype=None):
        t = torch.empty(shape, dtype=dtype or self.cdtype, device=self.device)
        nbytes = t.numel() * t.element_size()
        if nbytes % 16 == 0 and t.data_ptr() % 16 == 0:
            ops.fill(t, 0)
            _count()
        else:
            t.zero_()
        return t

    def seed(self):
        """A fresh Philox seed per dropout site (masks are regenerated in the backward, never stored)."""
        self._site += 1
        return (self._seed + 0x9E3779B97F4A7C15 * self._site) & 0xFFFFFFFFFFFFFFFF  # uint64 wrap, as on the device

    def p(self, prob):
        return float(prob) if self.training else 0.0

    # parameter views
    def cw(self, *ps):
        return self.store.cw(*ps)

    def w32(self, *ps):
        return self.store.w32(*ps)

    def g32(self, *ps):
        return self.store.g32(*ps)

    # kernels with launch accounting
    def seqs(self, st):
        """vb_row_seqs tuple for the rows of stream st ('t', 'v', 'all' = text rows then vision rows), or None."""
        if st is not None and getattr(self, "pack_on", False):
            # packed rows: each stream is ONE run of `count` valid rows followed by unused rows
            pk = self.pack
            if st == "t":
                return (pk.lim_t[0:1], self.Mt, self.Mt, None, 0)
            if st == "v":
                return (pk.lim_v[0:1], self.Mv, self.Mv, None, 0)
            return (pk.lim_t[0:1], self.Mt, self.Mt, pk.lim_v[0:1], self.Mv)
        lens = getattr(self, "lens", (None, None))
        if st is None or lens[0] is None or lens[1] is None or not SKIP_PADDING_LN:
            return None
        if st == "t":
            return (lens[0], self.T, self.Mt, None, 0)
        if st == "v":
            return (lens[1], self.V, self.Mv, None, 0)
        return (lens[0], self.T, self.Mt, lens[1], self.V)

    def row_limit(self, rows, padded=True):
        """Packed rows: the device-side row count (int32[1]) of the stream whose buffer has `rows` rows, else None."""
        if not getattr(self, "pack_on", False):
            return None
        pk = self.pack
        lim = pk.lim_t if rows == self.Mt else pk.lim_v if rows == self.Mv else None
        assert lim is not None, "packed rows: a buffer of %d rows belongs to neither stream" % rows
        return lim[1:2] if padded else lim[0:1]

    def gemm(self, a, b, out, **k):
        _count()
        k.setdefault("backend", self.gemm_backend)
        if getattr(self, "pack_on", False):
            # Packed rows (bert_fwd): only the first `count` rows of each stream exist; the GEMM schedules whole tiles /
            # k-blocks up to the count, the count stays on the device.
            if k.get("epilogue") == EPI_ACCUM:          # weight gradient: the reduction runs over the rows
                k["k_limit"] = self.row_limit(a.shape[0])
            elif a.shape[0] == self.Mt + self.Mv:       # joint buffer: text segment, then vision segment
                k.update(m_limit=self.pack.lim_t[0:1], m_limit2=self.pack.lim_v[0:1], limit_seg_M=self.Mt)
                if self.row_hint:
                    k["m_hint"] = _pad256(self.row_hint[0]) + _pad256(self.row_hint[1])
            else:
                k["m_limit"] = self.row_limit(a.shape[0], padded=False)
                if self.row_hint:
                    k["m_hint"] = _pad256(self.row_hint[0] if a.shape[0] == self.Mt else self.row_hint[1])
        return ops.gemm(a, b, out, **k)


def model_store(model):
    parent = model.__dict__.get("_engine_parent") if hasattr(model, "__dict__") else None
    if parent is not None and parent() is not None:
        return model_store(parent())
    st = getattr(model, "_engine_state", None)
    if st is None:
        st = ParamStore(model)
        model._engine_state = st
    return st


# ====================================================================================================
# building blocks (forward returns outputs + a `saved` record; backward consumes it)
# ====================================================================================================
# Masked-LM decoder on the labelled rows only (heads_fwd); VOLTA_B200_MLM_COMPACT=0 restores the all-positions decoder.
MLM_COMPACT = os.environ.get("VOLTA_B200_MLM_COMPACT", "1") != "0"
# The same for the masked-region head (label == 1 rows).  Measured on B200 (A/B on one box, 30 steps each): 12.40 / 12.45
# ms with it, 12.41 / 12.47 ms without - the head is small (0.5 % of the FLOPs) and the plan / gather / scatter launches
# cost what the skipped rows save.  Off by default; VOLTA_B200_IMG_COMPACT=1 turns it on (covered by the CPU tests).
IMG_COMPACT = os.environ.get("VOLTA_B200_IMG_COMPACT", "0") == "1"
# Attention skips 16-row tiles of trailing padding (vb_attn_args::len_*); VOLTA_B200_SKIP_PADDING=0 computes them.
SKIP_PADDING = os.environ.get("VOLTA_B200_SKIP_PADDING", "1") != "0"
SKIP_PADDING_LN = os.environ.get("VOLTA_B200_SKIP_PADDING_LN", "1") != "0"   # the residual-block LayerNorms skip them too
# The weight gradients of the two streams of a sub-layer run as one launch with two reduction segments (linear_bwd)
WGRAD_GROUP = os.environ.get("VOLTA_B200_WGRAD_GROUP", "1") != "0"
# Weight-gradient GEMM on the side stream next to the input-gradient GEMM of the same Linear: -1 = auto (on below 2560
# rows, where neither GEMM fills the machine: 5.53 -> 5.41 ms/step at 32 samples per GPU; off above: 12.32 -> 12.38 ms at
# 256), 0 = never, 1 = issued before the dgrad, 2 = after it
WGRAD_SIDE = int(os.environ.get("VOLTA_B200_WGRAD_SIDE", "-1"))
PLAN_EMULATED = False   # tests/test_engine_emulated.py: take the tensor-core launch plan (row-split / grouped launches) under the CPU emulation of the C ABI
# The memset of the flat gradient buffer runs on a side stream under the forward (store.prezero_grads)
PREZERO_GRADS = os.environ.get("VOLTA_B200_PREZERO_GRADS", "1") != "0"
_SIDE_STREAMS = {}
_SIDE_STREAMS_ON = os.environ.get("VOLTA_B200_SIDE_STREAM", "1") != "0"


def _side_stream(device):
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def _lins(lin):
    return lin if isinstance(lin, (list, tuple)) else [lin]


def linear_fwd(rt, x, lin, epilogue=EPI_BIAS, out=None, lin2=None, split=0, **kw):
    """y = x W^T + b through vb_gemm; W may be several stacked Linears (fused QKV).
    lin2 / split: row-split launch - rows [0, split) use `lin`, rows [split, M) use `lin2` (the text and the vision
    rows of a two-stream sub-layer that share one hidden-state buffer)."""
    lins = _lins(lin)
    W = rt.cw(*[l.weight for l in lins])
    b = rt.w32(*[l.bias for l in lins]) if lins[0].bias is not None else None
    rows = x.shape[0]
    if out is None:
        out = rt.empty(rows, W.shape[0])
    if b is None and epilogue == EPI_BIAS:
        epilogue = EPI_NONE
    if lin2 is not None:
        l2 = _lins(lin2)
        kw.update(b2=rt.cw(*[l.weight for l in l2]), bias2=rt.w32(*[l.bias for l in l2]) if b is not None else None, split_M=split)
    rt.gemm(x, W, out, epilogue=epilogue, bias=b, **kw)
    return out


def linear_bwd(rt, x, dy, lin, need_dx=True, dx_epilogue=EPI_NONE, dx_out=None, bias_done=False, lin2=None, split=0, lim=None, **kw):
    """dW += dy^T x (fp32, accumulated into the flat gradient buffer), db += colsum(dy), dx = dy W (+epilogue).
    lin2 / split: the rows [split, M) belong to a second Linear - two weight-gradient launches over the two row ranges,
    ONE row-split launch for dx."""
    parts = [(lin, x, dy)] if lin2 is None else [(lin, x[:split], dy[:split]), (lin2, x[split:], dy[split:])]
    if lin2 is None and getattr(rt, "pack_on", False) and dy.shape[0] == rt.Mt + rt.Mv and lim is None:
        # packed rows, shared weights over the joint buffer: each stream is its own run of valid rows, so the reduction
        # of the weight gradient (and of the bias gradient) runs per stream, both accumulating into the same parameter
        parts = [(lin, x[:rt.Mt], dy[:rt.Mt]), (lin, x[rt.Mt:], dy[rt.Mt:])]
    # lim = int32 [count, count padded to 256] on the device (compacted head rows): only that many rows hold data
    wl = dict(k_limit=lim[1:2]) if lim is not None else {}
    rl = dict(row_limit=lim[1:2]) if lim is not None else {}
    if lim is not None:
        kw["m_limit"] = lim[0:1]
    # The bias gradients (column sums of dy) run on a side stream next to the weight-gradient / input-gradient GEMMs of
    # the same Linear: the wgrad kernel leaves registers and shared memory for one small block per SM, and both read dy
    # from L2.  The side stream joins before this function returns (dy may be freed by the caller afterwards).
    side = None
    big = dy.shape[0] >= 1024 and _SIDE_STREAMS_ON and dy.is_cuda
    if _lins(lin)[0].bias is not None and not bias_done and big:
        side = _side_stream(rt.device)
        side.wait_stream(torch.cuda.current_stream(rt.device))
        with torch.cuda.stream(side):
            for l_, x_, dy_ in parts:
                ops.colsum(dy_, rt.g32(*[l.bias for l in _lins(l_)]), **(rl or dict(row_limit=rt.row_limit(dy_.shape[0]))))
                _count()
    # the weight-gradient GEMM on the side stream as well (WGRAD_SIDE): its CTAs and the input-gradient GEMM's fill each
    # other's prologue / tail when neither fills the machine
    wside = (WGRAD_SIDE if WGRAD_SIDE >= 0 else (2 if dy.shape[0] < 2560 else 0)) if (big and need_dx) else 0
    if wside and side is None:
        side = _side_stream(rt.device)
        side.wait_stream(torch.cuda.current_stream(rt.device))
    # Both row ranges of a joint buffer in ONE weight-gradient launch (vb_gemm_args::k_seg / D2): the reduction over the
    # text rows accumulates into the first Linear's gradient, the reduction over the vision rows into the second's (the
    # same one for shared weights) - half the launches, and the slices of both fill the machine together.
    tc_ok = (rt.precision == "bf16" and rt.gemm_backend != _lib.GEMM_SIMT and dy.is_cuda) or PLAN_EMULATED
    grouped = (len(parts) == 2 and WGRAD_GROUP and lim is None and tc_ok and parts[0][2].shape[0] % 64 == 0 and
               all(l.weight.requires_grad for pt in parts for l in _lins(pt[0])))
    def weight_grads():
        if grouped:
            seg = parts[0][2].shape[0]
            kl = {}
            if getattr(rt, "pack_on", False):
                kl = dict(k_limit=rt.row_limit(seg), k_limit2=rt.row_limit(dy.shape[0] - seg))
            _count()
            ops.gemm(dy, x, rt.g32(*[l.weight for l in _lins(parts[0][0])]), out2=rt.g32(*[l.weight for l in _lins(parts[1][0])]),
                     k_seg=seg, a_kmajor=False, b_kmajor=False, epilogue=EPI_ACCUM, backend=rt.gemm_backend, **kl)
        for l_, x_, dy_ in parts:
            lins = _lins(l_)
            if not grouped and any(l.weight.requires_grad for l in lins):   # frozen Linears (fixed_layers / freeze_layers): no weight gradient
                rt.gemm(dy_, x_, rt.g32(*[l.weight for l in lins]), a_kmajor=False, b_kmajor=False, epilogue=EPI_ACCUM, **wl)
            if lins[0].bias is not None and not bias_done and side is None:
                ops.colsum(dy_, rt.g32(*[l.bias for l in lins]), **(rl or dict(row_limit=rt.row_limit(dy_.shape[0]))))
                _count()
    if wside == 1:
        with torch.cuda.stream(side):
            weight_grads()
    elif wside == 0:
        weight_grads()
    if not need_dx:
        if side is not None:
            torch.cuda.current_stream(rt.device).wait_stream(side)
        return None
    lins = _lins(lin)
    W = rt.cw(*[l.weight for l in lins])
    if dx_out is None:
        dx_out = rt.empty(dy.shape[0], W.shape[1])
    if lin2 is not None:
        kw.update(b2=rt.cw(*[l.weight for l in _lins(lin2)]), split_M=split)
    rt.gemm(dy, W, dx_out, a_kmajor=True, b_kmajor=False, epilogue=dx_epilogue, **kw)
    if wside == 2:
        with torch.cuda.stream(side):
            weight_grads()
    if side is not None:
        torch.cuda.current_stream(rt.device).wait_stream(side)
    return dx_out


def ln_fwd(rt, x, ln, y=None, out_scale=1.0, accum=False, save=True, seqs=None):
    y = rt.empty(*x.shape) if y is None else y
    mean = rt.empty(x.shape[0], dtype=torch.float32) if save else None
    rstd = rt.empty(x.shape[0], dtype=torch.float32) if save else None
    ops.layernorm_fwd(x, rt.w32(ln.weight), rt.w32(ln.bias), y, mean, rstd, LN_EPS, out_scale, accum, seqs=seqs)
    _count()
    return y, (mean, rstd)


def ln_bwd(rt, dy, x, stats, ln, out_scale=1.0, dx=None, dx_drop=None, p=0.0, seed=0, dbias=None, offset=0, seqs=None):
    dx = rt.empty(*x.shape) if dx is None else dx
    ops.layernorm_bwd(dy, x, stats[0], stats[1], rt.w32(ln.weight), dx, rt.g32(ln.weight), rt.g32(ln.bias), out_scale,
                      dx_drop, p, seed, offset, dbias, seqs=seqs)
    _count()
    return dx


def dropout_(rt, x, p, seed):
    """In-place inverted dropout (embedding / pooled-output sites); the same call is its own backward."""
    if p > 0.0:
        ops.eltwise(ELT_DROPOUT, x, x, p=p, seed=seed)
        _count()
    return x


# ---- residual block tail:  u = dropout(dense(h)) + x ;  y = LN(u) -------------------------------------
def out_block_fwd(rt, h, x, dense, ln, p, y, st=None):
    """st: which stream's rows these are ('t' / 'v' / 'all'): the LayerNorm skips their trailing-padding rows."""
    seed = rt.seed() if p > 0 else 0
    u = linear_fwd(rt, h, dense, epilogue=EPI_BIAS_DROP_RES, aux_in=x, dropout_p=p, seed=seed)
    sq = rt.seqs(st)
    _, stats = ln_fwd(rt, u, ln, y=y, seqs=sq)
    return (h, u, stats, p, seed, sq)


def out_block_bwd(rt, dy, saved, dense, ln, dh_epilogue=EPI_NONE, dh_aux=None):
    """Returns (du, dh): du is the gradient of the residual input, dh = dense-input gradient (+epilogue)."""
    h, u, stats, p, seed, sq = saved
    du = rt.empty(*u.shape)
    du_drop = rt.empty(*u.shape) if p > 0 else None
    ln_bwd(rt, dy, u, stats, ln, dx=du, dx_drop=du_drop, p=p, seed=seed, dbias=rt.g32(dense.bias), seqs=sq)
    dz = du_drop if p > 0 else du
    dh = linear_bwd(rt, h, dz, dense, dx_epilogue=dh_epilogue, bias_done=True, aux_in=dh_aux)
    return du, dh


# ---- the same tail for BOTH streams of a two-stream sub-layer in one row-split GEMM ------------------------------
def split_ok(rt, x, p_t, p_v, *lin_pairs):
    """Can the text and vision rows of this sub-layer share row-split GEMM launches?  They must live in one buffer,
    the text rows must end on a 256-row tile edge, the streams must use equally shaped Linears and one dropout rate."""
    tc = (rt.precision == "bf16" and rt.gemm_backend != _lib.GEMM_SIMT) or PLAN_EMULATED
    if x.all is None or not tc or os.environ.get("VOLTA_B200_ROW_SPLIT", "1") == "0":
        return False
    if rt.Mt % 256 != 0 or rt.Mv == 0 or p_t != p_v:
        return False
    return all(a.weight.shape == b.weight.shape and (a.bias is None) == (b.bias is None) for a, b in lin_pairs)


def out_block_fwd2(rt, h, x, dense_t, ln_t, dense_v, ln_v, p, y):
    seed = rt.seed() if p > 0 else 0
    Mt = rt.Mt
    u = linear_fwd(rt, h, dense_t, epilogue=EPI_BIAS_DROP_RES, aux_in=x, dropout_p=p, seed=seed, lin2=dense_v, split=Mt)
    # one LayerNorm launch for both streams: rows [Mt, ...) take the vision module's gamma / beta
    sq = rt.seqs("all")
    mean, rstd = rt.empty(u.shape[0], dtype=torch.float32), rt.empty(u.shape[0], dtype=torch.float32)
    ops.layernorm_fwd(u, rt.w32(ln_t.weight), rt.w32(ln_t.bias), y, mean, rstd, LN_EPS, seqs=sq,
                      second=(rt.w32(ln_v.weight), rt.w32(ln_v.bias)), split_row=Mt)
    _count()
    return (h, u, (mean, rstd), p, seed, sq)


def out_block_bwd2(rt, dy, saved, dense_t, ln_t, dense_v, ln_v, dh_epilogue=EPI_NONE, dh_aux=None):
    h, u, (mean, rstd), p, seed, sq = saved
    Mt, H = rt.Mt, u.shape[1]
    du = rt.empty(*u.shape)
    du_drop = rt.empty(*u.shape) if p > 0 else None
    # one launch for both streams; the dropout mask of the row-split GEMM is indexed by the JOINT row, as here
    ops.layernorm_bwd(dy, u, mean, rstd, rt.w32(ln_t.weight), du, rt.g32(ln_t.weight), rt.g32(ln_t.bias), 1.0, du_drop, p, seed, 0,
                      rt.g32(dense_t.bias), seqs=sq,
                      second=(rt.w32(ln_v.weight), rt.g32(ln_v.weight), rt.g32(ln_v.bias), rt.g32(dense_v.bias)), split_row=Mt)
    _count()
    dz = du_drop if p > 0 else du
    dh = linear_bwd(rt, h, dz, dense_t, dx_epilogue=dh_epilogue, bias_done=True, aux_in=dh_aux, lin2=dense_v, split=Mt)
    return du, dh


# ====================================================================================================
# encoder sub-layers
# ====================================================================================================
def _streams(sub_has_t, sub_has_v, shared, x):
    """Yields (tag, rows-view selector) for the GEMM groups of a sub-layer."""
    if sub_has_t and sub_has_v and shared and x.all is not None:
        return ["all"]
    return [s for s, on in (("t", sub_has_t), ("v", sub_has_v)) if on]


def _grad_pair_like(rt, x):
    """Gradient buffers for a sub-layer input: joint (one [Mt+Mv, H] buffer) whenever the input was joint,
    because the producer of `x` may be a shared sub-layer that back-propagates over all rows at once."""
    if x.all is not None:
        return Pair.alloc(rt, x.t.shape[0], x.t.shape[1], x.v.shape[0], x.v.shape[1])
    return Pair(None, None, None)


def _pass_through(dy, dst):
    """A stream a sub-layer does not touch passes its gradient through (copied only into a joint buffer)."""
    if dst is None:
        return dy
    ops.eltwise(ELT_COPY, dy, dst)
    _count()
    return dst


def attention_fwd(rt, layer, x, masks):
    sa, so = layer.attention_self, layer.attention_output
    groups = _streams(sa.has_text, sa.has_vision, sa.share_layer, x)
    A_t, A_v = sa.all_head_size, sa.v_all_head_size
    joint = groups == ["all"]
    qkv, ctx = {}, {}
    split = (not joint and sa.has_text and sa.has_vision and A_t == A_v and
             split_ok(rt, x, rt.p(so.dropout.p), rt.p(so.v_dropout.p), (sa.query, sa.v_query), (sa.key, sa.v_key),
                      (sa.value, sa.v_value), (so.dense, so.v_dense)))
    if joint or split:
        # one GEMM over all rows: shared weights (joint) or a row-split launch with each stream's own weights
        q_all = linear_fwd(rt, x.all, [sa.query, sa.key, sa.value],
                           lin2=[sa.v_query, sa.v_key, sa.v_value] if split else None, split=rt.Mt)
        qkv["t"], qkv["v"] = q_all[:rt.Mt], q_all[rt.Mt:]
        if split:
            qkv["all"] = q_all
        c_all = rt.empty(rt.Mt + rt.Mv, A_t)
        ctx["t"], ctx["v"], ctx["all"] = c_all[:rt.Mt], c_all[rt.Mt:], c_all
    else:
        if sa.has_text:
            qkv["t"] = linear_fwd(rt, x.t, [sa.query, sa.key, sa.value])
            ctx["t"] = rt.empty(rt.Mt, A_t)
        if sa.has_vision:
            qkv["v"] = linear_fwd(rt, x.v, [sa.v_query, sa.v_key, sa.v_value])
            ctx["v"] = rt.empty(rt.Mv, A_v)
    heads = sa.num_attention_heads if sa.has_text else sa.v_num_attention_heads
    d = sa.attention_head_size if sa.has_text else sa.v_attention_head_size
    if sa.has_text and sa.has_vision and (sa.has_tv or sa.has_vt or joint):
        assert sa.num_attention_heads == sa.v_num_attention_heads and sa.attention_head_size == sa.v_attention_head_size
    lse_t = rt.empty(rt.B, heads, rt.T, dtype=torch.float32) if sa.has_text else None
    lse_v = rt.empty(rt.B, sa.v_num_attention_heads, rt.V, dtype=torch.float32) if sa.has_vision else None
    p_t, p_v = rt.p(sa.dropout.p) if sa.has_text else 0.0, rt.p(sa.v_dropout.p) if sa.has_vision else 0.0
    seed = rt.seed() if (p_t > 0 or p_v > 0) else 0
    cross = sa.has_tv or sa.has_vt
    calls = []
    same_heads = (sa.num_attention_heads == sa.v_num_attention_heads and sa.attention_head_size == sa.v_attention_head_size)
    if cross or joint or not (sa.has_text and sa.has_vision) or same_heads:
        # one launch covers both streams; disabled (query, key) blocks are skipped inside the kernel
        calls.append((sa.pair_mask, qkv.get("t"), qkv.get("v"), heads, d))
    else:
        # independent self-attentions (tt + vv): the streams may even differ in head size (vilbert_base)
        calls.append((1, qkv["t"], None, sa.num_attention_heads, sa.attention_head_size))
        calls.append((8, None, qkv["v"], sa.v_num_attention_heads, sa.v_attention_head_size))
    for pm, qt, qv, h_, d_ in calls:
        ops.attention_fwd(qt, qv, masks[0] if qt is not None else None, masks[1] if qv is not None else None,
                          ctx.get("t") if qt is not None else None, ctx.get("v") if qv is not None else None,
                          lse_t, lse_v, rt.B, rt.T, rt.V, h_, d_, pm, p_t, p_v, seed, rt.attn_backend,
                          len_t=getattr(rt, "lens", (None, None))[0], len_v=getattr(rt, "lens", (None, None))[1], **_pack_cu(rt))
        _count()
    _zero_pad_rows(rt, ctx)
    if getattr(rt, "attn_taps", None) is not None:
        rt.attn_taps.append(_attention_data(rt, sa, qkv, masks) if sa.visualization else (None, None))
    y = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1]) if (sa.has_text and sa.has_vision) else None
    saved_out = {}
    if joint:
        saved_out["all"] = out_block_fwd(rt, ctx["all"], x.all, so.dense, so.LayerNorm, rt.p(so.dropout.p), y.all, st="all")
    elif split:
        saved_out["split"] = out_block_fwd2(rt, ctx["all"], x.all, so.dense, so.LayerNorm, so.v_dense, so.v_LayerNorm,
                                            rt.p(so.dropout.p), y.all)
    else:
        if y is None:  # only one stream is touched: the other passes through unchanged (encoders.py:380-396)
            y = Pair(rt.empty(*x.t.shape) if sa.has_text else x.t, rt.empty(*x.v.shape) if sa.has_vision else x.v, None)
        if sa.has_text:
            saved_out["t"] = out_block_fwd(rt, ctx["t"], x.t, so.dense, so.LayerNorm, rt.p(so.dropout.p), y.t, st="t")
        if sa.has_vision:
            saved_out["v"] = out_block_fwd(rt, ctx["v"], x.v, so.v_dense, so.v_LayerNorm, rt.p(so.v_dropout.p), y.v, st="v")
    saved = dict(x=x, qkv=qkv, ctx=ctx, lse=(lse_t, lse_v), calls=calls, seed=seed, p=(p_t, p_v), out=saved_out,
                 joint=joint, split=split, masks=masks, lens=getattr(rt, "lens", (None, None)))
    return y, saved


def _attention_data(rt, sa, qkv, masks):
    """(t_attn_data, v_attn_data) of BertGatedSelfAttention with config.visualization (volta/encoders.py:342-356): the
    probabilities of the enabled (query stream, key stream) blocks after the joint softmax (eval semantics: no dropout),
    queries and keys as [B, heads, L, d].  The fused kernels never materialise P; vb_attention_probs recomputes it."""
    B, T, V = rt.B, rt.T, rt.V
    Lt, Lv = (T if "t" in qkv else 0), (V if "v" in qkv else 0)
    heads = sa.num_attention_heads if sa.has_text else sa.v_num_attention_heads
    d = sa.attention_head_size if sa.has_text else sa.v_attention_head_size
    same = (not (sa.has_text and sa.has_vision)) or (sa.num_attention_heads == sa.v_num_attention_heads and
                                                      sa.attention_head_size == sa.v_attention_head_size)
    out = {}
    calls = [(sa.pair_mask, qkv.get("t"), qkv.get("v"), heads, d)] if same else \
        [(1, qkv["t"], None, sa.num_attention_heads, sa.attention_head_size),
         (8, None, qkv["v"], sa.v_num_attention_heads, sa.v_attention_head_size)]
    for pm, qt, qv, h_, d_ in calls:
        # a stream no enabled pair touches has no columns (vb_attention_probs: text uses tt|tv|vt, vision tv|vt|vv)
        lt, lv = (T if (qt is not None and pm & 7) else 0), (V if (qv is not None and pm & 14) else 0)
        pt = rt.empty(B, h_, T, lt + lv, dtype=torch.float32) if (qt is not None and pm & 3) else None
        pv = rt.empty(B, h_, V, lt + lv, dtype=torch.float32) if (qv is not None and pm & 12) else None
        ops.attention_probs(qt, qv, masks[0] if qt is not None else None, masks[1] if qv is not None else None, pt, pv,
                            B, T, V, h_, d_, pm)
        _count()
        if pt is not None:
            out["t"] = (pt, lt, h_, d_)
        if pv is not None:
            out["v"] = (pv, lt, h_, d_)

    def heads_view(q, L, h_, d_, which):
        A = h_ * d_
        return q[:, which * A:(which + 1) * A].reshape(B, L, h_, d_).permute(0, 2, 1, 3)
    # both dicts always exist; a stream the sub-layer does not touch has None in every slot (encoders.py:343-354)
    t_data, v_data = (dict(intra_attn=None, inter_attn=None, queries=None, keys=None) for _ in range(2))
    if sa.has_text:
        pt, lt, h_, d_ = out["t"]
        t_data = {"intra_attn": pt[..., :lt] if sa.has_tt else None, "inter_attn": pt[..., lt:] if sa.has_tv else None,
                  "queries": heads_view(qkv["t"], T, h_, d_, 0), "keys": heads_view(qkv["t"], T, h_, d_, 1)}
    if sa.has_vision:
        pv, lt, h_, d_ = out["v"]
        v_data = {"intra_attn": pv[..., lt:] if sa.has_vv else None, "inter_attn": pv[..., :lt] if sa.has_vt else None,
                  "queries": heads_view(qkv["v"], V, h_, d_, 0), "keys": heads_view(qkv["v"], V, h_, d_, 1)}
    return t_data, v_data


def attention_bwd(rt, layer, dy, s):
    """dy: Pair of gradients w.r.t. the sub-layer outputs; returns the Pair w.r.t. its inputs."""
    sa, so = layer.attention_self, layer.attention_output
    x, qkv, ctx = s["x"], s["qkv"], s["ctx"]
    du, dctx = {}, {}
    if s["joint"]:
        du_all, dc_all = out_block_bwd(rt, dy.all, s["out"]["all"], so.dense, so.LayerNorm)
        du["all"], dctx["t"], dctx["v"] = du_all, dc_all[:rt.Mt], dc_all[rt.Mt:]
    elif s["split"]:
        du_all, dc_all = out_block_bwd2(rt, dy.all, s["out"]["split"], so.dense, so.LayerNorm, so.v_dense, so.v_LayerNorm)
        du["all"], dctx["t"], dctx["v"] = du_all, dc_all[:rt.Mt], dc_all[rt.Mt:]
    else:
        if sa.has_text:
            du["t"], dctx["t"] = out_block_bwd(rt, dy.t, s["out"]["t"], so.dense, so.LayerNorm)
        if sa.has_vision:
            du["v"], dctx["v"] = out_block_bwd(rt, dy.v, s["out"]["v"], so.v_dense, so.v_LayerNorm)
    dqkv, to_clear = {}, {}
    if s["joint"] or s["split"]:
        dq_all, fresh = _dqkv_buffer(rt, rt.Mt + rt.Mv, qkv["t"].shape[1])
        dqkv["t"], dqkv["v"], dqkv["all"] = dq_all[:rt.Mt], dq_all[rt.Mt:], dq_all
        if fresh:
            to_clear["all"] = dq_all
    else:
        for k in qkv:
            dqkv[k], fresh = _dqkv_buffer(rt, *qkv[k].shape)
            if fresh:
                to_clear[k] = dqkv[k]
    lse_t, lse_v = s["lse"]
    for pm, qt, qv, h_, d_ in s["calls"]:
        ops.attention_bwd(qt, qv, s["masks"][0] if qt is not None else None, s["masks"][1] if qv is not None else None,
                          ctx.get("t") if qt is not None else None, ctx.get("v") if qv is not None else None,
                          lse_t, lse_v, rt.B, rt.T, rt.V, h_, d_, pm, s["p"][0], s["p"][1], s["seed"], rt.attn_backend,
                          dctx_t=dctx.get("t") if qt is not None else None, dctx_v=dctx.get("v") if qv is not None else None,
                          dqkv_t=dqkv.get("t") if qt is not None else None, dqkv_v=dqkv.get("v") if qv is not None else None,
                          len_t=s["lens"][0], len_v=s["lens"][1], **_pack_cu(rt))
        _count()
    _zero_pad_rows(rt, to_clear)
    # dx = dqkv . Wqkv + du   (the residual gradient is folded into the dgrad epilogue)
    if s["joint"] or s["split"]:
        dx_all = linear_bwd(rt, x.all, dqkv["all"], [sa.query, sa.key, sa.value], dx_epilogue=EPI_ADD, aux_in=du["all"],
                            lin2=[sa.v_query, sa.v_key, sa.v_value] if s["split"] else None, split=rt.Mt)
        return Pair(dx_all[:rt.Mt], dx_all[rt.Mt:], dx_all)
    dx = _grad_pair_like(rt, x)
    if sa.has_text:
        dx.t = linear_bwd(rt, x.t, dqkv["t"], [sa.query, sa.key, sa.value], dx_epilogue=EPI_ADD, aux_in=du["t"], dx_out=dx.t)
    else:
        dx.t = _pass_through(dy.t, dx.t)
    if sa.has_vision:
        dx.v = linear_bwd(rt, x.v, dqkv["v"], [sa.v_query, sa.v_key, sa.v_value], dx_epilogue=EPI_ADD, aux_in=du["v"], dx_out=dx.v)
    else:
        dx.v = _pass_through(dy.v, dx.v)
    return dx


def _act_epilogues(name):
    if name == "gelu":
        return EPI_BIAS_GELU, EPI_MUL  # the forward epilogue saves gelu'(x); the backward epilogue multiplies by it
    if name == "relu":
        return EPI_BIAS_RELU, EPI_DRELU
    if name == "swish":
        # no shipped config uses it (ACT2FN, volta/encoders.py:157): the Linear runs with the plain bias epilogue and two
        # element-wise passes form x*sigmoid(x) and its derivative; the backward multiplies the derivative in like GELU's
        return "swish", EPI_MUL
    raise ValueError("hidden_act=%r is not one of gelu / relu / swish (volta/encoders.py:157)" % (name,))


def _act_linear_fwd(rt, x, lin, act, h, pre, **kw):
    """h = act(x W^T + b); pre = what the backward epilogue needs (the derivative for gelu / swish; h itself for relu)."""
    epi_f, _ = _act_epilogues(act)
    if epi_f == EPI_BIAS_GELU:
        linear_fwd(rt, x, lin, epilogue=EPI_BIAS_GELU, out=h, aux_out=pre, **kw)
        return pre
    if epi_f == "swish":
        linear_fwd(rt, x, lin, epilogue=EPI_BIAS, out=pre, **kw)      # pre-activation, then activation / derivative in place
        ops.eltwise(_lib.ELT_SWISH, pre, h)
        ops.eltwise(_lib.ELT_DSWISH, pre, pre)
        _count(2)
        return pre
    linear_fwd(rt, x, lin, epilogue=EPI_BIAS_RELU, out=h, **kw)
    return h  # relu'(pre) = (h > 0)


def _ffn_stream_fwd(rt, x, inter, dense2, ln, p, y, act, second=None, st=None):
    """second = (inter_v, dense2_v, ln_v): both streams in row-split launches over the joint buffer."""
    I = inter.weight.shape[0]
    h = rt.empty(x.shape[0], I)
    sp = dict(lin2=second[0], split=rt.Mt) if second else {}
    pre = _act_linear_fwd(rt, x, inter, act, h, rt.empty(x.shape[0], I) if act != "relu" else None, **sp)
    if second:
        return (x, pre, out_block_fwd2(rt, h, x, dense2, ln, second[1], second[2], p, y))
    return (x, pre, out_block_fwd(rt, h, x, dense2, ln, p, y, st=st))


def _ffn_stream_bwd(rt, dy, saved, inter, dense2, ln, act, dx_out=None, second=None):
    x, pre, out_saved = saved
    _, epi_b = _act_epilogues(act)
    if second:
        du, dpre = out_block_bwd2(rt, dy, out_saved, dense2, ln, second[1], second[2], dh_epilogue=epi_b, dh_aux=pre)
        return linear_bwd(rt, x, dpre, inter, dx_epilogue=EPI_ADD, aux_in=du, dx_out=dx_out, lin2=second[0], split=rt.Mt)
    du, dpre = out_block_bwd(rt, dy, out_saved, dense2, ln, dh_epilogue=epi_b, dh_aux=pre)
    return linear_bwd(rt, x, dpre, inter, dx_epilogue=EPI_ADD, aux_in=du, dx_out=dx_out)


def ffn_fwd(rt, layer, x):
    it, out = layer.intermediate, layer.output
    groups = _streams(it.has_language, it.has_vision, it.share_layer, x)
    saved = {"x": x}
    if groups == ["all"]:
        y = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1])
        saved["all"] = _ffn_stream_fwd(rt, x.all, it.dense, out.dense, out.LayerNorm, rt.p(out.dropout.p), y.all, it.hidden_act, st="all")
    elif (it.has_language and it.has_vision and it.hidden_act == it.v_hidden_act and
          split_ok(rt, x, rt.p(out.dropout.p), rt.p(out.v_dropout.p), (it.dense, it.v_dense), (out.dense, out.v_dense))):
        y = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1])
        saved["split"] = _ffn_stream_fwd(rt, x.all, it.dense, out.dense, out.LayerNorm, rt.p(out.dropout.p), y.all, it.hidden_act,
                                         second=(it.v_dense, out.v_dense, out.v_LayerNorm))
    else:
        if it.has_language and it.has_vision:
            y = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1])
        else:
            y = Pair(rt.empty(*x.t.shape) if it.has_language else x.t, rt.empty(*x.v.shape) if it.has_vision else x.v, None)
        if it.has_language:
            saved["t"] = _ffn_stream_fwd(rt, x.t, it.dense, out.dense, out.LayerNorm, rt.p(out.dropout.p), y.t, it.hidden_act, st="t")
        if it.has_vision:
            saved["v"] = _ffn_stream_fwd(rt, x.v, it.v_dense, out.v_dense, out.v_LayerNorm, rt.p(out.v_dropout.p), y.v, it.v_hidden_act, st="v")
    return y, saved


def ffn_bwd(rt, layer, dy, saved):
    it, out = layer.intermediate, layer.output
    if "all" in saved:
        dx = _ffn_stream_bwd(rt, dy.all, saved["all"], it.dense, out.dense, out.LayerNorm, it.hidden_act)
        return Pair(dx[:rt.Mt], dx[rt.Mt:], dx)
    if "split" in saved:
        dx = _ffn_stream_bwd(rt, dy.all, saved["split"], it.dense, out.dense, out.LayerNorm, it.hidden_act,
                             second=(it.v_dense, out.v_dense, out.v_LayerNorm))
        return Pair(dx[:rt.Mt], dx[rt.Mt:], dx)
    dx = _grad_pair_like(rt, saved["x"])
    if "t" in saved:
        dx.t = _ffn_stream_bwd(rt, dy.t, saved["t"], it.dense, out.dense, out.LayerNorm, it.hidden_act, dx.t)
    else:
        dx.t = _pass_through(dy.t, dx.t)
    if "v" in saved:
        dx.v = _ffn_stream_bwd(rt, dy.v, saved["v"], it.v_dense, out.v_dense, out.v_LayerNorm, it.v_hidden_act, dx.v)
    else:
        dx.v = _pass_through(dy.v, dx.v)
    return dx


# ====================================================================================================
# embeddings
# ====================================================================================================
_POS_CACHE = {}


def _pos_ids(rt, offset=0):
    key = (rt.device, rt.B, rt.T, offset)
    if key not in _POS_CACHE:
        _POS_CACHE[key] = (torch.arange(offset, rt.T + offset, dtype=torch.int64).repeat(rt.B)).to(rt.device)
    return _POS_CACHE[key]


def _to_compute(rt, feat):
    """Region features arrive fp32 from the loader; the GEMM operand is the compute dtype."""
    feat = feat.reshape(-1, feat.shape[-1])
    if feat.dtype == rt.cdtype:
        return feat
    out = rt.empty(*feat.shape)
    ops.eltwise(ELT_COPY, feat, out)
    _count()
    return out


def _text_gather(rt, emb, ids, seg, out, pos_offset=0):
    rt._pos_offset = pos_offset
    idx = [ids.reshape(-1), _pos_ids(rt, pos_offset), seg.reshape(-1)]
    tables = [emb.word_embeddings.weight, emb.position_embeddings.weight, emb.token_type_embeddings.weight]
    ops.embed_gather_fwd([rt.w32(t) for t in tables], idx, [0, 0, 0], out)
    _count()
    return (tables, idx)


def _text_gather_bwd(rt, saved, dout):
    """Scatter of the embedding-sum gradient.  Word rows: atomics (the padding row gets none, as with
    nn.Embedding(padding_idx=0)); positions: every sample hits the same T rows, so their gradient is a column sum
    of dout viewed as [B, T*H]; token types: 2-3 rows hit by every token -> per-key column sums."""
    tables, idx = saved
    word, pos, typ = tables
    H = dout.shape[1]
    ops.embed_gather_bwd([rt.g32(word)], [idx[0]], [0], dout, skip=[0])
    gpos = rt.g32(pos)
    off = rt._pos_offset
    ops.colsum(dout.view(rt.B, rt.T * H), gpos[off:off + rt.T].view(-1))
    ntypes = typ.shape[0]
    if ntypes <= 4:
        ops.colsum_by_key(dout, idx[2], ntypes, rt.g32(typ))
    else:
        ops.embed_gather_bwd([rt.g32(typ)], [idx[2]], [0], dout)
    _count(3)


def _region_proj_fwd(rt, lin, feat_c, out):
    linear_fwd(rt, feat_c, lin, epilogue=EPI_BIAS, out=out)


def _box_fwd(rt, lin, loc, out, accum):
    ops.box_project_fwd(loc, rt.w32(lin.weight), rt.w32(lin.bias), out, accum)
    _count()


def _box_bwd(rt, lin, loc, dout):
    ops.box_project_bwd(loc, dout, rt.g32(lin.weight), rt.g32(lin.bias))
    _count()


def embeddings_fwd(rt, bert, ids, feat, loc, seg):
    cfg = bert.config
    kind = cfg.image_embeddings
    H, Hv = cfg.hidden_size, cfg.v_hidden_size
    x = Pair.alloc(rt, rt.Mt, H, rt.Mv, Hv)
    feat_c = _to_compute(rt, feat)
    loc2 = loc.reshape(-1, loc.shape[-1]).contiguous().float()
    pos_off = bert.embeddings.padding_idx + 1 if hasattr(bert.embeddings, "padding_idx") and not bert.shared_embeddings else 0
    s = dict(kind=kind, feat=feat_c, loc=loc2)
    if kind in ("vilbert", "lxmert"):
        e, ve = bert.embeddings, bert.v_embeddings
        st = rt.empty(rt.Mt, H)
        s["text"] = _text_gather(rt, e, ids, seg, st, pos_off)
        _, s["t_stats"] = ln_fwd(rt, st, e.LayerNorm, y=x.t)
        s["t_sum"] = st
        s["t_drop"] = (rt.p(e.dropout.p), rt.seed())
        dropout_(rt, x.t, *s["t_drop"])
        a = rt.empty(rt.Mv, Hv)
        _region_proj_fwd(rt, ve.image_embeddings, feat_c, a)
        if kind == "vilbert":
            _box_fwd(rt, ve.image_location_embeddings, loc2, a, True)
            _, s["v_stats"] = ln_fwd(rt, a, ve.LayerNorm, y=x.v)
            s["v_sum"] = a
        else:
            l = rt.empty(rt.Mv, Hv)
            _box_fwd(rt, ve.image_location_embeddings, loc2, l, False)
            _, s["img_stats"] = ln_fwd(rt, a, ve.ImgLayerNorm, y=x.v, out_scale=0.5)
            _, s["loc_stats"] = ln_fwd(rt, l, ve.LocLayerNorm, y=x.v, out_scale=0.5, accum=True)
            s["img_sum"], s["loc_sum"] = a, l
        s["v_drop"] = (rt.p(ve.dropout.p), rt.seed())
        dropout_(rt, x.v, *s["v_drop"])
    elif kind == "visualbert":
        e = bert.embeddings
        assert x.all is not None
        u = Pair.alloc(rt, rt.Mt, H, rt.Mv, Hv)
        s["text"] = _text_gather(rt, e, ids, seg, u.t)
        _region_proj_fwd(rt, e.projection, feat_c, u.v)
        vt = [e.position_embeddings_visual.weight, e.token_type_embeddings_visual.weight]
        ops.embed_gather_fwd([rt.w32(t) for t in vt], [None, None], [0, 1], u.v, accum=True)
        _count()
        s["v_tables"] = vt
        _, s["stats"] = ln_fwd(rt, u.all, e.LayerNorm, y=x.all)
        s["sum"] = u
        s["drop"] = (rt.p(e.dropout.p), rt.seed())
        dropout_(rt, x.all, *s["drop"])
    elif kind == "uniter":
        e = bert.embeddings
        st = rt.empty(rt.Mt, H)
        s["text"] = _text_gather(rt, e, ids, seg, st)
        _, s["t_stats"] = ln_fwd(rt, st, e.LayerNorm, y=x.t)
        s["t_sum"] = st
        s["t_drop"] = (rt.p(e.dropout.p), rt.seed())
        dropout_(rt, x.t, *s["t_drop"])
        a, l, z = rt.empty(rt.Mv, Hv), rt.empty(rt.Mv, Hv), rt.empty(rt.Mv, Hv)
        _region_proj_fwd(rt, e.image_embeddings, feat_c, a)
        _box_fwd(rt, e.image_location_embeddings, loc2, l, False)
        _, s["img_stats"] = ln_fwd(rt, a, e.image_layer_norm, y=z)
        _, s["loc_stats"] = ln_fwd(rt, l, e.image_location_layer_norm, y=z, accum=True)
        ops.embed_gather_fwd([rt.w32(e.token_type_embeddings.weight)], [None], [1], z, accum=True)
        _count()
        _, s["v_stats"] = ln_fwd(rt, z, e.v_LayerNorm, y=x.v)
        s["img_sum"], s["loc_sum"], s["v_sum"] = a, l, z
        s["v_drop"] = (rt.p(e.v_dropout.p), rt.seed())
        dropout_(rt, x.v, *s["v_drop"])
    elif kind == "vl-bert":
        from .vlbert import vlbert_embeddings_fwd
        vlbert_embeddings_fwd(rt, bert, ids, feat, loc2, seg, x, s)
    else:
        raise ValueError("unknown image_embeddings %r" % kind)
    return x, s


def embeddings_bwd(rt, bert, dx, s):
    kind = s["kind"]
    feat_c, loc2 = s["feat"], s["loc"]
    if kind in ("vilbert", "lxmert"):
        e, ve = bert.embeddings, bert.v_embeddings
        dropout_(rt, dx.t, *s["t_drop"])
        ds = ln_bwd(rt, dx.t, s["t_sum"], s["t_stats"], e.LayerNorm)
        _text_gather_bwd(rt, s["text"], ds)
        dropout_(rt, dx.v, *s["v_drop"])
        if kind == "vilbert":
            da = ln_bwd(rt, dx.v, s["v_sum"], s["v_stats"], ve.LayerNorm, dbias=rt.g32(ve.image_embeddings.bias))
            _box_bwd(rt, ve.image_location_embeddings, loc2, da)
        else:
            da = ln_bwd(rt, dx.v, s["img_sum"], s["img_stats"], ve.ImgLayerNorm, out_scale=0.5,
                        dbias=rt.g32(ve.image_embeddings.bias))
            dl = ln_bwd(rt, dx.v, s["loc_sum"], s["loc_stats"], ve.LocLayerNorm, out_scale=0.5)
            _box_bwd(rt, ve.image_location_embeddings, loc2, dl)
        linear_bwd(rt, feat_c, da, ve.image_embeddings, need_dx=False, bias_done=True)
    elif kind == "visualbert":
        e = bert.embeddings
        dropout_(rt, dx.all, *s["drop"])
        du = ln_bwd(rt, dx.all, s["sum"].all, s["stats"], e.LayerNorm)
        _text_gather_bwd(rt, s["text"], du[:rt.Mt])
        dv = du[rt.Mt:]
        linear_bwd(rt, feat_c, dv, e.projection, need_dx=False)
        ops.colsum(dv, rt.g32(e.position_embeddings_visual.weight)[0])
        ops.colsum(dv, rt.g32(e.token_type_embeddings_visual.weight)[1])
        _count(2)
    elif kind == "uniter":
        e = bert.embeddings
        dropout_(rt, dx.t, *s["t_drop"])
        ds = ln_bwd(rt, dx.t, s["t_sum"], s["t_stats"], e.LayerNorm)
        _text_gather_bwd(rt, s["text"], ds)
        dropout_(rt, dx.v, *s["v_drop"])
        dz = ln_bwd(rt, dx.v, s["v_sum"], s["v_stats"], e.v_LayerNorm)
        ops.colsum(dz, rt.g32(e.token_type_embeddings.weight)[1])
        _count()
        da = ln_bwd(rt, dz, s["img_sum"], s["img_stats"], e.image_layer_norm, dbias=rt.g32(e.image_embeddings.bias))
        dl = ln_bwd(rt, dz, s["loc_sum"], s["loc_stats"], e.image_location_layer_norm)
        _box_bwd(rt, e.image_location_embeddings, loc2, dl)
        linear_bwd(rt, feat_c, da, e.image_embeddings, need_dx=False, bias_done=True)
    elif kind == "vl-bert":
        from .vlbert import vlbert_embeddings_bwd
        vlbert_embeddings_bwd(rt, bert, dx, s)


# ====================================================================================================
# BertModel: embeddings -> encoder -> poolers
# ====================================================================================================
def _int_mask(rt, mask, B, L):
    if mask is None:
        return torch.ones(B, L, dtype=torch.int32, device=rt.device)
    return mask.reshape(B, L).to(device=rt.device, dtype=torch.int32).contiguous()


def _pool_fwd(rt, pooler, x, L, rows=None):
    """relu(W x[:, 0] + b): the [CLS] / global rows are read in place through the leading dimension."""
    H = x.shape[1]
    a = x.view(rt.B, L * H)[:, :H] if rows is None else rows
    out = rt.empty(rt.B, pooler.dense.weight.shape[0])
    linear_fwd(rt, a, pooler.dense, epilogue=EPI_BIAS_RELU, out=out)
    return out, a


def _pool_bwd(rt, pooler, dpooled, pooled, a, dx, L):
    """dx[:, 0] += (dpooled * (pooled > 0)) W, in place through the leading dimension."""
    dpre = rt.empty(*pooled.shape)
    ops.eltwise(ELT_DRELU, dpooled, dpre, b=pooled)
    _count()
    H = dx.shape[1]
    rows0 = dx.view(rt.B, L * H)[:, :H]
    linear_bwd(rt, a, dpre, pooler.dense, dx_epilogue=EPI_ADD, dx_out=rows0, aux_in=rows0)


# ---- packed rows ----------------------------------------------------------------------------------------------------
# The reference computes every padded position (SURVEY.md 8a' item 3: finite garbage nothing reads).  Between the
# embeddings and the heads the valid rows of each stream are PACKED to the front of its buffer: the GEMMs schedule only
# the tiles below the device-side count (vb_gemm_args::m_limit / m_limit2 / k_limit), the attention kernels address
# sample b at cu[b] (vb_attn_args::cu_*), the LayerNorms stop at the count (vb_row_seqs).  Valid rows, losses and
# parameter gradients are what the padded layout gives; shapes, launches and CUDA graphs stay static.
PACK_ROWS = os.environ.get("VOLTA_B200_PACK_ROWS", "1") != "0"
# Shared sub-layers (ctrl_lxmert's cross layers; every layer of the single-stream models) run ONE GEMM over the joint
# buffer: packed, it holds two runs of valid rows (vb_gemm_args::m_limit / m_limit2 / limit_seg_M forward and dgrad;
# the weight / bias gradients reduce per run, linear_bwd).  Measured on a B200 (B=256, round 2): ctrl_visualbert_base
# 19.95 k -> 24.6 k samples/s, ctrl_lxmert 17.1 k -> 21.6 k.  VOLTA_B200_PACK_SHARED=0 switches it off.
PACK_SHARED = os.environ.get("VOLTA_B200_PACK_SHARED", "1") != "0"
# The vision rows of a joint buffer must start on a 256-row tile edge.  When B*T is not a multiple of 256 (the reference
# recipe: global batch 256 over 8 GPUs = 32 x 36 rows; BASELINE.json configs[0]: batch 8) the packed region pads the
# TEXT segment of its buffers up to the next edge (rt.Mt grows for the duration of the packed encoder; the padding rows
# are dropped rows of the plan, kept zero like every row between a count and its tile edge).  Measured (B=32, T=38):
# 4.92 k -> 5.61 k samples/s.  VOLTA_B200_PACK_ANY_ROWS=0 switches it off.
PACK_ANY_ROWS = os.environ.get("VOLTA_B200_PACK_ANY_ROWS", "1") != "0"


def _pad256(n):
    return (n + 255) // 256 * 256


class _Pack:
    pass


def _pack_supported(rt, bert, keep_layers):
    if not (PACK_ROWS and SKIP_PADDING) or keep_layers or rt.precision != "bf16":
        return False
    if rt.gemm_backend == _lib.GEMM_SIMT or rt.attn_backend == _lib.ATTN_SIMT or not torch.cuda.is_available():
        return False
    if (rt.Mt % 256 != 0 and not PACK_ANY_ROWS) or _pad256(rt.Mt) == rt.Mv or rt.Mv == 0 or rt.T + rt.V > 160:
        return False                                                   # 160: what one tensor-core attention block holds
    if bert.config.fusion_method == "vl-bert_vqa":
        return False
    for n, layer in enumerate(bert.encoder.layer):
        if bert.encoder.num2type[n] == "attn":
            sa = layer.attention_self
            if (sa.share_layer and not PACK_SHARED) or (sa.has_text and sa.attention_head_size not in (64, 128)) or \
                    (sa.has_vision and sa.v_attention_head_size not in (64, 128)):
                return False
        elif layer.intermediate.share_layer and not PACK_SHARED:
            return False
    return True


def _make_pack(rt):
    """Packing plan of both streams from the per-sample lengths (rows below the length are kept, holes included)."""
    pk = _Pack()
    for tag, ln, L in (("t", rt.lens[0], rt.T), ("v", rt.lens[1], rt.V)):
        rows = rt.B * L
        keep = (torch.arange(L, device=rt.device, dtype=torch.int32)[None, :] < ln[:, None]).reshape(-1)
        keys = keep.to(torch.int64) - 1                                  # 0 = keep, -1 = drop
        if tag == "t" and rows % 256 != 0:                               # text segment padded to the next tile edge
            keys = torch.cat([keys, torch.full((_pad256(rows) - rows,), -1, dtype=torch.int64, device=rt.device)])
            rows = _pad256(rows)
        lim = rt.empty(2, dtype=torch.int32)
        src, pos = rt.empty(rows, dtype=torch.int32), rt.empty(rows, dtype=torch.int32)
        ops.compact_plan(keys, 256, lim, src, pos, rt.empty(rows, dtype=torch.int64))
        cu = (torch.cumsum(ln, 0, dtype=torch.int32) - ln).contiguous()  # first packed row of every sample
        setattr(pk, "lim_" + tag, lim); setattr(pk, "src_" + tag, src); setattr(pk, "pos_" + tag, pos); setattr(pk, "cu_" + tag, cu)
        _count()
    return pk


def _pack_pair(rt, x):
    """Valid rows of both streams to the front of their segment; the text segment has rt.Mt rows (B*T, or B*T padded to a
    256-row edge - _enter_packed)."""
    pk = rt.pack
    y = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1])
    ops.compact_gather(x.t, pk.src_t, pk.lim_t, y.t)
    ops.compact_gather(x.v, pk.src_v, pk.lim_v, y.v)
    _count(2)
    return y


def _enter_packed(rt):
    rt.rows_t = rt.Mt
    rt.Mt = _pad256(rt.Mt)


def _leave_packed(rt):
    rt.pack_on = False
    rt.Mt = rt.rows_t


def _unpack_pair(rt, x):
    pk = rt.pack
    y = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1])
    ops.compact_scatter(x.t, pk.pos_t, y.t)
    ops.compact_scatter(x.v, pk.pos_v, y.v)
    _count(2)
    return y


def _dqkv_buffer(rt, rows, cols):
    """Gradient buffer of the fused Q/K/V projection.  On packed rows ONE buffer serves every sub-layer of a backward
    (each consumes it before the next one writes): the attention kernels write whole samples, so the tail between the
    count and its padding, cleared once, stays zero.  Returns (buffer, needs_its_tail_cleared)."""
    if not getattr(rt, "pack_on", False):
        return rt.empty(rows, cols), True
    cache = rt.pack.__dict__.setdefault("dq_cache", {})
    if (rows, cols) in cache:
        return cache[(rows, cols)], False
    cache[(rows, cols)] = rt.empty(rows, cols)
    return cache[(rows, cols)], True


def _pack_cu(rt):
    if not getattr(rt, "pack_on", False):
        return {}
    return dict(cu_t=rt.pack.cu_t, cu_v=rt.pack.cu_v)


def _zero_pad_rows(rt, bufs):
    """Rows between a stream's count and its 256-row padding are read by the tile-granular weight-gradient GEMMs: the
    attention kernels write per sample, so those rows of their outputs are zeroed here."""
    if not getattr(rt, "pack_on", False):
        return
    lt, lv = rt.pack.lim_t, rt.pack.lim_v
    if bufs.get("all") is not None:      # both streams live in one buffer: one launch clears both tails
        ops.zero_rows(bufs["all"], lt[0:1], lt[1:2], second=(lv[0:1], lv[1:2], rt.Mt))
        _count()
        return
    for tag, lim in (("t", lt), ("v", lv)):
        if bufs.get(tag) is not None:
            ops.zero_rows(bufs[tag], lim[0:1], lim[1:2])
            _count()


def bert_fwd(rt, bert, ids, feat, loc, seg, mask_t, mask_v, keep_layers=False, want_attn=False):
    cfg = bert.config
    rt.attn_taps = [] if want_attn else None
    x, s_emb = embeddings_fwd(rt, bert, ids, feat, loc, seg)
    masks = (_int_mask(rt, mask_t, rt.B, rt.T), _int_mask(rt, mask_v, rt.B, rt.V))
    rt.lens = (None, None)
    if SKIP_PADDING:
        # per-sample valid lengths, once per forward: the attention kernels skip whole tiles of trailing padding
        rt.lens = tuple(ops.mask_lengths(m, rt.empty(rt.B, dtype=torch.int32)) for m in masks)
        _count(2)
    rt.pack, rt.pack_on = None, False
    if _pack_supported(rt, bert, keep_layers or want_attn):     # attention data are [B, heads, L, L]: padded layout
        rt.pack = _make_pack(rt)
        _enter_packed(rt)
        x = _pack_pair(rt, x)
        rt.pack_on = True
    last_plan.clear()
    last_plan.update(packed=rt.pack is not None, row_split=0, joint=0, mlm_compact=False, rows_t=rt.Mt, rows_v=rt.Mv)
    tape, layers = [], []
    for n, layer in enumerate(bert.encoder.layer):
        if bert.encoder.num2type[n] == "attn":
            x, saved = attention_fwd(rt, layer, x, masks)
        else:
            x, saved = ffn_fwd(rt, layer, x)
        last_plan["row_split"] += int(bool(saved.get("split")))
        last_plan["joint"] += int(bool(saved.get("joint") or saved.get("all")))
        tape.append(saved if rt.grad_enabled else None)
        if keep_layers:
            layers.append(x)
    if rt.pack is not None:
        _leave_packed(rt)
        x = _unpack_pair(rt, x)
    pooled_t = pooled_v = None
    s_pool = {}
    fm = cfg.fusion_method
    if fm == "vl-bert_vqa":
        end = ((ids != 0).sum(1) - 2).clamp(min=0) + torch.arange(rt.B, device=rt.device) * rt.T
        rows = rt.empty(rt.B, x.t.shape[1])
        ops.gather_rows(x.t, end, rows)
        _count()
        pooled_t, a = _pool_fwd(rt, bert.t_pooler, x.t, rt.T, rows=rows)
        s_pool["t"] = (pooled_t, a, end)
    elif fm != "none":
        pooled_t, a = _pool_fwd(rt, bert.t_pooler, x.t, rt.T)
        s_pool["t"] = (pooled_t, a, None)
    if fm not in ("none", "text", "vl-bert_vqa"):
        pooled_v, a = _pool_fwd(rt, bert.v_pooler, x.v, rt.V)
        s_pool["v"] = (pooled_v, a, None)
    return x, pooled_t, pooled_v, layers, dict(emb=s_emb, tape=tape, pool=s_pool, x=x, pack=rt.pack, lens=rt.lens)


def bert_bwd(rt, bert, s, dx, dpooled_t, dpooled_v):
    """dx: Pair of gradients w.r.t. the final hidden states (modified in place)."""
    if dpooled_t is not None and "t" in s["pool"]:
        pooled, a, end = s["pool"]["t"]
        if end is None:
            _pool_bwd(rt, bert.t_pooler, dpooled_t, pooled, a, dx.t, rt.T)
        else:
            dpre = rt.empty(*pooled.shape)
            ops.eltwise(ELT_DRELU, dpooled_t, dpre, b=pooled)
            drows = linear_bwd(rt, a, dpre, bert.t_pooler.dense)
            ops.scatter_rows_add(drows, end, dx.t)
            _count(2)
    if dpooled_v is not None and "v" in s["pool"]:
        pooled, a, _ = s["pool"]["v"]
        _pool_bwd(rt, bert.v_pooler, dpooled_v, pooled, a, dx.v, rt.V)
    if hasattr(bert, "t_pooler"):
        _notify_progress(rt, bert.t_pooler)
    rt.pack, rt.pack_on = s.get("pack"), False
    if rt.pack is not None:       # the encoder ran on packed rows: so does its backward
        rt.lens = s["lens"]
        _enter_packed(rt)
        dx = _pack_pair(rt, dx)
        rt.pack_on = True
    for n in range(len(bert.encoder.layer) - 1, -1, -1):
        layer = bert.encoder.layer[n]
        if bert.encoder.num2type[n] == "attn":
            dx = attention_bwd(rt, layer, dx, s["tape"][n])
        else:
            dx = ffn_bwd(rt, layer, dx, s["tape"][n])
        s["tape"][n] = None  # free activations as the backward retires them
        _notify_progress(rt, layer)
    if rt.pack is not None:
        _leave_packed(rt)
        dx = _unpack_pair(rt, dx)
    embeddings_bwd(rt, bert, dx, s["emb"])


# ====================================================================================================
# pre-training heads + losses
# ====================================================================================================
def _lim_seqs(lim, rows):
    """vb_row_seqs for a buffer whose first lim[0] rows hold data (compacted head rows)."""
    return None if lim is None else (lim[0:1], rows, rows, None, 0)


def _transform_fwd(rt, tr, x, has_ln=True, lim=None):
    """lim: int32 [count, padded count] on the device - only the first `count` rows of x hold data."""
    H = tr.dense.weight.shape[0]
    h = rt.empty(x.shape[0], H)
    ml = dict(m_limit=lim[0:1]) if lim is not None else {}
    pre = _act_linear_fwd(rt, x, tr.dense, tr.hidden_act, h, rt.empty(x.shape[0], H) if tr.hidden_act != "relu" else None, **ml)
    if not has_ln:
        return h, (x, pre, h, None)
    y, stats = ln_fwd(rt, h, tr.LayerNorm, seqs=_lim_seqs(lim, x.shape[0]))
    return y, (x, pre, h, stats)


def _transform_bwd(rt, tr, dy, saved, dx_out, accumulate_dx=False, lim=None):
    x, pre, h, stats = saved
    _, epi_b = _act_epilogues(tr.hidden_act)
    assert lim is None or stats is not None     # the LayerNorm backward is what zeroes the rows beyond the count
    dh = ln_bwd(rt, dy, h, stats, tr.LayerNorm, seqs=_lim_seqs(lim, h.shape[0])) if stats is not None else dy
    dpre = rt.empty(*dh.shape)
    ops.eltwise(ELT_MUL if epi_b == EPI_MUL else ELT_DRELU, dh, dpre, b=pre)
    _count()
    if accumulate_dx:
        linear_bwd(rt, x, dpre, tr.dense, dx_epilogue=EPI_ADD, dx_out=dx_out, aux_in=dx_out, lim=lim)
    else:
        linear_bwd(rt, x, dpre, tr.dense, dx_out=dx_out, lim=lim)


def heads_fwd(rt, model, x, pooled_t, pooled_v, labels, want_logits=False):
    """BertPreTrainingHeads + the three criteria (reference: volta/encoders.py:766-784, 1079-1109)."""
    cfg, cls = model.config, model.cls
    s = {}
    V_, C_ = cfg.vocab_size, None
    # --- MLM: transform -> tied decoder
    E = rt.cw(cls.predictions.decoder.weight)
    ldv = _pad8(V_)
    mlm_logits = rt.empty(rt.Mt, ldv)
    cmp_ = None
    if MLM_COMPACT and not want_logits and labels.get("mlm") is not None:
        # The loss ignores every position whose label is -1 (CrossEntropyLoss(ignore_index=-1), reference:
        # volta/encoders.py:1030, 1095-1099) - ~90 % of the B*T rows.  Pack the labelled rows first and run the decoder /
        # criterion on `count` rows; the count stays on the device (static launches, CUDA-graph friendly).  Same loss,
        # same gradients as the reference's all-positions decoder (:686-696).
        lab_full = labels["mlm"].reshape(-1).to(torch.int64).contiguous()
        limits = rt.empty(2, dtype=torch.int32)          # [count, count rounded up to 256]
        src_index, pos = rt.empty(rt.Mt, dtype=torch.int32), rt.empty(rt.Mt, dtype=torch.int32)
        lab_c = rt.empty(rt.Mt, dtype=torch.int64)
        ops.compact_plan(lab_full, 256, limits, src_index, pos, lab_c)
        xc = rt.empty(rt.Mt, x.t.shape[1])
        ops.compact_gather(x.t, src_index, limits, xc)
        _count(2)
        hn, s["mlm_tr"] = _transform_fwd(rt, cls.predictions.transform, xc, lim=limits)   # transform of the labelled rows only
        rt.gemm(hn, E, mlm_logits, epilogue=EPI_BIAS, bias=rt.w32(cls.predictions.bias), N=V_, m_limit=limits[0:1])
        cmp_ = (limits, pos, lab_c)
        last_plan["mlm_compact"] = True
    else:
        hn, s["mlm_tr"] = _transform_fwd(rt, cls.predictions.transform, x.t)
        rt.gemm(hn, E, mlm_logits, epilogue=EPI_BIAS, bias=rt.w32(cls.predictions.bias), N=V_)
    s["mlm"] = (hn, mlm_logits)
    s["mlm_compact"] = cmp_
    # --- ITM
    itm_logits, pooled = None, None
    fm = cfg.fusion_method
    if fm != "none":
        if fm in ("sum", "mul"):
            pooled = rt.empty(*pooled_t.shape)
            ops.eltwise(ELT_ADD if fm == "sum" else ELT_MUL, pooled_t, pooled, b=pooled_v)
            _count()
        elif fm in ("text", "vl-bert_vqa"):
            pooled = rt.empty(*pooled_t.shape)
            ops.eltwise(ELT_COPY, pooled_t, pooled)
            _count()
        else:
            raise AssertionError("unknown fusion_method %r" % fm)
        s["pool_drop"] = (rt.p(cls.dropout.p), rt.seed())
        dropout_(rt, pooled, *s["pool_drop"])
        if cls.has_itm:
            itm_logits = rt.empty(rt.B, 8)[:, :2]
            linear_fwd(rt, pooled, cls.bi_seq_relationship, out=itm_logits)
        s["itm"] = (pooled, itm_logits, pooled_t, pooled_v)
    # --- MRM
    ip = cls.imagePredictions
    img_logits = {}
    img_cmp = None
    active = [ix for ix, w in cfg.visual_target_weights.items() if w > 0]
    if (MLM_COMPACT and IMG_COMPACT and not want_logits and active == ["0"] and "0" in ip.decoder_dict and ip.transform.has_ln and
            labels.get("image_label") is not None and labels.get("image_cls") is not None):
        # Masked-region head on the masked regions only (label == 1; kl_1601 ignores the rest, volta/losses.py:16-22):
        # the same compaction as the masked-LM head, over the region index r = b*per + j <-> hidden row b*V + j + skip.
        per = labels["image_label"].shape[1]
        skip = 1 if cfg.add_global_imgfeat == "first" else 0
        R = rt.B * per
        keys = (labels["image_label"].reshape(-1) == 1).to(torch.int64) - 1
        lim = rt.empty(2, dtype=torch.int32)
        src_r, pos_r = rt.empty(R, dtype=torch.int32), rt.empty(R, dtype=torch.int32)
        ops.compact_plan(keys.contiguous(), 256, lim, src_r, pos_r, rt.empty(R, dtype=torch.int64))
        src_h = (src_r + torch.div(src_r, per, rounding_mode="floor") * (rt.V - per) + skip).to(torch.int32)
        xc = rt.empty(R, x.v.shape[1])
        ops.compact_gather(x.v, src_h, lim, xc)
        hv, s["img_tr"] = _transform_fwd(rt, ip.transform, xc, lim=lim)
        dec = ip.decoder_dict["0"]
        n_cls = dec.weight.shape[0]
        lg = rt.empty(R, _pad8(n_cls))
        linear_fwd(rt, hv, dec, out=lg, N=n_cls, m_limit=lim[0:1])
        tgt = labels["image_cls"].reshape(-1, labels["image_cls"].shape[-1]).float().contiguous()
        tgt_c = rt.empty(R, tgt.shape[1], dtype=torch.float32)
        ops.compact_gather(tgt, src_r, lim, tgt_c)
        lab_c = (torch.arange(R, device=rt.device, dtype=torch.int32) < lim[0:1]).to(torch.int64) * 2 - 1   # 1 below the count, else -1
        _count(3)
        img_logits["0"] = lg
        img_cmp = dict(lim=lim, pos_r=pos_r, per=per, skip=skip, R=R, lab_c=lab_c, tgt_c=tgt_c)
    else:
        hv, s["img_tr"] = _transform_fwd(rt, ip.transform, x.v, has_ln=ip.transform.has_ln)
        for ix, dec in ip.decoder_dict.items():
            n_cls = dec.weight.shape[0]
            lg = rt.empty(rt.Mv, _pad8(n_cls))
            linear_fwd(rt, hv, dec, out=lg, N=n_cls)
            img_logits[ix] = lg
    s["img"] = (hv, img_logits)
    s["img_compact"] = img_cmp
    # --- criteria
    sums = rt.zeros(8, dtype=torch.float32)  # [sum_mlm, sum_img, sum_itm, -, cnt_mlm, cnt_img, cnt_itm, -]
    losses = rt.empty(4, dtype=torch.float32)
    s["sums"] = sums
    if labels.get("mlm") is not None:
        lse = rt.empty(rt.Mt, dtype=torch.float32)
        if cmp_ is not None:
            lab = cmp_[2]
            ops.xent_fwd(mlm_logits, lab, V_, lse, sums[0:1], sums[4:5], row_limit=cmp_[0][0:1])
        else:
            lab = labels["mlm"].reshape(-1).to(torch.int64).contiguous()
            ops.xent_fwd(mlm_logits, lab, V_, lse, sums[0:1], sums[4:5])
        _count()
        s["mlm_loss"] = (lab, lse)
    if labels.get("image_label") is not None:
        # masked-region criteria (reference: volta/encoders.py:1079-1087 + volta/losses.py); they share the denominator
        # max(#label==1, 1), so all accumulate into one (sum, count) pair and exactly one of them counts the regions
        per = labels["image_label"].shape[1]
        skip = 1 if cfg.add_global_imgfeat == "first" else 0
        lab = labels["image_label"].reshape(-1).to(torch.int64).contiguous()
        s["img_losses"] = []
        for ix, w in cfg.visual_target_weights.items():
            if w <= 0:
                continue
            n_cls = pre_vis_targets[ix]
            rec = None
            if ix == "0":
                if labels.get("image_cls") is None:
                    continue
                tgt = labels["image_cls"].reshape(-1, labels["image_cls"].shape[-1]).float().contiguous()
                rec = dict(kind="kl", tgt=tgt, w=w)
            elif ix in ("3", "4", "6"):
                names = {"3": ("obj_labels", "obj_confs"), "4": ("attr_labels", "attr_confs"), "6": ("obj_labels", None)}[ix]
                if labels.get(names[0]) is None or (names[1] is not None and labels.get(names[1]) is None):
                    continue
                conf = labels[names[1]].reshape(-1).float().contiguous() if names[1] else None
                rec = dict(kind=_lib.REGION_XENT, idx=labels[names[0]].reshape(-1).to(torch.int64).contiguous(), conf=conf, vec=None, w=w)
            elif ix in ("1", "5"):
                if labels.get("image_feat") is None:
                    continue
                if skip != 0 or labels["image_feat"].shape[1] != per:
                    raise ValueError("feature-regression targets need as many regions as predictions (no global feature), "
                                     "as in the reference (volta/losses.py:25-33)")
                vec = labels["image_feat"].reshape(-1, n_cls).float().contiguous()
                rec = dict(kind=_lib.REGION_MSE if ix == "1" else _lib.REGION_HUBER, idx=None, conf=None, vec=vec, w=w / n_cls)
            else:
                raise NotImplementedError("visual target 2 (nce_2048, volta/losses.py:36-84) draws its negatives with the host "
                                          "RNG inside the criterion; it is not on the device path")
            lse = rt.empty(lab.numel(), dtype=torch.float32)
            first = not s["img_losses"]
            if img_cmp is not None:      # compacted rows: prediction row r is logits row r
                R = img_cmp["R"]
                ops.kl_fwd(img_logits[ix], R, R, 0, img_cmp["lab_c"], img_cmp["tgt_c"], n_cls, w, lse, sums[1:2], sums[5:6])
                _count()
                rec.update(ix=ix, per=R, per_ld=R, skip=0, lab=img_cmp["lab_c"], tgt=img_cmp["tgt_c"], lse=lse, n_cls=n_cls)
                s["img_losses"].append(rec)
                continue
            if rec["kind"] == "kl":
                # kl_fwd always counts: make it the counting criterion by running it first
                ops.kl_fwd(img_logits[ix], per, rt.V, skip, lab, rec["tgt"], n_cls, w, lse, sums[1:2], sums[5:6] if first else rt.zeros(1, dtype=torch.float32))
            else:
                ops.region_loss_fwd(rec["kind"], img_logits[ix], per, rt.V, skip, lab, rec["idx"], rec["conf"], rec["vec"], n_cls,
                                    rec["w"], lse, sums[1:2], sums[5:6], first)
            _count()
            rec.update(ix=ix, per=per, skip=skip, lab=lab, lse=lse, n_cls=n_cls)
            s["img_losses"].append(rec)
    if labels.get("itm") is not None and itm_logits is not None:
        lab = labels["itm"].reshape(-1).to(torch.int64).contiguous()
        lse = rt.empty(rt.B, dtype=torch.float32)
        ops.xent_fwd(itm_logits, lab, 2, lse, sums[2:3], sums[6:7])
        _count()
        s["itm_loss"] = (lab, lse)
    ops.loss_finalize(sums[0:4], sums[4:8], losses)
    _count()
    return losses, (mlm_logits, img_logits, itm_logits, pooled), s


def heads_bwd(rt, model, s, dlosses, x):
    """dlosses: fp32 [4] device tensor of upstream gradients (mlm, img, itm).  Returns (dx Pair, dpooled_t, dpooled_v)."""
    cfg, cls = model.config, model.cls
    sums = s["sums"]
    dx = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1])
    # --- MLM
    hn, mlm_logits = s["mlm"]
    if "mlm_loss" in s:
        lab, lse = s["mlm_loss"]
        cmp_ = s.get("mlm_compact")
        E = cls.predictions.decoder
        dlog = mlm_logits[:, :cfg.vocab_size]
        dhn = rt.empty(*hn.shape)
        if cmp_ is not None:
            limits, pos, _ = cmp_
            n, npad = limits[0:1], limits[1:2]     # rows [count, npad) are zero rows: they complete the last tile / k-block
            ops.xent_bwd(mlm_logits, lab, cfg.vocab_size, lse, sums[4:5], dlosses[0:1], mlm_logits, row_limit=npad)  # in place
            rt.gemm(dlog, hn, rt.g32(E.weight), a_kmajor=False, b_kmajor=False, epilogue=EPI_ACCUM, k_limit=npad)
            ops.colsum(dlog, rt.g32(cls.predictions.bias), row_limit=npad)
            # d(hidden) = dlogits . E over K = 30522 with only ~count/256 x 3 output tiles: split the reduction (fp32
            # accumulate) so the whole machine works on it, then round to the activation dtype
            dhc32 = rt.zeros(*hn.shape, dtype=torch.float32)
            rt.gemm(dlog, rt.cw(E.weight), dhc32, a_kmajor=True, b_kmajor=False, epilogue=EPI_ACCUM, split_k=16, m_limit=n)
            ops.eltwise(ELT_COPY, dhc32, dhn)
            dxc = rt.empty(*dx.t.shape)
            _transform_bwd(rt, cls.predictions.transform, dhn, s["mlm_tr"], dxc, lim=limits)
            ops.compact_scatter(dxc, pos, dx.t)
            _count(4)
        else:
            ops.xent_bwd(mlm_logits, lab, cfg.vocab_size, lse, sums[4:5], dlosses[0:1], mlm_logits)  # in place
            rt.gemm(dlog, hn, rt.g32(E.weight), a_kmajor=False, b_kmajor=False, epilogue=EPI_ACCUM)
            ops.colsum(dlog, rt.g32(cls.predictions.bias))
            _count(2)
            rt.gemm(dlog, rt.cw(E.weight), dhn, a_kmajor=True, b_kmajor=False)
            _transform_bwd(rt, cls.predictions.transform, dhn, s["mlm_tr"], dx.t)
    else:
        ops.fill(dx.t, 0)
    # --- MRM: every active criterion back-propagates through its own decoder into the shared transform output
    hv, img_logits = s["img"]
    if s.get("img_losses"):
        dhv = None
        for rec in s["img_losses"]:
            ix, lg, n_cls = rec["ix"], img_logits[rec["ix"]], rec["n_cls"]
            if rec["kind"] == "kl":
                ops.kl_bwd(lg, rec["per"], rec.get("per_ld", rt.V), rec["skip"], rec["lab"], rec["tgt"], n_cls, rec["w"], rec["lse"], sums[5:6],
                           dlosses[1:2], lg)
            else:
                ops.region_loss_bwd(rec["kind"], lg, rec["per"], rt.V, rec["skip"], rec["lab"], rec["idx"], rec["conf"], rec["vec"], n_cls,
                                    rec["w"], rec["lse"], sums[5:6], dlosses[1:2], lg)  # in place
            _count()
            dec = cls.imagePredictions.decoder_dict[ix]
            lim = s["img_compact"]["lim"] if s.get("img_compact") else None
            if dhv is None:
                dhv = linear_bwd(rt, hv, lg[:, :n_cls], dec, lim=lim)
            else:
                linear_bwd(rt, hv, lg[:, :n_cls], dec, dx_epilogue=EPI_ADD, dx_out=dhv, aux_in=dhv, lim=lim)
        if s.get("img_compact"):
            ic = s["img_compact"]
            dxc = rt.empty(ic["R"], dx.v.shape[1])
            _transform_bwd(rt, cls.imagePredictions.transform, dhv, s["img_tr"], dxc, lim=ic["lim"])
            pos_h = torch.full((rt.B, rt.V), -1, dtype=torch.int32, device=rt.device)
            pos_h[:, ic["skip"]:ic["skip"] + ic["per"]] = ic["pos_r"].view(rt.B, ic["per"])
            ops.compact_scatter(dxc, pos_h.reshape(-1), dx.v)
            _count()
        else:
            _transform_bwd(rt, cls.imagePredictions.transform, dhv, s["img_tr"], dx.v)
    else:
        ops.fill(dx.v, 0)
    # --- ITM
    dpt = dpv = None
    if "itm_loss" in s:
        pooled, itm_logits, pooled_t, pooled_v = s["itm"]
        lab, lse = s["itm_loss"]
        ops.xent_bwd(itm_logits, lab, 2, lse, sums[6:7], dlosses[2:3], itm_logits)
        _count()
        dpooled = linear_bwd(rt, pooled, itm_logits, cls.bi_seq_relationship)
        dropout_(rt, dpooled, *s["pool_drop"])
        fm = cfg.fusion_method
        if fm == "mul":
            dpt, dpv = rt.empty(*dpooled.shape), rt.empty(*dpooled.shape)
            ops.eltwise(ELT_MUL, dpooled, dpt, b=pooled_v)
            ops.eltwise(ELT_MUL, dpooled, dpv, b=pooled_t)
            _count(2)
        elif fm == "sum":
            dpt = dpv = dpooled
        else:
            dpt = dpooled
    return dx, dpt, dpv


# ====================================================================================================
# autograd boundary
# ====================================================================================================
def _anchor(model):
    for p in model.parameters():
        if p.requires_grad:
            return p
    return None


def _check_inputs(input_txt, input_imgs, image_loc):
    ops.require_cuda(input_txt, input_imgs, image_loc)
    if input_txt.dim() != 2 or input_imgs.dim() != 3 or image_loc.dim() != 3:
        raise ValueError("expected input_txt [B,T], input_imgs [B,V,F], image_loc [B,V,num_locs]")
    _lib.check(_lib.load().vb_device_check(), "vb_device_check")


class _PretrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, model, args):
        (ids, feat, loc, seg, mask_t, mask_v, mlm_labels, image_label, image_cls, itm_label, obj_labels, obj_confs,
         attr_labels, attr_confs) = args
        B, T = ids.shape
        V = feat.shape[1]
        rt = Runtime(model, ids.device, B, T, V, model.training, grad=anchor is not None)
        if anchor is not None and PREZERO_GRADS:
            rt.store.prezero_grads()
        seg = torch.zeros_like(ids) if seg is None else seg
        x, pt, pv, _, s_bert = bert_fwd(rt, model.bert, ids.contiguous(), feat.contiguous(), loc, seg.contiguous(), mask_t, mask_v)
        labels = dict(mlm=mlm_labels, image_label=image_label, image_cls=image_cls, itm=itm_label, obj_labels=obj_labels,
                      obj_confs=obj_confs, attr_labels=attr_labels, attr_confs=attr_confs, image_feat=feat)
        losses, logits, s_heads = heads_fwd(rt, model, x, pt, pv, labels)
        ctx.rt, ctx.model, ctx.s_bert, ctx.s_heads, ctx.x = rt, model, s_bert, s_heads, x
        ctx.logits = logits
        return losses[0:1], losses[1:2], losses[2:3]

    @staticmethod
    def backward(ctx, g_mlm, g_img, g_itm):
        rt, model = ctx.rt, ctx.model
        rt.store.prepare_grads()
        dl = rt.zeros(4, dtype=torch.float32)
        for i, g in enumerate((g_mlm, g_img, g_itm)):
            if g is not None:
                ops.eltwise(ELT_COPY, g.reshape(1, 1).float(), dl[i:i + 1].view(1, 1))
                _count()
        with backward_sm_limit(model):
            dx, dpt, dpv = heads_bwd(rt, model, ctx.s_heads, dl, ctx.x)
            # head gradients are final (the tied decoder/embedding weight lives with the embeddings, at offset 0)
            _notify_progress(rt, model.cls, skip_below=_frontier(rt.store, model.bert.encoder) or 0)
            bert_bwd(rt, model.bert, ctx.s_bert, dx, dpt, dpv)
        _notify_grads_ready(model)
        ctx.s_bert = ctx.s_heads = ctx.x = ctx.logits = None
        return None, None, None


_GRAD_HOOKS = {}
_PROGRESS_HOOKS = {}


_PROGRESS_CUTS = {}
_BACKWARD_SM_LIMIT = {}   # id(model) -> SMs the GEMMs may use during the backward (data-parallel: the rest is NCCL's)


def set_backward_sm_limit(model, sms):
    _BACKWARD_SM_LIMIT[id(model)] = int(sms)


class backward_sm_limit:
    """Context: GEMM launches issued (or captured) inside use at most the model's backward SM limit."""

    def __init__(self, model):
        self.sms = _BACKWARD_SM_LIMIT.get(id(model_store(model).model), 0)

    def __enter__(self):
        if self.sms:
            _lib.load().vb_set_gemm_sm_limit(self.sms)

    def __exit__(self, *a):
        if self.sms:
            _lib.load().vb_set_gemm_sm_limit(0)


def register_progress_hook(model, fn, cut_fn_factory=None):
    """fn(frontier_offset): every gradient at flat offset >= frontier_offset is final (backward runs from the
    heads towards the embeddings, i.e. from the end of the flat buffer to its start).
    cut_fn_factory() -> cut(frontier) -> bool tells the graph executor at which notifications the hook will act
    (issue a collective), i.e. where the captured backward has to be cut into separately replayed segments."""
    _PROGRESS_HOOKS[id(model)] = fn
    _PROGRESS_CUTS[id(model)] = cut_fn_factory


def progress_cut_fn(model):
    f = _PROGRESS_CUTS.get(id(model_store(model).model))
    return f() if f is not None else None


def _frontier(store, module, skip_below=0):
    key = ("frontier", id(module), skip_below)
    cache = store.__dict__.setdefault("_frontier_cache", {})
    if key not in cache:
        offs = [store.offsets[id(p)][0] for p in module.parameters() if store.offsets[id(p)][0] >= skip_below]
        cache[key] = min(offs) if offs else None
    return cache[key]


def _notify_progress(rt, module, skip_below=0):
    if rt.quiet:
        return
    fn = rt.progress_override or _PROGRESS_HOOKS.get(id(rt.store.model))
    if fn is None:
        return
    off = _frontier(rt.store, module, skip_below)
    if off is not None:
        fn(off)


def register_grad_ready_hook(model, fn):
    """Called once per backward after all parameter gradients are written (used by the data-parallel wrapper)."""
    _GRAD_HOOKS[id(model)] = fn


def _notify_grads_ready(model):
    """Hooks are keyed by the model that owns the flat store (the object the data-parallel wrapper wraps): a BertModel
    inside BertForVLPreTraining / BertForVLTasks resolves to its parent."""
    fn = _GRAD_HOOKS.get(id(model_store(model).model))
    if fn is not None:
        fn()


def pretraining_forward(model, input_ids, image_feat, image_loc, token_type_ids, attention_mask, image_attention_mask,
                        masked_lm_labels, image_label, image_cls, obj_labels, obj_confs, attr_labels, attr_confs,
                        image_attrs, next_sentence_label, output_all_attention_masks):
    _check_inputs(input_ids, image_feat, image_loc)
    args = (input_ids, image_feat, image_loc, token_type_ids, attention_mask, image_attention_mask, masked_lm_labels,
            image_label, image_cls, next_sentence_label, obj_labels, obj_confs, attr_labels, attr_confs)
    no_labels = masked_lm_labels is None and image_label is None and next_sentence_label is None
    if no_labels:
        with torch.no_grad():
            B, T = input_ids.shape
            rt = Runtime(model, input_ids.device, B, T, image_feat.shape[1], model.training)
            seg = torch.zeros_like(input_ids) if token_type_ids is None else token_type_ids
            x, pt, pv, _, _ = bert_fwd(rt, model.bert, input_ids.contiguous(), image_feat.contiguous(), image_loc,
                                       seg.contiguous(), attention_mask, image_attention_mask,
                                       want_attn=bool(output_all_attention_masks))
            _, (mlm, img, itm, pooled), _ = heads_fwd(rt, model, x, pt, pv, {})
        V_ = model.config.vocab_size
        img = {ix: v.view(B, -1, v.shape[-1])[..., :pre_vis_targets[ix]].float() for ix, v in img.items()}
        # the reference returns ([], []) when no attention data were requested (encoders.py:865-875, 1114)
        attn = ([a for a, _ in rt.attn_taps], [b for _, b in rt.attn_taps]) if output_all_attention_masks else ([], [])
        return (mlm.view(B, T, -1)[..., :V_].float(), img, None if itm is None else itm.float(), attn,
                None if pooled is None else pooled.float())
    anchor = _anchor(model) if torch.is_grad_enabled() else None
    if anchor is None:
        with torch.no_grad():
            return _PretrainFn.apply(None, model, args)
    from . import graphs
    out = graphs.maybe_run(model, anchor, args)
    if out is not None:
        return out
    return _PretrainFn.apply(anchor, model, args)


class _BertFn(torch.autograd.Function):
    """BertModel alone: differentiable w.r.t. parameters through (seq_t, seq_v, pooled_t, pooled_v)."""

    @staticmethod
    def forward(ctx, anchor, model, args, keep_layers, attn_holder=None):
        ids, feat, loc, seg, mask_t, mask_v = args
        B, T = ids.shape
        rt = Runtime(model, ids.device, B, T, feat.shape[1], model.training, grad=anchor is not None)
        seg = torch.zeros_like(ids) if seg is None else seg
        x, pt, pv, layers, s = bert_fwd(rt, model, ids.contiguous(), feat.contiguous(), loc, seg.contiguous(), mask_t, mask_v,
                                        keep_layers=keep_layers, want_attn=attn_holder is not None)
        if attn_holder is not None:      # attention data for inspection: not differentiable, handed over on the side
            attn_holder["taps"] = rt.attn_taps
        ctx.rt, ctx.model, ctx.s = rt, model, s
        ctx.n_layers = len(layers)
        outs = [x.t.view(B, T, -1), x.v.view(B, rt.V, -1)]
        dummy = rt.empty(1)
        outs += [pt if pt is not None else dummy, pv if pv is not None else dummy]
        ctx.has_pool = (pt is not None, pv is not None)
        for l in layers:
            outs += [l.t.view(B, T, -1), l.v.view(B, rt.V, -1)]
        ctx.mark_non_differentiable(*outs[4:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_t, g_v, g_pt, g_pv, *rest):
        rt, model, s = ctx.rt, ctx.model, ctx.s
        rt.store.prepare_grads()
        x = s["x"]
        dx = Pair.alloc(rt, rt.Mt, x.t.shape[1], rt.Mv, x.v.shape[1])
        for g, d in ((g_t, dx.t), (g_v, dx.v)):
            if g is None:
                ops.fill(d, 0)
            else:
                ops.eltwise(ELT_COPY, g.reshape(d.shape).contiguous(), d)
            _count()

        def conv(g, on):
            if g is None or not on:
                return None
            out = rt.empty(*g.shape)
            ops.eltwise(ELT_COPY, g.contiguous(), out)
            return out
        bert_bwd(rt, model, s, dx, conv(g_pt, ctx.has_pool[0]), conv(g_pv, ctx.has_pool[1]))
        _notify_grads_ready(model)
        ctx.s = None
        return None, None, None, None, None


def _run_bert(model, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask, keep_layers,
              attn_holder=None):
    _check_inputs(input_txt, input_imgs, image_loc)
    args = (input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask)
    anchor = _anchor(model) if torch.is_grad_enabled() else None
    if anchor is None:
        with torch.no_grad():
            outs = _BertFn.apply(None, model, args, keep_layers, attn_holder)
    else:
        outs = _BertFn.apply(anchor, model, args, keep_layers, attn_holder)
    cfg = model.config
    seq_t, seq_v = outs[0], outs[1]
    pt = outs[2] if cfg.fusion_method != "none" else None
    pv = outs[3] if cfg.fusion_method not in ("none", "text", "vl-bert_vqa") else None
    layers = outs[4:]
    return seq_t, seq_v, pt, pv, layers


def bert_model_forward(model, input_txt, input_imgs, image_loc, token_type_ids, attention_mask, image_attention_mask,
                       output_all_encoded_layers, output_all_attention_masks):
    """BertModel.forward contract (reference: volta/encoders.py:954-1017).  Hidden states come back in the
    compute dtype (bf16 unless the fp32 validation mode is on)."""
    holder = {} if output_all_attention_masks else None
    seq_t, seq_v, pt, pv, layers = _run_bert(model, input_txt, input_imgs, image_loc, token_type_ids, attention_mask,
                                             image_attention_mask, output_all_encoded_layers, holder)
    # per attention sub-layer: the (t_attn_data, v_attn_data) dicts of the reference when config.visualization is set,
    # None otherwise (encoders.py:342-356, 873-875); ([], []) when not requested
    attn = ([a for a, _ in holder["taps"]], [b for _, b in holder["taps"]]) if holder is not None else ([], [])
    if output_all_encoded_layers:
        return list(layers[0::2]), list(layers[1::2]), pt, pv, attn
    return seq_t, seq_v, pt, pv, attn


# ====================================================================================================
# small differentiable ops for the task heads (thin autograd.Functions over single kernels)
# ====================================================================================================
def _cd(model_or_dtype):
    return model_or_dtype


class LayerNormFn(torch.autograd.Function):
    """Stand-alone LayerNorm on a CUDA tensor - the reference-facing replacement of apex
    `fused_layer_norm_cuda.forward_affine/backward_affine` (apex/apex/normalization/fused_layer_norm.py:12-37)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        ops.require_cuda(x, weight, bias)
        xc = x.contiguous()
        y = torch.empty_like(xc)
        rows = xc.numel() // xc.shape[-1]
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        w32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        ops.layernorm_fwd(xc, w32, b32, y, mean, rstd, eps)
        _count()
        ctx.save_for_backward(xc, mean, rstd, w32)
        ctx.wdtype = weight.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, mean, rstd, w32 = ctx.saved_tensors
        dx = torch.empty_like(xc)
        dg = torch.zeros_like(w32)
        db = torch.zeros_like(w32)
        ops.layernorm_bwd(dy.contiguous().to(xc.dtype), xc, mean, rstd, w32, dx, dg, db)
        _count()
        return dx, dg.to(ctx.wdtype), db.to(ctx.wdtype), None


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ops.require_cuda(x)
        xc = x.contiguous()
        y = torch.empty_like(xc)
        ops.eltwise(ELT_GELU, xc, y)
        _count()
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        dx = torch.empty_like(xc)
        ops.eltwise(ELT_DGELU, dy.contiguous(), dx, b=xc)
        _count()
        return dx


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) for the small task heads; weights are used in the activation dtype."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        ops.require_cuda(x, weight)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w = weight.detach().to(x2.dtype).contiguous()
        N = w.shape[0]
        out = torch.empty(x2.shape[0], _pad8(N), dtype=x2.dtype, device=x.device)[:, :N]
        b32 = bias.detach().float().contiguous() if bias is not None else None
        epi = (EPI_BIAS_RELU if relu else EPI_BIAS) if b32 is not None else EPI_NONE
        ops.gemm(x2, w, out, epilogue=epi, bias=b32)
        _count()
        ctx.save_for_backward(x2, w, out if relu else None)
        ctx.meta = (x.shape, weight.dtype, bias is not None, relu)
        return out.reshape(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, dy):
        x2, w, out = ctx.saved_tensors
        xshape, wdtype, has_bias, relu = ctx.meta
        N = w.shape[0]
        d2 = torch.empty(x2.shape[0], _pad8(N), dtype=x2.dtype, device=x2.device)[:, :N]
        if relu:
            ops.eltwise(ELT_DRELU, dy.reshape(-1, N).contiguous().to(x2.dtype), d2, b=out)
        else:
            ops.eltwise(ELT_COPY, dy.reshape(-1, N).contiguous(), d2)
        dW = torch.zeros(w.shape, dtype=torch.float32, device=w.device)
        ops.gemm(d2, x2, dW, a_kmajor=False, b_kmajor=False, epilogue=EPI_ACCUM)
        db = None
        if has_bias:
            db = torch.zeros(N, dtype=torch.float32, device=w.device)
            ops.colsum(d2, db)
        dx = torch.empty_like(x2)
        ops.gemm(d2, w, dx, a_kmajor=True, b_kmajor=False)
        _count(4)
        return dx.reshape(xshape), dW.to(wdtype), None if db is None else db.to(wdtype), None


class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        y = torch.empty_like(x.contiguous())
        ops.eltwise(ELT_DROPOUT, x.contiguous(), y, p=p, seed=seed)
        _count()
        ctx.ps = (p, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        dx = torch.empty_like(dy.contiguous())
        ops.eltwise(ELT_DROPOUT, dy.contiguous(), dx, p=ctx.ps[0], seed=ctx.ps[1])
        _count()
        return dx, None, None


class MulFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        y = torch.empty_like(a)
        ops.eltwise(ELT_MUL, a, y, b=b)
        _count()
        ctx.save_for_backward(a, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        dy = dy.contiguous()
        da, db = torch.empty_like(a), torch.empty_like(b)
        ops.eltwise(ELT_MUL, dy, da, b=b)
        ops.eltwise(ELT_MUL, dy, db, b=a)
        _count(2)
        return da, db


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a = a.contiguous()
        y = torch.empty_like(a)
        ops.eltwise(ELT_ADD, a, y, b=b.contiguous())
        _count()
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def _dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    return DropoutFn.apply(x, float(p), int(torch.randint(0, 2 ** 62, (1,)).item()))


def _apply_head(mod, x, training):
    """Runs a task classifier (nn.Linear / SimpleClassifier / the 2-layer V-logit Sequential) with our kernels."""
    from .encoders import BertLayerNorm, GeLU, SimpleClassifier
    from torch import nn
    if isinstance(mod, SimpleClassifier):
        return _apply_head(mod.logit_fc, x, training)
    if isinstance(mod, nn.Sequential):
        for m in mod:
            x = _apply_head(m, x, training)
        return x
    if isinstance(mod, nn.Linear):
        return LinearFn.apply(x, mod.weight, mod.bias, False)
    if isinstance(mod, GeLU):
        return GeluFn.apply(x)
    if isinstance(mod, BertLayerNorm):
        return mod(x)
    if isinstance(mod, nn.Dropout):
        return _dropout(x, mod.p, training)
    raise TypeError("unsupported head module %r" % type(mod))


def tasks_forward(model, input_txt, input_imgs, image_loc, task_id, token_type_ids, attention_mask,
                  image_attention_mask, output_all_encoded_layers, output_all_attention_masks):
    """BertForVLTasks.forward (reference: volta/encoders.py:1158-1206)."""
    seq_t, seq_v, pt, pv, all_masks = bert_model_forward(model.bert, input_txt, input_imgs, image_loc, token_type_ids,
                                                         attention_mask, image_attention_mask,
                                                         output_all_encoded_layers, output_all_attention_masks)
    fm = model.fusion_method
    if fm == "sum":
        pooled = _dropout(AddFn.apply(pt, pv), model.dropout.p, model.training)
    elif fm == "mul":
        pooled = _dropout(MulFn.apply(pt, pv), model.dropout.p, model.training)
    elif fm in ("text", "vl-bert_vqa"):
        pooled = _dropout(pt, model.dropout.p, model.training)
    elif fm == "none":
        pooled = None
    else:
        raise ValueError("Invalid fusion method: %s" % fm)
    ttype = model.task_cfg[task_id]["type"]
    clf = model.clfs_dict[task_id]
    if ttype.startswith("V-logit"):
        sv = seq_v[-1] if output_all_encoded_layers else seq_v
        mask = _int_mask_like(image_attention_mask, sv)
        logits = _apply_head(clf, _dropout(sv, model.dropout.p, model.training), model.training)
        vil = AddFn.apply(logits, mask.unsqueeze(2).to(logits.dtype))
    elif ttype == "VL-binary-classifier":
        vil = _apply_head(clf, pooled.reshape(-1, pooled.shape[1] * 2), model.training)
    else:
        vil = _apply_head(clf, pooled, model.training)
    return vil, None, None, all_masks


def _int_mask_like(mask, sv):
    B, V = sv.shape[0], sv.shape[1]
    if mask is None:
        return torch.zeros(B, V, device=sv.device)
    out = torch.empty(B, V, dtype=torch.float32, device=sv.device)
    ones = mask.reshape(B, V).to(torch.float32)
    # (1 - m) * -10000 : additive mask of padded regions (reference: volta/encoders.py:1198-1199)
    ops.eltwise(_lib.ELT_SCALE, ones, out, alpha=10000.0)
    ops.eltwise(_lib.ELT_AXPY, torch.full_like(out, 1.0), out, alpha=-10000.0)
    _count(2)
    return out


# ====================================================================================================
# stand-alone sub-layer call (module API parity for BertGatedAttention / BertGatedFeedForward)
# ====================================================================================================
class _SublayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, owner, layer, kind, xt, xv, masks):
        B, T, V = xt.shape[0], xt.shape[1], xv.shape[1]
        rt = Runtime(owner, xt.device, B, T, V, layer.training, grad=True)
        x = Pair.alloc(rt, rt.Mt, xt.shape[2], rt.Mv, xv.shape[2])
        ops.eltwise(ELT_COPY, xt.reshape(rt.Mt, -1).contiguous(), x.t)
        ops.eltwise(ELT_COPY, xv.reshape(rt.Mv, -1).contiguous(), x.v)
        _count(2)
        if kind == "attn":
            y, saved = attention_fwd(rt, layer, x, masks)
        else:
            y, saved = ffn_fwd(rt, layer, x)
        ctx.rt, ctx.layer, ctx.kind, ctx.saved = rt, layer, kind, saved
        ctx.shapes = (xt.shape, xv.shape, xt.dtype)
        return y.t.view(B, T, -1).to(xt.dtype), y.v.view(B, V, -1).to(xv.dtype)

    @staticmethod
    def backward(ctx, gt, gv):
        rt = ctx.rt
        rt.store.prepare_grads()
        (st, sv, dtype) = ctx.shapes
        dy = Pair.alloc(rt, rt.Mt, st[2], rt.Mv, sv[2])
        for g, d in ((gt, dy.t), (gv, dy.v)):
            if g is None:
                ops.fill(d, 0)
            else:
                ops.eltwise(ELT_COPY, g.reshape(d.shape).contiguous(), d)
        dx = attention_bwd(rt, ctx.layer, dy, ctx.saved) if ctx.kind == "attn" else ffn_bwd(rt, ctx.layer, dy, ctx.saved)
        return None, None, None, None, dx.t.view(st).to(dtype), dx.v.view(sv).to(dtype), None


def standalone_sublayer(layer, kind, xt, xv, mask_t_add, mask_v_add):
    """Runs one gated sub-layer on its own.  `layer` must belong to a model (its parameters are flattened per
    layer otherwise).  Masks follow the reference: additive [bs,1,1,L] with -10000 on padded keys."""
    ops.require_cuda(xt, xv)
    owner = getattr(layer, "_owner", None)
    if owner is None:
        owner = _LayerOwner(layer)
        layer._owner = owner
    masks = None
    if kind == "attn":
        B, T, V = xt.shape[0], xt.shape[1], xv.shape[1]
        mt = torch.ones(B, T, dtype=torch.int32, device=xt.device) if mask_t_add is None else (mask_t_add.reshape(B, T) > -1.0).to(torch.int32)
        mv = torch.ones(B, V, dtype=torch.int32, device=xt.device) if mask_v_add is None else (mask_v_add.reshape(B, V) > -1.0).to(torch.int32)
        masks = (mt.contiguous(), mv.contiguous())
    anchor = _anchor(layer)
    return _SublayerFn.apply(anchor, owner, layer, kind, xt, xv, masks)


class _LayerOwner:
    """Gives a bare sub-layer the store interface of a model."""

    def __init__(self, layer):
        self.layer = layer
        self._engine_state = None
        self.precision = "bf16"
        self.training = layer.training

    def parameters(self):
        return self.layer.parameters()

    def _fused_groups(self):
        from .encoders import BertPreTrainedModel
        return BertPreTrainedModel._fused_groups(self.layer)

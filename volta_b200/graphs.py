"""CUDA-graph execution of the pre-training step (forward graph + backward graph segments).

Why: the step is ~650 kernel launches of 10-60 us each.  Issued one by one from Python (ctypes call, argument
struct, two cuTensorMapEncodeTiled per GEMM) the host needs ~37 us per launch, so once the kernels got faster the
step became launch-bound (24 ms of host work for ~19 ms of GPU work).  The schedule is static - shapes, pointers and
kernel arguments repeat every step - so it is captured once per input signature and replayed:

  * inputs are copied into static buffers, the outputs (three losses) are read from a static buffer;
  * every dropout site's seed is `site constant + *seed_dev` (vb_set_seed_pointer): the host draws one seed per
    step from torch's CPU generator - exactly what the eager path does - and uploads it, so eager and replayed
    steps produce identical numbers for the same generator state;
  * the backward is captured right after the forward, on the same thread, by calling the hand-written backward
    directly (autograd only triggers the replay).  It is cut into SEGMENTS at the points where the data-parallel
    wrapper closes a gradient bucket, so the bucket's all-reduce is issued between two replays and overlaps the
    rest of the backward exactly as in the eager path (reference: apex/apex/parallel/distributed.py:358-405).

The first call with a new signature runs eagerly (it fills the host-side caches and the lazily set kernel
attributes); the second call captures.  A forward issued while the previous forward's backward is still pending, or
any call without labels / without grad, takes the eager path.
"""
import gc
import os

import torch

from . import _lib, engine, ops
from .engine import ELT_COPY, Runtime


def enabled(model):
    if os.environ.get("VOLTA_B200_GRAPHS", "1") == "0":
        return False
    return getattr(model, "_use_cuda_graphs", True)


def _sig(t):
    return None if t is None else (tuple(t.shape), t.dtype)


class GraphedPretrainStep:
    def __init__(self, model, args):
        self.model = model
        self.device = args[0].device
        self.static_in = [None if a is None else a.detach().clone() for a in args]
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.dl = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.pending = False
        self.fwd_graph = None
        self.bwd_segments = []   # [(graph, frontier offset or None)]
        self.n_fwd = self.n_bwd = 0
        self.replays = 0         # completed forward+backward replays (tests assert the graph path was the one measured)
        self._capture()
        self.plan = dict(engine.last_plan)

    # ------------------------------------------------------------------------------------------ capture
    def _capture(self):
        model = self.model
        (ids, feat, loc, seg, mask_t, mask_v, mlm_labels, image_label, image_cls, itm_label, obj_labels, obj_confs,
         attr_labels, attr_confs) = self.static_in
        B, T = ids.shape
        V = feat.shape[1]
        store = engine.model_store(model)
        store.ensure(self.device, model.precision)
        store.prepare_grads()           # grad32 must exist (and keep its address) before it is baked into the graph
        # The packed plan schedules GEMM tiles up to device-side row counts, but the tile WIDTH is a launch-time choice
        # that is baked into the graph: give the cost model the valid-row counts of this batch (one host read at capture
        # time, never during replays) - later batches of the same shape have similar counts.
        ht = int((mask_t != 0).sum()) if mask_t is not None else B * T
        hv = int((mask_v != 0).sum()) if mask_v is not None else B * V
        store.row_hint = (B, T, V, max(ht, 1), max(hv, 1))
        lib = _lib.load()
        pool = torch.cuda.graph_pool_handle()
        stream = torch.cuda.Stream(device=self.device)
        stream.wait_stream(torch.cuda.current_stream(self.device))
        lib.vb_set_seed_pointer(self.seed_dev.data_ptr())
        # Nothing may free device memory or destroy a graph while the capture runs: collect dead objects now (models of
        # earlier calls hold graphs in reference cycles) and keep the cyclic collector off until the capture has ended.
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.stream(stream):
                n0 = engine.launch_count
                g = torch.cuda.CUDAGraph()
                g.capture_begin(pool=pool, capture_error_mode="thread_local")
                rt = Runtime(model, self.device, B, T, V, model.training, grad=True, seed_base=0)
                rt.quiet = True
                seg_ = torch.zeros_like(ids) if seg is None else seg
                x, pt, pv, _, s_bert = engine.bert_fwd(rt, model.bert, ids, feat, loc, seg_, mask_t, mask_v)
                labels = dict(mlm=mlm_labels, image_label=image_label, image_cls=image_cls, itm=itm_label, obj_labels=obj_labels,
                              obj_confs=obj_confs, attr_labels=attr_labels, attr_confs=attr_confs, image_feat=feat)
                losses, _, s_heads = engine.heads_fwd(rt, model, x, pt, pv, labels)
                g.capture_end()
                self.fwd_graph, self.losses = g, losses
                self.n_fwd = engine.launch_count - n0

                # backward: segments end where the progress hook's owner wants to act (bucket closed)
                n0 = engine.launch_count
                cut = engine.progress_cut_fn(model)
                state = {"g": torch.cuda.CUDAGraph()}
                state["g"].capture_begin(pool=pool, capture_error_mode="thread_local")

                def on_progress(frontier):
                    if cut is not None and cut(frontier):
                        state["g"].capture_end()
                        self.bwd_segments.append((state["g"], frontier))
                        state["g"] = torch.cuda.CUDAGraph()
                        state["g"].capture_begin(pool=pool, capture_error_mode="thread_local")
                rt.quiet = False
                rt.progress_override = on_progress
                with engine.backward_sm_limit(model):   # grid sizes are baked into the captured launches
                    dx, dpt, dpv = engine.heads_bwd(rt, model, s_heads, self.dl, x)
                    engine._notify_progress(rt, model.cls, skip_below=engine._frontier(rt.store, model.bert.encoder) or 0)
                    engine.bert_bwd(rt, model.bert, s_bert, dx, dpt, dpv)
                state["g"].capture_end()
                self.bwd_segments.append((state["g"], None))
                self.n_bwd = engine.launch_count - n0
                engine.launch_count -= self.n_fwd + self.n_bwd  # capture launches nothing; replays are counted
        finally:
            lib.vb_set_seed_pointer(None)
            if gc_was_on:
                gc.enable()
        torch.cuda.current_stream(self.device).wait_stream(stream)
        self._keep = (rt, s_bert, s_heads, x)  # activations live in the graphs' private pool
        self._ptrs = self._store_ptrs()

    def _store_ptrs(self):
        st = engine.model_store(self.model)
        ptrs = tuple(None if t is None else t.data_ptr() for t in (st.flat32, st.flat16, st.grad32, st._trash))
        return ptrs + (tuple(p.requires_grad for p in st.params),)   # freezing a layer re-routes its gradient writes

    def valid(self):
        """The graphs hold raw pointers into the flat parameter / gradient buffers."""
        engine.model_store(self.model).ensure(self.device, self.model.precision)  # also refreshes the bf16 operands
        return self._ptrs == self._store_ptrs()

    # ------------------------------------------------------------------------------------------ replay
    def forward(self, args, seed):
        for dst, src in zip(self.static_in, args):
            if dst is not None and src is not dst:
                dst.copy_(src, non_blocking=True)
        self.seed_dev.copy_(torch.tensor([seed], dtype=torch.int64))
        if engine.PREZERO_GRADS:
            engine.model_store(self.model).prezero_grads()
        self.fwd_graph.replay()
        engine.launch_count += self.n_fwd
        self.pending = True
        return self.losses.clone()

    def backward(self, grads):
        model = self.model
        store = engine.model_store(model)
        store.prepare_grads()
        gs = [torch.zeros(1, dtype=torch.float32, device=self.device) if g is None else g.reshape(1).float() for g in grads]
        self.dl[:3].copy_(torch.cat(gs))
        hook = engine._PROGRESS_HOOKS.get(id(store.model))
        for g, frontier in self.bwd_segments:
            g.replay()
            if frontier is not None and hook is not None:
                hook(frontier)
        engine.launch_count += self.n_bwd
        self.replays += 1
        self.pending = False
        engine._notify_grads_ready(model)


class _GraphedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, step, args, seed):
        out = step.forward(args, seed)
        ctx.step = step
        return out[0:1], out[1:2], out[2:3]

    @staticmethod
    def backward(ctx, g_mlm, g_img, g_itm):
        ctx.step.backward((g_mlm, g_img, g_itm))
        return None, None, None, None


def maybe_run(model, anchor, args):
    """Returns the three losses through the graph path, or None when the call must run eagerly."""
    if not enabled(model) or anchor is None or not model.training:
        return None
    key = (tuple(_sig(a) for a in args), model.precision, engine._PROGRESS_CUTS.get(id(engine.model_store(model).model)) is not None)
    steps = model.__dict__.setdefault("_graph_steps", {})
    step = steps.get(key)
    if step is None:
        steps[key] = "seen"        # first call with this signature: eager (fills caches, sets kernel attributes)
        return None
    if step == "seen":
        if any(isinstance(s, GraphedPretrainStep) and s.pending for s in steps.values()):
            return None
        step = steps[key] = GraphedPretrainStep(model, args)
    if step.pending:
        return None                # a second forward before the first one's backward: do not overwrite activations
    if not step.valid():           # parameters were re-laid out (load_state_dict, .to()): capture again
        step = steps[key] = GraphedPretrainStep(model, args)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return _GraphedFn.apply(anchor, step, args, seed)

"""CUDA-graph capture of the whole training step (second half).

The reference's step is (train.py:187-203)

    Z = brain_encoder(X, subject_idxs); loss = loss_func(Y, Z)
    optimizer.zero_grad(); loss.backward(); optimizer.step()

-- about 150 kernel launches through this library plus a few dozen allocator / autograd operations on one Python
thread.  `GraphedTrainStep` captures all of it (encoder forward, CLIP loss, backward, fused Adam) into ONE CUDA graph
and replays it per step:

    step = GraphedTrainStep(brain_encoder, loss_func, optimizer, X, Y, subject_idxs)   # optimizer: sd_b200.optim.FusedAdam
    for X, Y, subject_idxs in loader:
        loss = step(X, Y, subject_idxs)        # device scalar; loss.item() when the value is needed (train.py:196)

What varies from step to step lives in static buffers the graph reads:
  * X, Y are copied into static device tensors (device-to-device, or host-to-device from pinned memory);
  * the subject ids, their sort order / group offsets (the grouped per-subject GEMMs) and the spatial-dropout centre
    (drawn from numpy's global RNG like models.py:81, so a seeded run makes the same draws) are staged in pinned
    memory and copied by a node of the graph;
  * the Adam table (per-parameter step size and bias correction, which depend on the step count) is rebuilt on the host
    each step in pinned memory and copied by a node of the graph.  Subjects absent from a batch get an entry of length
    zero: their weights, moments and step counts are left untouched, exactly what `grad None` means to torch.optim.Adam
     -- inside a graph their gradient tensors exist (and are zero) because the launch sequence is fixed.
Semantics are those of the eager path (tests/test_gpu_zz_graph.py compares parameters after several steps).
Single device; BatchNorm running statistics and num_batches_tracked are updated by the kernels as in eager mode.
"""
import ctypes
import math

import numpy as np
import torch

from . import _native as nat
from . import engine, ops
from .optim import FusedAdam


class StaticInputs:
    """Pinned staging + device copy of the per-step integers: [ids (B) | order (B) | offsets (S+1) | dropout centre]."""

    def __init__(self, B, S, device):
        self.B, self.S = B, S
        n = 2 * B + S + 2
        self.host = torch.zeros(n, dtype=torch.int32).pin_memory()
        self.dev = torch.zeros(n, dtype=torch.int32, device=device)
        self._copied = False

    def prepare(self, ids, centre):
        """the step's integers as a numpy array (not yet visible to the graph)"""
        tab = np.zeros(self.host.numel(), dtype=np.int32)
        t = engine.subject_tables_host(ids, self.S)
        tab[:t.size] = t
        tab[2 * self.B + self.S + 1] = int(centre)
        return tab

    def publish(self, tab):
        """into the pinned staging buffer the graph's copy node reads (the previous replay must have consumed it)"""
        self.host.numpy()[:] = tab

    def stage(self, ids, centre):
        self.publish(self.prepare(ids, centre))

    def begin_step(self):
        self._copied = False

    def _ensure_copied(self):
        if not self._copied:                     # ONE copy per step, issued by whichever stage asks first: a kernel that reads
            # the pinned buffer itself (a memcpy node would queue on the H2D copy engine behind the next batch's transfer)
            with ops.stream_scope():
                nat.call("sd_copy_small", self.dev.data_ptr(), self.host.data_ptr(), self.host.numel() * 4, ops._st())
            self._copied = True

    def subject_tables(self, B, S, device):
        assert (B, S) == (self.B, self.S), "GraphedTrainStep: batch size / subject count changed"
        self._ensure_copied()
        d = self.dev
        return d[:B], d[B:2 * B], d[2 * B:2 * B + S + 1]

    def dropout_mask(self, dropout, device):
        self._ensure_copied()
        centre = self.dev[2 * self.B + self.S + 1:2 * self.B + self.S + 2].long()
        return dropout._mask_table(device).index_select(0, centre).squeeze(0)


class GraphedTrainStep:
    def __init__(self, encoder, loss_fn, optimizer, X, Y, subject_idxs, warmup=3):
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs sd_b200.optim.FusedAdam (its update is one capturable launch)")
        if len(optimizer.param_groups) != 1:
            raise NotImplementedError("GraphedTrainStep supports a single parameter group")
        pipe = encoder.pipeline()
        if pipe.reducer is not None or getattr(loss_fn, "process_group", None) is not None:
            raise NotImplementedError("GraphedTrainStep is single-device (collectives are not captured)")
        ops.require_cuda(X, "X")
        ops.require_cuda(Y, "Y")
        self.enc, self.crit, self.opt, self.pipe = encoder, loss_fn, optimizer, pipe
        self.device = X.device
        self.S = encoder.num_subjects
        self.B = X.shape[0]
        self.X = X.detach().clone().contiguous()
        self.Y = Y.detach().clone().contiguous()
        self.static = StaticInputs(self.B, self.S, self.device)
        self.group = optimizer.param_groups[0]
        self.params = [p for p in self.group["params"] if p.requires_grad]
        self.subject_param_index = {}
        for s, layer in enumerate(encoder.subject_block.subject_layer):
            self.subject_param_index[id(layer.weight)] = s
        self.dropout = encoder.subject_block.spatial_attention.spatial_dropout
        self._graph = None
        self._capture(subject_idxs, warmup)

    # ---- one eager / captured step over the static buffers -----------------------------------------------
    def _stage(self, subject_idxs):
        ids = engine.normalize_subject_ids(subject_idxs, self.S)
        if len(ids) != self.B:
            raise ValueError("GraphedTrainStep: expected %d subject ids, got %d" % (self.B, len(ids)))
        centre = np.random.randint(self.dropout.num_channels) if self.enc.training else 0      # models.py:81
        self.static.stage(ids, centre)
        return ids

    def _body(self, ids):
        self.static.begin_step()
        self.pipe.static = self.static
        try:
            Z = self.enc(self.X, ids)
            loss = self.crit(self.Y, Z)
            loss.backward()
        finally:
            self.pipe.static = None
        return loss

    def _capture(self, subject_idxs, warmup):
        # autograd binds every parameter's gradient accumulator to the stream that was current when the accumulator was
        # created, and keeps it alive as long as ANY graph that used it is alive (e.g. a loss tensor of an earlier
        # eager step that the caller still holds).  Accumulators bound to the default (legacy) stream would pull that
        # stream into the capture and invalidate it, so stale graphs are collected first and the warm-up below -- on a
        # side stream, as for any whole-network capture -- creates fresh ones.
        import gc
        gc.collect()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        rng = np.random.get_state()
        snapshot = self._snapshot()
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):             # lazy initialisation (weight-pack tables, kernel attributes, caches)
                ids = self._stage(subject_idxs)
                for p in self.params:
                    p.grad = None
                self._body(ids)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        # the warm-up steps must leave no trace: BatchNorm buffers, RNG stream
        self._restore(snapshot)
        np.random.set_state(rng)

        for p in self.params:
            p.grad = None
        ids = self._stage(subject_idxs)
        np.random.set_state(rng)
        self._adam_init()
        self._graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self._graph):
                self.loss = self._body(ids)
                self._adam_enqueue()
        except RuntimeError as e:
            if "legacy stream" in str(e) or "capture" in str(e).lower():
                raise RuntimeError(
                    "GraphedTrainStep: stream capture failed (%s).  The usual cause: a tensor produced by an earlier EAGER "
                    "training step of these modules (typically the last `loss`, or Z) is still alive; its autograd graph "
                    "keeps the parameters' gradient accumulators bound to the default stream.  Drop those tensors (del "
                    "loss, Z) before constructing GraphedTrainStep." % str(e).splitlines()[0]) from e
            raise
        self._restore(snapshot)                         # (capturing does not execute, but keep the contract obvious)
        self.loss = self.loss.detach()

    def _snapshot(self):
        return [b.detach().clone() for b in self.enc.buffers()]

    def _restore(self, snap):
        with torch.no_grad():
            for b, s in zip(self.enc.buffers(), snap):
                b.copy_(s)

    # ---- fused Adam inside the graph ------------------------------------------------------------------------
    def _adam_init(self):
        self.adam_params = []
        self._grads = {}
        for p in self.params:
            st = self.opt.state[p]
            if not st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            self.adam_params.append(p)
        k = len(self.adam_params)
        nbytes = (k * ctypes.sizeof(nat.AdamEntry) + 3) // 4 * 4
        self._adam_host = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        self._adam_dev = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self._adam_max_n = max((2 if p.is_complex() else 1) * p.numel() for p in self.adam_params)

    def _adam_enqueue(self):
        """inside the capture: gradients exist now; record their addresses, copy the table, launch the update"""
        real = FusedAdam._real
        self._entries = []
        for p in self.adam_params:
            if p.grad is None:
                raise RuntimeError("GraphedTrainStep: a parameter received no gradient during capture")
            g = p.grad
            if not g.is_contiguous():
                raise RuntimeError("GraphedTrainStep: non-contiguous gradient")
            st = self.opt.state[p]
            self._entries.append((real(p).data_ptr(), real(g).data_ptr(), real(st["exp_avg"]).data_ptr(),
                                  real(st["exp_avg_sq"]).data_ptr(), real(p).numel()))
        beta1, beta2 = self.group["betas"]
        with torch.cuda.device(self.device), ops.stream_scope():
            nat.call("sd_copy_small", self._adam_dev.data_ptr(), self._adam_host.data_ptr(), self._adam_host.numel(), ops._st())
            nat.call("sd_adam_step", self._adam_dev.data_ptr(), len(self._entries), min(1024, (self._adam_max_n + 1023) // 1024),
                     float(beta1), float(beta2), float(self.group["eps"]), float(self.group["weight_decay"]), ops._st())

    def _adam_table(self, ids):
        """host side of the update for THIS step: step counts, bias corrections, absent subjects skipped.  Built (numpy,
        vectorised over the ~100 entries) in a shadow table; __call__ moves it into the pinned buffer once the previous
        replay has finished with that."""
        beta1, beta2 = self.group["betas"]
        lr = self.group["lr"]
        if getattr(self, "_adam_shadow", None) is None:
            k = len(self._entries)
            dt = np.dtype([("param", "<u8"), ("grad", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"),
                           ("step_size", "<f4"), ("bc2", "<f4")])
            assert dt.itemsize == ctypes.sizeof(nat.AdamEntry)
            self._adam_shadow = np.zeros(k, dtype=dt)
            for i, e in enumerate(self._entries):
                self._adam_shadow[i] = (e[0], e[1], e[2], e[3], e[4], 0.0, 1.0)
            self._adam_n = np.array([e[4] for e in self._entries], dtype=np.int64)
            self._adam_subject = np.array([self.subject_param_index.get(id(p), -1) for p in self.adam_params], dtype=np.int64)
            self._adam_steps = np.array([float(self.opt.state[p]["step"]) for p in self.adam_params], dtype=np.float64)
        present = np.zeros(self.S + 1, dtype=bool)
        present[np.unique(ids)] = True
        present[self.S] = True                                           # index -1: not a per-subject weight
        active = present[self._adam_subject]
        self._adam_steps += active                                       # grad None in eager mode: step count untouched
        t = self._adam_shadow
        t["n"] = np.where(active, self._adam_n, 0)                       # ... and a zero-length entry: nothing is written
        t["step_size"] = lr / (1.0 - beta1 ** self._adam_steps)
        t["bc2"] = np.sqrt(1.0 - beta2 ** self._adam_steps)
        self._steps_dirty = True

    def sync_optimizer_state(self):
        """write the step counts kept by the graph back into optimizer.state (state_dict() / checkpointing)"""
        if getattr(self, "_adam_shadow", None) is not None:
            for p, st in zip(self.adam_params, self._adam_steps):
                self.opt.state[p]["step"].fill_(float(st))

    # ---- public ---------------------------------------------------------------------------------------------
    def __call__(self, X, Y, subject_idxs):
        """Replay one training step on (X, Y, subject_idxs); returns the loss as a device scalar (valid until the next
        call).  X / Y may be device tensors or pinned host tensors of the captured shapes."""
        if tuple(X.shape) != tuple(self.X.shape) or tuple(Y.shape) != tuple(self.Y.shape):
            raise ValueError("GraphedTrainStep: input shapes differ from the captured ones")
        # host work of this step first (it overlaps the previous replay still running on the GPU) ...
        ids = engine.normalize_subject_ids(subject_idxs, self.S)
        if len(ids) != self.B:
            raise ValueError("GraphedTrainStep: expected %d subject ids, got %d" % (self.B, len(ids)))
        centre = np.random.randint(self.dropout.num_channels) if self.enc.training else 0      # models.py:81
        tab = self.static.prepare(ids, centre)
        self._adam_table(ids)
        # ... then the previous replay must have consumed the pinned staging buffers before they are rewritten
        torch.cuda.current_stream(self.device).synchronize()
        self.static.publish(tab)
        ctypes.memmove(self._adam_host.data_ptr(), self._adam_shadow.ctypes.data, self._adam_shadow.nbytes)
        if X.data_ptr() != self.X.data_ptr():
            self.X.copy_(X, non_blocking=True)
        if Y.data_ptr() != self.Y.data_ptr():
            self.Y.copy_(Y, non_blocking=True)
        self._graph.replay()
        return self.loss

    @property
    def inputs(self):
        """the static device tensors (X, Y): fill them directly (e.g. with an asynchronous H2D copy on another stream,
        ordered before the replay) and pass them back to __call__ to skip the extra device-to-device copy"""
        return self.X, self.Y

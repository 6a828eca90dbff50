"""Data-parallel plumbing (one process per GPU, torch.distributed over NCCL).

The reference is single-device (train.py:31); batch sharding is this build's
addition:

  * every rank runs the full encoder on its shard of the batch;
  * CLIP negatives span the global batch: the speech rows x (input data, no
    gradient) are all-gathered, each rank scores them against its local brain
    rows z, column statistics stay local and only the (M,2) row statistics
    (max, sum-exp) and two scalars cross ranks -- no gradient exchange for dz;
  * parameter gradients are summed across ranks (the loss is the global-batch
    loss, so SUM, not average), launched per stage from inside backward so the
    all-reduce overlaps the remaining backward kernels;
  * optional SyncBN (per-channel statistics all-reduced) so the result equals
    the single-process run on the concatenated batch;
  * subject presence (which per-subject weights receive a gradient at all) is
    agreed on the host over a gloo side-channel, so no device sync is needed.

Everything here is host logic over torch.distributed and runs unchanged on the
gloo backend with CPU tensors (tests/test_dist_cpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def world_rank(group=None):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_rows(x, group):
    world, _ = world_rank(group)
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


class _RawDeviceBuffer:
    """zero-copy view of a raw device allocation for torch.as_tensor (CUDA array interface)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class _PeerArrival:
    """Work-like handle of one copy-engine gather: wait() orders the current stream after the arrival of every peer's rows."""

    def __init__(self, flags_ptr, world, expected, device):
        self.flags_ptr, self.world, self.expected, self.device = flags_ptr, world, expected, device

    def wait(self):
        from . import _native as nat
        with torch.cuda.device(self.device):
            nat.call("sd_peer_wait_flags", self.flags_ptr, self.world, self.expected, torch.cuda.current_stream().cuda_stream)


class PeerGather:
    """All-gather of the bf16 speech rows (and their squared norms) on the COPY ENGINES: every rank owns a receive buffer
    (two parity slots of [world x rows x D bf16 | world x rows fp32 norms | world flag words], allocated by the native
    library so that it has a CUDA IPC handle), peers map it once and each step PUSH their rows, norms and -- last, on the
    same stream -- an arrival flag into their part of it with peer cudaMemcpyAsync on side streams.  No SM is occupied and
    no NCCL kernel takes part: the transfer overlaps the persistent tcgen05 grids of the encoder forward without costing
    them a wave (an NCCL all-gather kernel holds SMs for the whole 1.3 GB transfer at 8 GPUs, and even a tiny NCCL
    collective used as the arrival fence is starved of an SM by back-to-back 148-CTA grids -- both measured,
    tools/dist_phases.py).  The consumer's side of the fence is one polling kernel in stream order
    (sd_peer_wait_flags).  Slot reuse is safe with two slots because consecutive steps are separated by a collective
    every rank takes part in (row-statistics exchange in the loss, gradient all-reduce in backward)."""

    NSTREAMS = 2      # concurrent peer copies: 2 keep the pushes inside backward without slowing the conv grids (measured at N=8:
                      # 4 streams finish sooner but cost the concurrent conv launches 7 %)
    NCONST = 1024

    def __init__(self, group, host_group, device):
        self.group, self.host_group, self.device = group, host_group, torch.device(device)
        self.world, self.rank = world_rank(group)
        self.shape = None
        self.base = None             # own receive buffer (device pointer)
        self.peers = None            # device pointers of every rank's receive buffer, mapped into this process
        self.step = 0
        import os
        self.NSTREAMS = max(1, int(os.environ.get("SD_B200_DP_STREAMS", self.NSTREAMS)))   # concurrent peer copies
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.NSTREAMS)]
        self.ready = torch.cuda.Event()
        self.consts = torch.arange(1, self.NCONST + 1, dtype=torch.int32, device=self.device)    # flag values to copy from
        self.timing = None           # tools/dist_phases.py: list of (start, end) timing events of the pushes, per call

    def _release(self, collective=False):
        """Unmap the peers' buffers, then free the own one.  An exporter must not free memory an importer still has
        mapped, so when every rank releases together (re-allocation for a new shape, DataParallel.close) a host barrier
        separates the two halves; a lone release (object finalisation at process teardown) just lets go."""
        from . import _native as nat
        had = self.base is not None
        if self.peers is not None:
            for r, p in enumerate(self.peers):
                if r != self.rank and p:
                    nat.call("sd_ipc_close_handle", p)
        if collective and had:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.host_group)
        if self.base:
            nat.call("sd_peer_free", self.base)
        self.base = self.peers = self.shape = None

    def _setup(self, rows, D):
        import ctypes
        from . import _native as nat
        torch.cuda.synchronize(self.device)
        self._release(collective=True)       # (every rank re-allocates at the same step: the shard shape changed everywhere)
        self.rows_bytes = self.world * rows * D * 2
        self.norm_off = (self.rows_bytes + 255) // 256 * 256
        self.flag_off = self.norm_off + (self.world * rows * 4 + 255) // 256 * 256
        self.slot_bytes = self.flag_off + 256
        base = ctypes.c_void_p()
        nat.call("sd_peer_alloc", ctypes.byref(base), 2 * self.slot_bytes)
        self.base = base.value
        hb = ctypes.create_string_buffer(nat.lib().sd_ipc_handle_bytes())
        nat.call("sd_ipc_get_handle", self.base, hb)
        mine = (rows, D, hb.raw)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.host_group)
        for r, (r_rows, r_D, _) in enumerate(everyone):
            if (r_rows, r_D) != (rows, D):
                raise RuntimeError("sd_b200 DataParallel: rank %d holds %d x %d speech rows, rank %d holds %d x %d -- the "
                                   "batch must be sharded evenly" % (self.rank, rows, D, r, r_rows, r_D))
        self.peers = []
        for r, (_, _, h) in enumerate(everyone):
            if r == self.rank:
                self.peers.append(self.base)
            else:
                out = ctypes.c_void_p()
                nat.call("sd_ipc_open_handle", ctypes.create_string_buffer(h, len(h)), ctypes.byref(out))
                self.peers.append(out.value)
        flat = torch.as_tensor(_RawDeviceBuffer(self.base, 2 * self.slot_bytes), device=self.device)
        flat.zero_()                                   # flag words start at 0 (never a valid flag value)
        self.row_views, self.norm_views = [], []
        for i in range(2):
            sl = flat[i * self.slot_bytes:(i + 1) * self.slot_bytes]
            self.row_views.append(sl[:self.rows_bytes].view(torch.bfloat16).view(self.world * rows, D))
            self.norm_views.append(sl[self.norm_off:self.norm_off + self.world * rows * 4].view(torch.float32))
        self.shape = (rows, D)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.host_group)     # nobody pushes before everybody has mapped and zeroed

    def gather(self, xb, n2):
        """xb (rows, D) bf16, n2 (rows,) fp32, both produced on the current stream -> (all rows, all norms, [arrival])."""
        from . import _native as nat
        rows, D = xb.shape
        if self.shape != (rows, D):
            self._setup(rows, D)
        slot = self.step & 1
        expected = self.step % self.NCONST + 1
        self.step += 1
        nbytes = rows * D * 2
        base = slot * self.slot_bytes
        self.ready.record()
        t_ev = None
        if self.timing is not None:
            t_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        flag_src = self.consts.data_ptr() + 4 * (expected - 1)
        for k in range(self.world):               # k = 0 is the local slot; peers in ring order so that no two ranks
            peer = (self.rank + k) % self.world   # push into the same destination at the same time
            st = self.streams[k % self.NSTREAMS]
            if k < self.NSTREAMS:
                st.wait_event(self.ready)
                if k == 0 and t_ev is not None:
                    t_ev[0].record(st)
            if k < self.NSTREAMS:             # the side streams read xb / n2 after this call returns: the caching allocator
                xb.record_stream(st)          # must not hand their memory out again before those streams are past the copies
                n2.record_stream(st)
            dst = self.peers[peer] + base
            nat.call("sd_memcpy_async", dst + self.rank * nbytes, xb.data_ptr(), nbytes, st.cuda_stream)
            nat.call("sd_memcpy_async", dst + self.norm_off + self.rank * rows * 4, n2.data_ptr(), rows * 4, st.cuda_stream)
            nat.call("sd_memcpy_async", dst + self.flag_off + self.rank * 4, flag_src, 4, st.cuda_stream)   # after the data
        if t_ev is not None:
            for i in range(1, min(self.NSTREAMS, self.world)):
                ev = torch.cuda.Event()
                ev.record(self.streams[i])
                self.streams[0].wait_event(ev)
            t_ev[1].record(self.streams[0])
            self.timing.append(t_ev)
        arrival = _PeerArrival(self.base + base + self.flag_off, self.world, expected, self.device)
        return self.row_views[slot], self.norm_views[slot], [arrival]

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


class PeerMailbox:
    """Latency-bound exchanges through peer memory (sd_peer_exchange): every rank owns a mailbox of two parity slots x
    world rows x CAP bytes, IPC-mapped by its peers; one kernel stores the payload into every peer's mailbox over NVLink,
    publishes the epoch and waits for the peers'.  Used for the SyncBN statistics, the BatchNorm-backward sums, the CLIP row
    statistics and the loss partials instead of small NCCL all-reduces."""

    CAP = 64 * 1024

    def __init__(self, group, host_group, device):
        import ctypes
        from . import _native as nat
        self.group, self.host_group, self.device = group, host_group, torch.device(device)
        self.world, self.rank = world_rank(group)
        self.epoch = 0
        nbytes = 2 * self.world * self.CAP + 256
        base = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            nat.call("sd_peer_alloc", ctypes.byref(base), nbytes)
            self.base = base.value
            torch.as_tensor(_RawDeviceBuffer(self.base, nbytes), device=self.device).zero_()
            torch.cuda.synchronize(self.device)
            hb = ctypes.create_string_buffer(nat.lib().sd_ipc_handle_bytes())
            nat.call("sd_ipc_get_handle", self.base, hb)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, hb.raw, group=host_group)
            self.peers = []
            for r, h in enumerate(everyone):
                if r == self.rank:
                    self.peers.append(self.base)
                else:
                    out = ctypes.c_void_p()
                    nat.call("sd_ipc_open_handle", ctypes.create_string_buffer(h, len(h)), ctypes.byref(out))
                    self.peers.append(out.value)
            self.peers_dev = torch.tensor(self.peers, dtype=torch.int64, device=self.device)
            torch.cuda.synchronize(self.device)
        dist.barrier(group=host_group)

    def fits(self, t):
        return t.is_cuda and t.is_contiguous() and (t.numel() * t.element_size()) % 4 == 0 and t.numel() * t.element_size() <= self.CAP

    def exchange(self, t):
        """t (contiguous, <= CAP bytes) of every rank -> (world, *t.shape) on every rank, in stream order"""
        from . import _native as nat
        self.epoch += 1
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        with torch.cuda.device(self.device):
            nat.call("sd_peer_exchange", t.data_ptr(), t.numel() * t.element_size(), self.peers_dev.data_ptr(), self.rank, self.world,
                     self.CAP, self.epoch, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out

    def all_reduce_sum_(self, t):
        """t <- sum over ranks of t (fp32 / fp64), summed in rank order: bit-identical on every rank"""
        from . import _native as nat
        g = self.exchange(t)
        with torch.cuda.device(self.device):
            nat.call("sd_sum_rows", g.data_ptr(), t.data_ptr(), self.world, t.numel(), int(t.dtype == torch.float64),
                     torch.cuda.current_stream().cuda_stream)
        return t

    def close(self, collective=False):
        """see PeerGather._release"""
        from . import _native as nat
        if self.base is None:
            return
        for r, p in enumerate(self.peers):
            if r != self.rank and p:
                nat.call("sd_ipc_close_handle", p)
        if collective:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.host_group)
        nat.call("sd_peer_free", self.base)
        self.base, self.peers = None, []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_MAILBOX = {}          # id(group) -> PeerMailbox (installed by DataParallel next to the copy-engine gather)


def small_all_reduce_sum_(t, group):
    """In-place SUM all-reduce of a small fp32 / fp64 tensor: through the peer mailbox when there is one, else NCCL / gloo."""
    mb = _MAILBOX.get(id(group))
    if mb is not None and t.dtype in (torch.float32, torch.float64) and mb.fits(t):
        return mb.all_reduce_sum_(t)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def global_row_lse(row_stat, group):
    """(M,2) per-rank (max_j, sum_j exp(l - max)) over local columns -> (M,) log-sum-exp over the columns of all ranks."""
    from . import _native as nat
    world, _ = world_rank(group)
    if row_stat.is_cuda:
        mb = _MAILBOX.get(id(group)) if world > 1 else None
        if world == 1 or (mb is not None and mb.fits(row_stat)):
            g = row_stat.contiguous() if world == 1 else mb.exchange(row_stat.contiguous())
            out = torch.empty((row_stat.shape[0],), dtype=torch.float32, device=row_stat.device)
            with torch.cuda.device(row_stat.device):
                nat.call("sd_clip_merge_row_stats", g.data_ptr(), world, row_stat.shape[0], out.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
            return out
    if world > 1:
        row_stat = merge_row_stats(row_stat, group)
    return row_stat[:, 0] + torch.log(row_stat[:, 1])


_PEER_GATHER = {}      # id(group) -> PeerGather (installed by DataParallel when the copy-engine path is usable)


def gather_speech_rows(x2d, group, allow_bf16=True, async_op=False):
    """All-gather the speech rows of every rank.  In the bf16 mode (and when no gradient flows to them) the rows are
    rounded to bf16 ONCE on the owning rank together with their squared norms: the gather then moves, and the two
    CLIP GEMMs re-read, half the bytes.  Returns (rows, norms-or-None[, works])."""
    from . import ops
    world, _ = world_rank(group)
    works = []
    if world > 1 and allow_bf16 and x2d.is_cuda and ops.clip_bf16_ok(x2d):
        with torch.cuda.device(x2d.device), ops.stream_scope():
            if x2d.dtype == torch.bfloat16:        # shipped in bf16 already: only the norms are missing
                xb, n2 = x2d.contiguous(), None
                n2 = ops.rownorm2_bf16(xb)
            else:
                xb, n2 = ops.cast_rows_bf16(x2d)
        pg = _PEER_GATHER.get(id(group))
        if pg is not None:
            with torch.cuda.device(x2d.device):
                rows, norms, works = pg.gather(xb, n2)
            if not async_op:
                for w in works:
                    w.wait()
                return rows, norms
            return rows, norms, works, (xb, n2)
        rows = torch.empty((world * xb.shape[0], xb.shape[1]), dtype=xb.dtype, device=xb.device)
        norms = torch.empty((world * n2.shape[0],), dtype=n2.dtype, device=n2.device)
        works.append(dist.all_gather_into_tensor(norms, n2, group=group, async_op=async_op))
        works.append(dist.all_gather_into_tensor(rows, xb, group=group, async_op=async_op))
        keep = (xb, n2)
    else:
        rows = torch.empty((world * x2d.shape[0], x2d.shape[1]), dtype=x2d.dtype, device=x2d.device)
        norms = None
        works.append(dist.all_gather_into_tensor(rows, x2d.contiguous(), group=group, async_op=async_op))
        keep = (x2d,)
    if async_op:
        return rows, norms, works, keep
    return rows, norms


def all_reduce_sum(t, group):
    return small_all_reduce_sum_(t.contiguous(), group)


def merge_row_stats(row_stat, group):
    """(M,2) per-rank (max_j, sum_j exp(l - max)) over local columns ->
    the same statistics over the columns of all ranks."""
    m_local = row_stat[:, 0].contiguous()
    m_glob = m_local.clone()
    dist.all_reduce(m_glob, op=dist.ReduceOp.MAX, group=group)
    s = (row_stat[:, 1] * torch.exp(m_local - m_glob)).contiguous()
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return torch.stack([m_glob, s], dim=1)


def gather_host_ints(values, host_group):
    """all-gather a small int64 numpy vector over the host (gloo) side channel."""
    return gather_host_ints_async(values, host_group)()


def gather_host_ints_async(values, host_group):
    """Start the all-gather of a small int64 numpy vector over the host (gloo) side channel and return a function that
    waits for it and yields the per-rank vectors.  The encoder forward starts the exchange of the subject ids and only
    backward (which needs the union to decide which per-subject weights get a gradient) collects it, so the host never
    blocks in the middle of enqueueing the forward."""
    world, _ = world_rank(host_group)
    t = torch.from_numpy(np.asarray(values, dtype=np.int64).copy())
    if world == 1:
        return lambda: [t.numpy()]
    outs = [torch.empty_like(t) for _ in range(world)]
    work = dist.all_gather(outs, t, group=host_group, async_op=True)

    def collect():
        work.wait()
        return [o.numpy() for o in outs]
    return collect


class GradReducer:
    """Sums parameter gradients across ranks while backward is still running (hooked into engine.Pipeline).  Finished
    stages are coalesced into buckets of at least `bucket_bytes` (stages finish in reverse parameter order, so their
    slices of the GradPool are adjacent): every all-reduce is an SM-resident NCCL kernel running next to the persistent
    one-CTA-per-SM conv / wgrad grids, and each of them costs those grids a second wave -- few, large messages keep that
    to a couple of launches per step.  bucket_bytes = 0: one all-reduce per stage; None: SD_B200_DP_BUCKET_MB (default
    16 MB -> three buckets for the 37 MB of cfg2 gradients)."""

    def __init__(self, group, bucket_bytes=None):
        import os
        self.group = group
        self.pending = []
        if bucket_bytes is None:
            bucket_bytes = int(float(os.environ.get("SD_B200_DP_BUCKET_MB", "16")) * (1 << 20))
        self.bucket_bytes = bucket_bytes
        self.lo = self.hi = None
        self.flat = None
        self.launched = 0

    def _flush(self):
        if self.lo is None or self.hi <= self.lo:
            self.lo = self.hi = None
            return
        work = dist.all_reduce(self.flat[self.lo:self.hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append(work)
        self.launched += 1
        self.lo = self.hi = None

    def stage_done(self, flat, lo, hi):
        """flat[lo:hi] (one stage's parameter gradients, contiguous in the GradPool -- absent subjects included as
        zeros, so the message size is the same on every rank) is final."""
        if hi <= lo:
            return
        if self.lo is not None and not (hi == self.lo or lo == self.hi):
            self._flush()                      # not adjacent to the open bucket
        self.flat = flat
        self.lo = lo if self.lo is None else min(lo, self.lo)
        self.hi = hi if self.hi is None else max(hi, self.hi)
        if (self.hi - self.lo) * flat.element_size() >= self.bucket_bytes:
            self._flush()

    def finish(self):
        self._flush()
        for work in self.pending:
            work.wait()
        self.pending = []
        self.flat = None


class DataParallel:
    """Wire a BrainEncoder + CLIPLoss pair for batch-sharded training.

        dp = DataParallel(encoder, loss_fn, group=None, sync_bn=True)
        Z = encoder(X_local, ids_local); loss = loss_fn(Y_local, Z); loss.backward()

    `loss` is the global-batch loss on every rank; parameter .grad's are the
    global-batch gradients (identical on every rank)."""

    def __init__(self, encoder, loss_fn, group=None, sync_bn=True, host_group=None, reserve_sms=None, peer_gather=None,
                 bucket_bytes=None, broadcast=True):
        """reserve_sms: SMs left free for the NCCL kernels that run concurrently with the step (speech-row gather
        during the encoder forward, gradient all-reduces during backward).  The conv / wgrad / CLIP kernels are
        persistent, one CTA per SM: if a collective holds even one SM, the CTAs that cannot be placed start only
        when others finish and the launch takes two waves.  Default: SD_B200_DP_RESERVE_SMS or 0 (no cap)."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        import os
        if reserve_sms is None:
            reserve_sms = int(os.environ.get("SD_B200_DP_RESERVE_SMS", "0"))
        self.reserve_sms = int(reserve_sms)
        if self.reserve_sms > 0 and torch.cuda.is_available():
            from . import _native as nat
            sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            nat.call("sd_set_sm_limit", max(2, sms - self.reserve_sms))
        self.group = group if group is not None else dist.group.WORLD
        self.host_group = host_group
        if host_group is None and dist.get_backend(self.group) != "gloo":
            self.host_group = dist.new_group(backend="gloo")
        elif host_group is None:
            self.host_group = self.group
        pipe = encoder.pipeline()
        self.pipe = pipe
        pipe.reducer = GradReducer(self.group, bucket_bytes)
        pipe.bn_group = self.group if sync_bn else None
        pipe.host_group = self.host_group
        loss_fn.process_group = self.group
        self.loss_fn = loss_fn
        loss_fn._prefetched = None
        self.sync_bn = bool(sync_bn)
        if broadcast:
            # replicas must start identical (like DDP): parameters and buffers of rank 0 win.  With sync_bn=False the
            # BatchNorm running statistics then evolve per rank (each rank normalises with its own shard's statistics);
            # a checkpoint holds the saving rank's -- documented, not averaged.
            with torch.no_grad():
                for t in list(encoder.parameters()) + list(encoder.buffers()) + list(loss_fn.parameters()):
                    if t.is_complex():
                        dist.broadcast(torch.view_as_real(t.data), 0, group=self.group)
                    else:
                        dist.broadcast(t.data, 0, group=self.group)
        # speech-row exchange on the copy engines (CUDA IPC peer buffers) instead of an SM-resident NCCL kernel
        if peer_gather is None:
            peer_gather = os.environ.get("SD_B200_DP_GATHER", "peer") != "nccl"
        self.peer = self.mailbox = None
        dev = next(encoder.parameters()).device
        if peer_gather and dev.type == "cuda" and dist.get_world_size(self.group) > 1:
            try:
                self.peer = PeerGather(self.group, self.host_group, dev)
                _PEER_GATHER[id(self.group)] = self.peer
                self.mailbox = PeerMailbox(self.group, self.host_group, dev)
                _MAILBOX[id(self.group)] = self.mailbox
            except Exception as e:                                  # pragma: no cover
                import warnings
                warnings.warn("sd_b200: copy-engine peer gather unavailable (%s); using the NCCL all-gather" % e)

    def close(self):
        """Collective: release the peer-memory buffers (every rank must call it) and detach from the modules."""
        if self.peer is not None:
            self.peer._release(collective=True)
            _PEER_GATHER.pop(id(self.group), None)
            self.peer = None
        if self.mailbox is not None:
            self.mailbox.close(collective=True)
            _MAILBOX.pop(id(self.group), None)
            self.mailbox = None
        self.pipe.reducer = None
        self.pipe.bn_group = self.pipe.host_group = None
        self.loss_fn.process_group = None

    def prefetch_targets(self, Y, during_backward=None):
        """Start the all-gather of the speech embeddings for the coming step NOW (they are input data, known
        before the encoder runs), so the transfer over NVLink overlaps the encoder forward instead of sitting
        between the encoder and the loss.  Call right before `encoder(X, ids)`; `loss_fn(Y, Z)` with the same
        Y then picks the gathered tensor up.  Optional: without it the loss gathers synchronously.

        Pipelined use (a loader that has the NEXT batch ready): call it for the next step's Y between this step's
        `loss_fn(...)` and `loss.backward()` with during_backward=k: the exchange is then started from inside backward,
        right after the k-th stage from the end has been enqueued (0 = the final 1x1 convs, whose backward and the CLIP
        gradient before it are the HBM-bound part of backward; the conv blocks that follow are tensor-bound and lose
        less to the concurrent copy-engine traffic)."""
        world, _ = world_rank(self.group)
        if world == 1:
            return
        if during_backward is not None:
            pipe = self.pipe
            idx = len(pipe.stages) - 1 - int(during_backward)
            pipe.backward_hooks.setdefault(idx, []).append(lambda: self.prefetch_targets(Y))
            return
        from . import ops
        y2 = Y.reshape(Y.shape[0], -1)
        if y2.dtype == torch.bfloat16 and ops.clip_bf16_ok(y2):
            y2 = y2.contiguous()
        elif y2.dtype != torch.float32 or not y2.is_contiguous():
            y2 = y2.float().contiguous()
        rows, norms, works, keep = gather_speech_rows(y2, self.group, not Y.requires_grad, async_op=True)
        self.loss_fn._prefetched = (Y.data_ptr(), tuple(Y.shape), works, rows, norms, keep)
        # the temperature gradient is a partial sum per rank: the loss Function already
        # all-reduces `partial`, so dtemp is global -- nothing more to do for it.

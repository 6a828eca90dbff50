"""Host-side executor of the encoder hot path.

A module (BrainEncoder, or any of its sub-modules used stand-alone) is run as a
*pipeline* of stages over the internal channels-last activation layout.  One
`torch.autograd.Function` spans the whole pipeline, so PyTorch sees a single
graph node: forward enqueues every kernel of the encoder, backward enqueues the
matching dgrad/wgrad kernels in reverse and hands parameter gradients back to
autograd (absent subjects get `None`, exactly like the reference's ModuleList:
SURVEY.md §7.3 item 5).

Reference being replaced: speech_decoding/models.py:45-65 (SpatialAttention),
:77-86 (SpatialDropout), :111-117 (SubjectBlock), :152-166 (ConvBlock),
:191-196 (BrainEncoder).
"""
import ctypes

import numpy as np
import torch

from . import _native as nat
from . import ops
from .ops import rup8

FUSE_EVAL_BN = True     # inference: fold eval-mode BatchNorm + GELU into the producing conv epilogue (tests flip it)


# --------------------------------------------------------------------------------------------------
# packed weight shadows
# --------------------------------------------------------------------------------------------------
class WeightPack:
    """bf16/fp32 shadows of Conv1d weights in the two MMA layouts (forward `wf`
    (G,taps,Np,Kp) and data-gradient `wd` (G,taps,Kp,Np), taps reversed),
    refreshed from the fp32 master parameters by ONE kernel launch per step."""

    def __init__(self):
        self.specs = []          # (name, [params], N, K, taps)
        self.bufs = {}           # name -> (wf, wd)
        self._sig = None
        self._table = None

    def add(self, name, params, N, K, taps):
        self.specs.append((name, list(params), N, K, taps))

    def refresh(self, device, dtype):
        split = ops.tensor_core_fp32() == 3 and dtype == torch.float32
        sig = (str(device), dtype, split) + tuple(p.data_ptr() for s in self.specs for p in s[1])
        if sig != self._sig:
            entries = []
            self.bufs = {}
            code = nat.SD_F32 if dtype == torch.float32 else nat.SD_BF16
            # every shadow lives in ONE flat buffer (3xTF32: the packed weights are split into hi / lo planes by a
            # single launch over it)
            sizes = [(len(params) * taps * rup8(N) * rup8(K)) for _, params, N, K, taps in self.specs]
            self._flat = torch.empty(2 * sum(sizes), dtype=dtype, device=device)
            self._planes = torch.empty((2, 2 * sum(sizes)), dtype=dtype, device=device) if split else None
            off = 0
            for (name, params, N, K, taps), n in zip(self.specs, sizes):
                G, Np, Kp = len(params), rup8(N), rup8(K)
                wf = self._flat[off:off + n].view(G, taps, Np, Kp)
                wd = self._flat[off + n:off + 2 * n].view(G, taps, Kp, Np)
                if split:
                    hf = self._planes[0, off:off + n].view(G, taps, Np, Kp)
                    hd = self._planes[0, off + n:off + 2 * n].view(G, taps, Kp, Np)
                    hf._sd_lo = self._planes[1, off:off + n].view(G, taps, Np, Kp)
                    hd._sd_lo = self._planes[1, off + n:off + 2 * n].view(G, taps, Kp, Np)
                    self.bufs[name] = (hf, hd)
                else:
                    self.bufs[name] = (wf, wd)
                off += 2 * n
                for g, p in enumerate(params):
                    if p.dtype != torch.float32 or not p.is_contiguous():
                        raise RuntimeError("sd_b200: parameter %s must be contiguous fp32" % name)
                    entries.append(nat.PackEntry(p.data_ptr(), wf[g].data_ptr(), wd[g].data_ptr(),
                                                 N, K, taps, Np, Kp, code))
            arr = (nat.PackEntry * len(entries))(*entries)
            raw = np.frombuffer(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr)), dtype=np.uint8).copy()
            self._table = torch.from_numpy(raw).to(device)
            self._n = len(entries)
            self._max_tiles = max(((e.Np + 31) // 32) * ((e.Kp + 31) // 32) for e in entries)
            self._sig = sig
        nat.call("sd_pack_weights", self._table.data_ptr(), self._n, self._max_tiles, ops._st())
        if split:
            nat.call("sd_tf32_split", self._flat.data_ptr(), self._planes[0].data_ptr(), self._planes[1].data_ptr(),
                     self._flat.numel(), ops._st())

    def wf(self, name):
        return self.bufs[name][0]

    def wd(self, name):
        return self.bufs[name][1]


class GradPool:
    """One flat fp32 buffer for every parameter gradient of a pipeline (densely packed in parameter
    order): a single memset per backward instead of one per tensor, and data-parallel all-reduces can
    work on contiguous slices.  Gradients handed to autograd are views into it."""

    def __init__(self, params):
        self.offsets = {}
        off = 0
        for p in params:
            n = p.numel() * (2 if p.is_complex() else 1)
            self.offsets[p] = (off, n)
            off += n
        self.total = off
        self.flat = None

    def new(self, device):
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)

    def view(self, p):
        off, n = self.offsets[p]
        t = self.flat[off:off + n]
        return t.view(*p.shape, 2) if p.is_complex() else t.view(p.shape)

    def stacked(self, params):
        """(len(params), *shape) view over consecutive same-shaped parameters."""
        off0, n = self.offsets[params[0]]
        for i, p in enumerate(params):
            assert self.offsets[p] == (off0 + i * n, n), "parameters are not consecutive in the pool"
        return self.flat[off0:off0 + n * len(params)].view(len(params), *params[0].shape)

    def span(self, params):
        lo = min(self.offsets[p][0] for p in params)
        hi = max(self.offsets[p][0] + self.offsets[p][1] for p in params)
        return lo, hi


class ScratchPool:
    """Bump allocator over one zero-filled fp64 buffer (BatchNorm statistics / reduction scratch)."""

    def __init__(self, n, device):
        self.buf = torch.zeros(max(n, 1), dtype=torch.float64, device=device)
        self.off = 0

    def take(self, n):
        t = self.buf[self.off:self.off + n]
        self.off += n
        return t


# --------------------------------------------------------------------------------------------------
# stages
# --------------------------------------------------------------------------------------------------
class Stage:
    scratch = 0          # fp64 scratch elements needed per forward / backward

    def params(self):
        return []

    def register(self, pack: WeightPack):
        pass


class SpatialAttentionStage(Stage):
    """X (B,C,T) fp32 -> (B,T,D1p): Fourier softmax weights, dropout mask, channel mix."""

    def __init__(self, sa):
        self.m = sa

    def params(self):
        return [self.m.z]

    def forward(self, run, X, sv):
        m = self.m
        B, C, T = X.shape
        assert C == m.spatial_dropout.num_channels                                  # models.py:78
        D1, K2 = m.z.shape
        Xt = ops.nct_to_btc(X, run.dtype)
        if not m.training:
            mask = None
        elif run.static is not None:      # CUDA-graph replay: the centre drawn by the host this step, read on the device
            mask = run.static.dropout_mask(m.spatial_dropout, X.device)
        else:
            mask = m.spatial_dropout.draw_mask(X.device)                              # models.py:81-83
        z_ri = torch.view_as_real(m.z.detach())
        w_soft, w_packed = ops.sa_weights_fwd(z_ri, m.cos, m.sin, mask, D1, K2, C, run.dtype)
        out = torch.empty((B, T, rup8(D1)), dtype=run.dtype, device=X.device)
        ops.conv_fwd(Xt, w_packed, K=C, N=D1, out=out)
        if sv is not None:
            sv.update(Xt=Xt, mask=mask, w_soft=w_soft, C=C)
        return out

    def backward(self, run, sv, dout, grads, need_dx):
        m = self.m
        D1, K2 = m.z.shape
        C = sv["C"]
        if need_dx:
            raise NotImplementedError("sd_b200: gradient w.r.t. the sensor input X is not implemented "
                                      "(the reference never needs it: train.py:187-203)")
        dwm = torch.zeros((D1, C), dtype=torch.float32, device=dout.device)
        ops.conv_wgrad(dout, sv["Xt"], dwm, K=C, N=D1, strides=(0, C, 1, 0))
        dz = run.gpool.view(m.z)
        tt = getattr(m, "_sd_tables_T", None)          # transposed tables, cached (they are constant buffers)
        if tt is None or tt[0].data_ptr() != m.cos.data_ptr():
            tt = (m.cos, m.cos.t().contiguous(), m.sin.t().contiguous())
            m._sd_tables_T = tt
        ops.sa_weights_bwd(dwm, sv["w_soft"], sv["mask"], tt[1], tt[2], K2, out=dz)
        grads[m.z] = torch.view_as_complex(dz)
        return None


def subject_tables_host(ids, S):
    """int32 [ids (B) | sample order sorted by subject (B) | group offsets (S+1)]: what the grouped GEMMs index with"""
    order = np.argsort(ids, kind="stable").astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(ids, minlength=S))]).astype(np.int32)
    return np.concatenate([ids.astype(np.int32), order, offsets])


class SubjectStage(Stage):
    """1x1 conv (bias) + per-sample subject 1x1 (grouped GEMM)   models.py:113-116."""

    def __init__(self, blk):
        self.m = blk

    def params(self):
        return [self.m.conv.weight, self.m.conv.bias] + [l.weight for l in self.m.subject_layer]

    def register(self, pack):
        D1 = self.m.D1
        self.key = "sb%d" % id(self.m)
        pack.add(self.key + ".conv", [self.m.conv.weight], D1, D1, 1)
        pack.add(self.key + ".subj", [l.weight for l in self.m.subject_layer], D1, D1, 1)

    def forward(self, run, x, sv):
        m = self.m
        D1, S = m.D1, len(m.subject_layer)
        B = x.shape[0]
        ids = run.subject_ids
        if ids is None or len(ids) != B:
            raise RuntimeError("sd_b200: subject_idxs must have one entry per sample")
        if run.static is not None:
            # CUDA-graph replay: ids / order / offsets of THIS step were staged in pinned memory by the host; the copy
            # below is a node of the graph and re-reads the staging buffer on every replay
            widx, d_order, d_off = run.static.subject_tables(B, S, x.device)
        else:
            host = subject_tables_host(ids, S)
            dev = torch.from_numpy(host).to(x.device, non_blocking=True)
            widx, d_order, d_off = dev[:B], dev[B:2 * B], dev[2 * B:]
        h1 = torch.empty_like(x)
        ops.conv_fwd(x, run.pack.wf(self.key + ".conv"), K=D1, N=D1, bias=m.conv.bias, out=h1)
        h2 = torch.empty_like(x)
        ops.conv_fwd(h1, run.pack.wf(self.key + ".subj"), K=D1, N=D1, widx=widx, G=S, out=h2)
        if sv is not None:
            # which per-subject weights get a gradient: those seen in the batch (absent -> None like the reference's
            # ModuleList); under a CUDA graph all of them do (zeros for the absent ones) and the optimizer step skips the
            # absent entries instead (sd_b200.graph)
            present = np.unique(ids) if run.static is None else np.arange(S)
            collect = None
            if run.host_group is not None:       # data parallel: a subject has a gradient if ANY rank saw it
                from . import dist as sd_dist    # (exchange started here, collected in backward: the host never blocks)
                collect = sd_dist.gather_host_ints_async(ids, run.host_group)
            sv.update(x=x, h1=h1, widx=widx, order=d_order, offsets=d_off, present=present, collect=collect)
        return h2

    def backward(self, run, sv, dout, grads, need_dx):
        m = self.m
        D1, S = m.D1, len(m.subject_layer)
        dws = run.gpool.stacked([l.weight for l in m.subject_layer])
        ops.conv_wgrad(dout, sv["h1"], dws, K=D1, N=D1, order=sv["order"], offsets=sv["offsets"], G=S,
                       strides=(D1 * D1, D1, 1, 0))
        if sv["collect"] is not None:
            sv["present"] = np.unique(np.concatenate(sv["collect"]()))
        for s in sv["present"]:                      # absent subjects keep grad None
            grads[m.subject_layer[int(s)].weight] = dws[int(s)]
        dh1 = torch.empty_like(dout)
        ops.conv_fwd(dout, run.pack.wd(self.key + ".subj"), K=D1, N=D1, widx=sv["widx"], G=S, out=dh1)
        dwc, dbc = run.gpool.view(m.conv.weight), run.gpool.view(m.conv.bias)
        ops.conv_wgrad(dh1, sv["x"], dwc, K=D1, N=D1, dbias=dbc)
        grads[m.conv.weight], grads[m.conv.bias] = dwc, dbc
        dx = torch.empty_like(dout)
        ops.conv_fwd(dh1, run.pack.wd(self.key + ".conv"), K=D1, N=D1, out=dx)
        return dx


class ConvBlockStage(Stage):
    """models.py:152-166."""

    @property
    def scratch(self):
        return 4 * rup8(self.m.D2)

    def __init__(self, blk):
        self.m = blk

    def params(self):
        m = self.m
        return [m.conv0.weight, m.conv0.bias, m.batchnorm0.weight, m.batchnorm0.bias,
                m.conv1.weight, m.conv1.bias, m.batchnorm1.weight, m.batchnorm1.bias,
                m.conv2.weight, m.conv2.bias]

    def register(self, pack):
        m = self.m
        self.key = "cb%d" % id(m)
        pack.add(self.key + ".c0", [m.conv0.weight], m.D2, m.in_channels, 3)
        pack.add(self.key + ".c1", [m.conv1.weight], m.D2, m.D2, 3)
        pack.add(self.key + ".c2", [m.conv2.weight], 2 * m.D2, m.D2, 3)

    @staticmethod
    def _bn_config(bn):
        """momentum / eps / mode of ONE BatchNorm1d module (models.py:135,143 use the defaults, but a loaded or edited
        module is honoured); configurations the kernels do not implement raise instead of being silently ignored."""
        if bn.momentum is None:
            raise NotImplementedError("sd_b200: BatchNorm1d(momentum=None) (cumulative moving average) is not supported")
        if not bn.track_running_stats or bn.running_mean is None or bn.running_var is None:
            raise NotImplementedError("sd_b200: BatchNorm1d(track_running_stats=False) is not supported")
        if not bn.affine:
            raise NotImplementedError("sd_b200: BatchNorm1d(affine=False) is not supported")
        return float(bn.momentum), float(bn.eps), bool(bn.training)

    def _bn(self, run, bn, stats, n, ss):
        momentum, eps, training = self._bn_config(bn)
        if training and stats is None:
            raise RuntimeError("sd_b200: training-mode BatchNorm needs the batch statistics of its producing conv")
        if training and run.bn_group is not None:
            import torch.distributed as dist
            from . import dist as sd_dist
            sd_dist.small_all_reduce_sum_(stats, run.bn_group)      # SyncBN: peer-memory exchange (NCCL without a mailbox)
            n = n * dist.get_world_size(run.bn_group)
        ops.bn_finalize(stats, bn.num_features, rup8(bn.num_features), n, bn.weight, bn.bias, bn.running_mean,
                        bn.running_var, bn.num_batches_tracked, momentum, eps, training, ss)
        return training

    def forward(self, run, x, sv):
        m = self.m
        B, T, _ = x.shape
        D2, Cin = m.D2, m.in_channels
        D2p, N2p = rup8(D2), rup8(2 * D2)
        d0, d1 = m.conv0.dilation[0], m.conv1.dilation[0]
        dev, dt = x.device, run.dtype
        train = m.batchnorm0.training
        if m.batchnorm1.training != train:
            raise NotImplementedError("sd_b200: the two BatchNorm layers of a ConvBlock must be in the same mode")
        stats = run.scratch.take(4 * D2p).view(2, 2 * D2p) if train else None
        ss = torch.empty((2, 4 * D2p), dtype=torch.float32, device=dev)
        if (not train and sv is None and dt == torch.bfloat16 and FUSE_EVAL_BN and D2 % 8 == 0
                and ops.get_impl() != "simt"):        # (the fused epilogue lives in the tensor-core kernel)
            # inference (eval mode, nothing saved for backward): BatchNorm is a per-channel affine known up front,
            # so BN + GELU ride in the epilogue of the conv that feeds them -- no y0/y1 tensors,
            # no separate normalisation pass
            self._bn(run, m.batchnorm0, None, B * T, ss[0])
            self._bn(run, m.batchnorm1, None, B * T, ss[1])
            u0 = torch.empty((B, T, D2p), dtype=dt, device=dev)
            ops.conv_fwd(x, run.pack.wf(self.key + ".c0"), K=Cin, N=D2, taps=3, dil=d0, bias=m.conv0.bias,
                         res=x if m.k != 0 else None, out=u0, act=nat.ACT_GELU, affine=ss[0])
            u1 = torch.empty_like(u0)
            ops.conv_fwd(u0, run.pack.wf(self.key + ".c1"), K=D2, N=D2, taps=3, dil=d1, bias=m.conv1.bias, res=u0,
                         out=u1, act=nat.ACT_GELU, affine=ss[1])
            out = torch.empty((B, T, D2p), dtype=dt, device=dev)
            ops.conv_fwd(u1, run.pack.wf(self.key + ".c2"), K=D2, N=2 * D2, taps=3, dil=m.conv2.dilation[0],
                         bias=m.conv2.bias, out=out, act=nat.ACT_GLU)      # y2 is not kept either
            return out
        y0 = torch.empty((B, T, D2p), dtype=dt, device=dev)
        ops.conv_fwd(x, run.pack.wf(self.key + ".c0"), K=Cin, N=D2, taps=3, dil=d0, bias=m.conv0.bias,
                     res=x if m.k != 0 else None, out=y0, stats=stats[0] if train else None)
        self._bn(run, m.batchnorm0, stats[0] if train else None, B * T, ss[0])
        u0 = torch.empty_like(y0)
        ops.bn_gelu_fwd(y0, ss[0], u0)
        y1 = torch.empty_like(y0)
        ops.conv_fwd(u0, run.pack.wf(self.key + ".c1"), K=D2, N=D2, taps=3, dil=d1, bias=m.conv1.bias, res=u0,
                     out=y1, stats=stats[1] if train else None)
        self._bn(run, m.batchnorm1, stats[1] if train else None, B * T, ss[1])
        u1 = torch.empty_like(y0)
        ops.bn_gelu_fwd(y1, ss[1], u1)
        y2 = torch.empty((B, T, N2p), dtype=dt, device=dev)
        out = torch.empty((B, T, D2p), dtype=dt, device=dev)
        ops.conv_fwd(u1, run.pack.wf(self.key + ".c2"), K=D2, N=2 * D2, taps=3, dil=m.conv2.dilation[0],
                     bias=m.conv2.bias, out=out, preact=y2, act=nat.ACT_GLU)
        if sv is not None:
            sv.update(x=x, y0=y0, u0=u0, y1=y1, u1=u1, y2=y2, ss=ss, train=train)
        return out

    def backward(self, run, sv, dout, grads, need_dx):
        m = self.m
        D2, Cin = m.D2, m.in_channels
        D2p = rup8(D2)
        d0, d1, d2 = m.conv0.dilation[0], m.conv1.dilation[0], m.conv2.dilation[0]
        dev = dout.device
        ss, train = sv["ss"], sv["train"]
        red = run.scratch.take(4 * D2p).view(2, 2 * D2p)
        zl = run.gpool.view

        # conv2 + GLU
        dy2 = torch.empty_like(sv["y2"])
        ops.glu_bwd(dout, sv["y2"], dy2, D2)
        dw2, db2 = zl(m.conv2.weight), zl(m.conv2.bias)
        ops.conv_wgrad(dy2, sv["u1"], dw2, K=D2, N=2 * D2, taps=3, dil=d2, dbias=db2)
        # The data-gradient convs that feed a BatchNorm backward carry its reduce pass in their epilogue (bf16 tensor-core
        # path): they store g = du * gelu'(bn(y)) and accumulate sum g, sum g*y, so only the apply pass remains
        fuse = ops.bn_bwd_fusable(sv["y1"])
        du1 = torch.empty_like(sv["u1"])
        ops.conv_fwd(dy2, run.pack.wd(self.key + ".c2"), K=2 * D2, N=D2, taps=3, dil=d2, out=du1,
                     **(dict(bnr_y=sv["y1"], bnr_ss=ss[1], stats=red[1]) if fuse else {}))
        del dy2
        # bn1 + gelu
        dg1, dbt1 = zl(m.batchnorm1.weight), zl(m.batchnorm1.bias)
        ops.bn_gelu_bwd(du1, sv["y1"], ss[1], red[1], dg1, dbt1, D2, train, run.bn_group, g_ready=fuse)
        dy1 = du1
        dw1, db1 = zl(m.conv1.weight), zl(m.conv1.bias)
        # bias of a conv that feeds a training-mode BatchNorm: its gradient sum_rows(dy) is identically zero (BN
        # backward removes the per-channel mean of dy; autograd gets rounding noise) -> left at the pool's zero and
        # the kernel skips its bias MMAs.  Eval-mode BN (running statistics) keeps the real bias gradient.
        ops.conv_wgrad(dy1, sv["u0"], dw1, K=D2, N=D2, taps=3, dil=d1, dbias=None if train else db1)
        du0 = torch.empty_like(dy1)
        ops.conv_fwd(dy1, run.pack.wd(self.key + ".c1"), K=D2, N=D2, taps=3, dil=d1, res=dy1, out=du0,
                     **(dict(bnr_y=sv["y0"], bnr_ss=ss[0], stats=red[0]) if fuse else {}))
        # bn0 + gelu
        dg0, dbt0 = zl(m.batchnorm0.weight), zl(m.batchnorm0.bias)
        ops.bn_gelu_bwd(du0, sv["y0"], ss[0], red[0], dg0, dbt0, D2, train, run.bn_group, g_ready=fuse)
        dy0 = du0
        dw0, db0 = zl(m.conv0.weight), zl(m.conv0.bias)
        ops.conv_wgrad(dy0, sv["x"], dw0, K=Cin, N=D2, taps=3, dil=d0, dbias=None if train else db0)
        grads.update({m.conv2.weight: dw2, m.conv2.bias: db2, m.batchnorm1.weight: dg1, m.batchnorm1.bias: dbt1,
                      m.conv1.weight: dw1, m.conv1.bias: db1, m.batchnorm0.weight: dg0, m.batchnorm0.bias: dbt0,
                      m.conv0.weight: dw0, m.conv0.bias: db0})
        if not need_dx:
            return None
        dx = torch.empty_like(sv["x"])
        ops.conv_fwd(dy0, run.pack.wd(self.key + ".c0"), K=D2, N=Cin, taps=3, dil=d0,
                     res=dy0 if m.k != 0 else None, out=dx)
        return dx


class FinalStage(Stage):
    """gelu(conv_final1) -> gelu(conv_final2) -> Z (B,F,T) fp32 contiguous   models.py:194-195."""

    def __init__(self, enc):
        self.m = enc

    def params(self):
        m = self.m
        return [m.conv_final1.weight, m.conv_final1.bias, m.conv_final2.weight, m.conv_final2.bias]

    def register(self, pack):
        m = self.m
        self.key = "fin%d" % id(m)
        pack.add(self.key + ".f1", [m.conv_final1.weight], 2 * m.D2, m.D2, 1)
        pack.add(self.key + ".f2", [m.conv_final2.weight], m.F, 2 * m.D2, 1)

    def forward(self, run, x, sv):
        m = self.m
        B, T, _ = x.shape
        N1, Fo = 2 * m.D2, m.F
        dev, dt = x.device, run.dtype
        p1 = torch.empty((B, T, rup8(N1)), dtype=dt, device=dev)
        u = torch.empty_like(p1)
        ops.conv_fwd(x, run.pack.wf(self.key + ".f1"), K=m.D2, N=N1, bias=m.conv_final1.bias, out=u, preact=p1,
                     act=nat.ACT_GELU)
        keep = sv is not None
        p2 = torch.empty((B, T, rup8(Fo)), dtype=dt, device=dev) if keep else None
        Z = torch.empty((B, Fo, T), dtype=torch.float32, device=dev)
        zn2 = torch.zeros((B,), dtype=torch.float32, device=dev)     # sum of Z^2 per sample, fused into the epilogue
        ops.conv_fwd(u, run.pack.wf(self.key + ".f2"), K=N1, N=Fo, bias=m.conv_final2.bias, out=Z, preact=p2,
                     rownorm2=zn2, act=nat.ACT_GELU, out_mode=nat.OUT_NCT_F32)
        run.out_norm2 = zn2
        if keep:
            sv.update(x=x, p1=p1, u=u, p2=p2)
        return Z

    def backward(self, run, sv, dZ, grads, need_dx):
        m = self.m
        N1, Fo = 2 * m.D2, m.F
        dZ = dZ.contiguous().float()
        dp2 = ops.gelu_bwd_nct(dZ, sv["p2"], Fo)
        dw2, db2 = run.gpool.view(m.conv_final2.weight), run.gpool.view(m.conv_final2.bias)
        ops.conv_wgrad(dp2, sv["u"], dw2, K=N1, N=Fo, dbias=db2)
        du = torch.empty_like(sv["u"])
        ops.conv_fwd(dp2, run.pack.wd(self.key + ".f2"), K=Fo, N=N1, out=du)
        del dp2
        ops.gelu_bwd(du, sv["p1"])
        dw1, db1 = run.gpool.view(m.conv_final1.weight), run.gpool.view(m.conv_final1.bias)
        ops.conv_wgrad(du, sv["x"], dw1, K=m.D2, N=N1, dbias=db1)
        grads.update({m.conv_final2.weight: dw2, m.conv_final2.bias: db2,
                      m.conv_final1.weight: dw1, m.conv_final1.bias: db1})
        if not need_dx:
            return None
        dx = torch.empty_like(sv["x"])
        ops.conv_fwd(du, run.pack.wd(self.key + ".f1"), K=N1, N=m.D2, out=dx)
        return dx


class ToBTC(Stage):
    """(B,C,T) fp32 -> internal layout (stand-alone ConvBlock entry)."""

    def forward(self, run, X, sv):
        if sv is not None:
            sv["C"] = X.shape[1]
        return ops.nct_to_btc(X, run.dtype)

    def backward(self, run, sv, dout, grads, need_dx):
        return ops.btc_to_nct(dout, sv["C"]) if need_dx else None


class ToNCT(Stage):
    """internal layout -> (B,C,T) fp32 (stand-alone sub-module exit)."""

    def __init__(self, C):
        self.C = C

    def forward(self, run, x, sv):
        return ops.btc_to_nct(x, self.C)

    def backward(self, run, sv, dout, grads, need_dx):
        return ops.nct_to_btc(dout.contiguous().float(), run.dtype)


# --------------------------------------------------------------------------------------------------
# pipeline + autograd node
# --------------------------------------------------------------------------------------------------
class Pipeline:
    def __init__(self, stages):
        self.stages = stages
        self.pack = WeightPack()
        for s in stages:
            s.register(self.pack)
        self.subject_ids = None
        self.dtype = None
        self.gpool = None
        self.scratch = None
        self.out_norm2 = None       # squared row norms of the pipeline output when the last stage produced them
        self._scratch_n = sum(st.scratch for st in stages)
        self.static = None          # graph.StaticInputs while a CUDA-graph step is being captured / replayed
        self.backward_hooks = {}    # stage index -> [one-shot callables] run right after that stage's backward is enqueued
        self.reducer = None         # dist.GradReducer: per-stage gradient all-reduce (data parallel)
        self.bn_group = None        # process group for SyncBN statistics, or None
        self.host_group = None      # gloo side channel for host-side metadata

    def params(self):
        out = []
        for s in self.stages:
            out.extend(s.params())
        return out

    def run(self, X, subject_ids=None):
        ops.require_cuda(X, "input")
        if X.dtype != torch.float32 and not (X.dtype == torch.bfloat16 and not X.requires_grad):
            X = X.float()            # (a bf16 input is consumed as it is: the layout kernel reads it directly)
        X = X.contiguous()
        self.subject_ids = subject_ids
        params = self.params()
        for p in params:
            ops.require_cuda(p, "parameter")
        # grad mode is switched off inside Function.forward, so decide here whether to keep activations
        self._keep = torch.is_grad_enabled() and (X.requires_grad or any(p.requires_grad for p in params))
        return _PipelineFn.apply(self, X, *params)

    # called from the autograd node
    def _forward(self, X, keep):
        self.dtype = ops.dtypes()[0]
        self.pack.refresh(X.device, self.dtype)
        saved = [dict() if keep else None for _ in self.stages]
        self.scratch = ScratchPool(self._scratch_n, X.device)
        h = X
        for s, sv in zip(self.stages, saved):
            h = s.forward(self, h, sv)
        return h, saved

    def _backward(self, saved, dout, need_dx):
        grads = {}
        g = dout
        n = len(self.stages)
        params = self.params()
        if self.gpool is None or set(self.gpool.offsets) != set(params):
            self.gpool = GradPool(params)
        self.gpool.new(dout.device)
        self.scratch = ScratchPool(self._scratch_n, dout.device)
        for i in range(n - 1, -1, -1):
            sg = {}
            g = self.stages[i].backward(self, saved[i], g, sg, need_dx or i > 0)
            saved[i] = None          # release activations as we go
            for hook in self.backward_hooks.pop(i, ()):
                hook()               # (data parallel: the deferred exchange of the next step's speech rows starts here)
            if self.reducer is not None and self.stages[i].params():
                # async all-reduce of this stage's contiguous slice overlaps the remaining backward
                self.reducer.stage_done(self.gpool.flat, *self.gpool.span(self.stages[i].params()))
            grads.update(sg)
        if self.reducer is not None:
            self.reducer.finish()
        return g, grads


class _PipelineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pipe, X, *params):
        keep = pipe._keep
        with torch.cuda.device(X.device), ops.stream_scope():
            pipe.out_norm2 = None
            out, saved = pipe._forward(X, keep)
        ctx.pipe, ctx.saved, ctx.params = pipe, saved, params
        ctx.precision = ops.get_precision()
        return out

    @staticmethod
    def backward(ctx, dout):
        pipe = ctx.pipe
        if ctx.saved is None or ctx.saved[0] is None:
            raise RuntimeError("sd_b200: backward called twice or without saved activations")
        prev = ops.get_precision()
        ops.set_precision(ctx.precision)
        try:
            with torch.cuda.device(dout.device), ops.stream_scope():
                pipe.dtype = ops.dtypes()[0]
                dx, grads = pipe._backward(ctx.saved, dout, ctx.needs_input_grad[1])
        finally:
            ops.set_precision(prev)
        ctx.saved = None
        out = [None, dx]
        for i, p in enumerate(ctx.params):
            g = grads.get(p) if ctx.needs_input_grad[2 + i] else None
            out.append(g)
        return tuple(out)


def normalize_subject_ids(subject_idxs, num_subjects):
    """Accept what the reference accepts (CPU IntTensor / LongTensor / list / ndarray,
    models.py:114-116 indexes a ModuleList with each element) -> int64 ndarray."""
    if isinstance(subject_idxs, torch.Tensor):
        ids = subject_idxs.detach().cpu().numpy()
    else:
        ids = np.asarray(subject_idxs)
    ids = ids.astype(np.int64).reshape(-1)
    if ids.size and (ids.min() < -num_subjects or ids.max() >= num_subjects):
        raise IndexError("index %d is out of range" % int(ids.max() if ids.max() >= num_subjects else ids.min()))
    return np.where(ids < 0, ids + num_subjects, ids)

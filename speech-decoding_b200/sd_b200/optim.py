"""Multi-tensor Adam: a drop-in for the reference's
`torch.optim.Adam(list(brain_encoder.parameters()) + list(loss_func.parameters()), lr=args.lr)` (train.py:161-163)
whose `step()` is ONE kernel launch over every parameter that has a gradient (sd_adam_step), instead of the dozen
foreach launches of the stock optimizer.  Same update rule, same per-parameter step counts (the weights of subjects
absent from a batch have `grad None` and are skipped, exactly like torch -- SURVEY 7.3.5), and a `state_dict()` in
torch.optim.Adam's format (`step`, `exp_avg`, `exp_avg_sq`), so optimizer checkpoints interchange."""
import ctypes
import math

import torch

from . import _native as nat
from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("FusedAdam: amsgrad / maximize are not supported")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False, fused=None))
        # entry tables travel through a small ring of pinned staging buffers, each guarded by an event: the host may
        # run several steps ahead of the GPU, and a staging buffer must not be rewritten before its copy has executed
        self._ring = []            # [host pinned uint8, device uint8, event]
        self._turn = 0

    def __getstate__(self):
        # pinned staging buffers and CUDA events are per-process scratch: never pickled / deep-copied
        d = self.__dict__.copy()
        d["_ring"], d["_turn"] = [], 0
        return d

    @staticmethod
    def _real(t):
        return torch.view_as_real(t) if t.is_complex() else t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # validate every parameter BEFORE touching any state: a step that raises must not leave step counts advanced
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                ops.require_cuda(p, "parameter")
                if p.dtype not in (torch.float32, torch.complex64) or p.grad.is_sparse:
                    raise TypeError("FusedAdam supports dense float32 / complex64 parameters")
                if not p.is_contiguous():
                    raise ValueError("FusedAdam needs contiguous parameters")
        ring_size = max(4, len(self.param_groups))       # one staging buffer per group and step in flight
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            entries, device, max_n = [], None, 0
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step = float(st["step"])
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                pr, gr, m, v = self._real(p), self._real(g), self._real(st["exp_avg"]), self._real(st["exp_avg_sq"])
                n = pr.numel()
                entries.append((pr.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), n,
                                group["lr"] / (1.0 - beta1 ** step), math.sqrt(1.0 - beta2 ** step), g))
                device = p.device
                max_n = max(max_n, n)
            if not entries:
                continue
            k = len(entries)
            nbytes = k * ctypes.sizeof(nat.AdamEntry)
            if len(self._ring) < ring_size:
                self._ring.append(None)
            self._turn = (self._turn + 1) % len(self._ring)
            slot = self._ring[self._turn]
            if slot is None or slot[0].numel() < nbytes or slot[1].device != device:
                host = torch.empty(max(nbytes, 8192), dtype=torch.uint8).pin_memory()
                slot = [host, torch.empty(host.numel(), dtype=torch.uint8, device=device), torch.cuda.Event()]
                self._ring[self._turn] = slot
            else:
                slot[2].synchronize()        # the copy that last used this staging buffer has executed
            host, dev_table, event = slot
            table = (nat.AdamEntry * k).from_address(host.data_ptr())
            for i, e in enumerate(entries):
                table[i] = nat.AdamEntry(e[0], e[1], e[2], e[3], e[4], e[5], e[6])
            with torch.cuda.device(device), ops.stream_scope():
                dev_table[:nbytes].copy_(host[:nbytes], non_blocking=True)
                event.record(torch.cuda.current_stream(device))
                nat.call("sd_adam_step", dev_table.data_ptr(), k, min(1024, (max_n + 1023) // 1024), float(beta1), float(beta2),
                         float(group["eps"]), float(group["weight_decay"]), ops._st())
            del entries                      # (keeps contiguous grad copies alive until the launch is enqueued)
        return loss

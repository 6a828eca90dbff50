"""GPU: sd_b200.optim.FusedAdam against torch.optim.Adam.  Kept in its own file, last in
collection order, so that the hot-path parity files run first under `pytest -x`."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b, floor=1e-30):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), floor)


def test_fused_adam_matches_torch_adam():
    """sd_b200.optim.FusedAdam (one launch) against torch.optim.Adam over several steps: float and
    complex parameters, a parameter that misses a step (grad None, like an absent subject's weight), weight decay,
    and an optimizer state_dict that loads into the stock optimizer."""
    from sd_b200.optim import FusedAdam
    torch.manual_seed(5)
    shapes = [(640, 320, 3), (320,), (270, 1024), (1,), (7, 5, 1)]
    for wd in (0.0, 1e-2):
        ref = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
        ref[2] = torch.randn(270, 1024, dtype=torch.cfloat, device=DEV).requires_grad_(True)
        mine = [p.detach().clone().requires_grad_(True) for p in ref]
        o_ref = torch.optim.Adam(ref, lr=3e-3, weight_decay=wd)
        o_mine = FusedAdam(mine, lr=3e-3, weight_decay=wd)
        for it in range(6):
            for a, b in zip(ref, mine):
                g = torch.randn_like(a) * (10.0 ** (it - 3))
                a.grad, b.grad = g.clone(), g.clone()
            if it in (1, 4):                          # this parameter gets no gradient on two of the steps
                ref[4].grad = None
                mine[4].grad = None
            o_ref.step()
            o_mine.step()
        for a, b in zip(ref, mine):
            assert rel(torch.view_as_real(b) if b.is_complex() else b, torch.view_as_real(a) if a.is_complex() else a) < 2e-6
        assert float(o_mine.state[mine[4]]["step"]) == 4.0 and float(o_ref.state[ref[4]]["step"]) == 4.0
        o2 = torch.optim.Adam(mine, lr=3e-3, weight_decay=wd)
        o2.load_state_dict(o_mine.state_dict())       # same state layout as torch.optim.Adam
        assert torch.equal(o2.state[mine[0]]["exp_avg"], o_mine.state[mine[0]]["exp_avg"])

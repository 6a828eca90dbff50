"""CPU: pin oracle/restate.py against the unmodified reference's outputs
(tests/golden/*.npz) and, when /root/reference is present, live."""
import numpy as np
import pytest
import torch

from oracle import ref_import, restate
from tests import golden_util as G


@pytest.mark.parametrize("name", G.names())
def test_restatement_matches_golden_train_step(name):
    g = G.load(name)
    c = g["cfg"]
    sd = {k: v.clone() for k, v in g["sd0"].items()}
    mask = restate.dropout_mask(g["loc"], float(c["d_drop"]), int(g["drop_center"]))
    out = restate.train_step(sd, g["X"], g["Y"], g["ids"].tolist(), g["temp"], mask,
                             reduction=str(c["reduction"]))
    assert G.rel_err(out["Z"], g["Z"]) < 1e-5
    assert G.rel_err(out["logits"], g["logits"]) < 1e-5
    assert G.rel_err(out["loss"], g["loss"]) < 1e-5
    assert G.rel_err(out["dZ"], g["dZ"]) < 1e-4
    assert G.rel_err(out["dtemp"], g["dtemp"]) < 1e-4
    for k, gr in g["grad"].items():
        assert out["grads"][k] is not None, k
        layer_scale = float(gr.abs().max())
        if k.endswith("bias"):      # appendix A.4: zero-by-construction biases
            wk = k[:-4] + "weight"
            layer_scale = max(layer_scale, float(g["grad"][wk].abs().max()))
        assert G.rel_err(out["grads"][k], gr, floor=layer_scale) < 2e-4, k
    for k in g["absent_grads"]:
        assert out["grads"][k] is None, k
    for k, v in g["sd1"].items():
        assert G.rel_err(sd[k].float(), v.float(), floor=1e-6) < 1e-5, k


@pytest.mark.parametrize("name", G.names())
def test_restatement_eval_slow_and_classifier(name):
    g = G.load(name)
    c = g["cfg"]
    sd = {k: v.clone() for k, v in g["sd0"].items()}
    sd.update({k: v.clone() for k, v in g["sd1"].items()})
    with torch.no_grad():
        Ze = restate.encoder_forward(sd, g["X"], g["ids"].tolist(), train=False)
        assert G.rel_err(Ze, g["Z_eval"]) < 1e-5
        le = restate.clip_loss(g["Y"], Ze, g["temp"], reduction=str(c["reduction"]))
        assert G.rel_err(le, g["loss_eval"]) < 1e-5
        ls = restate.clip_loss(g["Y"], g["Z"], g["temp"], reduction=str(c["reduction"]), fast=False)
        assert G.rel_err(ls, g["loss_slow"]) < 1e-5
        top1, top10, _, _ = restate.classifier(g["Z"], g["Y"])
        assert abs(top1 - float(g["top1"])) < 1e-6
        assert abs(top10 - float(g["top10"])) < 1e-6


@pytest.mark.parametrize("name", G.names())
def test_clip_closed_form_matches_golden(name):
    g = G.load(name)
    loss, L, dy, dtemp = restate.clip_loss_closed_form(g["Y"].numpy(), g["Z"].numpy(), float(g["temp"]),
                                                      reduction=str(g["cfg"]["reduction"]))
    assert abs(loss - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    assert np.abs(L - g["logits"].numpy()).max() / np.abs(L).max() < 1e-5
    ref = g["dZ"].numpy().reshape(dy.shape)
    assert np.abs(dy - ref).max() / np.abs(ref).max() < 1e-4
    assert abs(dtemp - float(g["dtemp"])) / max(abs(float(g["dtemp"])), 1e-6) < 1e-4


def test_fourier_tables_match_reference_buffers():
    for name in G.names():
        g = G.load(name)
        cos, sin = restate.fourier_tables(int(g["cfg"]["K"]), g["loc"])
        assert torch.allclose(cos, g["sd0"]["subject_block.spatial_attention.cos"], atol=1e-6)
        assert torch.allclose(sin, g["sd0"]["subject_block.spatial_attention.sin"], atol=1e-6)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_restatement_matches_live_reference():
    torch.manual_seed(11)
    np.random.seed(11)
    args = restate.make_args(D1=20, D2=24, F_=40, K=4, d_drop=0.2, num_subjects=5, num_channels=17,
                             last4layers=False, layout_seed=5)
    M, L = ref_import.load(lambda a: restate.synthetic_layout(a.num_channels, a.layout_seed))
    enc, crit = M.BrainEncoder(args), L.CLIPLoss(args)
    X, Y = torch.randn(12, 17, 50), torch.randn(12, 40, 50)
    ids = torch.randint(0, 5, (12,))
    sd = {k: v.clone() for k, v in enc.state_dict().items()}
    st = np.random.get_state(); center = np.random.randint(17); np.random.set_state(st)
    enc.train()
    Z = enc(X, ids); loss = crit(Y, Z); loss.backward()
    mask = restate.dropout_mask(restate.synthetic_layout(17, 5), 0.2, center)
    out = restate.train_step(sd, X, Y, ids.tolist(), crit.temp.detach(), mask)
    assert G.rel_err(out["Z"], Z) < 1e-5
    assert G.rel_err(out["loss"], loss) < 1e-5
    for n, p in enc.named_parameters():
        if p.grad is None:
            assert out["grads"][n] is None
        elif not n.endswith("bias"):
            assert G.rel_err(out["grads"][n], p.grad) < 2e-4, n

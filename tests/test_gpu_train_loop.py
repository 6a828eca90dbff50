"""GPU: the call sequence of the reference's train.py (train.py:148-259) against the drop-in, on synthetic
loaders: construct -> .to(device) -> Adam over encoder+loss params -> [X,Y to device; forward; loss;
Classifier under no_grad; loss.item(); zero_grad; backward; step] -> eval loop -> state_dict round trip.
The loss must go down on a fixed batch and absent subjects must be skipped by Adam (grad None)."""
import numpy as np
import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_reference_training_loop_runs_on_the_dropin(precision, tmp_path):
    import sd_b200
    from speech_decoding.models import BrainEncoder, Classifier       # train.py:22
    from speech_decoding.utils.loss import CLIPLoss, MSELoss           # train.py:24 (star import)
    sd_b200.set_precision(precision)
    device = "cuda:0"                                                  # train.py:31
    torch.manual_seed(0); np.random.seed(0)
    args = restate.make_args(D1=48, D2=64, F_=96, K=6, num_subjects=5, num_channels=24, last4layers=False)
    args.lr, args.epochs = 3e-3, 12
    B, C, T = 16, 24, 120
    X = torch.randn(B, C, T).clamp(-20, 20); Y = torch.randn(B, 96, T)
    subject_idxs = torch.IntTensor([0, 1, 2, 0, 1, 2, 3, 3, 0, 1, 2, 3, 0, 1, 2, 3])   # subject 4 never appears
    brain_encoder = BrainEncoder(args).to(device)                      # train.py:148
    classifier = Classifier(args)                                      # train.py:150
    loss_func = CLIPLoss(args).to(device)                              # train.py:155
    loss_func.train()                                                  # train.py:156
    optimizer = torch.optim.Adam(list(brain_encoder.parameters()) + list(loss_func.parameters()), lr=float(args.lr))  # :161-163
    w_absent0 = brain_encoder.subject_block.subject_layer[4].weight.detach().clone()
    losses = []
    for epoch in range(args.epochs):                                   # train.py:166
        brain_encoder.train()                                          # train.py:174
        Xd, Yd = X.to(device), Y.to(device)                            # train.py:187
        Z = brain_encoder(Xd, subject_idxs)                            # train.py:189 (ids stay on the CPU)
        loss = loss_func(Yd, Z)                                        # train.py:191
        with torch.no_grad():
            top1, top10 = classifier(Z, Yd)                            # train.py:193-194
        losses.append(loss.item())                                     # train.py:196
        optimizer.zero_grad()                                          # train.py:201
        loss.backward()                                                # train.py:202
        optimizer.step()                                               # train.py:203
        assert 0.0 <= top1 <= 1.0 and 0.0 <= top10 <= 1.0
    assert losses[-1] < losses[0], losses
    assert torch.equal(brain_encoder.subject_block.subject_layer[4].weight.detach().cpu(), w_absent0.cpu())
    brain_encoder.eval()                                               # train.py:211
    with torch.no_grad():                                              # train.py:213-233
        Z = brain_encoder(X.to(device), subject_idxs)
        test_loss = loss_func(Y.to(device), Z)
        assert torch.isfinite(test_loss) and isinstance(loss_func.temp.item(), float)   # train.py:242
    path = tmp_path / "model_last.pt"
    torch.save(brain_encoder.state_dict(), path)                       # train.py:259
    again = BrainEncoder(args).to(device)
    again.load_state_dict(torch.load(path))
    again.eval()
    with torch.no_grad():
        assert torch.allclose(again(X.to(device), subject_idxs), Z)

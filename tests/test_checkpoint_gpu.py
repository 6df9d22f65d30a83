"""GPU: checkpoint / resume in the reference's file layout (thumos14/train.py:106-131) and `get_grad_norm` (:132-139).
A resumed Trainer holds bit-identical parameters / Adam moments / loss state, and the 'optimizer' entry of the training
state file drives a stock torch.optim.Adam to the same update as the fused Adam kernel."""
import pytest
import torch

from opental_b200 import engine, ops

pytestmark = pytest.mark.gpu


def _batch(B=1, frames=256):
    clips = torch.stack([engine.normalise_clip(engine.synthetic_clip_u8(i, 0, frames)) for i in range(B)]).cuda()
    targets = [engine.synthetic_targets(i, 0).cuda() for i in range(B)]
    scores = torch.stack([engine.synthetic_scores(t.cpu(), frames) for t in targets]).cuda()
    return clips, targets, scores


def _trainer(epoch):
    net, crit = engine.build_opental(epoch=epoch)
    return net, crit, engine.Trainer(net, crit)


def test_resume_is_bit_exact_and_matches_torch_adam(tmp_path):
    clips, targets, scores = _batch()
    net, crit, tr = _trainer(11)
    for _ in range(2):
        tr.step(clips, targets, scores)
    # get_grad_norm: the reference's per-parameter formulation vs the three flat reductions
    ref_norm = torch.sqrt(sum(p.grad.detach().norm(2) ** 2 for p in net.parameters() if p.grad is not None and p.requires_grad))
    assert abs(float(tr.grad_norm()) - float(ref_norm)) <= 1e-5 * float(ref_norm)
    ckpt, state = str(tmp_path / "checkpoint"), str(tmp_path / "train_state")
    model_file, state_file = tr.save_checkpoint(2, ckpt, state)
    sd = torch.load(model_file)
    assert list(sd) == list(net.state_dict()) and all(v.device.type == "cpu" for v in sd.values())
    st = torch.load(state_file, weights_only=False)
    assert set(st) >= {"optimizer", "state"} and len(st["optimizer"]["state"]) == 161

    net2, crit2, tr2 = _trainer(1)
    with torch.no_grad():
        for p in net2.parameters():
            p.add_(0.01)                                    # make sure resume really overwrites (views: padding untouched)
    assert tr2.resume(2, ckpt, state) == 3
    assert tr2.step_count == 2 and crit2.cls_loss.epoch == 11
    assert torch.equal(crit2.cls_loss.weight_accum, crit.cls_loss.weight_accum)
    for (wa, _), (wb, _), sa, sb in zip(tr.groups, tr2.groups, tr.state, tr2.state):
        assert torch.equal(wa, wb) and torch.equal(sa["m"], sb["m"]) and torch.equal(sa["v"], sb["v"])
    for a, b in zip(net.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)
    # the next step of both sees the same forward (wgrad accumulates with float atomics: the updates agree to rounding)
    cost_a = tr.step(clips, targets, scores)[0]
    cost_b = tr2.step(clips, targets, scores)[0]
    assert torch.equal(cost_a, cost_b)
    for (wa, _), (wb, _) in zip(tr.groups, tr2.groups):
        assert float((wa - wb).abs().mean()) <= 1e-2 * tr.lr

    # stock torch.optim.Adam loaded from the file vs the fused kernel, same gradients, same starting point
    net3, crit3, tr3 = _trainer(1)
    tr3.resume(2, ckpt, state)
    opt = torch.optim.Adam(net3.parameters(), lr=1.0)
    opt.load_state_dict(st["optimizer"])
    tr3.zero_grad()
    tr3.forward_backward(clips, targets, scores)
    start = [w.clone() for w, _ in tr3.groups]
    opt.step()
    w_torch = [w.clone() for w, _ in tr3.groups]
    for (w, g), w0, s in zip(tr3.groups, start, tr3.state):
        w.copy_(w0)
        ops.adam_step(w, g, s["m"], s["v"], lr=tr3.lr, betas=tr3.betas, eps=tr3.eps, weight_decay=tr3.wd, step=3)
    for (w, _), wt, w0 in zip(tr3.groups, w_torch, start):
        assert float((wt - w0).abs().max()) > 1e-6          # torch really stepped the views
        assert torch.allclose(w, wt, rtol=3e-7, atol=3e-8)

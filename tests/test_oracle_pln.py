"""CPU pins of the PLN oracle: closed-form gradient vs autograd, and the gathered-loss rule."""
import torch

from oracle import pln as opln
from osr_b200 import synth

KW = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)


def test_closed_form_gradient_matches_autograd():
    pi = synth.make_pln_inputs(256, seed=3)
    emb = (pi.roi_features @ pi.enc_w.t()).double().requires_grad_(True)
    reps = pi.reps.double().requires_grad_(True)
    loss = opln.pln_loss_from_emb(emb, reps, pi.gt_classes, pi.ious.double(), **KW)
    loss.backward()
    ge, gr = opln.pln_loss_grad_closed_form(emb.detach(), reps.detach(), pi.gt_classes, pi.ious.double(), **KW)
    assert (emb.grad - ge).abs().max() < 1e-10
    assert (reps.grad - gr).abs().max() < 1e-10


def test_loss_ignores_background_unknown_and_low_iou():
    pi = synth.make_pln_inputs(64, seed=4)
    emb = pi.roi_features @ pi.enc_w.t()
    l1 = opln.pln_loss_from_emb(emb, pi.reps, pi.gt_classes, pi.ious, **KW)
    fg = (pi.gt_classes < 20) & (pi.ious > 0.5)
    emb2 = emb.clone(); emb2[~fg] = 123.0
    l2 = opln.pln_loss_from_emb(emb2, pi.reps, pi.gt_classes, pi.ious, **KW)
    assert torch.equal(l1, l2)
    # normaliser is ALL rows (prototype_learning_network.py:187), not the fg count
    l3 = opln.pln_loss_from_emb(emb[:32], pi.reps, pi.gt_classes[:32], pi.ious[:32], **KW)
    assert l3 != l1


def test_gathered_rule_equals_ddp_mean():
    """(1/W) sum_r L_r == gathered loss with r_norm = W*R_loc and center_weight = W."""
    W, Rl = 4, 48
    pis = [synth.make_pln_inputs(Rl, seed=10 + r) for r in range(W)]
    reps = pis[0].reps
    embs = [p.roi_features @ pis[0].enc_w.t() for p in pis]
    per_rank = [opln.pln_loss_from_emb(e, reps, p.gt_classes, p.ious, **KW) for e, p in zip(embs, pis)]
    ddp_mean = sum(per_rank) / W
    g = opln.pln_loss_from_emb(torch.cat(embs), reps, torch.cat([p.gt_classes for p in pis]),
                               torch.cat([p.ious for p in pis]), r_norm=W * Rl, center_weight=W, **KW)
    torch.testing.assert_close(g, ddp_mean, rtol=1e-6, atol=1e-7)


def test_inference_unknown_threshold():
    pi = synth.make_pln_inputs(40, seed=5)
    rec, pred = opln.pln_inference(pi.roi_features, pi.enc_w, pi.enc_b, pi.dec_w, pi.dec_b, pi.reps,
                                   num_known_classes=20, unk_thr=0.9, unknown_id=80)
    assert rec.shape == (40, 1024)
    assert ((pred == 80) | ((pred >= 0) & (pred < 20))).all()

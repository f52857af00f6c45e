"""Generates tests/golden/golden_ref_pln_dist_v1.npz by EXECUTING THE UNMODIFIED REFERENCE PLN
(/root/reference/openset_rcnn/modeling/roi_heads/prototype_learning_network.py) with MODEL.PLN.DISTANCE_TYPE = 'L1' and
'L2' (and one 'COS' case with two representatives per class): ``PLN.loss`` (:117-187) with autograd gradients, and
``PLN.inference`` (:189-226).  Same mechanism as make_golden_ref.py (detectron2 stand-in in sys.modules, 'cuda' -> 'cpu'
device rewriting); nothing is imported from ``oracle/`` or from the product package.

alpha / beta / unk_thr are set per case to the (rounded) medians of the intra / inter / nearest distances of the case's
inputs - plain attributes of the reference module - so that every hinge (intra, inter, prototype separation) and both
inference outcomes (known / unknown) occur: the script asserts it.

Run:  python tests/golden/make_golden_pln_dist.py     (deterministic; commit the .npz with this script)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import d2shim  # noqa: E402
from make_golden_ref import REF_ROOT, CudaToCpu  # noqa: E402

CASES = [   # tag, distance_type, reps_per_class
    ("l1", "L1", 1),
    ("l2", "L2", 1),
    ("l2r2", "L2", 2),
    ("l1r2", "L1", 2),
    ("cosr2", "COS", 2),
]


def _dists(pln, emb, dt, rpc, K):
    eh = torch.nn.functional.normalize(emb)
    rh = torch.nn.functional.normalize(pln.representatives)
    d = {"L1": torch.cdist(eh, rh, p=1.0), "L2": torch.cdist(eh, rh), "COS": 1 - eh @ rh.t()}[dt]
    return d.reshape(emb.shape[0], K, rpc).min(dim=2)[0]


def main():
    ref = d2shim.import_reference(REF_ROOT)
    from detectron2.structures import Instances
    out = {}
    R, FC, EMB, NUM_CLASSES, NUM_KNOWN = 96, 64, 256, 81, 20
    out["cases"] = np.array([c[0] for c in CASES])
    for ci, (tag, dt, rpc) in enumerate(CASES):
        g = torch.Generator().manual_seed(4100 + ci)
        torch.manual_seed(900 + ci)
        with CudaToCpu():
            pln = ref.prototype_learning_network.PLN(
                num_classes=NUM_CLASSES, num_known_classes=NUM_KNOWN, feature_dim=FC, embedding_dim=EMB,
                distance_type=dt, reps_per_class=rpc, alpha=0.1, beta=0.9, loss_weight=0.5,
                dataset_name="voc_2007_train", iou_threshold=0.5, unk_thr=0.23, opendet_benchmark=True)
        with torch.no_grad():   # a livelier encoder than N(0, 0.01): embeddings spread over the sphere
            pln.encoder.weight.copy_(torch.randn(EMB, FC, generator=g) * 0.3)
            pln.encoder.bias.copy_(torch.randn(EMB, generator=g) * 0.1)
        x = torch.randn(R, FC, generator=g)
        # two images; labels: known classes, some background (80) and some unknown-ish ids, ious around the 0.5 threshold
        cls = torch.randint(0, NUM_KNOWN, (R,), generator=g)
        cls = torch.where(torch.rand(R, generator=g) < 0.2, torch.full_like(cls, 80), cls)
        cls = torch.where(torch.rand(R, generator=g) < 0.1, torch.full_like(cls, 45), cls)
        ious = torch.rand(R, generator=g) * 0.7 + 0.3
        props = []
        for lo, hi in ((0, 40), (40, R)):
            p = Instances((100, 100))
            p.gt_classes = cls[lo:hi].clone()
            p.ious = ious[lo:hi].clone()
            props.append(p)
        feats = [torch.randn(30, FC, generator=g), torch.randn(21, FC, generator=g)]
        with torch.no_grad():   # thresholds in the middle of this case's distance distributions
            dmin = _dists(pln, pln.encoder(x), dt, rpc, NUM_KNOWN)
            fg = (cls < NUM_KNOWN) & (ious > 0.5)
            yc = cls.clamp(max=NUM_KNOWN - 1)
            intra = dmin[torch.arange(R), yc][fg]
            dm = dmin.clone()
            dm[torch.arange(R), yc] = 1000
            inter = dm.min(dim=1)[0][fg]
            near = _dists(pln, pln.encoder(torch.cat(feats)), dt, rpc, NUM_KNOWN).min(dim=1)[0]
            alpha, beta = round(float(intra.median()), 3), round(float(inter.median()), 3)
            unk_thr = round(float(near.median()), 3)
            fa, fb = float((intra > alpha).float().mean()), float((inter < beta).float().mean())
            assert 0.2 < fa < 0.8 and 0.2 < fb < 0.8, (tag, fa, fb)
        pln.alpha, pln.beta, pln.unk_thr = alpha, beta, unk_thr
        xg = x.clone().requires_grad_(True)
        with CudaToCpu():
            emb, rec, loss = pln.loss(xg, props)
        emb.retain_grad()
        loss.backward()
        # inference on two images of fresh features
        fgi = []
        for f in feats:
            q = Instances((100, 100))
            q.features = f.clone()
            fgi.append(q)
        with CudaToCpu(), torch.no_grad():
            res = pln.inference(fgi)
        pred = torch.cat([r.pred_classes for r in res])
        fu = float((pred == 80).float().mean())
        assert 0.1 < fu < 0.9, (tag, fu)
        print(f"{tag}: alpha {alpha} beta {beta} unk_thr {unk_thr} loss {float(loss.detach()):.6f}  intra-active {fa:.2f}  inter-active {fb:.2f}  unknown {fu:.2f}")
        o = {
            "params": np.array([rpc, alpha, beta, unk_thr, 0.5, 0.5], dtype=np.float64),   # rpc, alpha, beta, unk_thr, loss_weight, iou_thr
            "dist": np.array(dt),
            "x": x.numpy(), "enc_w": pln.encoder.weight.detach().clone().numpy(),
            "enc_b": pln.encoder.bias.detach().clone().numpy(), "dec_w": pln.decoder.weight.detach().clone().numpy(),
            "dec_b": pln.decoder.bias.detach().clone().numpy(), "reps": pln.representatives.detach().clone().numpy(),
            "gt_classes": cls.numpy(), "ious": ious.numpy(), "split": np.array([40, R - 40]),
            "emb": emb.detach().clone().numpy(), "rec": rec.detach().clone().numpy(), "loss": loss.detach().clone().numpy(),
            "grad_emb": emb.grad.clone().numpy(), "grad_reps": pln.representatives.grad.clone().numpy(),
            "grad_enc_w": pln.encoder.weight.grad.clone().numpy(), "grad_x": xg.grad.clone().numpy(),
            "inf_feats0": feats[0].numpy(), "inf_feats1": feats[1].numpy(),
            "inf_pred0": res[0].pred_classes.numpy(), "inf_pred1": res[1].pred_classes.numpy(),
            "inf_rec0": res[0].features.numpy(), "inf_rec1": res[1].features.numpy(),
        }
        for k, v in o.items():
            out[f"{tag}_{k}"] = v
    path = os.path.join(HERE, "golden_ref_pln_dist_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

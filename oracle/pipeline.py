"""CPU restatement of one RoI-path step (the reference's torch/torchvision code path) - used ONLY as the
``cpu_baseline`` / ``--impl reference`` arm of bench.py and by smoke()/tests as the checker.

Follows the same stage list as osr_b200/pipeline.py: find_top_rpn_proposals (as shipped) -> matcher labels of all kept
proposals + the labelled 512-per-image sample (label_and_sample_proposals, osrcnn_roi_heads.py:136-230, without the
ground-truth append and without Instances) -> ROIPooler forward (4 x torchvision roi_align) -> PLN.loss forward with the
sampled classes / IoUs -> backward through PLN and ROIPooler.
"""
from __future__ import annotations

import time
from typing import Dict

import torch
import torch.nn.functional as F

from . import pln as opln
from . import roi_align as ora
from . import rpn as orpn
from . import sampling as osamp
from .structures import Boxes, pairwise_iou


class CpuRoiPath:
    def __init__(self, head_outputs, feats, pln_inputs, grad_pooled, sample_idx_per_image, *, pre_nms_topk=2000,
                 rois_per_image=512, num_known=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5,
                 pool_scales=(0.25, 0.125, 0.0625, 0.03125), targets=None, keys=None, num_classes=81, positive_fraction=0.25):
        """``targets``: list of ``(gt_boxes (G,4), gt_classes (G))`` per image - S2 is then the reference's labelled sampling
        (matcher + ``subsample_labels``; ``keys`` = per-image random keys replays the CUDA sampler's draw, else
        ``torch.randperm``) and the prototype loss consumes the sampled classes / IoUs.  Without targets: the fixed
        ``sample_idx_per_image`` rows and the synthetic labels of ``pln_inputs`` (the pre-sampler form)."""
        self.ho = head_outputs
        self.feats = [f.detach().clone().requires_grad_(True) for f in feats]
        self.pi = pln_inputs
        self.grad_pooled = grad_pooled
        self.sample_idx = sample_idx_per_image
        self.targets, self.keys, self.num_classes, self.pos_frac = targets, keys, num_classes, positive_fraction
        self.pre_k = pre_nms_topk
        self.rpi = rois_per_image
        self.kw = dict(num_known_classes=num_known, alpha=alpha, beta=beta, loss_weight=loss_weight,
                       iou_threshold=iou_threshold)
        self.pooler = ora.ROIPooler(7, pool_scales, 0, "ROIAlignV2")
        self.anchors = [Boxes(a) for a in head_outputs.anchors]

    def step(self) -> Dict[str, float]:
        t0 = time.perf_counter()
        props = orpn.predict_proposals(self.anchors, self.ho.deltas, self.ho.centerness, self.ho.image_sizes,
                                       pre_nms_topk=self.pre_k, post_nms_topk=self.pre_k, training=True,
                                       mode="as_shipped", topk_impl="torch")
        t1 = time.perf_counter()
        pi = self.pi
        if self.targets is None:
            sample_idx = self.sample_idx
            labels, ious = pi.gt_classes, pi.ious
        else:
            sample_idx, cls_l, iou_l = [], [], []
            for n, (p, (gb, gc)) in enumerate(zip(props, self.targets)):
                m = pairwise_iou(Boxes(gb), p.proposal_boxes)                                  # osrcnn_roi_heads.py:187-189
                midx, mlab = osamp.matcher(m, self.kw["iou_threshold"])                         # :190
                miou = m[midx, torch.arange(m.shape[1])]                                        # :193
                cls = gc[midx].clone()
                cls[mlab == 0] = self.num_classes
                if self.keys is not None:
                    k = self.keys[n][:len(p)]
                    rp = osamp.keyed_randperm([k[(cls != -1) & (cls != self.num_classes)], k[cls == self.num_classes]])
                else:
                    rp = torch.randperm
                fg, bg = osamp.subsample_labels(cls, self.rpi, self.pos_frac, self.num_classes, rp)   # :195-197
                s = torch.cat([fg, bg])
                sample_idx.append(s); cls_l.append(cls[s]); iou_l.append(miou[s])
            labels, ious = torch.cat(cls_l), torch.cat(iou_l)
        boxes = [Boxes(p.proposal_boxes.tensor[i]) for p, i in zip(props, sample_idx)]
        t2 = time.perf_counter()
        pooled = self.pooler.forward(self.feats, boxes)
        t3 = time.perf_counter()
        M = pooled.shape[0]
        emb = F.linear(pi.roi_features[:M], pi.enc_w, pi.enc_b).requires_grad_(True)
        reps = pi.reps.detach().clone().requires_grad_(True)
        loss = opln.pln_loss_from_emb(emb, reps, labels, ious, **self.kw)
        g_emb, g_reps = torch.autograd.grad(loss, [emb, reps])
        t4 = time.perf_counter()
        g_feats = torch.autograd.grad(pooled, self.feats, self.grad_pooled[:M])
        t5 = time.perf_counter()
        self.last = dict(props=props, sample_idx=sample_idx, labels=labels, ious=ious, pooled=pooled, loss=loss, g_emb=g_emb, g_reps=g_reps, g_feats=g_feats)
        return {"s1_proposals": t1 - t0, "s2_sample_glue": t2 - t1, "s3_roialign_fwd": t3 - t2,
                "s5_pln_fwd_bwd": t4 - t3, "s3_roialign_bwd": t5 - t4, "total": t5 - t0}

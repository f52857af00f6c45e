"""Registry swap-in: when detectron2 AND the reference package are importable, the reference's
own classes are subclassed with the osr_b200 operators mixed in and registered under new names, so a config switches
to the B200 path with two keys and no edit of the reference's files:

    MODEL.PROPOSAL_GENERATOR.NAME: "OsrClsFreeRPN"        # instead of "ClsFreeRPN"   (classification_free_rpn.py:165-166)
    MODEL.ROI_HEADS.NAME:          "OsrOpensetROIHeads"   # instead of "OpensetROIHeads" (osrcnn_roi_heads.py:26-27)

    import osr_b200.registry as R; R.register_with_detectron2()     # once, before build_model(cfg)

What changes relative to the reference classes (everything else - heads, losses, box predictor, softmax classifier - is
inherited untouched):
  * ``ClsFreeRPN.predict_proposals``            -> ``osr_b200.proposals.predict_proposals``      (:558-610 + find_top_proposals.py)
  * ``OpensetROIHeads.label_and_sample_proposals`` -> ``osr_b200.sampling.label_and_sample_proposals`` (:136-230)
  * ``box_pooler``                               -> ``osr_b200.poolers.ROIPooler``                (built :108-113, called :306)
  * ``dml``                                      -> ``osr_b200.pln.PLN``                          (built :125, called :315 / :325)
The mixins only read attributes the reference classes define; they carry no detectron2 import themselves.
"""
from __future__ import annotations

from typing import List

import torch


class OsrProposalMixin:
    """``predict_proposals`` of ``ClsFreeRPN`` on the fused select-decode kernel.  ``proposal_mode`` = "as_shipped" (the
    reference as it is: NMS commented out) or "nominal" (stock detectron2)."""

    proposal_mode = "as_shipped"

    def predict_proposals(self, anchors, pred_anchor_deltas, pred_centerness, image_sizes):
        from .proposals import predict_proposals
        return predict_proposals(anchors, pred_anchor_deltas, pred_centerness, image_sizes,
                                 nms_thresh=self.nms_thresh[self.training], pre_nms_topk=self.pre_nms_topk[self.training],
                                 post_nms_topk=self.post_nms_topk[self.training], min_box_size=self.min_box_size,
                                 training=self.training, mode=self.proposal_mode)


class OsrRoiHeadsMixin:
    """``label_and_sample_proposals`` of ``OpensetROIHeads`` on the fused matching kernel + batched sampler, and the
    B200 pooler / PLN as ``box_pooler`` / ``dml``."""

    @torch.no_grad()
    def label_and_sample_proposals(self, proposals, targets):
        from .sampling import label_and_sample_proposals
        thr = self.proposal_matcher.thresholds[1]           # Matcher([thr], [0, 1]): (-inf, thr, +inf)
        out = label_and_sample_proposals(proposals, targets, num_classes=self.num_classes,
                                         batch_size_per_image=self.batch_size_per_image,
                                         positive_fraction=self.positive_fraction, iou_threshold=thr,
                                         proposal_append_gt=self.proposal_append_gt)
        try:   # the reference logs the fg / bg sample counts (osrcnn_roi_heads.py:225-228)
            from detectron2.utils.events import get_event_storage
            storage = get_event_storage()
            cls = torch.stack([(x.gt_classes == self.num_classes).sum() for x in out]).float()
            n = torch.tensor([float(len(x)) for x in out], device=cls.device)
            stats = torch.stack(((n - cls).mean(), cls.mean())).tolist()
            storage.put_scalar("roi_head/num_fg_samples", stats[0])
            storage.put_scalar("roi_head/num_bg_samples", stats[1])
        except Exception:  # noqa: BLE001 - no event storage outside a training loop
            pass
        return out

    @classmethod
    def _init_box_head(cls, cfg, input_shape):
        from .pln import PLN
        from .poolers import ROIPooler
        ret = super()._init_box_head(cfg, input_shape)
        old = ret["box_pooler"]
        scales = tuple(1.0 / input_shape[k].stride for k in cfg.MODEL.ROI_HEADS.IN_FEATURES)
        ret["box_pooler"] = ROIPooler(output_size=old.output_size, scales=scales,
                                      sampling_ratio=cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO,
                                      pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE)
        ref_pln = ret["dml"]
        kw = PLN.from_config(cfg)
        if not kw["opendet_benchmark"]:
            kw["known_class_ids"] = [int(v) for v in ref_pln.class_id.tolist()]
        pln = PLN(**kw)
        pln.load_state_dict(ref_pln.state_dict())     # same parameter names and shapes: keep the reference's initial values
        ret["dml"] = pln
        return ret


_REGISTERED = {}


def register_with_detectron2(names=("OsrClsFreeRPN", "OsrOpensetROIHeads")) -> dict:
    """Create and register the two subclasses; returns ``{name: class}``.  Raises ImportError when detectron2 or the
    reference package (``openset_rcnn``) is missing - there is nothing to swap into then."""
    if _REGISTERED:
        return dict(_REGISTERED)
    from detectron2.modeling import PROPOSAL_GENERATOR_REGISTRY
    try:
        from detectron2.modeling import ROI_HEADS_REGISTRY
    except ImportError:
        from detectron2.modeling.roi_heads.roi_heads import ROI_HEADS_REGISTRY
    from openset_rcnn.modeling.proposal_generator.classification_free_rpn import ClsFreeRPN
    from openset_rcnn.modeling.roi_heads.osrcnn_roi_heads import OpensetROIHeads

    rpn_cls = type(names[0], (OsrProposalMixin, ClsFreeRPN), {"__doc__": "ClsFreeRPN with the osr_b200 proposal stage."})
    heads_cls = type(names[1], (OsrRoiHeadsMixin, OpensetROIHeads),
                     {"__doc__": "OpensetROIHeads with the osr_b200 sampling glue, ROIPooler and PLN."})
    PROPOSAL_GENERATOR_REGISTRY.register(rpn_cls)
    ROI_HEADS_REGISTRY.register(heads_cls)
    _REGISTERED.update({names[0]: rpn_cls, names[1]: heads_cls})
    return dict(_REGISTERED)

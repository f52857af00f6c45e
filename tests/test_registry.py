"""Registry swap-in (osr_b200/registry.py) against the UNMODIFIED reference classes, imported through the detectron2
stand-in of tests/golden/d2shim.py.  Needs /root/reference (build container only; skipped on the GPU box)."""
import os
import sys

import pytest
import torch

REF = os.environ.get("OSR_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "openset_rcnn")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import d2shim
    return d2shim.import_reference(REF)


def test_register_with_detectron2_subclasses_the_reference_classes(ref):
    import d2shim
    from osr_b200 import registry
    got = registry.register_with_detectron2()
    rpn_cls, heads_cls = got["OsrClsFreeRPN"], got["OsrOpensetROIHeads"]
    assert d2shim.PROPOSAL_GENERATOR_REGISTRY.get("OsrClsFreeRPN") is rpn_cls
    assert d2shim.ROI_HEADS_REGISTRY.get("OsrOpensetROIHeads") is heads_cls
    assert d2shim.PROPOSAL_GENERATOR_REGISTRY.get("ClsFreeRPN") is ref.classification_free_rpn.ClsFreeRPN   # the original stays
    assert issubclass(rpn_cls, ref.classification_free_rpn.ClsFreeRPN)
    assert issubclass(heads_cls, ref.osrcnn_roi_heads.OpensetROIHeads)
    # method resolution: ours first, everything else inherited from the reference
    assert rpn_cls.predict_proposals is registry.OsrProposalMixin.predict_proposals
    assert rpn_cls.losses is ref.classification_free_rpn.ClsFreeRPN.losses
    assert heads_cls.label_and_sample_proposals.__wrapped__ is registry.OsrRoiHeadsMixin.label_and_sample_proposals.__wrapped__
    assert heads_cls._forward_box is ref.osrcnn_roi_heads.OpensetROIHeads._forward_box
    # the subclass constructs like the reference class (explicit-argument form of @configurable) ...
    rpn = rpn_cls(in_features=["p2"], head=torch.nn.Identity(), anchor_generator=None, anchor_matcher=None,
                  objectness_anchor_matcher=None, box2box_transform=d2shim.Box2BoxTransformLinear(True),
                  batch_size_per_image=256, positive_fraction=0.5, objectness_positive_fraction=1.0,
                  pre_nms_topk=(2000, 1000), post_nms_topk=(2000, 1000), nms_thresh=(1.0, 1.0))
    assert rpn.pre_nms_topk[True] == 2000 and rpn.proposal_mode == "as_shipped"
    # ... and its proposal stage refuses CPU tensors instead of silently falling back
    a = [d2shim.Boxes(torch.zeros(4, 4))]
    with pytest.raises(Exception):
        rpn.predict_proposals(a, [torch.zeros(1, 4, 4)], [torch.zeros(1, 4)], [(8, 8)])
    assert registry.register_with_detectron2() == got    # idempotent

"""CPU oracle for the Openset-RCNN RoI hot path.  TEST INFRASTRUCTURE ONLY.

This package is a device-agnostic, pure-torch restatement of the reference's
post-backbone RoI path (CF-RPN proposal stage -> ROIPooler/ROIAlignV2 -> PLN
loss) and of its neighbours (``sampling``: proposal<->GT matching + labelled
sampling; ``rcnn_inference``: ROI-head inference post-processing).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
(``openset-rcnn_b200/osr_b200``) never does.

Parity pinning
--------------
The reference (``/root/reference``) is pure Python on top of detectron2 v0.6 + torchvision; detectron2 / fvcore are NOT
installable in this image and the reference ships no tests and no golden vectors.  The oracle is
nevertheless pinned to the REFERENCE'S OWN CODE: ``tests/golden/make_golden_ref.py`` imports the unmodified reference
files through a minimal detectron2 stand-in (``tests/golden/d2shim.py``, which imports nothing from this package) and
executes ``ClsFreeRPN.predict_proposals``, ``OpensetROIHeads.label_and_sample_proposals`` / ``_forward_box``, ``PLN.loss`` /
``inference``, ``OpensetFastRCNNOutputLayers.inference`` and ``SoftMaxClassifier.inference``; the outputs are committed as
``tests/golden/golden_ref_v1.npz`` and ``tests/test_golden_ref.py`` requires this oracle to reproduce them (bit-exact for
proposals, sampling, levels, pooled features and embeddings; 1e-6 / 1e-5 for the loss and its gradients).  The older
fixtures (``golden_v1.npz``, ``golden_v2.npz``) are generated the same way (no fixture comes from ``oracle/``).  The oracle

* restates the ~10 small detectron2 glue functions (each cites the reference file:line it serves and the detectron2
  function it restates), and
* calls the *real* torchvision / ATen ops for the kernels, and additionally carries loop-level restatements of those
  kernels (``roi_align_loops``, ``nms_loops``) that ``tests/test_oracle_*.py`` pin against the torchvision binary and
  against the closed-form vectors of SURVEY.md Appendix C.
"""

from . import structures, rpn, nms, roi_align, pln, bytes_model, pipeline, sampling, rcnn_inference  # noqa: F401

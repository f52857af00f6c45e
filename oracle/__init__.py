"""CPU oracle for the Openset-RCNN RoI hot path.  TEST INFRASTRUCTURE ONLY.

This package is a device-agnostic, pure-torch restatement of the reference's
post-backbone RoI path (CF-RPN proposal stage -> ROIPooler/ROIAlignV2 -> PLN
loss) and of its neighbours (``sampling``: proposal<->GT matching + labelled
sampling; ``rcnn_inference``: ROI-head inference post-processing).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
(``openset-rcnn_b200/osr_b200``) never does.

Parity pinning
--------------
The reference (``/root/reference``) is pure Python on top of detectron2 v0.6 +
torchvision; detectron2 / fvcore are NOT installable in this image, so the
reference itself cannot be imported.  What *is* here is the real arithmetic the
reference bottoms out in: torch 2.11 (``topk``, ``mm``) and the torchvision
0.26 binary (``roi_align``, ``nms``).  The oracle therefore

* restates the ~10 small detectron2 glue functions (each cites the reference
  file:line it serves and the detectron2 function it restates), and
* calls the *real* torchvision / ATen ops for the kernels, and additionally
  carries loop-level restatements of those kernels (``roi_align_loops``,
  ``nms_loops``) that ``tests/test_oracle_*.py`` pin against the torchvision
  binary and against the known-answer vectors in ``tests/golden/``.

The reference ships no tests and no golden vectors, so the
pins are: (1) outputs of the torchvision/ATen binaries run in this container,
committed under ``tests/golden/`` with the generating script, and (2) the
closed-form vectors of SURVEY.md Appendix C.
"""

from . import structures, rpn, nms, roi_align, pln, bytes_model, pipeline, sampling, rcnn_inference  # noqa: F401

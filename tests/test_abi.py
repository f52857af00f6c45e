"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/osr.h declares, and the
host-side argument checks work without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "osr.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(osr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from osr_b200 import _lib
    h = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/osr.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in osr_b200/_lib.py"
    assert _lib.missing_symbols() == []
    assert h.osr_version() == 1


def test_library_is_built_for_sm100a_only():
    import subprocess
    from osr_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_errors_are_reported_without_a_gpu():
    from osr_b200 import _lib
    h = _lib.lib()
    lv = (_lib.RpnLevel * 1)()
    lv[0].num_anchors = 0
    assert h.osr_rpn_kmax(lv, 1, 1000) == -1
    assert b"num_anchors" in h.osr_last_error()
    lv[0].num_anchors = 5000
    assert h.osr_rpn_kmax(lv, 1, 1000) == 1000
    assert h.osr_rpn_kmax(lv, 0, 1000) == -1
    assert h.osr_rpn_select_decode_workspace(lv, 1, 4, 1000) > 4 * 1000 * 24
    # NMS: unsupported segment length -> shape error, before any launch
    rc = h.osr_nms_segmented(None, None, 0, C.c_void_p(8), C.c_void_p(8), 1, 1 << 20, 0.5, 0, None, C.c_void_p(8), None,
                             C.c_void_p(256), 1 << 20, None)
    assert rc == -1 or rc == -2
    # ROIAlign: pooler resolution other than 7 is rejected
    fl = (_lib.FeatLevel * 1)()
    fl[0].data = 256; fl[0].H = 8; fl[0].W = 8; fl[0].scale = 0.25
    rc = h.osr_roi_align_fwd(fl, 1, 1, 16, C.c_void_p(256), 4, 14, 0, 1, 224, 4, 2, C.c_void_p(256), C.c_void_p(256),
                             None, 0, None)
    assert rc == -2 and b"pooler resolution" in h.osr_last_error()
    rc = h.osr_pln_loss_fwd(None, None, None, None, 4, 512, 20, 1, 0, 0.1, 0.9, 0.5, 0.5, 4.0, 1.0, None, None, None, None,
                            None, None, None, None, 0, None)
    assert rc == -2 and b"embedding dim" in h.osr_last_error()
    rc = h.osr_pln_loss_fwd(None, None, None, None, 4, 256, 20, 1, 7, 0.1, 0.9, 0.5, 0.5, 4.0, 1.0, None, None, None, None,
                            None, None, None, None, 0, None)
    assert rc == -1 and b"distance_type" in h.osr_last_error()
    assert h.osr_pln_workspace(8192, 256, 20, 1) >= 8 * 20 * 256 * 4
    assert h.osr_nms_workspace(1000, 2, 1000) > 2 * 1000 * 16 * 8


def test_product_ops_refuse_cpu_tensors():
    import torch
    from osr_b200 import _lib
    from osr_b200.nms import nms
    with pytest.raises(_lib.OsrError):
        nms(torch.zeros(3, 4), torch.zeros(3), 0.5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openset-rcnn_b200", "osr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports oracle/"

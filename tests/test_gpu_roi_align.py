"""GPU parity: osr_roi_align_fwd / _bwd (through ROIPooler) vs the oracle ROIPooler, which calls the real
torchvision roi_align.  Level ids bit-exact; features/gradients within fp32 tolerance (summation order
differs: separable regrouping vs per-sample accumulation) - rtol 1e-5, atol 1e-5 x scale (stated below).
"""
import pytest
import torch

from oracle import roi_align as ora
from oracle.structures import Boxes as OBoxes

pytestmark = pytest.mark.gpu

FWD_RTOL, FWD_ATOL = 1e-5, 2e-5   # features are N(0,1); outputs are means of <= O(10^3) products
BWD_RTOL, BWD_ATOL = 1e-4, 1e-4   # gradients accumulate up to ~hundreds of RoIs per pixel (fp32, different order)


@pytest.fixture(params=[0, 1, 2], ids=["bwd_reg_ffma2_two_buffers", "bwd_reg_ffma2_one_buffer", "bwd_smem_r1"])
def bwd_variant(request):
    """Every channels_last backward test runs on the three thread-per-channel kernels (include/osr.h OSR_TUNE_BWD_VARIANT):
    0 = register accumulators + packed fp32x2 FMAs, two staging buffers per warp (shipped), 1 = the same with one staging buffer,
    2 = shared-memory accumulators (round-1 kernel)."""
    from osr_b200 import _lib
    prev = _lib.set_tuning("bwd", request.param)
    yield request.param
    _lib.set_tuning("bwd", prev)


@pytest.fixture(params=[0, 4, 5], ids=["fwd_cta_per_roi", "fwd_persistent", "fwd_one_row_loop"])
def fwd_variant(request):
    """channels_last forward tests run on every implementation (include/osr.h OSR_TUNE_FWD_VARIANT): 0 = one CTA per RoI,
    two footprint rows per iteration (shipped), 4 = persistent CTAs that prefetch the next RoI's record and first rows,
    5 = one CTA per RoI with the one-row loop."""
    from osr_b200 import _lib
    prev = _lib.set_tuning("fwd", request.param)
    yield request.param
    _lib.set_tuning("fwd", prev)


@pytest.fixture(params=[True, False], ids=["nchw_staged", "nchw_native"])
def nchw_staging(request):
    """NCHW maps either go through the tiled NCHW->NHWC staging copy + the channels_last kernels (default) or through the
    NCHW-native kernels (``poolers.NCHW_STAGING = False``); the staging path only engages for C % 32 == 0."""
    from osr_b200 import poolers
    prev = poolers.NCHW_STAGING
    poolers.NCHW_STAGING = request.param
    yield request.param
    poolers.NCHW_STAGING = prev


def _pooler_pair():
    from osr_b200.poolers import ROIPooler
    from osr_b200 import synth
    ours = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    ref = ora.ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    return ours, ref


def _special_rois(h, w):
    """Edge cases: tiny, sub-pixel, zero-area, full image, elongated, touching borders, level boundaries."""
    r = [
        [10.0, 10.0, 10.5, 10.5], [0.0, 0.0, 3.0, 3.0], [5.0, 5.0, 5.0, 5.0], [0.0, 0.0, float(w), float(h)],
        [0.0, 100.0, float(w), 110.0], [200.0, 0.0, 210.0, float(h)], [w - 20.0, h - 20.0, float(w), float(h)],
        [0.0, 0.0, 112.0, 112.0], [0.0, 0.0, 111.99, 112.0], [50.0, 50.0, 274.0, 274.0], [50.0, 50.0, 273.9, 274.0],
        [100.0, 100.0, 548.0, 548.0], [100.0, 100.0, 547.9, 548.0], [30.0, 40.0, 37.0, 47.0], [1.0, 1.0, 29.0, 57.0],
        [w - 1.0, h - 1.0, float(w), float(h)], [0.0, h / 2.0, w / 3.0, h / 2.0 + 400.0],
    ]
    return torch.tensor(r, dtype=torch.float32)


@pytest.mark.parametrize("hw,n,per_img,C", [((800, 1333), 2, 300, 256), ((320, 480), 3, 200, 64), ((224, 224), 1, 64, 40)])
def test_forward_matches_torchvision(hw, n, per_img, C, nchw_staging):
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(n, hw, C, seed=3, device="cuda:0")
    rois = synth.make_rois(n, per_img, hw, seed=17)
    rois[0] = torch.cat([rois[0], _special_rois(*hw)])
    out, lvl = ours.forward_with_levels(feats, [OBoxes(r.cuda()) for r in rois])
    # oracle on the GPU (torchvision CUDA kernel) and on the CPU (torchvision CPU kernel)
    ref_gpu = ref.forward(feats, [OBoxes(r.cuda()) for r in rois])
    lvl_gpu = ref.level_assignments([OBoxes(r.cuda()) for r in rois])
    assert torch.equal(lvl.long(), lvl_gpu), "level assignment must be bit-exact vs torch on the same device"
    torch.testing.assert_close(out, ref_gpu, rtol=FWD_RTOL, atol=FWD_ATOL)
    ref_cpu = ref.forward([f.cpu() for f in feats], [OBoxes(r) for r in rois])
    # torchvision's CPU and CUDA kernels differ from each other by a few 1e-5 (different summation order);
    # the CPU comparison therefore uses a looser absolute tolerance
    torch.testing.assert_close(out.cpu(), ref_cpu, rtol=FWD_RTOL, atol=1e-4)
    assert (ref_gpu.cpu() - ref_cpu).abs().max() > 0  # (documenting that the two reference kernels are not bit-equal)


def test_forward_channels_last_input(fwd_variant):
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(2, (320, 480), 64, seed=5, device="cuda:0", channels_last=True)
    rois = synth.make_rois(2, 100, (320, 480), seed=18)
    out = ours.forward(feats, [OBoxes(r.cuda()) for r in rois])
    ref_gpu = ref.forward([f.contiguous() for f in feats], [OBoxes(r.cuda()) for r in rois])
    torch.testing.assert_close(out, ref_gpu, rtol=FWD_RTOL, atol=FWD_ATOL)


def test_golden_ramp_vectors():
    """x-ramp p4 map, RoI [160,160,480,480], scale 1/16 -> row = 9.5 + (k+0.5)*20/7."""
    from osr_b200.poolers import ROIPooler
    p = ROIPooler(7, (1.0 / 16,), 0, "ROIAlignV2")
    f = torch.arange(84, dtype=torch.float32).view(1, 1, 1, 84).expand(1, 3, 50, 84).contiguous().cuda()
    out = p.forward([f], [OBoxes(torch.tensor([[160.0, 160.0, 480.0, 480.0]]).cuda())])
    exp = torch.tensor([9.5 + (k + 0.5) * 20.0 / 7.0 for k in range(7)])
    torch.testing.assert_close(out[0, 0, 3].cpu(), exp, rtol=1e-6, atol=1e-5)
    assert torch.allclose(out[0, 1], out[0, 0])


def test_level_boundary_sweep_bit_exact():
    """All fp32 sizes within +-64 ulp of the level boundaries sqrt(area) in {112, 224, 448}: our level ids must
    equal torch's own op sequence on the same GPU."""
    from osr_b200.poolers import ROIPooler
    from osr_b200 import synth
    p = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    ref = ora.ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    boxes = []
    for b in (112.0, 224.0, 448.0):
        base = torch.tensor(b, dtype=torch.float32).view(torch.int32).item()
        for d in range(-64, 65):
            s = torch.tensor(base + d, dtype=torch.int32).view(torch.float32).item()
            boxes.append([0.0, 0.0, s, b])           # area = s*b -> sqrt within an ulp of the boundary
            boxes.append([3.0, 5.0, 3.0 + s, 5.0 + s])
    boxes = torch.tensor(boxes, dtype=torch.float32).cuda()
    feats = [torch.zeros(1, 8, 64 >> l, 64 >> l, device="cuda:0") for l in range(4)]
    _, lvl = p.forward_with_levels(feats, [OBoxes(boxes)])
    exp = ref.level_assignments([OBoxes(boxes)])
    assert torch.equal(lvl.long(), exp)


def test_empty_and_ragged_box_lists():
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(3, (224, 224), 16, seed=6, device="cuda:0")
    rois = [torch.empty(0, 4), synth.make_rois(1, 5, (224, 224), seed=1)[0], torch.empty(0, 4)]
    out = ours.forward(feats, [OBoxes(r.cuda()) for r in rois])
    exp = ref.forward(feats, [OBoxes(r.cuda()) for r in rois])
    assert out.shape == (5, 16, 7, 7)
    torch.testing.assert_close(out, exp, rtol=FWD_RTOL, atol=FWD_ATOL)
    out0 = ours.forward(feats, [OBoxes(torch.empty(0, 4).cuda()) for _ in range(3)])
    assert out0.shape == (0, 16, 7, 7)


def _grads(pooler, feats, box_lists, gout):
    feats = [f.detach().clone().requires_grad_(True) for f in feats]
    out = pooler.forward(feats, box_lists)
    out.backward(gout)
    return [f.grad for f in feats]


@pytest.mark.parametrize("hw,n,per_img,C", [((800, 1333), 2, 256, 64), ((320, 480), 3, 200, 48), ((224, 224), 1, 64, 20)])
def test_backward_matches_torchvision_autograd(hw, n, per_img, C, nchw_staging):
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(n, hw, C, seed=4, device="cuda:0")
    rois = synth.make_rois(n, per_img, hw, seed=19)
    rois[0] = torch.cat([rois[0], _special_rois(*hw)])
    boxes = [OBoxes(r.cuda()) for r in rois]
    M = sum(len(r) for r in rois)
    gout = torch.randn(M, C, 7, 7, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(1))
    g_ours = _grads(ours, feats, boxes, gout)
    g_ref = _grads(ref, feats, boxes, gout)
    for a, b in zip(g_ours, g_ref):
        assert a.shape == b.shape
        # scale-aware: |grad| grows with the number of overlapping RoIs; atol relative to the largest entry
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)


def test_backward_is_deterministic_and_dense():
    from osr_b200 import synth
    ours, _ = _pooler_pair()
    feats = synth.make_features(2, (320, 480), 32, seed=4, device="cuda:0")
    rois = synth.make_rois(2, 300, (320, 480), seed=20)
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(600, 32, 7, 7, device="cuda:0")
    g1 = _grads(ours, feats, boxes, gout)
    g2 = _grads(ours, feats, boxes, gout)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b), "backward must be run-to-run bit-identical"
        assert torch.isfinite(a).all()


def test_backward_is_adjoint_of_forward():
    """<pool(F), G> == <F, pool^T(G)> (linearity / adjoint property; size-independent)."""
    from osr_b200 import synth
    ours, _ = _pooler_pair()
    feats = synth.make_features(2, (800, 1333), 16, seed=8, device="cuda:0")
    rois = synth.make_rois(2, 512, (800, 1333), seed=21)
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(1024, 16, 7, 7, device="cuda:0")
    out = ours.forward(feats, boxes)
    grads = _grads(ours, feats, boxes, gout)
    lhs = (out.double() * gout.double()).sum()
    rhs = sum((f.double() * g.double()).sum() for f, g in zip(feats, grads))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-2, (float(lhs), float(rhs))


def test_backward_dense_tile_more_than_64_rois():
    """> kNB RoIs on one tile exercises the multi-batch path."""
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(1, (224, 224), 8, seed=9, device="cuda:0")
    g = torch.Generator().manual_seed(5)
    c = torch.rand(200, 2, generator=g) * 20 + 60
    wh = torch.rand(200, 2, generator=g) * 30 + 10
    rois = [torch.cat([c - wh / 2, c + wh / 2], dim=1)]
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(200, 8, 7, 7, device="cuda:0")
    for a, b in zip(_grads(ours, feats, boxes, gout), _grads(ref, feats, boxes, gout)):
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)


@pytest.mark.parametrize("channels_last", [True, False])
def test_backward_more_rois_per_image_than_one_scan_window(channels_last):
    """1300 RoIs per image (cfg 5 has 1024): the tile kernels collect a batch over several 512-index scan windows, and
    dense tiles continue the partial sums of their earlier batches."""
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(2, (320, 480), 32, seed=19, device="cuda:0", channels_last=channels_last)
    rois = synth.make_rois(2, 1300, (320, 480), seed=23)
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(2600, 32, 7, 7, device="cuda:0")
    a1 = _grads(ours, feats, boxes, gout)
    a2 = _grads(ours, feats, boxes, gout)
    for a, a_again, b in zip(a1, a2, _grads(ref, feats, boxes, gout)):
        assert torch.equal(a, a_again)   # deterministic
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)


def test_backward_empty_rois_gives_zero_grads():
    from osr_b200 import synth
    ours, _ = _pooler_pair()
    feats = synth.make_features(2, (224, 224), 8, seed=9, device="cuda:0")
    boxes = [OBoxes(torch.empty(0, 4).cuda()), OBoxes(torch.tensor([[10.0, 10.0, 50.0, 60.0]]).cuda())]
    gout = torch.randn(1, 8, 7, 7, device="cuda:0")
    grads = _grads(ours, feats, boxes, gout)
    for gl in grads:
        assert float(gl[0].abs().max()) == 0.0
    assert sum(float(gl[1].abs().sum()) for gl in grads) > 0


@pytest.mark.parametrize("hw,n,per_img,C", [((800, 1333), 2, 300, 256), ((320, 480), 3, 200, 64), ((224, 224), 1, 64, 40)])
def test_forward_channels_last_kernel_matches_torchvision(hw, n, per_img, C, fwd_variant):
    """channels_last maps take the dedicated NHWC kernel (bulk-copy ring, thread = channel)."""
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(n, hw, C, seed=3, device="cuda:0", channels_last=True)
    rois = synth.make_rois(n, per_img, hw, seed=17)
    rois[0] = torch.cat([rois[0], _special_rois(*hw)])
    boxes = [OBoxes(r.cuda()) for r in rois]
    out, lvl = ours.forward_with_levels(feats, boxes)
    ref_gpu = ref.forward([f.contiguous() for f in feats], boxes)
    assert torch.equal(lvl.long(), ref.level_assignments(boxes))
    torch.testing.assert_close(out, ref_gpu, rtol=FWD_RTOL, atol=FWD_ATOL)
    # and it agrees with our own NCHW kernel to fp32 summation-order noise
    out_nchw = ours.forward([f.contiguous() for f in feats], boxes)
    torch.testing.assert_close(out, out_nchw, rtol=FWD_RTOL, atol=FWD_ATOL)


def test_forward_persistent_kernel_is_bit_identical_to_cta_per_roi():
    """Both channels_last forwards execute the same arithmetic per RoI: outputs must be bit-equal, including RoIs that
    leave the fast path (oversized / degenerate footprints), images without RoIs and the bf16 output mode."""
    from osr_b200 import _lib, synth
    from osr_b200.poolers import ROIPooler
    ours = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    feats = synth.make_features(3, (800, 1333), 256, seed=31, device="cuda:0", channels_last=True)
    rois = synth.make_rois(3, 700, (800, 1333), seed=41)
    rois[1] = torch.empty(0, 4)
    rois[2] = torch.cat([rois[2], _special_rois(800, 1333)])
    boxes = [OBoxes(r.cuda()) for r in rois]
    packed = torch.cat([torch.cat([torch.full((len(r), 1), float(i)), r], dim=1) for i, r in enumerate(rois)]).cuda()
    offsets = torch.tensor([0, 700, 700, 700 + len(rois[2])], dtype=torch.int32, device="cuda:0")
    outs = {}
    for v in (0, 4, 5):
        prev = _lib.set_tuning("fwd", v)
        try:
            outs[v] = (ours.forward(feats, boxes), ours.pool_rois_bf16(feats, packed, offsets)[0])
        finally:
            _lib.set_tuning("fwd", prev)
    assert torch.equal(outs[5][0], outs[4][0]) and torch.equal(outs[5][1], outs[4][1])   # same one-row arithmetic
    # the two-row loop adds the rows of a pair in the same order, with the same roundings: bit-equal as well
    assert torch.equal(outs[0][0], outs[5][0]) and torch.equal(outs[0][1], outs[5][1])
    torch.testing.assert_close(outs[4][1].float(), outs[4][0], rtol=8e-3, atol=1e-6)   # bf16 = fp32 result rounded once


def test_backward_channels_last_grads(bwd_variant):
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(2, (320, 480), 32, seed=4, device="cuda:0", channels_last=True)
    rois = synth.make_rois(2, 150, (320, 480), seed=22)
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(300, 32, 7, 7, device="cuda:0")
    g_ours = _grads(ours, feats, boxes, gout)
    g_ref = _grads(ref, [f.contiguous() for f in feats], boxes, gout)
    for a, b in zip(g_ours, g_ref):
        assert a.is_contiguous(memory_format=torch.channels_last)
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)


@pytest.mark.parametrize("hw,n,per_img,C", [((800, 1333), 2, 256, 256), ((320, 480), 3, 200, 96), ((224, 224), 1, 64, 32),
                                             ((224, 224), 1, 64, 40)])
def test_backward_channels_last_kernel_matches_torchvision(hw, n, per_img, C, bwd_variant):
    """channels_last gradient maps with C % 32 == 0 take the thread-per-channel gather kernel (96 = a partly filled
    128-channel slab; 40 falls back to the pixel-per-thread kernel)."""
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(n, hw, C, seed=4, device="cuda:0", channels_last=True)
    rois = synth.make_rois(n, per_img, hw, seed=19)
    rois[0] = torch.cat([rois[0], _special_rois(*hw)])
    boxes = [OBoxes(r.cuda()) for r in rois]
    M = sum(len(r) for r in rois)
    gout = torch.randn(M, C, 7, 7, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(1))
    g_ours = _grads(ours, feats, boxes, gout)
    g_ref = _grads(ref, [f.contiguous() for f in feats], boxes, gout)
    g_nchw = _grads(ours, [f.contiguous() for f in feats], boxes, gout)
    for a, b, c in zip(g_ours, g_ref, g_nchw):
        assert a.is_contiguous(memory_format=torch.channels_last)
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)
        torch.testing.assert_close(a, c, rtol=BWD_RTOL, atol=BWD_ATOL * scale)


def test_backward_channels_last_dense_tile_tiny_rois_deterministic_adjoint(bwd_variant):
    """> kCNB RoIs on one tile (multi-batch path), sub-pixel bins (dense 7-bin fold), run-to-run bit-identical,
    and <pool(F), G> == <F, pool^T(G)>."""
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(1, (224, 224), 64, seed=9, device="cuda:0", channels_last=True)
    g = torch.Generator().manual_seed(5)
    c = torch.rand(240, 2, generator=g) * 20 + 60
    wh = torch.rand(240, 2, generator=g) * 30 + 10
    wh[200:] = torch.rand(40, 2, generator=g) * 6 + 0.5      # tiny RoIs: bins narrower than half a pixel
    rois = [torch.cat([c - wh / 2, c + wh / 2], dim=1)]
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(240, 64, 7, 7, device="cuda:0")
    g1 = _grads(ours, feats, boxes, gout)
    g2 = _grads(ours, feats, boxes, gout)
    g_ref = _grads(ref, [f.contiguous() for f in feats], boxes, gout)
    for a, a2, b in zip(g1, g2, g_ref):
        assert torch.equal(a, a2), "backward must be run-to-run bit-identical"
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)
    out = ours.forward(feats, boxes)
    lhs = (out.double() * gout.double()).sum()
    rhs = sum((f.double() * gg.double()).sum() for f, gg in zip(feats, g1))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-2, (float(lhs), float(rhs))


def test_channels_last_width_sweep_stage_geometry(bwd_variant):
    """Footprints 25..56 feature pixels wide at level 2: whole-row stages (<= 48 px, 2-3 stages) and the 32-column
    chunked path (> 48 px) of the channels_last forward, plus the matching backward tiles."""
    from osr_b200 import synth
    ours, ref = _pooler_pair()
    feats = synth.make_features(1, (256, 320), 64, seed=12, device="cuda:0", channels_last=True)
    widths = torch.arange(100.0, 226.0, 7.0)
    x1 = torch.linspace(3.3, 60.0, len(widths))
    boxes_t = torch.stack((x1, torch.full_like(x1, 20.5), x1 + widths, torch.full_like(x1, 20.5) + 40.0), dim=1)
    boxes = [OBoxes(boxes_t.cuda())]
    out, lvl = ours.forward_with_levels(feats, boxes)
    assert int(lvl.max()) == 0   # all on p2
    exp = ref.forward([f.contiguous() for f in feats], boxes)
    torch.testing.assert_close(out, exp, rtol=FWD_RTOL, atol=FWD_ATOL)
    gout = torch.randn(len(widths), 64, 7, 7, device="cuda:0")
    for a, b in zip(_grads(ours, feats, boxes, gout), _grads(ref, [f.contiguous() for f in feats], boxes, gout)):
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)


def test_nchw_staging_equals_channels_last_path_exactly():
    """NCHW maps staged through osr_nchw_to_nhwc / osr_nhwc_to_nchw run the SAME kernels on the same values as
    channels_last maps: outputs and gradients are bit-identical, gradients come back NCHW-contiguous."""
    from osr_b200 import poolers, synth
    ours, _ = _pooler_pair()
    assert poolers.NCHW_STAGING
    feats = synth.make_features(2, (320, 480), 64, seed=21, device="cuda:0")            # NCHW
    feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    for f in feats:
        assert torch.equal(poolers.nchw_to_channels_last(f), f) and poolers.nchw_to_channels_last(f).is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(poolers.channels_last_to_nchw(f.contiguous(memory_format=torch.channels_last)), f)
    rois = synth.make_rois(2, 120, (320, 480), seed=23)
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(240, 64, 7, 7, device="cuda:0")
    a = ours.forward([f.clone().requires_grad_(True) for f in feats], boxes)
    b = ours.forward([f.clone().requires_grad_(True) for f in feats_cl], boxes)
    assert torch.equal(a, b)
    ga = _grads(ours, feats, boxes, gout)
    gb = _grads(ours, feats_cl, boxes, gout)
    for x, y in zip(ga, gb):
        assert x.is_contiguous() and torch.equal(x, y)


@pytest.mark.parametrize("channels_last", [True, False])
def test_backward_prepare_then_prepared_equals_one_call(channels_last):
    """osr_roi_align_bwd_prepare (RoI-only tables, here on a side stream) + osr_roi_align_bwd_prepared == osr_roi_align_bwd,
    bit for bit."""
    from osr_b200 import synth
    ours, _ = _pooler_pair()
    feats = synth.make_features(2, (320, 480), 64, seed=8, device="cuda:0", channels_last=channels_last)
    rois = synth.make_rois(2, 300, (320, 480), seed=29)
    packed = torch.cat([torch.cat([torch.full((len(r), 1), float(i)), r], dim=1) for i, r in enumerate(rois)]).cuda()
    offsets = torch.tensor([0, 300, 600], dtype=torch.int32, device="cuda:0")
    gout = torch.randn(600, 64, 7, 7, device="cuda:0")
    ref = ours.backward_rois(gout, feats, packed, offsets)
    side = torch.cuda.Stream()
    ws = ours.alloc_backward_workspace(feats, packed)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ours.prepare_backward(feats, packed, offsets, out=ws)
    torch.cuda.current_stream().wait_stream(side)
    got = ours.backward_rois(gout, feats, packed, offsets, prepared=ws)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


@pytest.mark.parametrize("C", [256, 128])
def test_backward_dense_tiles_shared_by_two_ctas_equal_one_cta_per_tile(C):
    """RoIs clustered on a few boxes (what the labelled sampler produces): the coarse levels' dense tiles are worked by
    two CTAs (dynamic split, shipped); every other setting of OSR_TUNE_BWD_SPLIT - one CTA per tile, static splits, a
    threshold of one RoI - must give the same gradient bit for bit (each (pixel, channel) is summed by exactly one warp in
    RoI order whichever CTA owns it), and it must match torchvision."""
    from osr_b200 import _lib, synth
    ours, ref = _pooler_pair()
    hw = (512, 640)
    feats = synth.make_features(2, hw, C, seed=29, device="cuda:0", channels_last=True)
    g = torch.Generator().manual_seed(11)
    rois = []
    for n in range(2):
        ctr = torch.tensor([[200.0, 180.0], [420.0, 330.0]])[torch.randint(0, 2, (300,), generator=g)]
        size = torch.tensor([[360.0, 300.0], [90.0, 120.0]])[torch.randint(0, 2, (300,), generator=g)]
        c = ctr + torch.randn(300, 2, generator=g) * 12
        wh = size * (1 + 0.2 * torch.rand(300, 2, generator=g))
        b = torch.cat([c - wh / 2, c + wh / 2], dim=1)
        b[:, 0::2].clamp_(0, hw[1]); b[:, 1::2].clamp_(0, hw[0])
        rois.append(b)
    boxes = [OBoxes(r.cuda()) for r in rois]
    gout = torch.randn(600, C, 7, 7, device="cuda:0")
    base = _grads(ours, feats, boxes, gout)            # shipped: dynamic two-CTA split on the two coarsest levels
    for code in (-1, 0x2211, 0x0102211, 0x1E03320):
        prev = _lib.set_tuning("bwd_split", code)
        try:
            other = _grads(ours, feats, boxes, gout)
        finally:
            _lib.set_tuning("bwd_split", prev)
        for a, b in zip(base, other):
            assert torch.equal(a, b), hex(code & 0xffffffff)
    for a, b in zip(base, _grads(ref, feats, boxes, gout)):
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=BWD_RTOL, atol=BWD_ATOL * scale)

"""Box-head FC on tcgen05: ``osr_linear_bf16_fwd`` / ``osr_roi_align_fwd_bf16`` /
``box_head.FastRCNNConvFCHead`` against ``nn.Linear`` arithmetic.

Tolerances (bf16 tier, stated): the kernel's operands ARE bf16, so against a reference computed from the SAME
bf16-rounded operands in fp32 the only differences are the fp32 summation order and the output rounding: rtol 2e-3 for
fp32 outputs of K = 12544 sums, 1 bf16 ulp (2^-8 relative) for bf16 outputs.  Against the un-rounded fp32 ``nn.Linear``
the bf16 operand rounding shows: rtol 2e-2 / atol 2e-2 x output scale."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,K,N,relu,out_dtype", [(300, 12544, 1024, True, torch.bfloat16), (128, 1024, 1024, True, torch.float32),
                                                  (1, 64, 256, False, torch.float32), (257, 192, 512, False, torch.bfloat16)])
def test_linear_bf16_matches_torch(R, K, N, relu, out_dtype):
    from osr_b200.box_head import linear_bf16
    g = torch.Generator("cuda").manual_seed(R + K)
    a = torch.randn(R, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    out = linear_bf16(a, w, b, relu, out_dtype)
    ref = a.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp(min=0)
    assert out.dtype == out_dtype and out.shape == (R, N)
    if out_dtype == torch.float32:
        torch.testing.assert_close(out.double(), ref, rtol=2e-3, atol=1e-4)
    else:
        torch.testing.assert_close(out.double(), ref, rtol=2 ** -7, atol=1e-3)
    out2 = linear_bf16(a, w, None, relu, out_dtype)          # no bias
    ref2 = a.double() @ w.double().t()
    torch.testing.assert_close(out2.double(), ref2.clamp(min=0) if relu else ref2, rtol=2 ** -7, atol=1e-3)


def test_roi_align_bf16_output_is_the_rounded_fp32_output():
    from osr_b200 import synth
    from osr_b200.poolers import ROIPooler
    feats = synth.make_features(2, (320, 480), 64, seed=5, device="cuda:0", channels_last=True)
    rois = synth.make_rois(2, 150, (320, 480), seed=8)
    rois[0][0] = torch.tensor([5.0, 5.0, 5.0, 5.0])     # empty RoI: zero tile
    packed = torch.cat([torch.cat((torch.full((len(r), 1), float(n)), r), 1) for n, r in enumerate(rois)]).cuda()
    off = torch.tensor([0, 150, 300], dtype=torch.int32, device="cuda")
    p = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    ref, lvl = p.pool_rois(feats, packed, off)
    out, lvl2 = p.pool_rois_bf16(feats, packed, off)
    assert out.dtype == torch.bfloat16 and torch.equal(lvl, lvl2)
    assert torch.equal(out, ref.to(torch.bfloat16)), "bf16 output must be the round-to-nearest-even of the fp32 output"
    with pytest.raises(RuntimeError):
        p.pool_rois_bf16([f.contiguous() for f in feats], packed, off)   # NCHW maps: not supported for bf16 output


def test_box_head_matches_nn_linear_and_backpropagates():
    from osr_b200.box_head import FastRCNNConvFCHead
    torch.manual_seed(3)
    head = FastRCNNConvFCHead((64, 7, 7), fc_dims=[1024, 1024])
    sd = head.state_dict()
    assert set(sd) == {"fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"} and sd["fc1.weight"].shape == (1024, 64 * 49)
    x = torch.randn(200, 64, 7, 7, device="cuda")
    y = head(x.to(torch.bfloat16))
    ref = F.relu(F.linear(F.relu(F.linear(x.flatten(1), head.fc1.weight, head.fc1.bias)), head.fc2.weight, head.fc2.bias))
    assert y.dtype == torch.float32
    torch.testing.assert_close(y.detach(), ref.detach(), rtol=2e-2, atol=2e-2 * float(ref.detach().abs().max()))
    # gradients flow to the weights (library GEMMs) and agree with the fp32 head to bf16 accuracy
    gw = torch.autograd.grad(y.sum(), [head.fc2.weight, head.fc1.weight])
    gw_ref = torch.autograd.grad(ref.sum(), [head.fc2.weight, head.fc1.weight])
    for a, b in zip(gw, gw_ref):
        assert torch.isfinite(a).all()
        assert float((a - b).norm() / b.norm()) < 8e-2   # bf16 gradient operands (g and x rounded to 8 bits of mantissa)
    # the bf16 weight copy follows in-place updates of the parameter
    with torch.no_grad():
        head.fc2.weight.mul_(0.5)
    y2 = head(x.to(torch.bfloat16))
    torch.testing.assert_close(y2, 0.5 * y, rtol=2e-2, atol=1e-3)

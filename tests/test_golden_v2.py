"""Committed fixtures for the section 8(f) rows (tests/golden/golden_v2.npz, produced by tests/golden/make_golden_v2.py
with torch's CPU ops): the oracle must reproduce them on the CPU, the CUDA kernels must match them on the GPU
(matching: bit-exact; box decode: 1e-3 px / 1e-6 - expf and the reciprocal multiply differ by ulps between the CPU
and CUDA math paths)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import rcnn_inference as oinf, sampling as osamp
from oracle.structures import Boxes, pairwise_iou

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"))


def t(name):
    return torch.from_numpy(G[name])


def test_oracle_matching_reproduces_golden():
    for n in range(2):
        m = pairwise_iou(Boxes(t(f"match_gt{n}")), Boxes(t(f"match_props{n}")))
        idx, lab = osamp.matcher(m, 0.5)
        assert torch.equal(idx, t(f"match_idx{n}")) and torch.equal(lab, t(f"match_lab{n}"))
        assert torch.equal(m[idx, torch.arange(m.shape[1])], t(f"match_iou{n}"))
        assert float(t(f"match_iou{n}")[-1]) == 1.0   # the appended exact GT duplicate


def test_oracle_decode_reproduces_golden():
    dec = oinf.apply_deltas(t("dec_deltas"), t("dec_boxes_in"))
    b = Boxes(dec); b.clip((800, 1333))
    assert torch.equal(b.tensor, t("dec_boxes_clipped"))
    assert torch.equal(torch.sqrt(t("dec_ious")[:, 0] * t("dec_ctr")), t("dec_scores"))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port timed on the host cores) runs without a GPU and prints one JSON
    line with the keys the driver reads."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "roi_path_images_per_sec" and line["value"] > 0
    assert line["higher_is_better"] is True and line["unit"] == "images/s" and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


@pytest.mark.gpu
def test_gpu_matching_matches_golden():
    from osr_b200.sampling import match_proposals
    dev = "cuda:0"
    props = [t(f"match_props{n}") for n in range(2)]
    gts = [t(f"match_gt{n}") for n in range(2)]
    off = torch.tensor([0, props[0].shape[0], props[0].shape[0] + props[1].shape[0]], dtype=torch.int32, device=dev)
    goff = torch.tensor([0, gts[0].shape[0], gts[0].shape[0] + gts[1].shape[0]], dtype=torch.int32, device=dev)
    midx, miou, mlab, mcls = match_proposals(torch.cat(props).to(dev), off, torch.cat(gts).to(dev),
                                             torch.cat([t("match_cls0"), t("match_cls1")]).to(dev), goff,
                                             max(p.shape[0] for p in props), background_label=80)
    b0 = 0
    for n in range(2):
        sl = slice(b0, b0 + props[n].shape[0])
        assert torch.equal(midx[sl].cpu().long(), t(f"match_idx{n}"))
        assert torch.equal(miou[sl].cpu(), t(f"match_iou{n}"))
        assert torch.equal(mlab[sl].cpu().to(torch.int8), t(f"match_lab{n}"))
        b0 += props[n].shape[0]


@pytest.mark.gpu
def test_gpu_decode_matches_golden():
    from osr_b200.inference import inference
    from osr_b200.structures import Boxes as PBoxes, Instances
    dev = "cuda:0"
    p = Instances((800, 1333))
    p.set("proposal_boxes", PBoxes(t("dec_boxes_in").to(dev)))
    p.set("objectness_logits", t("dec_ctr").to(dev))
    R = t("dec_ctr").shape[0]
    feats = torch.arange(R, dtype=torch.float32, device=dev)[:, None]   # carries the original row index through
    res, _ = inference((t("dec_deltas").to(dev), t("dec_ious").to(dev)), [p], feats, score_thresh=-1.0, nms_thresh=1.0,
                       topk_per_image=-1)
    rows = res[0].get("features")[:, 0].long().cpu()
    assert sorted(rows.tolist()) == list(range(R))       # nothing dropped (threshold -1, NMS 1.0)
    torch.testing.assert_close(res[0].get("pred_boxes").tensor.cpu(), t("dec_boxes_clipped")[rows], rtol=1e-6, atol=1e-3)
    torch.testing.assert_close(res[0].get("scores").cpu(), t("dec_scores")[rows], rtol=1e-6, atol=1e-7)

"""The oracle against (a) fixtures produced by the reference's own code (oracle/gen_golden.py) and
(b) the installed third-party code the reference relies on (torchvision nms, HF ViTPose processor)."""
import numpy as np
import pytest
import torch

from oracle import crop as ocrop
from oracle import det as odet
from oracle import match as omatch
from oracle import pose as opose


@pytest.mark.parametrize("tag", ["nc1", "nc3"])
def test_head_decode_and_nms_match_reference(golden, tag):
    g = golden(f"det_{tag}.npz")
    levels = [torch.from_numpy(g[k]) for k in ("l0", "l1", "l2")]
    dec = odet.head_decode(levels)
    ref = torch.from_numpy(g["decoded"])
    assert dec.shape == ref.shape
    np.testing.assert_array_equal(dec.numpy(), ref.numpy())        # same torch calls -> bit-exact
    a, s = odet.make_anchors([tuple(l.shape[2:]) for l in levels], (8, 16, 32))
    np.testing.assert_array_equal(a.t().numpy(), g["anchors"])
    np.testing.assert_array_equal(s.t().numpy(), g["strides"])
    # NMS on the reference's own decoded tensor: bit-exact rows
    for fn in (odet.nms_greedy, odet.nms_greedy_np):
        dets = odet.non_max_suppression(ref, float(g["conf"]), float(g["iou"]), nms_fn=fn)
        assert [d.shape[0] for d in dets] == list(g["n"])
        np.testing.assert_array_equal(torch.cat(dets).numpy(), g["dets"])


def test_reference_head_module_fixture(golden):
    """The fixture of the reference's REAL ``Head`` (conv stacks + BatchNorm + DFL, nn.py:228-270): the layout restatement
    loads its state_dict strictly and reproduces its conv outputs; the oracle's decode and NMS reproduce its eval output
    and the reference's NMS rows bit for bit."""
    from oracle.refhead import RefShapedHead
    g = golden("ref_head.npz")
    head = RefShapedHead(int(g["nc"]), tuple(int(x) for x in g["filters"]))
    head.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    head.eval()
    feats = [torch.from_numpy(g[k]) for k in ("f0", "f1", "f2")]
    with torch.no_grad():
        cat = [torch.cat((b(x), c(x)), 1) for b, c, x in zip(head.box, head.cls, feats)]
    for c, k in zip(cat, ("l0", "l1", "l2")):
        np.testing.assert_array_equal(c.numpy(), g[k])
    dec = odet.head_decode(cat)
    np.testing.assert_array_equal(dec.numpy(), g["decoded"])
    dets = odet.non_max_suppression(dec, float(g["conf"]), float(g["iou"]))
    assert [d.shape[0] for d in dets] == g["n"].tolist() and min(g["n"]) >= 20
    np.testing.assert_array_equal(torch.cat(dets).numpy(), g["dets"])


def test_det_metrics_oracle_matches_reference(golden):
    """compute_metric / compute_ap restatement (oracle/detmetrics.py) against the reference's own outputs."""
    from oracle import detmetrics as dm
    g = golden("det_metrics.npz")
    for b in range(g["dets"].shape[0]):
        n, m = int(g["dcount"][b]), int(g["tcount"][b])
        c = dm.compute_metric(g["dets"][b, :n], g["targets"][b, :m], g["iou_v"])
        np.testing.assert_array_equal(c, g["correct"][b, :n])
    assert g["correct"].sum() > 100
    tp, fp, m_pre, m_rec, map50, mean_ap, _ = dm.compute_ap(g["cat_tp"], g["cat_conf"], g["cat_cls"], g["cat_target_cls"])
    np.testing.assert_array_equal(tp, g["tp"])
    np.testing.assert_array_equal(fp, g["fp"])
    np.testing.assert_array_equal(np.array([m_pre, m_rec, map50, mean_ap]), g["summary"])


def test_nms_restatement_matches_torchvision():
    import torchvision
    gen = torch.Generator().manual_seed(0)
    for trial in range(20):
        n = 200
        xy = torch.rand(n, 2, generator=gen) * 100
        wh = torch.rand(n, 2, generator=gen) * 40 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=gen)
        if trial % 4 == 0:
            scores = (scores * 8).round() / 8          # many exact ties -> stable order matters
        keep_tv = torchvision.ops.nms(boxes, scores, 0.5).numpy()
        for fn in (odet.nms_greedy, odet.nms_greedy_np):
            np.testing.assert_array_equal(fn(boxes.numpy(), scores.numpy(), 0.5), keep_tv)


def test_match_oracle(golden):
    g = golden("match.npz")
    emb, norm = omatch.backbone_tail(torch.from_numpy(g["pre"]))
    np.testing.assert_array_equal(emb.numpy(), g["emb"])
    np.testing.assert_array_equal(norm.numpy(), g["norm"])
    kernel = torch.from_numpy(g["kernel"])
    np.testing.assert_array_equal(omatch.l2_norm(kernel, 0).numpy(), g["kernel_l2"])
    probes = torch.from_numpy(g["probes"])
    pred, sim = omatch.match_top1(probes, omatch.enrol_gallery(kernel, quirk_q3=True))
    np.testing.assert_array_equal(pred.numpy(), g["pred_q3"])       # confirmed by the real validation_step
    np.testing.assert_allclose(sim.numpy(), g["sim_q3"], rtol=1e-6, atol=1e-7)
    pred, sim = omatch.match_top1(probes, omatch.enrol_gallery(kernel))
    np.testing.assert_array_equal(pred.numpy(), g["pred"])
    np.testing.assert_allclose(sim.numpy(), g["sim"], rtol=1e-6, atol=1e-7)
    gated, _ = omatch.match_top1(probes, omatch.enrol_gallery(kernel), threshold=0.4)
    known = g["true_ids"] >= 0
    np.testing.assert_array_equal(gated.numpy()[known], g["true_ids"][known])
    assert (gated.numpy()[~known] == -1).all()


def test_pose_live_oracle(golden):
    g = golden("pose_live.npz")
    hm, fl, perm = (torch.from_numpy(g[k]) for k in ("hm", "flipped", "perm"))
    avg = opose.flip_average(hm, fl, perm)
    np.testing.assert_array_equal(avg.numpy(), g["avg"])
    from importlib import import_module
    pairs = import_module("person-recognition-for-pose-estimation_b200.synth").COCO_FLIP_PAIRS
    np.testing.assert_array_equal(((hm + opose.flip_back_quirk_q1(fl, pairs)) * 0.5).numpy(), g["avg_q1"])
    c, s = opose.soft_argmax_decode(hm)
    np.testing.assert_array_equal(c.numpy(), g["coords_plain"])
    np.testing.assert_array_equal(s.numpy(), g["scores_plain"])
    c, s = opose.soft_argmax_decode(avg, torch.from_numpy(g["boxes_xyxy"]))
    np.testing.assert_array_equal(c.numpy(), g["coords_avg_box"])
    np.testing.assert_array_equal(s.numpy(), g["scores_avg_box"])


def test_pose_hf_oracle_fixture(golden):
    g = golden("pose_hf.npz")
    hm, fl, perm = (torch.from_numpy(g[k]) for k in ("hm", "flipped", "perm"))
    avg = opose.flip_average(hm, fl, perm).numpy()
    boxes = [[float(v) for v in b] for b in g["boxes"]]
    kp, sc, idx = opose.hf_dark_decode(avg, boxes)
    np.testing.assert_allclose(kp, g["keypoints"], rtol=1e-6, atol=1e-4)
    np.testing.assert_array_equal(sc, g["scores"])
    kp2, sc2, idx2 = opose.dark_decode_local(avg, boxes)
    np.testing.assert_array_equal(idx2, idx)
    np.testing.assert_allclose(kp2, kp, rtol=1e-5, atol=2e-3)
    pix = ocrop.crop_affine_hf(g["frame"][None], boxes, [0] * len(boxes))
    np.testing.assert_allclose(pix[:, :, ::4, ::4], g["crop_sub"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(pix.astype(np.float64).sum(axis=(2, 3)), g["crop_sum"], rtol=1e-6)


def test_oracle_against_installed_hf(synth):
    """Live comparison with the third-party code itself (same package on the GPU box)."""
    from transformers import VitPoseImageProcessor
    from transformers.models.vitpose.modeling_vitpose import VitPoseEstimatorOutput
    proc = VitPoseImageProcessor()
    hs = synth.make_heatmaps(3, 17, seed=31, negative_frac=0.1)
    avg = opose.flip_average(hs.heatmaps, hs.flipped, hs.perm)
    cs = synth.make_crop_set(1, 200, 260, per_frame=3, seed=32)
    boxes = [[float(v) for v in b] for b in cs.boxes]
    res = proc.post_process_pose_estimation(VitPoseEstimatorOutput(heatmaps=avg), boxes=[boxes])
    kp_hf = torch.stack([r["keypoints"] for r in res[0]]).numpy()
    sc_hf = torch.stack([r["scores"] for r in res[0]]).numpy()
    kp, sc, _ = opose.hf_dark_decode(avg.numpy(), boxes)
    np.testing.assert_allclose(kp, kp_hf, rtol=1e-6, atol=1e-4)
    np.testing.assert_array_equal(sc, sc_hf)
    pix_hf = proc.preprocess([cs.frames[0]], boxes=[boxes], do_rescale=False, return_tensors="pt")["pixel_values"]
    pix = ocrop.crop_affine_hf(cs.frames.numpy(), boxes, [0, 0, 0])
    np.testing.assert_allclose(pix, pix_hf.numpy(), rtol=1e-5, atol=2e-6)
    # default HF path (rescale 1/255 folded into mean/std)
    fr255 = (cs.frames * 255.0)
    pix_hf = proc.preprocess([fr255[0]], boxes=[boxes], return_tensors="pt")["pixel_values"]
    pix = ocrop.crop_affine_hf(fr255.numpy(), boxes, [0, 0, 0], rescale_factor=1 / 255)
    np.testing.assert_allclose(pix, pix_hf.numpy(), rtol=1e-5, atol=2e-5)


def test_oracle_crop_uint8_against_installed_hf(synth):
    from transformers import VitPoseImageProcessor
    proc = VitPoseImageProcessor()
    cs = synth.make_crop_set(1, 200, 256, per_frame=3, seed=33)
    fr8 = (cs.frames * 255.0).round().clamp(0, 255).to(torch.uint8)
    boxes = [[float(v) for v in b] for b in cs.boxes]
    pix_hf = proc.preprocess([fr8[0]], boxes=[boxes], return_tensors="pt")["pixel_values"]
    pix = ocrop.crop_affine_hf(fr8.numpy(), boxes, [0, 0, 0], rescale_factor=1 / 255)
    np.testing.assert_allclose(pix, pix_hf.numpy(), rtol=1e-5, atol=2e-5)


def test_quarter_offset_hand_made():
    hm = np.zeros((1, 2, 8, 6), np.float32)
    hm[0, 0, 3, 2] = 1.0; hm[0, 0, 3, 3] = 0.5; hm[0, 0, 2, 2] = 0.25     # peak (2,3): right>left, up>down
    hm[0, 1] = -1.0                                                        # all negative -> zeroed coords
    centers = np.array([[100.0, 200.0]]); scales = np.array([[60.0, 80.0]])
    preds, maxv, idx = opose.quarter_offset_decode(hm, centers, scales)
    assert idx[0, 0] == 3 * 6 + 2 and maxv[0, 0] == 1.0
    r = 60.0 / 6
    np.testing.assert_allclose(preds[0, 0], [(2 + 0.25) * r + 100 - 30, (3 - 0.25) * r + 200 - r * 4], rtol=1e-6)
    np.testing.assert_allclose(preds[0, 1], [100 - 30, 200 - r * 4], rtol=1e-6)


def test_coco_results_oracle_matches_reference_validation_step(golden):
    """a15 / 8f-3: the restated result loop against the rows the reference's own validation_step produced
    (oracle/gen_golden.py:gen_pose_results)."""
    from oracle import results as ores
    g = golden("pose_results.npz")
    res = ores.coco_results(g["coords"], g["scores"], g["boxes"], g["areas"], g["masks"], g["is_crowd"], g["image_ids"])
    assert [r["image_id"] for r in res] == g["res_image_id"].tolist()
    assert np.array_equal(np.array([r["keypoints"] for r in res]), g["res_keypoints"])       # bit-exact, incl. v
    assert np.array_equal(np.array([r["bbox"] for r in res]), g["res_bbox"])
    np.testing.assert_allclose([r["score"] for r in res], g["res_score"], rtol=1e-6)
    np.testing.assert_allclose([r["area"] for r in res], g["res_area"], rtol=0)
    # the soft-argmax oracle reproduces the coordinates the reference decoded (batch of 1: no channel swap)
    for b in range(3):
        avg = opose.flip_average(torch.from_numpy(g["hm"][b:b + 1]), torch.from_numpy(g["flipped"][b:b + 1]), None)
        c, s = opose.soft_argmax_decode(avg, torch.from_numpy(g["boxes"][b:b + 1, 0]))
        np.testing.assert_allclose(c[0].numpy(), g["coords"][b], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(s[0].numpy(), g["scores"][b], rtol=1e-5)


def test_oks_known_answers():
    """COCOeval.computeOks restatement (parity unpinned): closed-form cases."""
    from oracle import results as ores
    k = 17
    gt = np.zeros((k, 3))
    gt[:, 0] = np.arange(k) * 3.0
    gt[:, 1] = 50.0
    gt[:, 2] = 2
    assert ores.compute_oks(gt[:, :2], gt, area=1000.0) == 1.0
    # one joint off by d: oks = (k - 1 + exp(-d^2 / (2 * area * (2 sigma)^2))) / k
    pred = gt[:, :2].copy()
    pred[5, 0] += 4.0
    want = (k - 1 + np.exp(-16.0 / (2 * 1000.0 * (2 * float(np.float32(0.079))) ** 2))) / k   # sigmas are fp32 (datamodule.py:37-40)
    assert abs(ores.compute_oks(pred, gt, area=1000.0) - want) < 1e-12
    # unlabelled joints do not count
    gt2 = gt.copy()
    gt2[5, 2] = 0
    assert ores.compute_oks(pred, gt2, area=1000.0) == 1.0
    # nothing labelled: distance to the doubled box
    gt3 = gt.copy()
    gt3[:, 2] = 0
    inside = np.tile([[20.0, 20.0]], (k, 1))
    assert ores.compute_oks(inside, gt3, area=1000.0, gt_box_xywh=[10, 10, 20, 20]) == 1.0
    assert ores.compute_oks(inside + 500.0, gt3, area=1000.0, gt_box_xywh=[10, 10, 20, 20]) < 1e-6


def test_hf_float32_index_quirk_q6(synth):
    """Reference quirk Q6: HF's DARK taps are addressed through a float32 flat index (image_processing_vitpose.py:248-250),
    exact only for the first 2^24 / ((W+2)(H+2)) = 5 084 maps of a call.  The literal restatement (hf_dark_decode, which
    the GPU quirk mode is held to) and the per-joint restatement (dark_decode_local, the default kernel's target) agree
    bit for bit before that map and part ways after it."""
    hs = synth.make_heatmaps(44, 133, seed=5)                       # 5 852 maps: the last 768 are past the limit
    boxes = synth.make_crop_set(4, 480, 640, per_frame=11, seed=5).boxes.tolist()
    kp_hf, sc, _ = opose.hf_dark_decode(hs.heatmaps.numpy(), boxes)
    kp_loc, sc2, _ = opose.dark_decode_local(hs.heatmaps.numpy(), boxes)
    first = 2 ** 24 // (50 * 66)
    d = np.abs(kp_hf - kp_loc).reshape(-1, 2).max(1)
    v = (sc > 0).reshape(-1)
    np.testing.assert_array_equal(sc, sc2)
    assert d[:first][v[:first]].max() < 1e-3
    assert (d[first:][v[first:]] > 1e-3).mean() > 0.3

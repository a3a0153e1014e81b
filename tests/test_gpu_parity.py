"""Parity of the sm_100a kernels (through the C ABI) against the CPU oracle.  Needs a B200.

Bars (BASELINE.json north_star): bit-exact NMS keep indices, arg-max joint indices and matched
identity ids; keypoint / score / embedding floats within 1e-3 relative tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import crop as ocrop
from oracle import det as odet
from oracle import match as omatch
from oracle import pose as opose

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def _close(a, b, rtol=RTOL, atol=0.0, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} beyond tolerance, max err {err.max():.3e} at {np.argmax(err - tol)}"


@pytest.fixture(params=["persistent", "per_item"])
def crop_path(request, spp):
    """Both crop implementations: the persistent plan + stream kernels (spp_crop_affine*_ws, the default) and the
    one-CTA-per-item kernels (spp_crop_affine / _u8, no workspace)."""
    prev = spp.ops.CROP_USE_WORKSPACE
    spp.ops.CROP_USE_WORKSPACE = request.param == "persistent"
    prev_policy = spp._lib.lib().spp_crop_policy(2 if request.param == "persistent" else 0)     # also for uint8 frames
    yield request.param
    spp.ops.CROP_USE_WORKSPACE = prev
    spp._lib.lib().spp_crop_policy(prev_policy)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------------
# heatmap decode
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("k", [17, 5])
def test_heatmap_dark_vs_oracle(spp, synth, dev, flip, k):
    hs = synth.make_heatmaps(9, k, seed=100 + k, negative_frac=0.1)
    cs = synth.make_crop_set(1, 480, 640, per_frame=9, seed=7)
    boxes = [[float(v) for v in b] for b in cs.boxes]
    avg = opose.flip_average(hs.heatmaps, hs.flipped, hs.perm) if flip else hs.heatmaps
    kp_o, sc_o, idx_o = opose.hf_dark_decode(avg.numpy(), boxes)
    kp, sc, am = spp.heatmap_decode(hs.heatmaps.to(dev), hs.flipped.to(dev) if flip else None,
                                    hs.perm.to(dev) if flip else None, cs.boxes.to(dev), "dark", 11)
    np.testing.assert_array_equal(am.cpu().numpy(), idx_o)                 # arg-max indices: bit-exact
    np.testing.assert_array_equal(sc.cpu().numpy(), sc_o)                   # max of the averaged map: bit-exact
    valid = sc_o > 0
    _close(kp.cpu().numpy()[valid], kp_o[valid], what="DARK keypoints (valid joints)")
    # joints with score <= 0 follow HF's flat-index behaviour (reads the previous map's tail)
    _close(kp.cpu().numpy()[~valid], kp_o[~valid], atol=1e-2, what="DARK keypoints (score<=0 joints)")


def test_heatmap_dark_golden(spp, golden, dev):
    g = golden("pose_hf.npz")
    kp, sc, _ = spp.heatmap_decode(torch.from_numpy(g["hm"]).to(dev), torch.from_numpy(g["flipped"]).to(dev),
                                   torch.from_numpy(g["perm"]).to(dev), torch.from_numpy(g["boxes"]).to(dev), "dark", 11)
    np.testing.assert_array_equal(sc.cpu().numpy(), g["scores"])
    valid = g["scores"] > 0
    _close(kp.cpu().numpy()[valid], g["keypoints"][valid], what="DARK vs HF fixture")


def test_heatmap_dark_heatmap_space_and_kernel_sizes(spp, synth, dev):
    hs = synth.make_heatmaps(4, 17, seed=5, negative_frac=0.0)
    for kernel in (3, 7, 11, 17):
        coords, scores, idx = opose.argmax_predictions(hs.heatmaps.numpy())
        ref = opose.dark_refine_full(coords, hs.heatmaps.numpy(), kernel=kernel)
        kp, sc, am = spp.heatmap_decode(hs.heatmaps.to(dev), mode="dark", kernel=kernel)
        np.testing.assert_array_equal(am.cpu().numpy(), idx)
        _close(kp.cpu().numpy(), ref, atol=1e-3, what=f"DARK heatmap-space kernel={kernel}")


def test_heatmap_softargmax_vs_oracle_and_golden(spp, golden, dev):
    g = golden("pose_live.npz")
    hm, fl, perm = (torch.from_numpy(g[k]) for k in ("hm", "flipped", "perm"))
    boxes = torch.from_numpy(g["boxes_xyxy"])
    c, s = spp.get_keypoints_from_heatmaps(hm.to(dev))
    _close(c.cpu().numpy(), g["coords_plain"], what="softargmax coords (reference fixture)")
    _close(s.cpu().numpy(), g["scores_plain"], what="softargmax scores (reference fixture)")
    kp, sc, am = spp.heatmap_decode(hm.to(dev), fl.to(dev), perm.to(dev), boxes.to(dev), "softargmax", flags=1)
    _close(kp.cpu().numpy(), g["coords_avg_box"], what="softargmax+flip coords")
    _close(sc.cpu().numpy(), g["scores_avg_box"], what="softargmax+flip scores")
    np.testing.assert_array_equal(am.cpu().numpy(), torch.from_numpy(g["avg"]).flatten(2).argmax(2).numpy())
    out = spp.flip_test_keypoints(hm.to(dev), fl.to(dev), boxes.to(dev))
    ref = opose.backproject_live(torch.from_numpy(g["coords_avg_box"]), torch.from_numpy(g["scores_avg_box"]), boxes)
    _close(out.cpu().numpy(), ref.numpy(), what="flip_test_keypoints")


def test_heatmap_quarter_vs_oracle(spp, synth, dev):
    hs = synth.make_heatmaps(6, 17, seed=9, negative_frac=0.1)
    cs = synth.make_crop_set(1, 480, 640, per_frame=6, seed=3)
    cen, scl = zip(*[ocrop.center_scale_v2(b.tolist()) for b in cs.boxes])
    cen, scl = np.asarray(cen, np.float32), np.asarray(scl, np.float32)
    ref, mv, idx = opose.quarter_offset_decode(hs.heatmaps.numpy(), cen, scl)
    kp, sc, am = spp.heatmap_decode(hs.heatmaps.to(dev), None, None, cs.boxes.to(dev), "quarter")
    np.testing.assert_array_equal(am.cpu().numpy(), idx)
    _close(kp.cpu().numpy(), ref, atol=1e-3, what="quarter-offset keypoints")
    kp2, mv2 = spp.get_final_preds(hs.heatmaps.to(dev), torch.from_numpy(cen), torch.from_numpy(scl))
    _close(kp2.cpu().numpy(), ref, atol=1e-3, what="get_final_preds shim")
    np.testing.assert_array_equal(mv2.cpu().numpy()[..., 0], mv)


def test_heatmap_full_size_properties(spp, synth, dev):
    """cfg2 size (640 crops x 17 joints, flip test): arg-max indices against torch on the device-side
    average, and the planted sub-pixel centres are recovered."""
    hs = synth.make_heatmaps(640, 17, seed=0)
    hm, fl, perm = hs.heatmaps.to(dev), hs.flipped.to(dev), hs.perm.to(dev)
    kp, sc, am = spp.heatmap_decode(hm, fl, perm, None, "dark", 11)
    avg = (hm + fl[:, perm.long()].flip(-1)) * 0.5
    ref_max, ref_idx = avg.flatten(2).max(2)
    assert torch.equal(am.long(), ref_idx) and torch.equal(sc, ref_max)
    good = (~hs.negative).to(dev)
    err = (kp - hs.centres.to(dev)).norm(dim=-1)[good]
    assert float(err.median()) < 0.15 and float(err.quantile(0.99)) < 1.0


# ------------------------------------------------------------------------------------------------
# detection decode + NMS
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag", ["nc1", "nc3"])
def test_det_golden(spp, golden, dev, tag):
    g = golden(f"det_{tag}.npz")
    levels = [torch.from_numpy(g[k]).to(dev) for k in ("l0", "l1", "l2")]
    conf, iou = float(g["conf"]), float(g["iou"])
    dec = spp.head_forward(levels)
    _close(dec.cpu().numpy(), g["decoded"], atol=1e-3, what="head decode")
    # (1) the reference's signature on the reference's own decoded tensor: bit-exact rows
    dets = spp.non_max_suppression(torch.from_numpy(g["decoded"]).to(dev), conf, iou)
    assert [d.shape[0] for d in dets] == list(g["n"])
    np.testing.assert_array_equal(torch.cat(dets).cpu().numpy(), g["dets"])
    # (2) fused raw path: same keep set (bit-exact class / anchor keys), floats within tolerance
    ref_rows, ref_keys = odet.non_max_suppression(torch.from_numpy(g["decoded"]), conf, iou, return_index=True)
    res = spp.decode_nms(levels, conf_thres=conf, iou_thres=iou)
    assert res.count.tolist() == [r.shape[0] for r in ref_rows]
    for rows, keys, rr, rk in zip(res.to_list(), res.keys_list(), ref_rows, ref_keys):
        np.testing.assert_array_equal(keys.cpu().numpy(), rk.numpy())
        _close(rows.cpu().numpy(), rr.numpy(), atol=1e-3, what="fused decode+NMS rows")


@pytest.mark.parametrize("dense", [False, True])
def test_det_vs_oracle_bigger(spp, synth, dev, dense):
    hm = synth.make_head_maps(3, 384, 640, n_obj=12, nc=1, seed=4, dense=dense)
    dec_o = odet.head_decode(hm.levels)
    conf = 0.001 if not dense else 0.3
    rows_o, keys_o = odet.non_max_suppression(dec_o, conf, 0.65, return_index=True)
    near = sum(odet.near_threshold_pairs(r, 0.65) for r in rows_o)
    # NMS on the oracle's decoded tensor must be bit-exact regardless of borderline pairs
    res = spp.nms_decoded(dec_o.to(dev), conf, 0.65)
    for rows, keys, rr, rk in zip(res.to_list(), res.keys_list(), rows_o, keys_o):
        np.testing.assert_array_equal(keys.cpu().numpy(), rk.numpy())
        np.testing.assert_array_equal(rows.cpu().numpy(), rr.numpy())
    if near == 0 and not dense:
        res = spp.decode_nms([l.to(dev) for l in hm.levels], conf_thres=conf, iou_thres=0.65)
        for keys, rk in zip(res.keys_list(), keys_o):
            np.testing.assert_array_equal(keys.cpu().numpy(), rk.numpy())


def test_det_edge_cases(spp, synth, dev):
    # no candidates at all
    hm = synth.make_head_maps(2, 64, 64, n_obj=0, nc=1, seed=1)
    res = spp.decode_nms([l.to(dev) for l in hm.levels])
    assert res.count.tolist() == [0, 0] and all(r.shape == (0, 6) for r in res.to_list())
    # more than max_det survivors: distinct far-apart boxes, every anchor a candidate
    pred = torch.zeros(1, 5, 1000)
    pred[0, 0] = torch.arange(1000) * 50.0
    pred[0, 1] = 10.0
    pred[0, 2:4] = 20.0
    pred[0, 4] = torch.linspace(0.9, 0.1, 1000)
    out = spp.non_max_suppression(pred.to(dev))
    ref = odet.non_max_suppression(pred)
    assert out[0].shape == (300, 6)
    np.testing.assert_array_equal(out[0].cpu().numpy(), ref[0].numpy())
    # exact score ties resolve to the lower anchor first (stable order)
    pred[0, 4] = 0.5
    pred[0, 0] = (torch.arange(1000) // 2) * 50.0          # pairs of identical boxes
    out = spp.nms_decoded(pred.to(dev))
    ref_rows, ref_keys = odet.non_max_suppression(pred, return_index=True)
    np.testing.assert_array_equal(out.keys_list()[0].cpu().numpy(), ref_keys[0].numpy())


# ------------------------------------------------------------------------------------------------
# crop
# ------------------------------------------------------------------------------------------------

def test_crop_vs_oracle(spp, synth, dev, crop_path):
    cs = synth.make_crop_set(2, 360, 480, per_frame=6, seed=11)
    ref = ocrop.crop_affine_hf(cs.frames.numpy(), cs.boxes.tolist(), cs.frame_idx.tolist())
    out = spp.crop_affine(cs.frames.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev))
    _close(out.cpu().numpy(), ref, atol=1e-3, what="HF/UDP crop")
    assert float(np.abs(out.cpu().numpy() - ref).max()) < 2e-5
    ref = ocrop.crop_affine_v2(cs.frames.numpy(), cs.boxes.tolist(), cs.frame_idx.tolist())
    out = spp.crop_affine(cs.frames.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev), variant="gluoncv")
    assert float(np.abs(out.cpu().numpy() - ref).max()) < 2e-5


def test_crop_uint8_vs_oracle(spp, synth, dev, crop_path):
    cs = synth.make_crop_set(2, 360, 480, per_frame=6, seed=12)
    fr8 = (cs.frames * 255.0).round().clamp(0, 255).to(torch.uint8)
    ref = ocrop.crop_affine_hf(fr8.numpy(), cs.boxes.tolist(), cs.frame_idx.tolist(), rescale_factor=1 / 255)
    m, s = ocrop.fused_mean_std(rescale_factor=1 / 255)
    out = spp.crop_affine(fr8.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev), mean=m.tolist(), std=s.tolist())
    diff = np.abs(out.cpu().numpy() - ref)
    assert float(diff.max()) < 2e-5, f"{(diff > 2e-5).sum()} pixels differ (a flipped uint8 rounding shows up as ~0.017)"
    proc = spp.VitPoseImageProcessor()
    boxes = [[[float(v) for v in b] for b in cs.boxes[:6]], [[float(v) for v in b] for b in cs.boxes[6:]]]
    pix = proc.preprocess(fr8.to(dev), boxes)["pixel_values"]
    assert float(np.abs(pix.cpu().numpy() - ref).max()) < 2e-5
    # near-ties: a ~2x upscale of a low-amplitude image puts most interpolated values within the fast path's
    # guard band of x.5, so nearly every pixel takes the fp64 re-computation
    g = torch.Generator().manual_seed(5)
    lo8 = torch.randint(0, 8, (1, 3, 360, 480), generator=g, dtype=torch.uint8)
    tb = [[100.1625, 60.55, 57.3, 76.4], [200.55, 100.7333, 76.4, 101.8667], [-20.0875, 300.55, 57.3, 76.4]]
    ref = ocrop.crop_affine_hf(lo8.numpy(), tb, [0, 0, 0], rescale_factor=1 / 255)
    out = spp.crop_affine(lo8.to(dev), torch.tensor(tb, device=dev), torch.zeros(3, dtype=torch.int32, device=dev),
                          mean=m.tolist(), std=s.tolist())
    diff = np.abs(out.cpu().numpy() - ref)
    assert float(diff.max()) < 2e-5, f"{(diff > 2e-5).sum()} near-tie pixels differ"
    # unaligned frame width: bulk-TMA staging is not possible, the direct path must agree
    fr8b = fr8[:, :, :, :478].contiguous()
    ref = ocrop.crop_affine_hf(fr8b.numpy(), cs.boxes.tolist(), cs.frame_idx.tolist(), rescale_factor=1 / 255)
    out = spp.crop_affine(fr8b.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev), mean=m.tolist(), std=s.tolist())
    assert float(np.abs(out.cpu().numpy() - ref).max()) < 2e-5
    frb = cs.frames[:, :, :, :478].contiguous()
    ref = ocrop.crop_affine_hf(frb.numpy(), cs.boxes.tolist(), cs.frame_idx.tolist())
    out = spp.crop_affine(frb.to(dev), cs.boxes.to(dev), cs.frame_idx.to(dev))
    assert float(np.abs(out.cpu().numpy() - ref).max()) < 2e-5


def test_crop_golden_and_processor_shim(spp, golden, dev):
    g = golden("pose_hf.npz")
    proc = spp.VitPoseImageProcessor()
    boxes = [[[float(v) for v in b] for b in g["boxes"]]]
    pix = proc.preprocess([torch.from_numpy(g["frame"]).to(dev)], boxes, do_rescale=False)["pixel_values"].cpu().numpy()
    assert float(np.abs(pix[:, :, ::4, ::4] - g["crop_sub"]).max()) < 2e-5
    np.testing.assert_allclose(pix.astype(np.float64).sum(axis=(2, 3)), g["crop_sum"], rtol=1e-5)


def test_crop_then_decode_round_trip(spp, synth, dev):
    """Back-projection is the exact inverse of the crop warp (UDP): a keypoint at heatmap position
    (x, y) maps to the frame position that the crop sampled for input pixel (x*191/47, y*255/63)."""
    cs = synth.make_crop_set(1, 480, 640, per_frame=4, seed=2)
    hs = synth.make_heatmaps(4, 17, seed=8, negative_frac=0.0)
    kp_hm, _, _ = spp.heatmap_decode(hs.heatmaps.to(dev), mode="dark")
    kp_img, _, _ = spp.heatmap_decode(hs.heatmaps.to(dev), None, None, cs.boxes.to(dev), "dark")
    for i, b in enumerate(cs.boxes.tolist()):
        c, s = ocrop.box_to_center_and_scale(b)
        xs, ys = ocrop.source_coords(ocrop.warp_matrix(c, s), 192, 256)
        x_in = kp_hm[i, :, 0].cpu().numpy().astype(np.float64) * 191 / 47
        y_in = kp_hm[i, :, 1].cpu().numpy().astype(np.float64) * 255 / 63
        ex = xs[0] + x_in * (xs[1] - xs[0])
        ey = ys[0] + y_in * (ys[1] - ys[0])
        _close(kp_img[i, :, 0].cpu().numpy(), ex, rtol=1e-4, atol=1e-3, what="round trip x")
        _close(kp_img[i, :, 1].cpu().numpy(), ey, rtol=1e-4, atol=1e-3, what="round trip y")


# ------------------------------------------------------------------------------------------------
# gallery match
# ------------------------------------------------------------------------------------------------

def _bf16_gallery(ms):
    return ms.gallery.to(torch.bfloat16)


def test_l2_normalize_matches_reference_fixture(spp, golden, dev):
    g = golden("match.npz")
    emb, norm = spp.backbone_tail(torch.from_numpy(g["pre"]).to(dev))
    _close(emb.cpu().numpy(), g["emb"], rtol=1e-5, atol=1e-7, what="embedding")
    _close(norm.cpu().numpy(), g["norm"], rtol=1e-5, what="norm")
    kn = spp.l2_norm(torch.from_numpy(g["kernel"]).to(dev), axis=0)
    _close(kn.cpu().numpy(), g["kernel_l2"], rtol=1e-5, atol=1e-8, what="l2_norm axis 0")


@pytest.mark.parametrize("m,n", [(5, 100), (48, 300), (640, 10000), (130, 777)])
@pytest.mark.parametrize("simt", [True, False])
def test_match_vs_oracle(spp, synth, dev, m, n, simt):
    ms = synth.make_match_set(m, n, seed=m + n)
    gal = _bf16_gallery(ms)
    pred_o, sim_o = omatch.match_top1(ms.embeddings, gal.float(), threshold=0.4)
    gap = omatch.top2_gap(ms.embeddings, gal.float())
    ids, sims = spp.match_top1(ms.embeddings.to(dev), gal.to(dev), 0.4, _simt=simt)
    _close(sims.cpu().numpy(), sim_o.numpy(), rtol=RTOL, atol=1e-5, what="match similarity")
    decided = (gap > 1e-5) & ((sim_o - 0.4).abs() > 1e-5)
    np.testing.assert_array_equal(ids.cpu().numpy()[decided], pred_o.numpy()[decided])
    known = ms.true_ids >= 0
    np.testing.assert_array_equal(ids.cpu().numpy()[known], ms.true_ids.numpy()[known])
    # un-gated: fp32 arg-max ids
    pred_o, _ = omatch.match_top1(ms.embeddings, gal.float())
    ids, _ = spp.match_top1(ms.embeddings.to(dev), gal.to(dev), None, _simt=simt)
    np.testing.assert_array_equal(ids.cpu().numpy()[gap > 1e-5], pred_o.numpy()[gap > 1e-5])


def test_match_reference_fixture_and_keys(spp, golden, dev):
    g = golden("match.npz")
    kernel = torch.from_numpy(g["kernel"])
    gal = spp.Gallery.from_kernel(kernel.to(dev))
    ids, sims = gal.match(torch.from_numpy(g["probes"]).to(dev), threshold=0.4)
    known = g["true_ids"] >= 0
    np.testing.assert_array_equal(ids.cpu().numpy()[known], g["pred"][known])
    assert (ids.cpu().numpy()[~known] == -1).all()
    # bf16-only gallery: every gallery element carries a relative rounding error <= 2^-9, so |sim - sim_fp32| <= 2^-9
    # (Cauchy-Schwarz, unit vectors) = 2.2e-3 of a planted similarity ~0.9 in the worst case — DESIGN.md section 3.4
    _close(sims.cpu().numpy()[known], g["sim"][known], rtol=2.2e-3, what="similarity vs fp32-gallery fixture (bf16 gallery bound)")
    # with the fp32 rows kept at enrolment the re-score runs against them: the reference's own fp32 numbers at 1e-3
    gal32 = spp.Gallery.from_kernel(kernel.to(dev), keep_f32=True)
    ids32, sims32 = gal32.match(torch.from_numpy(g["probes"]).to(dev), threshold=0.4)
    np.testing.assert_array_equal(ids32.cpu().numpy()[known], g["pred"][known])
    _close(sims32.cpu().numpy()[known], g["sim"][known], rtol=RTOL, what="similarity vs fp32-gallery fixture (fp32 re-score)")
    # sharded gallery: per-shard keys + integer MAX == single-shard result
    rows = gal.rows
    full_ids, full_sims = spp.match_top1(torch.from_numpy(g["probes"]).to(dev), rows)
    k0 = spp.match_top1(torch.from_numpy(g["probes"]).to(dev), rows[:128].contiguous(), None, 0, want_keys=True)[2]
    k1 = spp.match_top1(torch.from_numpy(g["probes"]).to(dev), rows[128:].contiguous(), None, 128, want_keys=True)[2]
    ids2, sims2 = spp.match_unpack_keys(torch.maximum(k0, k1))
    assert torch.equal(ids2, full_ids) and torch.equal(sims2, full_sims)


# ------------------------------------------------------------------------------------------------
# match exactness: the id is the fp32 arg-max for ANY gallery content (near-duplicate enrolments, non-unit rows)
# ------------------------------------------------------------------------------------------------

def _equidistant_rows(target: torch.Tensor, count: int, g: torch.Generator) -> torch.Tensor:
    """``count`` unit rows at the SAME angle (45 degrees) from ``target`` in random orthogonal directions, rounded to bf16:
    for a probe near ``target`` their fp32 cosines agree to ~1e-4 — inside the bf16 score error, so the bf16 ranking of
    these rows is scrambled relative to the fp32 one."""
    u = torch.randn(count, target.numel(), generator=g)
    u = u - (u @ target)[:, None] * target[None]
    u = u / u.norm(dim=1, keepdim=True)
    return ((target[None] + u) / 2 ** 0.5).to(torch.bfloat16)


@pytest.mark.parametrize("n,slots", [(1000, list(range(100, 112))), (1000, list(range(250, 262))),
                                     (300000, [70000 + 83 * i for i in range(12)])])
@pytest.mark.parametrize("simt", [False, True])
def test_match_adversarial_close_enrolments(spp, synth, dev, n, slots, simt):
    """Many gallery rows of one chunk inside the bf16 error band of the winner (VERDICT r1 weak 1 / ADVICE): the
    epilogue's top-2 cannot hold them all and the bf16 ranking is not the fp32 one, so the dropped-score flag must send
    the chunk to the exact fp32 re-scan.  The set-up is checked to defeat a plain top-2-per-chunk search."""
    if simt and n > 10000:
        pytest.skip("SIMT cross-check only at small sizes")
    m = 96
    ms = synth.make_match_set(m, n, seed=31)
    for attempt in range(16):            # the rows' own bf16 rounding decides how adversarial a draw is: take the first good one
        g = torch.Generator().manual_seed(n + slots[0] + attempt)
        gal = ms.gallery.to(torch.bfloat16)
        target = torch.randn(512, generator=g)
        target = target / target.norm()
        gal[slots] = _equidistant_rows(target, len(slots), g)
        probes = ms.embeddings.clone()
        e = torch.randn(64, 512, generator=g)
        probes[:64] = (target[None] + 1e-3 * e / e.norm(dim=1, keepdim=True)) * 7.0     # 64 probes aimed between the close rows
        galf = gal.float()
        pred_o, sim_o = omatch.match_top1(probes, galf)
        # emulated candidate search (bf16 probe x bf16 gallery): how often is the fp32 winner outside the bf16 top-2?
        qb = torch.nn.functional.normalize(probes[:64]).to(torch.bfloat16).float()
        top2 = torch.tensor(slots)[(qb @ galf[slots].t()).topk(2, 1).indices]
        missed = int(((top2 == pred_o[:64, None]).sum(1) == 0).sum())
        if missed >= 10:
            break
    assert missed >= 10, f"set-up too easy: only {missed} probes would defeat a top-2-per-chunk search"
    gap = omatch.top2_gap(probes, galf)
    assert int(torch.isin(pred_o[:64], torch.tensor(slots)).sum()) == 64
    ids, sims = spp.match_top1(probes.to(dev), gal.to(dev), None, _simt=simt)
    decided = gap > 1e-6
    assert int(decided[:64].sum()) >= 48
    np.testing.assert_array_equal(ids.cpu().numpy()[decided], pred_o.numpy()[decided])
    _close(sims.cpu().numpy(), sim_o.numpy(), rtol=1e-5, atol=1e-6, what="similarity of the winner")
    # the result does not depend on how the gallery is cut into shards
    cut = slots[len(slots) // 2]
    k0 = spp.match_top1(probes.to(dev), gal[:cut].contiguous().to(dev), None, 0, want_keys=True)[2]
    k1 = spp.match_top1(probes.to(dev), gal[cut:].contiguous().to(dev), None, cut, want_keys=True)[2]
    ids2, sims2 = spp.match_unpack_keys(torch.maximum(k0, k1))
    assert torch.equal(ids2, ids) and torch.equal(sims2, sims)


def test_match_every_row_identical_and_exact_ties(spp, dev):
    """A gallery of identical rows: every chunk is flagged for every probe, the answer is id 0 (first maximum)."""
    g = torch.Generator().manual_seed(1)
    row = torch.randn(1, 512, generator=g)
    row = (row / row.norm()).to(torch.bfloat16)
    gal = row.repeat(700, 1).contiguous()
    probes = torch.randn(9, 512, generator=g)
    ids, sims = spp.match_top1(probes.to(dev), gal.to(dev))
    assert ids.tolist() == [0] * 9
    ids, _ = spp.match_top1(probes.to(dev), gal.to(dev), None, 1000)
    assert ids.tolist() == [1000] * 9


def test_match_fp32_gallery_rescore_and_non_unit_rows(spp, synth, dev):
    """(a) fp32 rows given: ids and sims of the reference's fp32 F.linear().max(1) on the UN-rounded gallery;
    (b) quirk-Q3 enrolment (rows not unit-norm): max_row_norm widens the re-score band."""
    ms = synth.make_match_set(200, 5000, seed=77)
    pred_o, sim_o = omatch.match_top1(ms.embeddings, ms.gallery, threshold=0.4)
    gap = omatch.top2_gap(ms.embeddings, ms.gallery)
    gal = spp.Gallery.from_rows(ms.gallery.to(dev), normalize=False, keep_f32=True)
    assert abs(gal.max_row_norm - 1.0) < 1e-2
    ids, sims = gal.match(ms.embeddings.to(dev), threshold=0.4)
    ok = (gap > 1e-6) & ((sim_o - 0.4).abs() > 1e-5)
    np.testing.assert_array_equal(ids.cpu().numpy()[ok], pred_o.numpy()[ok])
    _close(sims.cpu().numpy(), sim_o.numpy(), rtol=1e-5, atol=1e-6, what="fp32-gallery similarity")
    # (b) a [512, N] kernel normalised along the wrong axis: row norms spread over ~[0.2, 0.4] * sqrt(N / 512)
    kern = torch.randn(512, 3000, generator=torch.Generator().manual_seed(5)) * torch.linspace(0.5, 2.0, 3000)[None]
    rows = omatch.enrol_gallery(kern, quirk_q3=True)
    galq = spp.Gallery.from_kernel(kern.to(dev), quirk_q3=True)
    assert galq.max_row_norm == pytest.approx(float(galq.rows.float().norm(dim=1).max()), rel=1e-6)
    probes = rows[torch.arange(0, 3000, 50)] + 0.01 * torch.randn(60, 512, generator=torch.Generator().manual_seed(6))
    rb = galq.rows.float().cpu()
    pred_q, sim_q = omatch.match_top1(probes, rb)
    gq = omatch.top2_gap(probes, rb)
    ids, sims = galq.match(probes.to(dev))
    np.testing.assert_array_equal(ids.cpu().numpy()[gq > 1e-6], pred_q.numpy()[gq > 1e-6])
    _close(sims.cpu().numpy(), sim_q.numpy(), rtol=1e-5, atol=1e-6, what="quirk-Q3 gallery similarity")


# ------------------------------------------------------------------------------------------------
# gallery sharded over ranks, exchange through peer memory (virtual ranks on one GPU; 2 processes below)
# ------------------------------------------------------------------------------------------------

def test_peer_sharded_match_virtual_ranks(spp, synth, dev):
    """Three virtual ranks in one process, one exchange buffer and one stream each: the real kernels, flags and waits.
    Stages are enqueued rank-interleaved (all pushes, then all searches, then all reduces) so the test does not depend on
    the streams actually running concurrently.  Result = the single-GPU match on the whole gallery, bit for bit, over
    several steps (both parities) and with different probes per step."""
    d = spp.dist
    world, m, n = 3, 40, 3000
    torch.cuda.set_device(dev)
    groups = d.PeerGroup.virtual(world, m)
    ms = synth.make_match_set(world * m * 4, n, seed=12)
    gal = ms.gallery.to(torch.bfloat16).to(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    matchers = []
    for r in range(world):
        lo, hi = d.shard_bounds(n, world, r)
        matchers.append(d.PeerShardedMatcher(groups[r], gal[lo:hi].contiguous(), lo, threshold=0.4))
    for step in range(4):
        probes = [ms.embeddings[(step * world + r) * m:(step * world + r + 1) * m].contiguous().to(dev) for r in range(world)]
        torch.cuda.synchronize()
        for stages in (d.STAGE_PUSH, d.STAGE_WAIT | d.STAGE_SEARCH | d.STAGE_FINALIZE, d.STAGE_REDUCE):
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    matchers[r].match(probes[r], stages)
        torch.cuda.synchronize()
        for r in range(world):
            ref_ids, ref_sims = spp.match_top1(probes[r], gal, 0.4)
            assert torch.equal(matchers[r].ids, ref_ids), (step, r)
            assert torch.equal(matchers[r].sims, ref_sims), (step, r)
    # world = 1 degenerates to a local match through the same five kernels, in one call, and is graph-capturable
    g1 = d.PeerGroup.virtual(1, m)[0]
    m1 = d.PeerShardedMatcher(g1, gal, 0, threshold=0.4)
    pr = ms.embeddings[:m].contiguous().to(dev)
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        m1.match(pr)
    st.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        m1.match(pr)
    for _ in range(3):                      # odd and even parities replay through the same graph
        with torch.cuda.stream(st):
            graph.replay()
    st.synchronize()
    ref_ids, ref_sims = spp.match_top1(pr, gal, 0.4)
    assert torch.equal(m1.ids, ref_ids) and torch.equal(m1.sims, ref_sims)


def _peer_worker(rank, world, port, n, m, out):
    import importlib
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
    ms = spp.synth.make_match_set(world * m * 3, n, seed=17)
    gal = ms.gallery.to(torch.bfloat16)
    lo, hi = spp.dist.shard_bounds(n, world, rank)
    peers = spp.dist.PeerGroup(m)
    matcher = spp.dist.PeerShardedMatcher(peers, gal[lo:hi].contiguous().to(dev), lo, threshold=0.4)
    ok = True
    graph = None
    for step in range(3):
        mine = ms.embeddings[(step * world + rank) * m:(step * world + rank + 1) * m].contiguous().to(dev)
        ids, sims = matcher.match(mine)
        torch.cuda.synchronize()
        ref_ids, ref_sims = spp.match_top1(mine, gal.to(dev), 0.4)
        ok = ok and bool(torch.equal(ids, ref_ids) and torch.equal(sims, ref_sims))
    # captured into a CUDA graph and replayed (the way the pipeline runs it)
    st = torch.cuda.Stream(dev)
    dist.barrier()
    with torch.cuda.stream(st):
        matcher.match(mine)
    st.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        matcher.match(mine)
    dist.barrier()
    for _ in range(5):
        with torch.cuda.stream(st):
            graph.replay()
    st.synchronize()
    ok = ok and bool(torch.equal(matcher.ids, ref_ids) and torch.equal(matcher.sims, ref_sims))
    out[rank] = ok
    dist.barrier()
    peers.close()
    dist.destroy_process_group()


def test_peer_sharded_match_two_processes(dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_peer_worker, args=(world, port, 5000, 96, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


# ------------------------------------------------------------------------------------------------
# round-2 fixes: overflow flag, mirrored crops, result de-duplication
# ------------------------------------------------------------------------------------------------

def test_nms_candidate_overflow_flag(spp, dev):
    pred = torch.zeros(2, 5, 400)
    pred[:, 0] = torch.arange(400) * 50.0
    pred[:, 1] = 10.0
    pred[:, 2:4] = 20.0
    pred[0, 4] = torch.linspace(0.9, 0.1, 400)          # 400 candidates > max_candidates
    pred[1, 4, :3] = 0.5                                # 3 candidates
    res = spp.nms_decoded(pred.to(dev), max_candidates=64)
    assert res.overflowed().tolist() == [True, False]
    assert res.kept().tolist()[1] == 3 and 0 < res.kept().tolist()[0] <= 64
    assert res.count.tolist()[0] == ~res.kept().tolist()[0]
    # overflow with nothing kept is still flagged (count -1): max_det rows of an empty image cannot happen, so force it with max_det 1
    assert len(res.to_list()[1]) == 3


def test_crop_mirrored_boxes_vs_hf(spp, dev, crop_path):
    """Negative width AND height: HF / scipy produce the mirrored crop (the staged kernel assumed a
    non-decreasing source map).  Compared against the real HF preprocess."""
    from transformers import VitPoseImageProcessor
    g = torch.Generator().manual_seed(4)
    fr = torch.rand(1, 3, 360, 480, generator=g)
    boxes = [[150.0, 120.0, -60.0, -90.0], [50.0, 40.0, -30.0, 80.0], [60.0, 150.0, 40.0, -70.0], [400.0, 300.0, -300.0, -250.0],
             [100.0, 100.0, 80.0, 120.0]]
    want = VitPoseImageProcessor().preprocess([fr[0]], boxes=[boxes], do_rescale=False, return_tensors="pt")["pixel_values"]
    got = spp.crop_affine(fr.to(dev), torch.tensor(boxes, device=dev), torch.zeros(len(boxes), dtype=torch.int32, device=dev))
    assert float((got.cpu() - want).abs().max()) < 2e-5
    fr8 = (fr * 255).round().to(torch.uint8)
    want8 = VitPoseImageProcessor().preprocess([fr8[0]], boxes=[boxes], return_tensors="pt")["pixel_values"]
    got8 = spp.VitPoseImageProcessor().preprocess(fr8.to(dev), [boxes])["pixel_values"]
    assert float((got8.cpu() - want8).abs().max()) < 2e-5


def test_coco_results_skip_seen_image_ids(spp, golden, dev):
    """module.py:509-513: an image id already in the evaluation cache is skipped."""
    g = golden("pose_results.npz")
    args = (torch.from_numpy(g["coords"]).to(dev), torch.from_numpy(g["scores"]).to(dev), torch.from_numpy(g["boxes"]).to(dev),
            torch.from_numpy(g["areas"]), torch.from_numpy(g["masks"]), torch.from_numpy(g["is_crowd"]))
    ids = g["image_ids"].tolist()
    seen = set()
    first = spp.coco_keypoint_results(*args, ids, seen_image_ids=seen)
    assert seen == set(ids) and len(first) == len(g["res_image_id"])
    assert spp.coco_keypoint_results(*args, ids, seen_image_ids=seen) == []
    dup = spp.coco_keypoint_results(*args, [ids[0]] * len(ids))          # the same id repeated inside one batch: first only
    assert {r["image_id"] for r in dup} == {ids[0]} and len(dup) == sum(1 for r in first if r["image_id"] == ids[0])


# ------------------------------------------------------------------------------------------------
# face -> person association (not in the reference; parity against the builder's own oracle)
# ------------------------------------------------------------------------------------------------

def test_associate_vs_oracle(spp, synth, dev):
    from oracle import assoc as oassoc
    g = torch.Generator().manual_seed(5)
    b, fcap, pcap, cap = 6, 300, 300, 8
    face = spp.NmsResult(torch.zeros(b, fcap, 6), torch.zeros(b, dtype=torch.int32), torch.zeros(b, fcap, dtype=torch.int32))
    person = spp.NmsResult(torch.zeros(b, pcap, 6), torch.zeros(b, dtype=torch.int32), torch.zeros(b, pcap, dtype=torch.int32))
    ids = torch.full((b, fcap), -1, dtype=torch.int32)
    for i in range(b):
        npers, nface = int(torch.randint(0, 14, (1,), generator=g)), int(torch.randint(0, 20, (1,), generator=g))
        xy = torch.rand(npers, 2, generator=g) * torch.tensor([1000.0, 400.0])
        wh = torch.rand(npers, 2, generator=g) * torch.tensor([200.0, 300.0]) + 40
        person.dets[i, :npers, :4] = torch.cat([xy, xy + wh], 1)
        person.dets[i, :npers, 4] = torch.linspace(0.9, 0.2, npers) if npers else torch.zeros(0)
        person.count[i] = npers
        fxy = torch.rand(nface, 2, generator=g) * torch.tensor([1200.0, 600.0])
        if npers and nface:       # put most faces at the head of a random person
            own = torch.randint(0, npers, (nface,), generator=g)
            inside = torch.rand(nface, generator=g) < 0.8
            hx = person.dets[i, own, 0] + 0.5 * (person.dets[i, own, 2] - person.dets[i, own, 0])
            hy = person.dets[i, own, 1] + 0.1 * (person.dets[i, own, 3] - person.dets[i, own, 1])
            fxy = torch.where(inside[:, None], torch.stack([hx, hy], 1), fxy)
        fwh = torch.rand(nface, 2, generator=g) * 40 + 16
        face.dets[i, :nface, :4] = torch.cat([fxy - fwh / 2, fxy + fwh / 2], 1)
        face.count[i] = nface
        ids[i, :nface] = torch.where(torch.rand(nface, generator=g) < 0.7, torch.randint(0, 10000, (nface,), generator=g), -1).int()
    ref_boxes, ref_ident, ref_rows = oassoc.associate([face.dets[i, :face.count[i]].numpy() for i in range(b)],
                                                      [ids[i, :face.count[i]].numpy() for i in range(b)],
                                                      [person.dets[i, :person.count[i]].numpy() for i in range(b)], cap)
    fg = spp.NmsResult(face.dets.to(dev), face.count.to(dev), face.keys.to(dev))
    pg = spp.NmsResult(person.dets.to(dev), person.count.to(dev), person.keys.to(dev))
    boxes, ident, rows, count = spp.associate(fg, ids.to(dev), pg, cap)
    assert count.tolist() == [len(r) for r in ref_rows]
    assert sum(count.tolist()) > 5
    for i in range(b):
        k = len(ref_rows[i])
        np.testing.assert_array_equal(rows[i, :k].cpu().numpy(), ref_rows[i])
        np.testing.assert_array_equal(ident[i, :k].cpu().numpy(), ref_ident[i])
        np.testing.assert_array_equal(boxes[i, :k].cpu().numpy(), ref_boxes[i])
        assert (ident[i, k:] == -1).all() and (boxes[i, k:] == 0).all()


def test_pipeline_graph_vs_eager_and_device_selection(spp, synth, dev):
    """The CUDA-graphed 4-branch pipeline gives the same tensors as eager serial launches, and in
    select_on_device mode the crops / keypoints are those of the persons the association oracle picks."""
    from oracle import assoc as oassoc
    pipeline = spp.pipeline
    b, pf = 4, 6
    inp = pipeline.synthetic_inputs(b, 360, 480, pf, 17, seed=3)
    ms = synth.make_match_set(b * pf, 500, seed=9)
    inp.embeddings = ms.embeddings
    gal = ms.gallery.to(torch.bfloat16)
    a = pipeline.SelectivePosePipeline(inp, gal, dev, use_graph=True, concurrent=True)
    e = pipeline.SelectivePosePipeline(inp, gal, dev, use_graph=False, concurrent=False)
    for _ in range(3):
        a.step(); e.step()
    a.stream.synchronize(); e.stream.synchronize()
    for k in ("face_dets", "face_count", "person_dets", "person_count", "ids", "sims", "pixel_values", "keypoints", "scores", "argmax"):
        assert torch.equal(a.out[k], e.out[k]), k
    host = a.bind_host(inp) or a.run_host()
    a.stream.synchronize()
    assert torch.equal(host["keypoints"], a.out["keypoints"].cpu()) and torch.equal(host["ids"], a.out["ids"].cpu())
    # bench.py's configuration: bounded candidate list -> fused small-footprint detection kernels, heatmap decode first,
    # match GEMM on reserved SMs; and the round-1 order.  Same tensors, bit for bit.
    for kw in (dict(det_max_candidates=512), dict(det_max_candidates=512, match_sms=24), dict(det_max_candidates=512, det_fused=False),
               dict(heatmap_first=False), dict(det_max_candidates=512, det_after_heatmap=0), dict(det_max_candidates=512, det_after_heatmap=1),
               dict(det_max_candidates=512, crop_free_ctas=0)):
        v = pipeline.SelectivePosePipeline(inp, gal, dev, use_graph=True, concurrent=True, **kw)
        for _ in range(2):
            v.step()
        v.stream.synchronize()
        assert not bool(v.out["_face"].overflowed().any())
        for k in ("face_dets", "face_count", "person_dets", "person_count", "ids", "sims", "pixel_values", "keypoints", "scores", "argmax"):
            assert torch.equal(v.out[k], e.out[k]), (kw, k)
    # a bound that is too small is flagged, not silently truncated
    t = pipeline.SelectivePosePipeline(inp, gal, dev, use_graph=False, det_max_candidates=8)
    t.step(); t.stream.synchronize()
    assert bool(t.out["_face"].overflowed().any()) and bool((t.out["_face"].kept() <= 8).all())
    # crop boxes selected on the device
    s = pipeline.SelectivePosePipeline(inp, gal, dev, use_graph=True, select_on_device=True)
    s.step(); s.stream.synchronize()
    face_rows, person_rows = s.out["_face"].to_list(), s.out["_person"].to_list()
    ids = s.out["ids"].view(b, pf).cpu().numpy()
    fids = []
    for i in range(b):
        f = np.full(face_rows[i].shape[0], -1, np.int64)
        f[:min(pf, len(f))] = ids[i, :min(pf, len(f))]
        fids.append(f)
    rb, ri, rr = oassoc.associate([r.cpu().numpy() for r in face_rows], fids, [r.cpu().numpy() for r in person_rows], pf)
    assert s.out["sel_count"].tolist() == [len(x) for x in rr]
    sel = s.out["sel_boxes"].cpu().numpy()
    for i in range(b):
        np.testing.assert_array_equal(sel[i, :len(rr[i])], rb[i])
        np.testing.assert_array_equal(s.out["sel_ident"][i, :len(rr[i])].cpu().numpy(), ri[i])
    boxes = s.out["sel_boxes"].view(-1, 4)
    fidx = torch.arange(b, dtype=torch.int32, device=dev).repeat_interleave(pf)
    assert torch.equal(s.out["pixel_values"], spp.crop_affine(s.inp.frames, boxes, fidx))


# ------------------------------------------------------------------------------------------------
# multi-GPU: gallery sharded by rows, NCCL top-1 (value, index) reduction  (needs >= 2 GPUs)
# ------------------------------------------------------------------------------------------------

def _nccl_worker(rank, world, port, n, m, out):
    import importlib
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spp = importlib.import_module("person-recognition-for-pose-estimation_b200")
    ms = spp.synth.make_match_set(world * m, n, seed=17)
    gal = ms.gallery.to(torch.bfloat16)
    lo, hi = spp.dist.shard_bounds(n, world, rank)
    matcher = spp.dist.gpu_matcher(gal[lo:hi].contiguous().to(dev), lo, threshold=0.4)
    ids, sims = matcher.match(ms.embeddings[rank * m:(rank + 1) * m].to(dev))
    ref_ids, ref_sims = omatch.match_top1(ms.embeddings[rank * m:(rank + 1) * m], gal.float(), threshold=0.4)
    gap = omatch.top2_gap(ms.embeddings[rank * m:(rank + 1) * m], gal.float())
    ok = (gap > 1e-5) & ((ref_sims - 0.4).abs() > 1e-5)
    out[rank] = bool(torch.equal(ids.cpu()[ok], ref_ids[ok]) and torch.allclose(sims.cpu(), ref_sims, rtol=1e-3, atol=1e-5))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gallery_match_nccl(dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_nccl_worker, args=(world, port, 5000, 96, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


# ------------------------------------------------------------------------------------------------
# edge cases: empty and ragged inputs, other shapes, error behaviour
# ------------------------------------------------------------------------------------------------

def test_empty_inputs(spp, dev):
    kp, sc, am = spp.heatmap_decode(torch.zeros(0, 17, 64, 48, device=dev))
    assert kp.shape == (0, 17, 2) and sc.shape == (0, 17) and am.shape == (0, 17)
    out = spp.crop_affine(torch.zeros(1, 3, 32, 32, device=dev), torch.zeros(0, 4, device=dev), torch.zeros(0, dtype=torch.int32, device=dev))
    assert out.shape == (0, 3, 256, 192)
    ids, sims = spp.match_top1(torch.zeros(0, 512, device=dev), torch.zeros(4, 512, device=dev, dtype=torch.bfloat16))
    assert ids.shape == (0,) and sims.shape == (0,)
    dets = spp.non_max_suppression(torch.zeros(0, 5, 100, device=dev))
    assert dets == []
    # all-zero decoded tensor: no candidate survives conf > 0.001
    dets = spp.non_max_suppression(torch.zeros(3, 5, 100, device=dev))
    assert [d.shape for d in dets] == [(0, 6)] * 3


@pytest.mark.parametrize("k,h,w", [(133, 64, 48), (17, 96, 72), (3, 32, 24), (1, 16, 12)])
def test_heatmap_other_shapes(spp, synth, dev, k, h, w):
    hs = synth.make_heatmaps(3, k, h, w, seed=k + h, negative_frac=0.05, pairs=synth.wholebody_flip_pairs(k) if k > 17 else None)
    cs = synth.make_crop_set(1, 480, 640, per_frame=3, seed=1)
    avg = opose.flip_average(hs.heatmaps, hs.flipped, hs.perm)
    kp_o, sc_o, idx_o = opose.hf_dark_decode(avg.numpy(), cs.boxes.tolist())
    kp, sc, am = spp.heatmap_decode(hs.heatmaps.to(dev), hs.flipped.to(dev), hs.perm.to(dev), cs.boxes.to(dev), "dark", 11)
    np.testing.assert_array_equal(am.cpu().numpy(), idx_o)
    np.testing.assert_array_equal(sc.cpu().numpy(), sc_o)
    valid = sc_o > 0
    _close(kp.cpu().numpy()[valid], kp_o[valid], what=f"DARK keypoints K={k} {h}x{w}")
    c_o, s_o = opose.soft_argmax_decode(avg)
    c, s_, _ = spp.heatmap_decode(hs.heatmaps.to(dev), hs.flipped.to(dev), hs.perm.to(dev), None, "softargmax")
    _close(c.cpu().numpy(), c_o.numpy(), what="softargmax coords")
    _close(s_.cpu().numpy(), s_o.numpy(), what="softargmax scores")


@pytest.mark.parametrize("mode", ["dark", "softargmax", "quarter"])
@pytest.mark.parametrize("flip", [False, True])
def test_heatmap_bf16_maps(spp, synth, dev, mode, flip):
    """bf16 heatmaps: (a) identical, bit for bit, to the fp32 kernel on the widened maps; (b) against the
    ORACLE on the widened maps: arg-max / scores bit-exact, keypoints within 1e-3; (c) against the fp32 originals the
    rounding itself costs what tools/study_bf16_heatmaps.py measured (a few arg-max indices, <= ~0.1 px) — bounded here."""
    hs = synth.make_heatmaps(24, 17, seed=41, negative_frac=0.05)
    cs = synth.make_crop_set(2, 480, 640, per_frame=12, seed=3)
    hb, fb = hs.heatmaps.to(torch.bfloat16), hs.flipped.to(torch.bfloat16)
    boxes = cs.boxes.to(dev) if mode != "softargmax" else None
    args16 = (hb.to(dev), fb.to(dev) if flip else None, hs.perm.to(dev) if flip else None, boxes, mode, 11)
    args32 = (hb.float().to(dev), fb.float().to(dev) if flip else None, hs.perm.to(dev) if flip else None, boxes, mode, 11)
    kp16, sc16, am16 = spp.heatmap_decode(*args16)
    kp32, sc32, am32 = spp.heatmap_decode(*args32)
    assert torch.equal(am16, am32) and torch.equal(sc16, sc32) and torch.equal(kp16, kp32)
    if mode == "dark":
        avg = opose.flip_average(hb.float(), fb.float(), hs.perm) if flip else hb.float()
        kp_o, sc_o, idx_o = opose.hf_dark_decode(avg.numpy(), cs.boxes.tolist())
        np.testing.assert_array_equal(am16.cpu().numpy(), idx_o)
        np.testing.assert_array_equal(sc16.cpu().numpy(), sc_o)
        _close(kp16.cpu().numpy()[sc_o > 0], kp_o[sc_o > 0], what="DARK keypoints, bf16 maps vs oracle on the widened maps")
        # (c) vs the fp32 originals
        avg0 = opose.flip_average(hs.heatmaps, hs.flipped, hs.perm) if flip else hs.heatmaps
        kp0, sc0, idx0 = opose.hf_dark_decode(avg0.numpy(), cs.boxes.tolist())
        same = idx0 == idx_o
        assert same.mean() > 0.9
        good = same & (sc0 > 0)
        assert np.abs(kp16.cpu().numpy()[good] - kp0[good]).max() < 2.0        # image pixels; 0.1 heatmap px * ~4x crop scale
        np.testing.assert_allclose(sc16.cpu().numpy(), sc0, atol=5e-3)


def test_heatmap_ties_and_constant_maps(spp, dev):
    hm = torch.zeros(2, 3, 64, 48)
    hm[0, 0, 10, 5] = hm[0, 0, 10, 6] = hm[0, 0, 40, 1] = 2.0          # three equal maxima: the first one wins
    hm[0, 1] = 0.7                                                      # constant map: index 0
    hm[0, 2] = -1.0
    hm[1, 0, 63, 47] = 1.0                                              # last element
    hm[1, 1, 0, 0] = 1.0                                                # first element
    hm[1, 2] = float("-inf")                                            # np.argmax of all -inf is 0
    kp, sc, am = spp.heatmap_decode(hm.to(dev), mode="quarter")
    ref = hm.flatten(2).argmax(2)
    assert torch.equal(am.cpu().long(), ref)
    assert torch.equal(sc.cpu(), hm.flatten(2).max(2).values)


def test_nms_multilabel_many_classes(spp, dev):
    g = torch.Generator().manual_seed(2)
    b, nc, a = 2, 80, 600
    pred = torch.zeros(b, 4 + nc, a)
    pred[:, 0] = torch.rand(b, a, generator=g) * 600
    pred[:, 1] = torch.rand(b, a, generator=g) * 400
    pred[:, 2:4] = torch.rand(b, 2, a, generator=g) * 80 + 10
    pred[:, 4:] = torch.rand(b, nc, a, generator=g) ** 8                 # a few confident classes per anchor
    ref_rows, ref_keys = odet.non_max_suppression(pred, 0.25, 0.45, return_index=True)
    res = spp.nms_decoded(pred.to(dev), 0.25, 0.45)
    assert res.count.tolist() == [r.shape[0] for r in ref_rows]
    for rows, keys, rr, rk in zip(res.to_list(), res.keys_list(), ref_rows, ref_keys):
        np.testing.assert_array_equal(keys.cpu().numpy(), rk.numpy())
        np.testing.assert_array_equal(rows.cpu().numpy(), rr.numpy())


def test_nms_thousands_of_candidates(spp, dev):
    """More candidates than the shared-memory sort holds (8192): the workspace path."""
    g = torch.Generator().manual_seed(3)
    a = 12000
    pred = torch.zeros(1, 5, a)
    pred[0, 0] = torch.rand(a, generator=g) * 1200
    pred[0, 1] = torch.rand(a, generator=g) * 700
    pred[0, 2:4] = torch.rand(2, a, generator=g) * 60 + 8
    pred[0, 4] = torch.rand(a, generator=g) * 0.9 + 0.05
    ref_rows, ref_keys = odet.non_max_suppression(pred, 0.001, 0.65, return_index=True)
    res = spp.nms_decoded(pred.to(dev))
    np.testing.assert_array_equal(res.keys_list()[0].cpu().numpy(), ref_keys[0].numpy())
    np.testing.assert_array_equal(res.to_list()[0].cpu().numpy(), ref_rows[0].numpy())


def test_argument_errors(spp, dev):
    with pytest.raises((ValueError, RuntimeError)):
        spp.heatmap_decode(torch.zeros(1, 17, 64, 46, device=dev))          # width not a multiple of 4
    with pytest.raises((ValueError, RuntimeError)):
        spp.heatmap_decode(torch.zeros(1, 17, 64, 48, device=dev), kernel=10)
    with pytest.raises((ValueError, TypeError, RuntimeError)):
        spp.match_top1(torch.zeros(2, 256, device=dev), torch.zeros(4, 256, device=dev, dtype=torch.bfloat16))
    with pytest.raises(TypeError):
        spp.match_top1(torch.zeros(2, 512, device=dev), torch.zeros(4, 512, device=dev))   # gallery must be bf16
    with pytest.raises(ValueError):
        spp.non_max_suppression(torch.zeros(1, 4, 10, device=dev))           # no class channel
    with pytest.raises(ValueError):
        spp.crop_affine(torch.zeros(1, 1, 8, 8, device=dev), torch.zeros(1, 4, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))


def test_torch_ops_dispatch_to_the_kernels(spp, synth, dev):
    hs = synth.make_heatmaps(3, 17, seed=4)
    a = torch.ops.spp.heatmap_decode(hs.heatmaps.to(dev), hs.flipped.to(dev), hs.perm.to(dev), None, "dark", 11, 0)
    b = spp.heatmap_decode(hs.heatmaps.to(dev), hs.flipped.to(dev), hs.perm.to(dev), None, "dark", 11)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    hm = synth.make_head_maps(2, 128, 160, n_obj=3, seed=2)
    lv = [l.to(dev) for l in hm.levels]
    dets, cnt, keys = torch.ops.spp.decode_nms(lv, [8.0, 16.0, 32.0], 0.001, 0.65, 300)
    ref = spp.decode_nms(lv)
    assert torch.equal(dets, ref.dets) and torch.equal(cnt, ref.count) and torch.equal(keys, ref.keys)
    dec = torch.ops.spp.head_decode(lv, [8.0, 16.0, 32.0])
    d2, c2, k2 = torch.ops.spp.nms_decoded(dec, 0.001, 0.65, 300)
    assert torch.equal(c2, cnt) and torch.equal(k2, keys)


def test_gallery_save_load_shards(spp, synth, dev, tmp_path):
    ms = synth.make_match_set(32, 1000, seed=6)
    gal = spp.Gallery.from_rows(ms.gallery.to(dev))
    path = str(tmp_path / "gallery.bf16")
    gal.save(path)
    full = spp.Gallery.load(path, dev)
    assert torch.equal(full.rows, gal.rows)
    ids_full, sims_full = full.match(ms.embeddings.to(dev), threshold=0.4)
    keys = None
    for r in range(3):                                   # three shards, reduced with an integer max
        sh = spp.Gallery.load(path, dev, world=3, rank=r)
        k = spp.match_top1(ms.embeddings.to(dev), sh.rows, None, sh.id_offset, want_keys=True)[2]
        keys = k if keys is None else torch.maximum(keys, k)
    ids, sims = spp.match_unpack_keys(keys, 0.4)
    assert torch.equal(ids.long(), ids_full) and torch.equal(sims, sims_full)


# ------------------------------------------------------------------------------------------------
# COCO result rows + OKS (a15 / 8f-3)
# ------------------------------------------------------------------------------------------------

def test_coco_results_vs_reference_rows(spp, golden, dev):
    """heatmaps -> flip test (no channel swap: the reference's batch-of-1 behaviour) -> soft-argmax -> result rows,
    against the rows the reference's own validation_step emitted."""
    g = golden("pose_results.npz")
    hm, fl = torch.from_numpy(g["hm"]).to(dev), torch.from_numpy(g["flipped"]).to(dev)
    boxes = torch.from_numpy(g["boxes"]).to(dev)
    kp, sc, _ = spp.heatmap_decode(hm, fl, None, boxes[:, 0].contiguous(), "softargmax", flags=spp.ops.FLAG_SCALE_SCORE)
    _close(kp.cpu().numpy(), g["coords"], atol=1e-6, what="normalised coordinates")
    _close(sc.cpu().numpy(), g["scores"], what="scores")
    res = spp.coco_keypoint_results(kp, sc, boxes, torch.from_numpy(g["areas"]), torch.from_numpy(g["masks"]),
                                    torch.from_numpy(g["is_crowd"]), g["image_ids"].tolist())
    assert [r["image_id"] for r in res] == g["res_image_id"].tolist()
    got = np.array([r["keypoints"] for r in res]).reshape(len(res), -1, 3)
    want = g["res_keypoints"].reshape(len(res), -1, 3)
    _close(got[..., :2], want[..., :2], what="result keypoints")
    clear = np.abs(g["scores"][[0, 0, 1, 1, 2]] - 0.3) > 1e-3          # visibility flags away from the threshold: exact
    assert np.array_equal(got[..., 2][clear], want[..., 2][clear])
    _close([r["score"] for r in res], g["res_score"], what="instance score")
    assert np.array_equal(np.array([r["bbox"] for r in res]), g["res_bbox"])
    np.testing.assert_allclose([r["area"] for r in res], g["res_area"], rtol=0)
    # the kernel on the reference's own decoded coordinates: bit-exact rows
    res2 = spp.coco_keypoint_results(torch.from_numpy(g["coords"]).to(dev), torch.from_numpy(g["scores"]).to(dev), boxes,
                                     torch.from_numpy(g["areas"]), torch.from_numpy(g["masks"]), torch.from_numpy(g["is_crowd"]),
                                     g["image_ids"].tolist())
    assert np.array_equal(np.array([r["keypoints"] for r in res2]), g["res_keypoints"])
    np.testing.assert_allclose([r["score"] for r in res2], g["res_score"], rtol=1e-6)
    assert spp.coco_keypoint_results(kp, sc, boxes, torch.from_numpy(g["areas"]), torch.zeros(3, 3, dtype=torch.bool),
                                     torch.from_numpy(g["is_crowd"]), g["image_ids"].tolist()) == []


def test_pose_results_and_oks_vs_oracle(spp, dev):
    from oracle import results as ores
    g = torch.Generator().manual_seed(9)
    p, k = 300, 17
    gt = torch.rand(p, k, 3, generator=g) * 200
    gt[..., 2] = torch.randint(0, 3, (p, k), generator=g).float()
    gt[:7, :, 2] = 0                                                     # pairs without a labelled joint
    pred = gt[..., :2] + torch.randn(p, k, 2, generator=g) * 6
    area = torch.rand(p, generator=g) * 20000 + 500
    gbox = torch.cat([torch.rand(p, 2, generator=g) * 100, torch.rand(p, 2, generator=g) * 150 + 10], 1)
    sig = torch.tensor(ores.COCO_SIGMAS)
    oks = spp.pose_oks(pred.to(dev), gt.to(dev), area.to(dev), sig.to(dev), gbox.to(dev)).cpu().numpy()
    want = [ores.compute_oks(pred[i].numpy(), gt[i].numpy(), float(area[i]), ores.COCO_SIGMAS, gbox[i].numpy()) for i in range(p)]
    _close(oks, want, rtol=1e-5, atol=1e-7, what="OKS")
    sc = torch.rand(p, k, generator=g)
    rows, inst = spp.pose_results(pred.to(dev), sc.to(dev))            # already in image pixels: no boxes
    assert torch.equal(rows[..., :2].cpu(), pred)
    assert torch.equal(rows[..., 2].cpu(), torch.where(sc > 0.3, 2.0, 1.0))
    _close(inst.cpu().numpy(), sc.mean(1).numpy(), rtol=1e-6, what="instance score")
    rows3 = torch.cat([pred, sc[..., None]], -1).to(dev)                # [P, K, 3] predictions
    assert np.allclose(spp.pose_oks(rows3, gt.to(dev), area.to(dev), sig.to(dev), gbox.to(dev)).cpu().numpy(), oks)
    e_rows, e_inst = spp.pose_results(torch.zeros(0, k, 2, device=dev), torch.zeros(0, k, device=dev))
    assert e_rows.shape == (0, k, 3) and e_inst.shape == (0,)


# ------------------------------------------------------------------------------------------------
# detection evaluation on the device (8f-3): compute_metric / compute_ap
# ------------------------------------------------------------------------------------------------

def test_det_metrics_vs_reference_fixture(spp, golden, dev):
    """True-positive matrix bit-exact and AP summary to 1e-12 against the reference's own compute_metric / compute_ap
    (tests/golden/det_metrics.npz), driven as training/yolopt/main.py:210-234 drives them."""
    g = golden("det_metrics.npz")
    dets, dcount = torch.from_numpy(g["dets"]).to(dev), torch.from_numpy(g["dcount"]).to(dev)
    targets, tcount = torch.from_numpy(g["targets"]).to(dev), torch.from_numpy(g["tcount"]).to(dev)
    iou_v = torch.from_numpy(g["iou_v"])
    correct = spp.det_match_targets(dets, dcount, targets, tcount, iou_v.tolist())
    np.testing.assert_array_equal(correct.cpu().numpy(), g["correct"])
    # the reference's single-image signature
    for b in (0, 1, 3, 5):
        n, m = int(g["dcount"][b]), int(g["tcount"][b])
        c = spp.compute_metric(dets[b, :n], targets[b, :m], iou_v.to(dev))
        np.testing.assert_array_equal(c.cpu().numpy(), g["correct"][b, :n])
    res = spp.det_average_precision(torch.from_numpy(g["cat_tp"]).to(dev), torch.from_numpy(g["cat_conf"]).to(dev),
                                    torch.from_numpy(g["cat_cls"]).to(dev), torch.from_numpy(g["cat_target_cls"]).to(dev))
    np.testing.assert_array_equal(res["tp"].cpu().numpy(), g["tp"])
    np.testing.assert_array_equal(res["fp"].cpu().numpy(), g["fp"])
    got = np.array([res["m_pre"], res["m_rec"], res["map50"], res["mean_ap"]])
    np.testing.assert_allclose(got, g["summary"], rtol=1e-12, atol=0)
    tp, fp, m_pre, m_rec, map50, mean_ap = spp.compute_ap(g["cat_tp"], g["cat_conf"], g["cat_cls"], g["cat_target_cls"])
    np.testing.assert_allclose([m_pre, m_rec, map50, mean_ap], g["summary"], rtol=1e-12)


def test_det_metrics_vs_oracle_larger(spp, dev):
    """More detections than one sort block (2 048) and 7 classes, two of them without any detection: per-class AP matrix,
    curves' operating point and summary against the numpy restatement."""
    from oracle import detmetrics as dm
    g = torch.Generator().manual_seed(3)
    n, nt, nc = 9000, 700, 7
    conf = torch.rand(n, generator=g) * 0.98 + 0.01
    cls = torch.randint(0, 5, (n,), generator=g).float()                 # classes 5, 6 have labels but no detections
    tcls = torch.randint(0, nc, (nt,), generator=g).float()
    tp = (torch.rand(n, 10, generator=g) < (conf[:, None] * torch.linspace(0.9, 0.2, 10)[None])).cummin(1).values   # nested in the threshold
    want = dm.compute_ap(tp.numpy(), conf.numpy(), cls.numpy(), tcls.numpy())
    res = spp.det_average_precision(tp.to(dev), conf.to(dev), cls.to(dev), tcls.to(dev), nc_max=16)
    ex = want[6]
    np.testing.assert_array_equal(res["classes"].cpu().numpy(), ex["classes"].astype(np.int32))
    np.testing.assert_allclose(res["ap"].cpu().numpy(), ex["ap"], rtol=1e-12, atol=1e-15)
    assert res["index"] == ex["index"]
    np.testing.assert_array_equal(res["tp"].cpu().numpy(), want[0])
    np.testing.assert_array_equal(res["fp"].cpu().numpy(), want[1])
    np.testing.assert_allclose([res["m_pre"], res["m_rec"], res["map50"], res["mean_ap"]], list(want[2:6]), rtol=1e-12)
    # no detections at all: every AP is zero
    z = spp.det_average_precision(torch.zeros(0, 10, dtype=torch.bool, device=dev), torch.zeros(0, device=dev), torch.zeros(0, device=dev),
                                  tcls.to(dev), nc_max=16)
    assert float(z["ap"].abs().sum()) == 0.0 and z["classes"].numel() == nc


# ------------------------------------------------------------------------------------------------
# detection head consumed before the per-level cat (8f-1)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag", ["nc1", "nc3"])
def test_split_head_inputs_match_concatenated(spp, golden, dev, tag):
    g = golden(f"det_{tag}.npz")
    cat = [torch.from_numpy(g[k]).to(dev) for k in ("l0", "l1", "l2")]
    pairs = [(l[:, :64].contiguous(), l[:, 64:].contiguous()) for l in cat]
    conf = float(g["conf"])
    assert torch.equal(spp.head_decode(pairs), spp.head_decode(cat))
    _close(spp.head_decode(pairs).cpu().numpy(), g["decoded"], atol=1e-4, what="decoded head (split inputs) vs reference")
    a, b = spp.decode_nms(pairs, conf_thres=conf), spp.decode_nms(cat, conf_thres=conf)
    assert torch.equal(a.count, b.count) and torch.equal(a.keys, b.keys) and torch.equal(a.dets, b.dets)
    assert a.kept().tolist() == g["n"].tolist()
    with pytest.raises(ValueError):
        spp.decode_nms([pairs[0], cat[1], cat[2]])


def test_head_eval_forward_with_torch_convs(spp, dev):
    """A Head-shaped module (box / cls ModuleLists + stride, as nn.py:228-253): the shim runs its convs and
    replaces everything after them; compared with the plain torch restatement of nn.py:255-270."""
    torch.manual_seed(0)

    class TinyHead(torch.nn.Module):
        def __init__(self, nc=2, filters=(8, 16, 24)):
            super().__init__()
            self.box = torch.nn.ModuleList(torch.nn.Conv2d(f, 64, 1) for f in filters)
            self.cls = torch.nn.ModuleList(torch.nn.Conv2d(f, nc, 1) for f in filters)
            self.stride = torch.tensor([8.0, 16.0, 32.0])

    head = TinyHead().to(dev).eval()
    feats = [torch.randn(2, f, s, s + 4, device=dev) for f, s in ((8, 16), (16, 8), (24, 4))]
    with torch.no_grad():
        out = spp.head_eval_forward(head, feats)
        cat = [torch.cat((b(x), c(x)), 1) for b, c, x in zip(head.box, head.cls, feats)]
    ref = odet.head_decode([l.cpu() for l in cat])
    _close(out.cpu().numpy(), ref.numpy(), atol=1e-4, what="head_eval_forward")


def test_reference_head_module_drop_in(spp, golden, dev):
    """The reference's REAL detection head (state_dict saved from training/yolopt/nets/nn.py ``Head`` by
    oracle/gen_golden.py) loaded into a same-layout module: ``spp.head_eval_forward(head, feats)`` replaces
    ``head(feats)`` in eval mode (nn.py:255-270), ``spp.detect`` replaces ``non_max_suppression(head(feats))``."""
    from oracle.refhead import RefShapedHead
    g = golden("ref_head.npz")
    head = RefShapedHead(int(g["nc"]), tuple(int(x) for x in g["filters"]))
    head.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    head = head.to(dev).eval()
    head.stride = head.stride.to(dev)
    feats = [torch.from_numpy(g[k]).to(dev) for k in ("f0", "f1", "f2")]
    conf, iou = float(g["conf"]), float(g["iou"])
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the fixture is fp32 CPU convolution; TF32 would cost 1e-3 on the logits
    try:
        with torch.no_grad():
            out = spp.head_eval_forward(head, feats)
            pairs = spp.head_conv_outputs(head, feats)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert out.shape == g["decoded"].shape
    _close(out.cpu().numpy(), g["decoded"], atol=2e-3, what="head_eval_forward on the reference head vs the reference's eval output")
    for (bx, cl), k in zip(pairs, ("l0", "l1", "l2")):
        _close(torch.cat((bx, cl), 1).cpu().numpy(), g[k], rtol=1e-4, atol=1e-4, what="conv stacks of the loaded head")
    # post-processing on the reference's own conv outputs: same rows as the reference's non_max_suppression(head(x))
    lv = [torch.from_numpy(g[k]).to(dev) for k in ("l0", "l1", "l2")]
    dets = spp.detect(lv, conf, iou)
    assert [d.shape[0] for d in dets] == g["n"].tolist()
    _close(torch.cat(dets).cpu().numpy(), g["dets"], atol=1e-3, what="detect() rows vs the reference's NMS rows")
    assert np.array_equal(torch.cat(dets).cpu().numpy()[:, 5], g["dets"][:, 5])
    dets2 = spp.non_max_suppression(torch.from_numpy(g["decoded"]).to(dev), conf, iou)
    np.testing.assert_array_equal(torch.cat(dets2).cpu().numpy(), g["dets"])          # from the decoded tensor: bit-exact
    # end to end through the module on the GPU (cuDNN convolutions): same number of detections per image up to score ties
    dets3 = spp.detect(pairs, conf, iou)
    assert all(abs(d.shape[0] - n) <= max(3, n // 20) for d, n in zip(dets3, g["n"].tolist()))


@pytest.mark.parametrize("tag", ["nc1", "nc3"])
def test_fused_single_kernel_decode_nms_is_identical(spp, synth, golden, dev, tag):
    """spp_decode_nms_mode(1): one kernel per head (scan + candidate decode + sort + NMS in one CTA per image)."""
    g = golden(f"det_{tag}.npz")
    cat = [torch.from_numpy(g[k]).to(dev) for k in ("l0", "l1", "l2")]
    big = synth.make_head_maps(3, 736, 1280, n_obj=10, nc=1, seed=4)
    try:
        prev = spp.ops.set_decode_nms_mode("split")
        a = spp.decode_nms(cat, conf_thres=float(g["conf"]))
        a2 = spp.decode_nms([l.to(dev) for l in big.levels])
        assert spp.ops.set_decode_nms_mode("fused") == "split"
        b = spp.decode_nms(cat, conf_thres=float(g["conf"]))
        b2 = spp.decode_nms([l.to(dev) for l in big.levels])
    finally:
        spp.ops.set_decode_nms_mode(prev)
    for x, y in ((a, b), (a2, b2)):
        assert torch.equal(x.count, y.count) and torch.equal(x.keys, y.keys) and torch.equal(x.dets, y.dets)
    assert b.kept().tolist() == g["n"].tolist()


# ------------------------------------------------------------------------------------------------
# the benchmark workload itself (BASELINE.json configs[1], full size) against the reference's CPU calls
# ------------------------------------------------------------------------------------------------

def test_cfg2_full_size_against_cpu_reference_calls(spp, synth, dev):
    """cfg2 exactly as bench.py runs it (64 frames 1280x720, 640 faces / persons / crops, 10k-id gallery, 17 joints with
    flip test), CUDA-graphed pipeline vs the CPU path bench.py times as the baseline: torch Head decode + torchvision
    nms, F.normalize / F.linear / max, HF VitPoseImageProcessor.preprocess and post_process_pose_estimation."""
    import bench
    pipeline = spp.pipeline
    b, pf = 64, 10
    inp = pipeline.synthetic_inputs(b, 720, 1280, pf, 17, seed=0)
    ms = synth.make_match_set(b * pf, 10000, seed=1000)
    inp.embeddings = ms.embeddings
    pipe = pipeline.SelectivePosePipeline(inp, ms.gallery.to(torch.bfloat16), dev)
    pipe.step()
    pipe.stream.synchronize()
    ref = bench.cpu_reference_step(inp, ms.gallery, b, pf)
    # detections: same number of rows per frame, same rows (boxes / scores within 1e-3, conf-descending order)
    skipped = {}
    for name in ("face", "person"):
        rows = pipe.out["_" + name].to_list()
        checked = 0
        for got, want in zip(rows, ref[name]):
            if odet.near_threshold_pairs(want, 0.65):      # an IoU within rounding of the threshold may legitimately flip
                continue
            checked += 1
            assert got.shape == want.shape, name
            _close(got.cpu().numpy(), want.numpy(), atol=1e-3, what=f"{name} detections")
        skipped[name] = b - checked
        assert checked >= 0.9 * b, f"{name}: only {checked} of {b} frames free of borderline IoU pairs"
    # how many frames the raw-path comparison skipped (an IoU within 1e-6 of the threshold somewhere in
    # the frame); shown with `pytest -rP`, recorded in DESIGN.md section 4
    print(f"cfg2 raw-path detection parity: frames skipped for a borderline IoU pair: face {skipped['face']} / {b}, person {skipped['person']} / {b}")
    # identities: exact wherever the fp32 top-2 gap and the gate margin are not degenerate
    gap = omatch.top2_gap(ms.embeddings, ms.gallery.to(torch.bfloat16).float())
    ref_ids, ref_sims = omatch.match_top1(ms.embeddings, ms.gallery.to(torch.bfloat16).float(), threshold=0.4)
    ok = (gap > 1e-5) & ((ref_sims - 0.4).abs() > 1e-5)
    assert torch.equal(pipe.out["ids"].cpu()[ok].long(), ref_ids[ok].long())
    _close(pipe.out["sims"].cpu().numpy(), ref_sims.numpy(), atol=1e-5, what="similarities")
    # crops against HF's own scipy warp
    diff = (pipe.out["pixel_values"].cpu() - ref["pixel_values"]).abs().max()
    assert float(diff) < 2e-5, f"crop max abs diff {float(diff):.3e}"
    # poses against HF's own post-processing.  HF addresses its DARK taps through a float32 flat index (quirk Q6):
    # exact for the first 5 084 maps of a call = 299 crops here.  (a) HF called in chunks of 25 frames (250 crops) is
    # the intended computation -> the default kernel; (b) HF called once on all 640 crops -> the quirk flag.
    from transformers import VitPoseImageProcessor
    from transformers.models.vitpose.modeling_vitpose import VitPoseEstimatorOutput
    proc = VitPoseImageProcessor()
    avg = opose.flip_average(inp.heatmaps, inp.flipped, inp.perm)
    boxes = [[[float(v) for v in inp.boxes[f * pf + j]] for j in range(pf)] for f in range(b)]
    chunks = []
    for f0 in range(0, b, 25):
        f1 = min(b, f0 + 25)
        chunks += proc.post_process_pose_estimation(VitPoseEstimatorOutput(heatmaps=avg[f0 * pf:f1 * pf]), boxes=boxes[f0:f1], kernel_size=11)
    kp = pipe.out["keypoints"].cpu().numpy()
    sc = pipe.out["scores"].cpu().numpy()
    want_kp = np.stack([p["keypoints"].numpy() for img in chunks for p in img])
    want_sc = np.stack([p["scores"].numpy() for img in chunks for p in img])
    valid = want_sc > 0
    np.testing.assert_array_equal(sc, want_sc)                                   # max of the averaged map: bit-exact
    _close(kp[valid], want_kp[valid], what="keypoints (image pixels) vs HF in calls of <= 299 crops")
    once_kp = np.stack([p["keypoints"].numpy() for img in ref["poses"] for p in img])
    assert float(np.abs(once_kp - want_kp)[299:].max()) > 1.0, "HF's float32 index quirk should be visible from crop 299 on"
    np.testing.assert_array_equal(once_kp[:299], want_kp[:299])
    qk, qs, _ = spp.heatmap_decode(pipe.inp.heatmaps, pipe.inp.flipped, pipe.inp.perm, pipe.inp.boxes, "dark", 11,
                                   flags=spp.ops.FLAG_HF_F32_INDEX)
    _close(qk.cpu().numpy()[valid], once_kp[valid], what="keypoints with the HF float32-index quirk vs HF in ONE call of 640 crops")
    # (score <= 0 joints: both modes reproduce HF's flat-index reads there as well)
    _close(qk.cpu().numpy()[~valid], once_kp[~valid], atol=1e-2, what="score <= 0 joints, quirk mode")


def test_crop_uint8_720p_against_hf_preprocess(spp, dev, crop_path):
    """uint8 1280x720 frames (HF's default do_rescale=True input) through the real HF VitPoseImageProcessor.preprocess:
    every pixel of 160 crops, i.e. every uint8 rounding decision of the fast path / fp64 re-computation split."""
    from transformers import VitPoseImageProcessor
    pipeline = spp.pipeline
    b, pf = 16, 10
    inp = pipeline.synthetic_inputs(b, 720, 1280, pf, 17, seed=5)
    fr8 = (inp.frames * 255.0).round().clamp(0, 255).to(torch.uint8)
    boxes = [[[float(v) for v in inp.boxes[f * pf + j]] for j in range(pf)] for f in range(b)]
    want = VitPoseImageProcessor().preprocess([fr8[f] for f in range(b)], boxes=boxes, return_tensors="pt")["pixel_values"]
    got = spp.VitPoseImageProcessor().preprocess(fr8.to(dev), boxes)["pixel_values"].cpu()
    diff = (got - want).abs()
    assert float(diff.max()) < 2e-5, f"{int((diff > 2e-5).sum())} of {diff.numel()} pixels differ (one flipped rounding = 0.017)"


def test_crop_extreme_boxes_vs_oracle(spp, dev, crop_path):
    """Boxes far from the synthetic distribution: down-scales of 3-7x (one source band per output row, source windows as wide
    as the frame), 20x up-scales (many output rows per source row), boxes mostly or entirely outside the frame."""
    g = torch.Generator().manual_seed(21)
    frames = torch.rand(2, 3, 720, 1280, generator=g)
    boxes = [[0.0, 0.0, 1280.0, 720.0], [100.0, 10.0, 1100.0, 700.0], [-300.0, -200.0, 1900.0, 1100.0], [10.0, 300.0, 1260.0, 60.0],
             [640.0, 360.0, 8.0, 8.0], [5.5, 7.25, 3.0, 14.0], [1270.0, 710.0, 40.0, 60.0], [-500.0, -500.0, 100.0, 100.0],
             [1279.0, 0.0, 2.0, 719.0], [300.0, -50.0, 20.0, 900.0]]
    fidx = [0, 1, 0, 1, 0, 1, 0, 1, 0, 1]
    ref = ocrop.crop_affine_hf(frames.numpy(), boxes, fidx)
    out = spp.crop_affine(frames.to(dev), torch.tensor(boxes, device=dev), torch.tensor(fidx, dtype=torch.int32, device=dev))
    diff = np.abs(out.cpu().numpy() - ref)
    assert float(diff.max()) < 2e-5, f"max abs diff {float(diff.max()):.3e} at crop {int(np.argmax(diff.reshape(len(boxes), -1).max(1)))}"
    fr8 = (frames * 255.0).round().clamp(0, 255).to(torch.uint8)
    m, s = ocrop.fused_mean_std(rescale_factor=1 / 255)
    ref8 = ocrop.crop_affine_hf(fr8.numpy(), boxes, fidx, rescale_factor=1 / 255)
    out8 = spp.crop_affine(fr8.to(dev), torch.tensor(boxes, device=dev), torch.tensor(fidx, dtype=torch.int32, device=dev),
                           mean=m.tolist(), std=s.tolist())
    diff8 = np.abs(out8.cpu().numpy() - ref8)
    assert float(diff8.max()) < 2e-5, f"uint8: {(diff8 > 2e-5).sum()} pixels differ"
    # gluoncv-style variant on the same boxes
    refb = ocrop.crop_affine_v2(frames.numpy(), boxes, fidx)
    outb = spp.crop_affine(frames.to(dev), torch.tensor(boxes, device=dev), torch.tensor(fidx, dtype=torch.int32, device=dev), variant="gluoncv")
    assert float(np.abs(outb.cpu().numpy() - refb).max()) < 2e-5


@pytest.mark.gpu
def test_crop_persistent_equals_per_item_and_planned(spp, synth, dev):
    """The persistent plan + stream kernels, the same as two calls (crop_plan, then crop_affine(planned=True)) and the
    one-CTA-per-item kernels produce the same bits — fp32 and uint8 frames, boxes partly and wholly outside the frame, a
    mirrored box, more items than resident CTAs (tickets), an odd crop count."""
    cs = synth.make_crop_set(24, 360, 640, per_frame=11, seed=5)
    boxes = cs.boxes.clone()
    boxes[3] = torch.tensor([-500.0, -500.0, 40.0, 60.0])          # wholly outside: constant crop
    boxes[7] = torch.tensor([600.0, 300.0, 200.0, 150.0])          # hangs over the corner
    boxes[9] = torch.tensor([300.0, 200.0, -80.0, -120.0])         # mirrored
    fr, bx, fi = cs.frames.to(dev), boxes.to(dev), cs.frame_idx.to(dev)
    fr8 = (cs.frames * 255).round().to(torch.uint8).to(dev)
    for frames, kw in ((fr, {}), (fr8, {"mean": [123.7, 116.3, 103.5], "std": [58.4, 57.1, 57.4]})):
        spp.ops.CROP_USE_WORKSPACE = False
        try:
            ref = spp.crop_affine(frames, bx, fi, **kw)
        finally:
            spp.ops.CROP_USE_WORKSPACE = True
        prev_policy = spp._lib.lib().spp_crop_policy(2)          # the persistent kernels for uint8 frames too
        got = spp.crop_affine(frames, bx, fi, **kw)
        assert torch.equal(got, ref)
        ws = spp.ops.alloc_workspace(dev, spp.crop_workspace_bytes(bx.shape[0], 256, 192, frames.dtype == torch.uint8))
        spp.ops.crop_plan(bx, fi, frames.shape, frames.dtype == torch.uint8, ws)
        got2 = spp.crop_affine(frames, bx, fi, workspace=ws, planned=True, **kw)
        assert torch.equal(got2, ref)
        # bf16 output = the fp32 result rounded to nearest even, on both implementations
        gotb = spp.crop_affine(frames, bx, fi, out_dtype=torch.bfloat16, **kw)
        assert gotb.dtype == torch.bfloat16 and torch.equal(gotb, ref.to(torch.bfloat16))
        spp.ops.CROP_USE_WORKSPACE = False
        try:
            gotb2 = spp.crop_affine(frames, bx, fi, out_dtype=torch.bfloat16, **kw)
        finally:
            spp.ops.CROP_USE_WORKSPACE = True
        assert torch.equal(gotb2, gotb)
        # the fixed-signature entry points of include/spp.h (fp32 output) through ctypes
        import ctypes
        L = spp._lib.lib()
        u8 = frames.dtype == torch.uint8
        m3 = (ctypes.c_float * 3)(*[float(v) for v in kw.get("mean", (0.485, 0.456, 0.406))])
        s3 = (ctypes.c_float * 3)(*[float(v) for v in kw.get("std", (0.229, 0.224, 0.225))])
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        head = (vp(frames), frames.shape[0], frames.shape[2], frames.shape[3], vp(bx), vp(fi), bx.shape[0], 256, 192, m3, s3, 0)
        for name, tail in (("spp_crop_affine", ()), ("spp_crop_affine_ws", (vp(ws), ws.numel())), ("spp_crop_affine_run", (vp(ws), ws.numel()))):
            fn = getattr(L, name.replace("affine", "affine_u8") if u8 else name)
            o = torch.full_like(ref, float("nan"))
            if name.endswith("_run"):
                assert L.spp_crop_plan(1 if u8 else 0, frames.shape[0], frames.shape[2], frames.shape[3], vp(bx), vp(fi), bx.shape[0], 256, 192, 0,
                                       vp(ws), ws.numel(), st) == 0
            assert fn(*head, vp(o), *tail, st) == 0, spp._lib.lib().spp_last_error()
            assert torch.equal(o, ref), name
        spp._lib.lib().spp_crop_policy(prev_policy)


@pytest.mark.gpu
@pytest.mark.parametrize("out_hw", [(128, 96), (384, 288), (64, 48), (100, 75), (256, 200), (96, 640), (33, 17)])
def test_crop_output_sizes_vs_oracle(spp, synth, dev, crop_path, out_hw):
    """Output sizes other than ViTPose-B's 256x192 (the HF processor's `size` is configurable): one / three column chunks per
    warp row, widths that are not a multiple of the 96 columns a warp covers, odd sizes (8-byte table entries cannot be copied
    by bulk TMA: the persistent path hands over to the per-item kernel), a width beyond six column chunks (staging-free
    kernel) — fp32 and uint8 frames, on a frame whose rows are 16-byte aligned and on one whose rows are not."""
    from oracle import crop as ocrop
    for fw in (320, 322):
        cs = synth.make_crop_set(3, 200, fw, per_frame=4, seed=11 + fw)
        boxes = cs.boxes.clone()
        boxes[1] = torch.tensor([-30.0, 150.0, 120.0, 90.0])       # hangs over two edges
        want = ocrop.crop_affine_hf(cs.frames.numpy(), boxes.tolist(), cs.frame_idx.tolist(), out_hw=out_hw)
        got = spp.crop_affine(cs.frames.to(dev), boxes.to(dev), cs.frame_idx.to(dev), out_hw=out_hw)
        assert tuple(got.shape) == (boxes.shape[0], 3) + tuple(out_hw)
        assert float(np.abs(got.cpu().numpy() - want).max()) < 2e-5, (out_hw, fw)
        fr8 = (cs.frames * 255).round().to(torch.uint8)
        m = np.array([0.485, 0.456, 0.406], np.float32) * 255
        sd = np.array([0.229, 0.224, 0.225], np.float32) * 255
        want8 = ocrop.crop_affine_hf(fr8.numpy(), boxes.tolist(), cs.frame_idx.tolist(), out_hw=out_hw, mean=m, std=sd)
        got8 = spp.crop_affine(fr8.to(dev), boxes.to(dev), cs.frame_idx.to(dev), out_hw=out_hw, mean=m.tolist(), std=sd.tolist())
        d8 = np.abs(got8.cpu().numpy() - want8)
        assert float(d8.max()) < 2e-5, (out_hw, fw, float(d8.max()))      # a flipped uint8 rounding would show up as ~0.017


@pytest.mark.gpu
def test_crop_fuzz_both_implementations_agree(spp, dev):
    """Random and degenerate boxes (NaN, infinities, zero and negative sizes, boxes far outside the frame, sub-pixel and
    frame-sized boxes, out-of-range frame indices) through the persistent and the per-item kernels: both finish, agree bit for
    bit, and produce finite pixels — fp32 and uint8 frames, fp32 and bf16 output."""
    g = torch.Generator().manual_seed(123)
    fr = torch.rand(5, 3, 96, 160, generator=g)
    n = 400
    boxes = torch.empty(n, 4)
    boxes[:, 0] = torch.empty(n).uniform_(-200, 300, generator=g)
    boxes[:, 1] = torch.empty(n).uniform_(-150, 200, generator=g)
    boxes[:, 2] = torch.empty(n).uniform_(-50, 400, generator=g)
    boxes[:, 3] = torch.empty(n).uniform_(-50, 300, generator=g)
    special = [float("nan"), float("inf"), -float("inf"), 0.0, 1e-3, 1e6, -1e6, 1e30]
    for i in range(0, 120):
        boxes[i, int(torch.randint(0, 4, (1,), generator=g))] = special[i % len(special)]
    boxes[120] = torch.tensor([0.0, 0.0, 160.0, 96.0])
    boxes[121] = torch.tensor([10.5, 20.25, 0.5, 0.5])
    fidx = torch.randint(-2, 8, (n,), generator=g, dtype=torch.int32)          # clamped by the kernels, as the shim documents
    L = spp._lib.lib()
    for frames, kw in ((fr.to(dev), {}), ((fr * 255).round().to(torch.uint8).to(dev), {"mean": [120.0, 115.0, 100.0], "std": [58.0, 57.0, 57.0]})):
        for od in (torch.float32, torch.bfloat16):
            prev = L.spp_crop_policy(0)
            try:
                a = spp.crop_affine(frames, boxes.to(dev), fidx.to(dev), out_dtype=od, **kw)
                L.spp_crop_policy(2)
                b = spp.crop_affine(frames, boxes.to(dev), fidx.to(dev), out_dtype=od, **kw)
            finally:
                L.spp_crop_policy(prev)
            torch.cuda.synchronize()
            assert torch.equal(a, b), (frames.dtype, od)
            assert bool(torch.isfinite(a.float()).all())


@pytest.mark.gpu
def test_kernels_survive_nonfinite_inputs(spp, synth, dev):
    """NaN / +-inf in the backbone outputs (a diverged model) must not hang or fault a kernel: every op completes, counts stay
    in range, and the finite part of the batch is unaffected where the op is per-item."""
    g = torch.Generator().manual_seed(7)
    bad = torch.tensor([float("nan"), float("inf"), -float("inf")])
    # detection head: poison frame 1 of 3, frames 0 and 2 must decode as before
    hm = synth.make_head_maps(3, 320, 320, n_obj=4, nc=1, seed=3)
    clean = spp.decode_nms([l.to(dev) for l in hm.levels])
    lv = [l.clone() for l in hm.levels]
    for l in lv:
        idx = torch.randint(0, l[1].numel(), (l[1].numel() // 7,), generator=g)
        l[1].view(-1)[idx] = bad[torch.randint(0, 3, (idx.numel(),), generator=g)]
    for mc in (0, 512):
        res = spp.decode_nms([l.to(dev) for l in lv], max_candidates=mc)
        torch.cuda.synchronize()
        cnt = res.count.cpu()
        assert int(cnt[0]) == int(clean.count[0]) and int(cnt[2]) == int(clean.count[2])
        assert torch.equal(res.dets[0], clean.dets[0]) and torch.equal(res.dets[2], clean.dets[2])
        assert -301 <= int(cnt[1]) <= 300
    # heatmaps: poison some maps, the others decode as before
    hs = synth.make_heatmaps(6, 17, seed=21)
    boxes = torch.tensor([[10.0, 20.0, 100.0, 200.0]]).repeat(6, 1).to(dev)
    ref = spp.ops.heatmap_decode(hs.heatmaps.to(dev), hs.flipped.to(dev), hs.perm.to(dev), boxes, "dark", 11, 0)
    h2 = hs.heatmaps.clone()
    h2[2, 3].view(-1)[::5] = float("nan")
    h2[4, 0] = float("inf")
    h2[4, 1] = -float("inf")
    for mode in ("dark", "softargmax", "quarter"):
        out = spp.ops.heatmap_decode(h2.to(dev), hs.flipped.to(dev), hs.perm.to(dev), boxes, mode, 11, 0)
        torch.cuda.synchronize()
        assert out[0].shape == (6, 17, 2)
    out = spp.ops.heatmap_decode(h2.to(dev), hs.flipped.to(dev), hs.perm.to(dev), boxes, "dark", 11, 0)
    for p_ in (0, 1, 3, 5):
        assert torch.equal(out[0][p_], ref[0][p_]) and torch.equal(out[2][p_], ref[2][p_])
    # match: zero, NaN and inf probes beside ordinary ones
    ms = synth.make_match_set(16, 300, seed=9)
    gal = ms.gallery.to(torch.bfloat16).to(dev)
    ids0, sims0 = spp.match_top1(ms.embeddings.to(dev), gal, 0.4)
    emb = ms.embeddings.clone()
    emb[3] = 0.0
    emb[5, 7] = float("nan")
    emb[9, 0] = float("inf")
    ids, sims = spp.match_top1(emb.to(dev), gal, 0.4)
    torch.cuda.synchronize()
    keep = [i for i in range(16) if i not in (3, 5, 9)]
    assert torch.equal(ids[keep], ids0[keep]) and torch.equal(sims[keep], sims0[keep])
    assert int(ids[3]) == -1                                   # a zero probe has similarity 0 with every identity: below the gate


@pytest.mark.gpu
def test_crop_graph_replay_with_new_boxes_and_concurrent_streams(spp, synth, dev):
    """The persistent crop inside a CUDA graph: the plan kernel is part of the graph, so a replay after the boxes changed in
    place crops the NEW boxes (tables, band layout and the ticket counter are rebuilt every replay).  And two crops enqueued on
    two streams at once, each with its own workspace and ticket counter, do not disturb each other."""
    cs = synth.make_crop_set(12, 240, 320, per_frame=9, seed=31)          # 108 crops x 3 x 8 slabs = 2 592 items > 740 CTAs
    fr, bx, fi = cs.frames.to(dev), cs.boxes.to(dev).clone(), cs.frame_idx.to(dev)
    ws = spp.ops.alloc_workspace(dev, spp.crop_workspace_bytes(bx.shape[0]))
    out = torch.empty(bx.shape[0], 3, 256, 192, device=dev)
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        spp.crop_affine(fr, bx, fi, out=out, workspace=ws)                # warm-up outside the capture
    s.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph, stream=s):
        spp.crop_affine(fr, bx, fi, out=out, workspace=ws)
    gph.replay()
    torch.cuda.synchronize()
    first = out.clone()
    assert torch.equal(first, spp.crop_affine(fr, bx, fi))
    other = synth.make_crop_set(12, 240, 320, per_frame=9, seed=77).boxes.to(dev)
    bx.copy_(other)                                                        # same buffer the graph reads
    gph.replay()
    torch.cuda.synchronize()
    want = spp.crop_affine(fr, other, fi)
    assert torch.equal(out, want) and not torch.equal(out, first)
    # two streams at once
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    w1 = spp.ops.alloc_workspace(dev, spp.crop_workspace_bytes(bx.shape[0]))
    w2 = spp.ops.alloc_workspace(dev, spp.crop_workspace_bytes(bx.shape[0]))
    o1, o2 = torch.empty_like(out), torch.empty_like(out)
    torch.cuda.synchronize()
    for _ in range(3):
        with torch.cuda.stream(s1):
            spp.crop_affine(fr, other, fi, out=o1, workspace=w1)
        with torch.cuda.stream(s2):
            spp.crop_affine(fr, cs.boxes.to(dev), fi, out=o2, workspace=w2)
    torch.cuda.synchronize()
    assert torch.equal(o1, want) and torch.equal(o2, first)

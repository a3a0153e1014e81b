"""Generate tests/golden/*.npz by running THE REFERENCE'S OWN CODE (imported from /root/reference,
read-only) and the un-vendored third-party code it relies on (HF transformers ViTPose processor)
on seeded synthetic inputs.  Run in the build container only:

    python oracle/gen_golden.py

The GPU box has no /root/reference; tests read only the committed fixtures.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
synth = importlib.import_module("person-recognition-for-pose-estimation_b200.synth")


def _stub_modules():
    """pytorch_lightning / pycocotools / albumentations are not installed: stub just
    enough for the reference's Lightning modules to import."""
    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            self.hparams = types.SimpleNamespace()

        def log(self, *a, **k):
            pass

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = LightningModule
    pl.LightningDataModule = object
    sys.modules["pytorch_lightning"] = pl
    class _Permissive(types.ModuleType):
        """any attribute resolves to a dummy class (only used in type annotations / unused imports)"""
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return type(item, (), {})

    for name in ["pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "albumentations",
                 "albumentations.pytorch"]:
        if name not in sys.modules:
            sys.modules[name] = _Permissive(name)


class _Slice(torch.nn.Module):
    def __init__(self, a, b):
        super().__init__()
        self.a, self.b = a, b

    def forward(self, x):
        return x[:, self.a:self.b]


def gen_det():
    sys.path.insert(0, os.path.join(REF, "training"))
    from yolopt.nets.nn import Head
    import yolopt.util as yutil

    yutil.time = lambda: 0.0          # neutralise the wall-clock bail-out (util.py:133-134,166-167)
    for tag, nc, seed in (("nc1", 1, 11), ("nc3", 3, 12)):
        hm = synth.make_head_maps(2, 256, 320, n_obj=4, nc=nc, seed=seed, min_size=24, max_size=160)
        head = Head(nc=nc, filters=(64, 64, 64))
        head.stride = torch.tensor([8.0, 16.0, 32.0])
        head.box = torch.nn.ModuleList([_Slice(0, 64) for _ in range(3)])      # raw maps already hold
        head.cls = torch.nn.ModuleList([_Slice(64, 64 + nc) for _ in range(3)])  # cat(box(x), cls(x))
        head.eval()
        with torch.no_grad():
            decoded = head([l.clone() for l in hm.levels])                     # nn.py:255-270
        conf = 0.001 if nc == 1 else 0.25
        dets = yutil.non_max_suppression(decoded, conf, 0.65)                   # util.py:123-169
        np.savez_compressed(
            os.path.join(OUT, f"det_{tag}.npz"),
            l0=hm.levels[0].numpy(), l1=hm.levels[1].numpy(), l2=hm.levels[2].numpy(),
            decoded=decoded.numpy(), conf=np.float32(conf), iou=np.float32(0.65),
            n=np.array([d.shape[0] for d in dets]),
            dets=np.concatenate([d.numpy() for d in dets], 0),
            anchors=head.anchors.numpy(), strides=head.strides.numpy())
        print("det", tag, [d.shape[0] for d in dets])


def gen_ref_head():
    """the reference's REAL ``Head`` module (its conv stacks, BatchNorm statistics, DFL) — state_dict,
    input features, the per-level ``cat(box(x), cls(x))`` maps, the eval output (nn.py:255-270) and the NMS rows
    (util.py:123-169).  The GPU test loads the state_dict into oracle/refhead.RefShapedHead and runs
    ``spp.head_eval_forward`` / ``spp.detect`` on it."""
    sys.path.insert(0, os.path.join(REF, "training"))
    from yolopt.nets.nn import Head
    import yolopt.util as yutil
    from oracle import det as odet

    yutil.time = lambda: 0.0
    filters, nc = (16, 32, 64), 2
    for seed in range(20, 60):
        torch.manual_seed(seed)
        head = Head(nc=nc, filters=filters)
        head.stride = torch.tensor([8.0, 16.0, 32.0])
        for mod in head.modules():                      # non-trivial BatchNorm statistics and affine parameters
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0.0, 0.3)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.data.uniform_(0.7, 1.3)
                mod.bias.data.normal_(0.0, 0.2)
        head.eval()
        feats = [torch.randn(2, f, 256 // s, 320 // s) * 1.5 for f, s in zip(filters, (8, 16, 32))]
        with torch.no_grad():
            cat = [torch.cat((b(x), c(x)), 1) for b, c, x in zip(head.box, head.cls, feats)]      # nn.py:257
            decoded = head([f.clone() for f in feats])
        conf = 0.5                                      # random-init class logits sit around 0: about half of the anchors pass
        dets = yutil.non_max_suppression(decoded, conf, 0.65)
        if sum(odet.near_threshold_pairs(d, 0.65) for d in dets) == 0 and all(20 <= d.shape[0] for d in dets):
            break
    else:
        raise RuntimeError("no seed without borderline IoU pairs")
    sd = {"sd." + k: v.numpy() for k, v in head.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "ref_head.npz"), nc=np.int64(nc), filters=np.array(filters), conf=np.float32(conf),
                        iou=np.float32(0.65), f0=feats[0].numpy(), f1=feats[1].numpy(), f2=feats[2].numpy(),
                        l0=cat[0].numpy(), l1=cat[1].numpy(), l2=cat[2].numpy(), decoded=decoded.numpy(),
                        n=np.array([d.shape[0] for d in dets]), dets=np.concatenate([d.numpy() for d in dets], 0), **sd)
    print("ref_head seed", seed, [d.shape[0] for d in dets], "state_dict tensors", len(sd))


def make_eval_set(batch=8, seed=0, nc=3, max_targets=14, det_cap=300):
    """Synthetic evaluation batch: per image a few labelled boxes, detections = jittered copies of them (true positives of
    varying IoU, duplicates, wrong classes) + background boxes; rows conf-descending like NMS output; tie-free scores."""
    g = torch.Generator().manual_seed(seed)
    dets = torch.zeros(batch, det_cap, 6)
    dcount = torch.zeros(batch, dtype=torch.int32)
    targets = torch.zeros(batch, max_targets, 5)
    tcount = torch.zeros(batch, dtype=torch.int32)
    for b in range(batch):
        m = int(torch.randint(0 if b == 3 else 2, max_targets + 1, (1,), generator=g))
        xy = torch.rand(m, 2, generator=g) * 500
        wh = torch.rand(m, 2, generator=g) * 150 + 20
        cls = torch.randint(0, nc, (m,), generator=g).float()
        targets[b, :m] = torch.cat([cls[:, None], xy, xy + wh], 1)
        tcount[b] = m
        rows = []
        for t in range(m):
            for _ in range(int(torch.randint(0, 5, (1,), generator=g))):
                jit = (torch.rand(4, generator=g) - 0.5) * wh[t].repeat(2) * float(torch.rand(1, generator=g)) * 0.6
                c = cls[t] if float(torch.rand(1, generator=g)) < 0.85 else float((int(cls[t]) + 1) % nc)
                rows.append(torch.cat([torch.cat([xy[t], xy[t] + wh[t]]) + jit, torch.rand(1, generator=g) * 0.9 + 0.05,
                                       torch.tensor([float(c)])]))
        for _ in range(int(torch.randint(0, 12, (1,), generator=g))):
            p = torch.rand(2, generator=g) * 600
            rows.append(torch.cat([p, p + torch.rand(2, generator=g) * 120 + 10, torch.rand(1, generator=g) * 0.6 + 0.01,
                                   torch.randint(0, nc + 1, (1,), generator=g).float()]))      # class nc never appears in the labels
        if b == 5:
            rows = []                                                                           # labels but no detections
        if rows:
            r = torch.stack(rows)
            r = r[torch.argsort(r[:, 4], descending=True)][:det_cap]
            dets[b, :r.shape[0]] = r
            dcount[b] = r.shape[0]
    return dets, dcount, targets, tcount


def gen_det_metrics():
    """the reference's own ``compute_metric`` (util.py:99-120) per image and ``compute_ap`` (util.py:225-300)
    over the whole set, driven exactly as training/yolopt/main.py:210-234 drives them."""
    import warnings
    sys.path.insert(0, os.path.join(REF, "training"))
    import yolopt.util as yutil
    dets, dcount, targets, tcount = make_eval_set()
    iou_v = torch.linspace(0.5, 0.95, 10)                                    # main.py:193
    correct = np.zeros((dets.shape[0], dets.shape[1], 10), dtype=bool)
    metrics = []
    for b in range(dets.shape[0]):
        n, m = int(dcount[b]), int(tcount[b])
        output, cls = dets[b, :n], targets[b, :m, 0:1]
        metric = torch.zeros(n, 10, dtype=torch.bool)
        if n == 0:
            if m:
                metrics.append((metric, *torch.zeros((2, 0)), cls.squeeze(-1)))
            continue
        if m:
            metric = yutil.compute_metric(output[:, :6], targets[b, :m], iou_v)
        correct[b, :n] = metric.numpy()
        metrics.append((metric, output[:, 4], output[:, 5], cls.squeeze(-1)))
    cat = [torch.cat(x, dim=0).cpu().numpy() for x in zip(*metrics)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                                       # numpy.trapz deprecation
        tp, fp, m_pre, m_rec, map50, mean_ap = yutil.compute_ap(*cat)
    np.savez_compressed(os.path.join(OUT, "det_metrics.npz"), dets=dets.numpy(), dcount=dcount.numpy(), targets=targets.numpy(),
                        tcount=tcount.numpy(), iou_v=iou_v.numpy(), correct=correct, cat_tp=cat[0], cat_conf=cat[1], cat_cls=cat[2],
                        cat_target_cls=cat[3], tp=tp, fp=fp, summary=np.array([m_pre, m_rec, map50, mean_ap]))
    print("det_metrics", correct.sum(), tp, fp, m_pre, m_rec, map50, mean_ap)


def gen_match():
    sys.path.insert(0, os.path.join(REF, "libs"))
    import net_adaface
    import head_adaface

    torch.manual_seed(3)
    net = net_adaface.build_model("ir_18").eval()
    feats = {}
    net.output_layer.register_forward_hook(lambda m, i, o: feats.__setitem__("pre", o.detach().clone()))
    with torch.no_grad():
        emb, norm = net(torch.randn(6, 3, 112, 112))                            # net_adaface.py:324-337
    ms = synth.make_match_set(48, 300, seed=5)
    kernel = (ms.gallery * torch.empty(300, 1).uniform_(0.5, 2.0)).t().contiguous()   # [512, N] un-normalised
    kn = head_adaface.l2_norm(kernel, axis=0)                                   # head_adaface.py:39-42,79

    # the live match, training/lightning/face_recognition/module.py:119-145, via the real module
    _stub_modules()
    sys.path.insert(0, os.path.join(REF, "training"))
    from lightning.face_recognition.module import FaceRecognitionModule
    import torch.nn.functional as F

    def run(labels, k):
        fake = types.SimpleNamespace()
        fake.model = lambda images: (images, None)                              # embeddings pass-through
        fake.model.ada_face = types.SimpleNamespace(head=types.SimpleNamespace(kernel=k))
        fake.hparams = types.SimpleNamespace(s=64.0)
        fake.validation_step_outputs = []
        fake.log = lambda *a, **kw: None
        return float(FaceRecognitionModule.validation_step(fake, (ms.embeddings, labels), 0)["val_acc"])

    # Recover the module's predictions probe by probe: acc==1 iff label == its argmax.
    cos_q3 = F.linear(F.normalize(ms.embeddings), F.normalize(kernel).t())     # what :137-138 computes (quirk Q3)
    pred_q3 = (cos_q3 * 64.0).max(1)[1]
    assert run(pred_q3, kernel) == 1.0 and run((pred_q3 + 1) % 300, kernel) == 0.0
    # with a kernel that is already column-normalised the quirk is (almost) a no-op on the ranking
    cos = F.linear(F.normalize(ms.embeddings), kn.t())
    pred = (cos * 64.0).max(1)[1]
    np.savez_compressed(
        os.path.join(OUT, "match.npz"),
        pre=feats["pre"].numpy(), emb=emb.numpy(), norm=norm.numpy(),
        probes=ms.embeddings.numpy(), kernel=kernel.numpy(), kernel_l2=kn.numpy(),
        pred_q3=pred_q3.numpy(), sim_q3=cos_q3.gather(1, pred_q3[:, None]).squeeze(1).numpy(),
        pred=pred.numpy(), sim=cos.gather(1, pred[:, None]).squeeze(1).numpy(), true_ids=ms.true_ids.numpy())
    print("match ok; acc vs planted:", float((pred == ms.true_ids).float().mean()))


def gen_pose_live():
    _stub_modules()
    sys.path.insert(0, os.path.join(REF, "training"))
    from lightning.pose_estimation.module import PoseEstimationModule
    from lightning.pose_estimation.datamodule import COCO_FLIP_PAIRS

    class M(torch.nn.Module):
        def set_task(self, t):
            pass

    mod = PoseEstimationModule(M())
    hs = synth.make_heatmaps(4, 17, seed=21)
    boxes = torch.tensor([[10., 20., 110., 320.], [0., 0., 30., 40.], [5., 5., 400., 700.], [50., 60., 146., 156.]])
    # correct flip-back + average (module copy.py:465-472 semantics) computed with plain torch
    fb = hs.flipped.clone()
    for a, b in COCO_FLIP_PAIRS:
        fb[:, [a, b]] = fb[:, [b, a]]
    avg = (hs.heatmaps + fb.flip(-1)) * 0.5
    # the live module's own (buggy, Q1) flip-back, module.py:479-484
    q1 = torch.flip(hs.flipped.clone(), dims=[-1])
    for pair in COCO_FLIP_PAIRS:
        q1[:, pair] = q1[:, pair].flip(0)
    avg_q1 = (hs.heatmaps + q1) * 0.5
    c0, s0 = mod._get_keypoints_from_heatmaps(hs.heatmaps)                      # module.py:237-296
    c1, s1 = mod._get_keypoints_from_heatmaps(avg, boxes=boxes)
    np.savez_compressed(
        os.path.join(OUT, "pose_live.npz"),
        hm=hs.heatmaps.numpy(), flipped=hs.flipped.numpy(), perm=hs.perm.numpy(), boxes_xyxy=boxes.numpy(),
        avg=avg.numpy(), avg_q1=avg_q1.numpy(),
        coords_plain=c0.numpy(), scores_plain=s0.numpy(), coords_avg_box=c1.numpy(), scores_avg_box=s1.numpy())
    print("pose_live ok")


def gen_pose_hf():
    """Third-party HF code (the DARK/UDP oracle, a9/a14).  Same package on both boxes; the fixtures
    guard against silent behaviour changes and give the GPU tests a reference-independent target."""
    from transformers import VitPoseImageProcessor
    from transformers.models.vitpose.modeling_vitpose import VitPoseEstimatorOutput

    proc = VitPoseImageProcessor()
    hs = synth.make_heatmaps(5, 17, seed=22)
    perm = hs.perm.long()
    avg = (hs.heatmaps + hs.flipped[:, perm].flip(-1)) * 0.5
    cs = synth.make_crop_set(1, 240, 320, per_frame=5, seed=23)
    boxes = [[[float(v) for v in b] for b in cs.boxes]]
    res = proc.post_process_pose_estimation(VitPoseEstimatorOutput(heatmaps=avg), boxes=boxes, kernel_size=11)
    kp = torch.stack([r["keypoints"] for r in res[0]]).numpy()
    sc = torch.stack([r["scores"] for r in res[0]]).numpy()
    pix = proc.preprocess([cs.frames[0]], boxes=boxes, do_rescale=False, return_tensors="pt")["pixel_values"].numpy()
    np.savez_compressed(
        os.path.join(OUT, "pose_hf.npz"),
        hm=hs.heatmaps.numpy(), flipped=hs.flipped.numpy(), perm=hs.perm.numpy(), boxes=cs.boxes.numpy(),
        keypoints=kp, scores=sc, frame=cs.frames[0].numpy(), crop_sub=pix[:, :, ::4, ::4].copy(),
        crop_sum=pix.astype(np.float64).sum(axis=(2, 3)))
    print("pose_hf ok", kp.shape, pix.shape)


def gen_pose_results():
    """a15 / 8f-3: the reference's own validation_step result loop (module.py:451-560), run on a fake
    ``self`` whose model returns seeded heatmaps.  Batch size 1 per call: there the live flip-back
    (quirk Q1, ``.flip(0)`` over the batch axis) is the identity, i.e. the flipped heatmaps are mirrored
    but NOT channel-swapped (= the ``perm=None`` mode of the kernel)."""
    _stub_modules()
    sys.path.insert(0, os.path.join(REF, "training"))
    from lightning.pose_estimation.module import PoseEstimationModule

    hs = synth.make_heatmaps(3, 17, seed=31)
    g = torch.Generator().manual_seed(32)
    n_inst = 3
    xy = torch.rand(3, n_inst, 2, generator=g) * 300
    wh = torch.rand(3, n_inst, 2, generator=g) * 250 + 20
    boxes = torch.cat([xy, xy + wh], -1)                                           # [B, N, 4] xyxy
    areas = (wh[..., 0] * wh[..., 1]) * 0.6
    masks = torch.tensor([[True, True, True], [True, True, False], [True, False, False]])
    is_crowd = torch.tensor([[False, True, False], [False, False, False], [False, False, False]])
    image_ids = [101, 202, 303]
    preds = []
    coords, scores = [], []
    for b in range(3):
        calls = []

        def model(images, b=b, calls=calls):
            calls.append(1)
            hm = hs.heatmaps[b:b + 1] if len(calls) == 1 else hs.flipped[b:b + 1]
            return types.SimpleNamespace(heatmaps=hm.clone())

        fake = types.SimpleNamespace()
        fake.model = model
        fake._generate_target_heatmap = lambda c, v, a: (None, None)
        fake.heatmap_loss = lambda *a, **k: torch.zeros(())
        real = PoseEstimationModule._get_keypoints_from_heatmaps

        def get_kp(hm, boxes=None, fake=fake):
            c, s = real(fake, hm, boxes=boxes)
            coords.append(c.numpy().copy())
            scores.append(s.numpy().copy())
            return c, s

        fake._get_keypoints_from_heatmaps = get_kp
        fake._eval_cache = {"image_ids": set()}
        fake.keypoint_thresh = 0.3
        fake.eval_predictions = preds
        fake.log = lambda *a, **k: None
        fake.print = print
        batch = {"images": torch.zeros(1, 3, 8, 8), "keypoints": torch.zeros(1, n_inst, 17, 3), "boxes": boxes[b:b + 1],
                 "areas": areas[b:b + 1], "masks": masks[b:b + 1], "is_crowd": is_crowd[b:b + 1], "image_ids": [image_ids[b]]}
        PoseEstimationModule.validation_step(fake, batch, b)
    np.savez_compressed(
        os.path.join(OUT, "pose_results.npz"),
        hm=hs.heatmaps.numpy(), flipped=hs.flipped.numpy(), boxes=boxes.numpy(), areas=areas.numpy(), masks=masks.numpy(),
        is_crowd=is_crowd.numpy(), image_ids=np.array(image_ids),
        coords=np.concatenate(coords, 0), scores=np.concatenate(scores, 0),
        res_image_id=np.array([p["image_id"] for p in preds]), res_keypoints=np.array([p["keypoints"] for p in preds], np.float64),
        res_score=np.array([p["score"] for p in preds], np.float64), res_bbox=np.array([p["bbox"] for p in preds], np.float64),
        res_area=np.array([p["area"] for p in preds], np.float64))
    print("pose_results ok:", len(preds), "rows")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]               # e.g. `python oracle/gen_golden.py ref_head` regenerates one fixture
    if only:
        for name in only:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_det()
    gen_ref_head()
    gen_det_metrics()
    gen_match()
    gen_pose_live()
    gen_pose_hf()
    gen_pose_results()

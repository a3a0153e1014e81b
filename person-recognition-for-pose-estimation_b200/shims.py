"""Drop-in replacements that keep the reference's call signatures.

Every function here has the name, argument order, defaults and return types of the reference
function it replaces, and forwards to the sm_100a kernels in ``ops``.  Inputs must be CUDA tensors
(host tensors raise: there is no CPU fallback); variable-length results are produced on the device as
padded tensors + counts and only converted to the reference's Python ``list`` forms here.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .synth import COCO_FLIP_PAIRS, flip_perm

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


# ---- training/yolopt/util.py:123 -----------------------------------------------------------------
def non_max_suppression(outputs: torch.Tensor, confidence_threshold: float = 0.001, iou_threshold: float = 0.65
                        ) -> List[torch.Tensor]:
    """``non_max_suppression(outputs, confidence_threshold=0.001, iou_threshold=0.65)`` —
    training/yolopt/util.py:123-169.  ``outputs``: ``[B, 4+nc, A]`` as returned by ``Head.forward`` in
    eval mode; returns a list of ``[n_i, 6]`` tensors (x1, y1, x2, y2, conf, cls), n_i <= 300.
    Unlike the reference there is no wall-clock bail-out (util.py:166-167)."""
    return ops.nms_decoded(outputs, confidence_threshold, iou_threshold).to_list()


# ---- training/yolopt/nets/nn.py:255 (eval branch, after the conv stacks) ---------------------------
def head_forward(levels: Sequence[torch.Tensor], strides: Sequence[float] = (8, 16, 32)) -> torch.Tensor:
    """Eval-mode ``Head.forward`` on the per-level ``cat(box(x), cls(x))`` maps -> ``[B, 4+nc, A]``."""
    return ops.head_decode(levels, strides)


def detect(levels: Sequence[torch.Tensor], confidence_threshold: float = 0.001, iou_threshold: float = 0.65,
           strides: Sequence[float] = (8, 16, 32)) -> List[torch.Tensor]:
    """Fused fast path: ``non_max_suppression(Head.forward(levels))`` without materialising the decoded tensor."""
    return ops.decode_nms(levels, strides, confidence_threshold, iou_threshold).to_list()


def head_conv_outputs(head: torch.nn.Module, feats: Sequence[torch.Tensor]):
    """The reference ``Head``'s conv stacks WITHOUT the per-level ``torch.cat`` of nn.py:257:
    ``[(head.box[i](x[i]), head.cls[i](x[i]))]`` — the pairs ``head_forward`` / ``detect`` / ``ops.decode_nms``
    consume in place of the concatenated maps.  ``head`` is the reference's own module (its weights, its convs)."""
    return [(b(x).contiguous(), c(x).contiguous()) for b, c, x in zip(head.box, head.cls, feats)]


def head_eval_forward(head: torch.nn.Module, feats: Sequence[torch.Tensor]) -> torch.Tensor:
    """Drop-in for ``Head.forward(x)`` in eval mode (nn.py:255-270): same ``[B, 4+nc, A]`` result, conv stacks by
    the reference module, everything after them (cat, anchors, DFL, dist2bbox, sigmoid) by one kernel."""
    strides = [float(s) for s in head.stride.tolist()]
    return ops.head_decode(head_conv_outputs(head, feats), strides)


# ---- libs/head_adaface.py:39 / libs/net_adaface.py:334 ---------------------------------------------
def l2_norm(input: torch.Tensor, axis: int = 1) -> torch.Tensor:  # noqa: A002 (reference's argument name)
    """``l2_norm(input, axis=1)`` — libs/head_adaface.py:39-42."""
    if axis in (1, -1) and input.dim() == 2:
        return ops.l2_normalize(input, mode="backbone")[0]
    if axis == 0 and input.dim() == 2:
        return ops.l2_normalize(input.t().contiguous(), mode="backbone")[0].t()
    raise ValueError("l2_norm: only 2-D inputs with axis 0 or 1 are supported")


def backbone_tail(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``Backbone.forward`` lines libs/net_adaface.py:334-337: returns ``(x / ||x||, ||x||)``."""
    return ops.l2_normalize(x, mode="backbone")


class Gallery:
    """Enrolled identities: a row-normalised bf16 ``[N, 512]`` matrix resident in HBM.

    ``Gallery.from_kernel`` takes the AdaFace classifier kernel ``[512, N]`` whose columns are
    identities (libs/head_adaface.py:79 normalises it along axis 0); ``quirk_q3=True`` reproduces
    ``F.normalize(kernel)`` of training/lightning/face_recognition/module.py:137 (wrong axis).
    Enrolment is off the hot path."""

    def __init__(self, rows_bf16: torch.Tensor, id_offset: int = 0, rows_f32: Optional[torch.Tensor] = None,
                 max_row_norm: Optional[float] = None):
        if rows_bf16.dtype != torch.bfloat16 or rows_bf16.dim() != 2:
            raise TypeError("Gallery: expected a bfloat16 [N, 512] tensor")
        self.rows = rows_bf16.contiguous()
        self.id_offset = int(id_offset)
        # fp32 rows the bf16 copy was rounded from (optional): candidates are re-scored against them, so ids and
        # similarities are the reference's fp32 ones; without them they are exact for the bf16 gallery.
        self.rows_f32 = None if rows_f32 is None else rows_f32.float().contiguous()
        # Largest row norm: scales the band of bf16 scores that get the exact re-score (1 for normalised rows; the
        # quirk-Q3 enrolment does NOT give unit rows).  Measured once here — enrolment is off the hot path.
        if max_row_norm is None:
            src = self.rows_f32 if self.rows_f32 is not None else self.rows.float()
            max_row_norm = float(src.norm(dim=1).max()) if src.shape[0] else 1.0
            if self.rows_f32 is not None and self.rows.shape[0]:
                max_row_norm = max(max_row_norm, float(self.rows.float().norm(dim=1).max()))
        self.max_row_norm = float(max(max_row_norm, 1e-6))

    @classmethod
    def from_kernel(cls, kernel: torch.Tensor, quirk_q3: bool = False, keep_f32: bool = False) -> "Gallery":
        kn = torch.nn.functional.normalize(kernel) if quirk_q3 else kernel / torch.norm(kernel, 2, 0, True)
        rows = kn.t().contiguous()
        return cls(ops.to_bf16(rows), rows_f32=rows if keep_f32 else None)

    @classmethod
    def from_rows(cls, rows: torch.Tensor, normalize: bool = True, id_offset: int = 0, keep_f32: bool = False) -> "Gallery":
        if normalize:
            rows = ops.l2_normalize(rows, mode="normalize")[0]
        return cls(ops.to_bf16(rows), id_offset, rows_f32=rows if keep_f32 else None)

    def __len__(self) -> int:
        return self.rows.shape[0]

    # ---- storage: a plain [N, 512] bf16 matrix; a shard is a contiguous row range ----
    def save(self, path: str) -> None:
        """Raw little-endian bf16 rows (``N * 1024`` bytes) + a tiny JSON header beside it."""
        import json
        import numpy as np
        self.rows.cpu().view(torch.int16).numpy().tofile(path)
        with open(path + ".json", "w") as f:
            json.dump({"ids": int(self.rows.shape[0]), "dim": int(self.rows.shape[1]), "dtype": "bfloat16",
                       "id_offset": self.id_offset, "max_row_norm": self.max_row_norm,
                       "layout": "row-major, rows normalised at enrolment"}, f)

    @classmethod
    def load(cls, path: str, device, world: int = 1, rank: int = 0) -> "Gallery":
        """Memory-map the file and upload only this rank's contiguous row shard (``dist.shard_bounds``)."""
        import json
        import numpy as np
        from .dist import shard_bounds
        with open(path + ".json") as f:
            meta = json.load(f)
        lo, hi = shard_bounds(meta["ids"], world, rank)
        mm = np.memmap(path, dtype=np.int16, mode="r", shape=(meta["ids"], meta["dim"]))
        rows = torch.from_numpy(np.array(mm[lo:hi])).view(torch.bfloat16).to(device)
        return cls(rows, id_offset=meta.get("id_offset", 0) + lo, max_row_norm=meta.get("max_row_norm"))

    def match(self, embeddings: torch.Tensor, threshold: Optional[float] = None):
        """``pred = (F.linear(F.normalize(emb), G) * s).max(1)[1]`` (face_recognition/module.py:136-145) plus the
        optional similarity gate; returns ``(ids int64 [M], sims fp32 [M])``, ``ids == -1`` where gated."""
        ids, sims = ops.match_top1(embeddings, self.rows, threshold, self.id_offset, gallery_f32=self.rows_f32,
                                   max_row_norm=self.max_row_norm)
        return ids.long(), sims


# ---- HF VitPoseImageProcessor (transformers/models/vitpose/image_processing_vitpose.py) -------------
def _flatten_boxes(boxes, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """``list[list[list[float]]]`` (per image, per box, COCO x,y,w,h) or a ready ``[P,4]`` tensor."""
    if isinstance(boxes, torch.Tensor):
        raise TypeError("pass (boxes[P,4], frame_idx[P]) tensors through ops.crop_affine directly")
    flat, idx = [], []
    for i, per_image in enumerate(boxes):
        for b in per_image:
            flat.append([float(v) for v in b[:4]])
            idx.append(i)
    return (torch.tensor(flat, dtype=torch.float32, device=device).reshape(-1, 4),
            torch.tensor(idx, dtype=torch.int32, device=device))


class VitPoseImageProcessor:
    """Signature-compatible subset of HF ``VitPoseImageProcessor`` (preprocess + post-process)."""

    def __init__(self, size=None, do_rescale: bool = True, rescale_factor: float = 1 / 255, do_normalize: bool = True,
                 image_mean=IMAGENET_DEFAULT_MEAN, image_std=IMAGENET_DEFAULT_STD):
        self.size = size or {"height": 256, "width": 192}
        self.do_rescale, self.rescale_factor, self.do_normalize = do_rescale, rescale_factor, do_normalize
        self.image_mean, self.image_std = tuple(image_mean), tuple(image_std)

    def _mean_std(self, do_rescale, do_normalize):
        mean = torch.tensor(self.image_mean if do_normalize else (0.0, 0.0, 0.0), dtype=torch.float32)
        std = torch.tensor(self.image_std if do_normalize else (1.0, 1.0, 1.0), dtype=torch.float32)
        if do_rescale:   # image_processing_backends.py:301-305 — rescale folded into mean/std in fp32
            mean = mean * (1.0 / self.rescale_factor)
            std = std * (1.0 / self.rescale_factor)
        return mean.tolist(), std.tolist()

    def preprocess(self, images, boxes, do_rescale: Optional[bool] = None, do_normalize: Optional[bool] = None, **_):
        """HF:355-448.  ``images``: ``[B,3,H,W]`` float tensor or list of ``[3,H,W]`` tensors of equal size."""
        frames = images if isinstance(images, torch.Tensor) else torch.stack(list(images))
        b, idx = _flatten_boxes(boxes, frames.device)
        mean, std = self._mean_std(self.do_rescale if do_rescale is None else do_rescale,
                                   self.do_normalize if do_normalize is None else do_normalize)
        if frames.dtype != torch.uint8:
            frames = frames.float()
        pix = ops.crop_affine(frames, b, idx, (self.size["height"], self.size["width"]), mean, std, "hf")
        return {"pixel_values": pix}

    def post_process_pose_estimation(self, outputs, boxes, kernel_size: int = 11, threshold: Optional[float] = None,
                                     flipped_heatmaps: Optional[torch.Tensor] = None,
                                     flip_pairs: Optional[Sequence[Tuple[int, int]]] = COCO_FLIP_PAIRS,
                                     hf_index_quirk: bool = False):
        """HF:465-535.  ``outputs`` has ``.heatmaps [P,K,H,W]`` (or is that tensor).  Extension: pass the raw
        heatmaps of the mirrored crops as ``flipped_heatmaps`` and the flip test is fused into the decode.
        ``hf_index_quirk=True`` reproduces HF's float32 flat tap index (quirk Q6): identical to HF for calls of ANY size;
        the default agrees with HF for the first 2^24 / ((W+2)(H+2)) maps of a call (299 crops of 17 joints at 64x48) and
        returns the intended DARK refinement where HF's index has lost its low bits."""
        hm = outputs.heatmaps if hasattr(outputs, "heatmaps") else outputs
        b, _ = _flatten_boxes(boxes, hm.device)
        perm = None
        if flipped_heatmaps is not None:
            perm = flip_perm(hm.shape[1], flip_pairs).to(hm.device)
        kp, sc, _ = ops.heatmap_decode(hm, flipped_heatmaps, perm, b, "dark", kernel_size,
                                       flags=ops.FLAG_HF_F32_INDEX if hf_index_quirk else 0,
                                       crop_hw=(self.size["height"], self.size["width"]))
        kp, sc = kp.cpu(), sc.cpu()
        labels = torch.arange(0, hm.shape[1])
        # bbox field exactly as HF builds it: (cx, cy, scale_x, scale_y) pushed through coco_to_pascal_voc
        from .hostmath import hf_center_scale
        cs = hf_center_scale(b.cpu(), self.size["width"], self.size["height"])
        bbox = cs.clone()
        bbox[:, 2] = cs[:, 2] + cs[:, 0] - 1
        bbox[:, 3] = cs[:, 3] + cs[:, 1] - 1
        results, it = [], iter(range(hm.shape[0]))
        for per_image in boxes:
            image_results = []
            for _ in per_image:
                i = next(it)
                pose, score, lab = kp[i], sc[i], labels
                if threshold is not None:
                    keep = score > threshold
                    pose, score, lab = pose[keep], score[keep], lab[keep]
                image_results.append({"keypoints": pose, "scores": score, "labels": lab, "bbox": bbox[i]})
            results.append(image_results)
        return results


# ---- training/lightning/pose_estimation/module.py:237 ----------------------------------------------
def get_keypoints_from_heatmaps(heatmaps: torch.Tensor, boxes: Optional[torch.Tensor] = None
                                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``PoseEstimationModule._get_keypoints_from_heatmaps(heatmaps, boxes=None)`` — soft-argmax decode,
    returns ``(coords [B,K,2] in [0,1], scores [B,K])``; ``boxes`` ``[B,4]`` xyxy scale the scores."""
    kp, sc, _ = ops.heatmap_decode(heatmaps, None, None, boxes, "softargmax",
                                   flags=ops.FLAG_SCALE_SCORE if boxes is not None else 0)
    return kp, sc


def flip_test_keypoints(heatmaps: torch.Tensor, flipped_heatmaps: torch.Tensor, boxes: torch.Tensor,
                        keypoint_thresh: float = 0.3, flip_pairs=COCO_FLIP_PAIRS) -> torch.Tensor:
    """module.py:473-484 (with the correct channel swap) + :499 + :534-546 in one pass: returns COCO-style
    ``[B, K, 3]`` (x, y, v) in image pixels, v = 2 where score > thresh else 1."""
    perm = flip_perm(heatmaps.shape[1], flip_pairs).to(heatmaps.device)
    kp, sc, _ = ops.heatmap_decode(heatmaps, flipped_heatmaps, perm, boxes, "softargmax",
                                   flags=ops.FLAG_SCALE_SCORE | ops.FLAG_BACKPROJECT)
    v = torch.where(sc > keypoint_thresh, 2.0, 1.0)
    return torch.cat([kp, v[..., None]], -1)


# ---- gluoncv get_final_preds as called at pose_estimation/module_v2.py:216 --------------------------
def get_final_preds(batch_heatmaps: torch.Tensor, center: torch.Tensor, scale: torch.Tensor):
    """``get_final_preds(heatmaps, center, scale)`` -> ``(preds [N,K,2], maxvals [N,K,1])`` (arg-max +
    quarter offset + inverse affine; scale in pixels as produced by datamodule_v2.py:122)."""
    cs = torch.cat([center.float(), scale.float()], 1).to(batch_heatmaps.device).contiguous()
    kp, sc, _ = ops.heatmap_decode(batch_heatmaps, None, None, cs, "quarter", flags=4)
    return kp, sc[..., None]


# ---- training/lightning/pose_estimation/module.py:505-560 (the COCO result loop of validation_step) --------
COCO_SIGMAS = (.026, .025, .025, .035, .035, .079, .079, .072, .072, .062, .062, .107, .107, .087, .087, .089, .089)  # datamodule.py:37-40


def coco_keypoint_results(pred_coords: torch.Tensor, pred_scores: torch.Tensor, boxes: torch.Tensor, areas: torch.Tensor,
                          masks: torch.Tensor, is_crowd: torch.Tensor, image_ids: Sequence[int],
                          keypoint_thresh: float = 0.3, seen_image_ids: Optional[set] = None) -> list:
    """The reference builds its COCO predictions with a triple Python loop and ``.item()`` per value
    (module.py:505-560).  Same rows, one kernel: ``pred_coords [B,K,2]`` normalised, ``pred_scores [B,K]``,
    ``boxes [B,N,4]`` xyxy, ``areas [B,N]``, ``masks / is_crowd [B,N]`` bool.  As in the reference, every
    instance ``n < masks[b].sum()`` of image ``b`` that is not a crowd gets image ``b``'s keypoints scaled
    into its own box; ``bbox`` is the xyxy box as the reference emits it.  One D2H of the packed rows.
    ``seen_image_ids`` is the reference's ``self._eval_cache['image_ids']`` (module.py:509-513): an image whose id is
    already in the set is skipped, every processed id is added — pass the same set for every batch of a validation
    run (``None``: de-duplicate within this call only)."""
    dev = pred_coords.device
    m = masks.to("cpu").bool()
    crowd = is_crowd.to("cpu").bool()
    seen = seen_image_ids if seen_image_ids is not None else set()
    fresh = []
    for b in range(len(image_ids)):
        iid = image_ids[b]
        iid = iid.item() if hasattr(iid, "item") else iid
        if iid in seen:
            continue
        seen.add(iid)
        fresh.append(b)
    pairs = [(b, n) for b in fresh for n in range(int(m[b].sum())) if not bool(crowd[b, n])]
    if not pairs:
        return []
    bi = torch.tensor([b for b, _ in pairs], device=dev)
    ni = torch.tensor([n for _, n in pairs], device=dev)
    sel_boxes = boxes.to(dev).float()[bi, ni].contiguous()
    rows, inst = ops.pose_results(pred_coords.float()[bi].contiguous(), pred_scores.float()[bi].contiguous(), sel_boxes,
                                  keypoint_thresh)
    rows, inst, sel_boxes = rows.cpu(), inst.cpu(), sel_boxes.cpu()
    areas = areas.to("cpu")
    out = []
    for r, (b, n) in enumerate(pairs):
        kp = rows[r].tolist()
        out.append({"image_id": int(image_ids[b]), "category_id": 1,
                    "keypoints": [v if i < 2 else int(v) for xyv in kp for i, v in enumerate(xyv)],
                    "score": float(inst[r]), "bbox": [float(v) for v in sel_boxes[r]], "area": float(areas[b, n])})
    return out


# ---- training/yolopt/util.py:99 / :225 (evaluation: true-positive matrix and AP summary) ------------------------------
def compute_metric(output: torch.Tensor, target: torch.Tensor, iou_v: torch.Tensor) -> torch.Tensor:
    """``compute_metric(output, target, iou_v)`` — training/yolopt/util.py:99-120: ``output [n, 6]`` detections of one
    image, ``target [m, 5]`` (cls, x1, y1, x2, y2), ``iou_v [T]`` -> ``correct [n, T]`` bool."""
    n, m = output.shape[0], target.shape[0]
    if n == 0 or m == 0:
        return torch.zeros((n, iou_v.shape[0]), dtype=torch.bool, device=output.device)
    cnt = torch.tensor([n], dtype=torch.int32, device=output.device)
    tcn = torch.tensor([m], dtype=torch.int32, device=output.device)
    return ops.det_match_targets(output[None, :, :6].contiguous(), cnt, target[None].contiguous(), tcn, iou_v.tolist())[0]


def compute_ap(tp, conf, output, target, plot: bool = False, names=(), eps: float = 1e-16):
    """``compute_ap(tp, conf, output, target)`` — training/yolopt/util.py:225-300 without the plots; tensors (CUDA) or
    numpy arrays as the reference passes them.  Returns ``(tp, fp, m_pre, m_rec, map50, mean_ap)`` (numpy / floats)."""
    if plot:
        raise NotImplementedError("compute_ap: the PR / F1 plots of the reference are not reproduced")
    dev = tp.device if isinstance(tp, torch.Tensor) and tp.is_cuda else torch.device("cuda", torch.cuda.current_device())
    as_t = lambda x: (x if isinstance(x, torch.Tensor) else torch.as_tensor(x)).to(dev)
    res = ops.det_average_precision(as_t(tp), as_t(conf), as_t(output), as_t(target), eps=eps)
    return res["tp"].cpu().numpy(), res["fp"].cpu().numpy(), res["m_pre"], res["m_rec"], res["map50"], res["mean_ap"]

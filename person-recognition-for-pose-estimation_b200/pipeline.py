"""The post-backbone selective-pose pass assembled from the four ops.

The reference has no end-to-end inference function (scripts/modify_models.py:71-76 is a TODO); the
pieces live in the per-task ``validation_step``s.  ``SelectivePosePipeline`` strings their B200
replacements together for one batch of backbone outputs:

    face head maps   -> decode + NMS                      (yolopt/nets/nn.py:255-270, util.py:123-169)
    person head maps -> decode + NMS
    face embeddings  -> L2-norm + gallery top-1 + gate    (face_recognition/module.py:136-145)
    frames + person boxes -> 256x192 crops                (HF VitPoseImageProcessor.preprocess)
    heatmaps (+ mirrored) -> keypoints in frame pixels    (HF post_process_pose_estimation / module.py:237-296)

Everything stays on the device; the whole chain is captured once into a CUDA graph and replayed, so a
step costs one graph launch.  ``run_host`` is the same pass fed from pinned host buffers with the
result copied back (what ``bench.py`` reports as ``e2e``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import _lib, ops, synth


@dataclass
class StepInputs:
    """Backbone outputs for one batch (device or pinned-host tensors)."""
    face_levels: Sequence[torch.Tensor]      # 3 x [B, 64+nc, H_l, W_l]
    person_levels: Sequence[torch.Tensor]    # 3 x [B, 64+nc, H_l, W_l]
    embeddings: torch.Tensor                 # [M, 512]
    frames: torch.Tensor                     # [B, 3, H, W]
    boxes: torch.Tensor                      # [P, 4] COCO x,y,w,h
    frame_idx: torch.Tensor                  # [P] int32
    heatmaps: torch.Tensor                   # [P, K, 64, 48]
    flipped: Optional[torch.Tensor]          # [P, K, 64, 48] or None
    perm: Optional[torch.Tensor]             # [K] int32

    def tensors(self) -> Dict[str, torch.Tensor]:
        d = {f"face_l{i}": t for i, t in enumerate(self.face_levels)}
        d.update({f"person_l{i}": t for i, t in enumerate(self.person_levels)})
        d.update(embeddings=self.embeddings, frames=self.frames, boxes=self.boxes, frame_idx=self.frame_idx,
                 heatmaps=self.heatmaps)
        if self.flipped is not None:
            d["flipped"] = self.flipped
        if self.perm is not None:
            d["perm"] = self.perm
        return d

    @classmethod
    def from_tensors(cls, d: Dict[str, torch.Tensor]) -> "StepInputs":
        return cls([d[f"face_l{i}"] for i in range(3)], [d[f"person_l{i}"] for i in range(3)], d["embeddings"],
                   d["frames"], d["boxes"], d["frame_idx"], d["heatmaps"], d.get("flipped"), d.get("perm"))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors().values())


def synthetic_inputs(batch: int, height: int, width: int, per_frame: int, joints: int = 17, seed: int = 0,
                     flip: bool = True, nc: int = 1) -> StepInputs:
    """SURVEY.md §8d synthetic workload (CPU tensors): ``height``/``width`` is the frame size; the
    detection heads see it letterboxed up to a multiple of 32."""
    lh, lw = (height + 31) // 32 * 32, (width + 31) // 32 * 32
    face = synth.make_head_maps_fast(batch, lh, lw, n_obj=per_frame, nc=nc, seed=seed)
    person = synth.make_head_maps_fast(batch, lh, lw, n_obj=per_frame, nc=nc, seed=seed + 1)
    cs = synth.make_crop_set(batch, height, width, per_frame=per_frame, seed=seed + 2)
    p = batch * per_frame
    hs = synth.make_heatmaps(p, joints, seed=seed + 3)
    ms = synth.make_match_set(p, 16, seed=seed + 4)     # probes only; the gallery is built separately
    return StepInputs(face.levels, person.levels, ms.embeddings, cs.frames, cs.boxes, cs.frame_idx, hs.heatmaps,
                      hs.flipped if flip else None, hs.perm if flip else None)


class SelectivePosePipeline:
    """One CUDA-graphed pass of the glue path over fixed-shape inputs resident on ``device``."""

    def __init__(self, inputs: StepInputs, gallery_bf16: torch.Tensor, device: torch.device, threshold: float = 0.4,
                 conf_thres: float = 0.001, iou_thres: float = 0.65, decode_mode: str = "dark", use_graph: bool = True,
                 id_offset: int = 0, concurrent: bool = True, matcher=None, capture_collectives: bool = False,
                 select_on_device: bool = False, gallery_f32: Optional[torch.Tensor] = None, max_row_norm: float = 1.0,
                 det_max_candidates: int = 0, match_sms: int = 0, heatmap_first: bool = True, det_fused: Optional[bool] = None,
                 det_after_heatmap: Optional[int] = None, crop_free_ctas: int = 48):
        self.device = device
        self.threshold, self.conf, self.iou, self.mode = threshold, conf_thres, iou_thres, decode_mode
        self.id_offset = id_offset
        # det_max_candidates: bound on the candidates per image handed to the NMS (ops.decode_nms max_candidates).  <= 512
        # selects the small-footprint NMS kernel, whose CTAs fit beside the heatmap decode's, so the detection chains run
        # under it; an image with more candidates is flagged (NmsResult.overflowed()).  0 = every anchor may be one.
        # match_sms: SMs given to the match GEMM (it needs whole SMs); the heatmap decode runs on the others at the same
        # time.  0 = no split (each kernel takes the whole machine in turn).
        # heatmap_first: heatmap decode, then crop on the main stream (they are independent inputs of a step); the
        # latency-bound chains overlap the heatmap decode, whose CTAs leave registers and 20 KB of shared memory free.
        # det_fused: one kernel per head (a CTA per image: scan, candidate decode, sort, NMS) instead of three launches.
        # Default: fused exactly when the small-footprint configuration applies — 128 small CTAs then run under the heatmap
        # decode with no launch gaps and no waves; the three-launch form is faster when the chain has the machine to itself.
        self.det_max_candidates, self.match_sms = int(det_max_candidates), int(match_sms)
        self.det_fused = (0 < self.det_max_candidates <= 512) if det_fused is None else bool(det_fused)
        self.heatmap_first = heatmap_first and not select_on_device
        # det_after_heatmap: number of detection chains (0, 1, 2) that wait for the heatmap decode and run beside the crop instead.
        # Their candidate decode is ~450 k isolated 128-byte line fetches; beside the heatmap decode the two cost each other 40 us
        # (heatmap decode 100 us instead of 62), beside the persistent crop 27 us: 0.237-0.240 ms per step with 2 against 0.247
        # with 0 and 0.251 with 1 (measured with the round-2 kernels; with the round-1 crop it was a tie).
        # Default: 2 with the fused small-footprint detection kernels, 0 otherwise (cfg4: unbounded candidate lists, three launches
        # with a 112 KB NMS CTA; beside the crop they cost more than under the 3 ms heatmap decode: 4.39 against 4.28 ms per step).
        if det_after_heatmap is None:
            det_after_heatmap = 2 if self.det_fused else 0
        self.det_after_heatmap = int(det_after_heatmap) if self.heatmap_first else 0
        # crop_free_ctas: CTA slots the persistent crop kernel leaves free (SPP_LIMIT_CROP_FREE_CTAS) so that the match re-score,
        # which becomes ready while the crop holds the machine, runs beside it instead of after it.
        self.crop_free_ctas = int(crop_free_ctas) if concurrent else 0
        self.gallery = gallery_bf16.to(device).contiguous()
        self.gallery_f32 = None if gallery_f32 is None else gallery_f32.to(device).float().contiguous()
        self.max_row_norm = float(max_row_norm)
        self.inp = StepInputs.from_tensors({k: v.to(device).contiguous() for k, v in inputs.tensors().items()})
        self.out: Dict[str, torch.Tensor] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self._stream = torch.cuda.Stream(device)
        self.concurrent = concurrent
        # select_on_device: the crop list is not an input but comes from the detections — persons that contain
        # a face with a matched identity (ops.associate), at most `per_frame` per frame, no host round-trip.
        self.select_on_device = select_on_device
        if select_on_device:
            assert matcher is None or capture_collectives, "select_on_device needs the match inside the step"
            b = self.inp.frames.shape[0]
            self.sel_cap = self.inp.boxes.shape[0] // b
            assert self.sel_cap * b == self.inp.boxes.shape[0], "select_on_device needs the same crop capacity per frame"
            self._face_ids = torch.full((b, 300), -1, dtype=torch.int32, device=device)
            self._sel_frame_idx = torch.arange(b, dtype=torch.int32, device=device).repeat_interleave(self.sel_cap).contiguous()
        # multi-GPU: a dist.ShardedGalleryMatcher replaces the local match chain; its NCCL collectives run
        # eagerly on a side stream next to the graph (they are not captured)
        # The NCCL collectives run eagerly on a high-priority side stream beside the graph.
        # capture_collectives=True would capture them into the graph instead — NOT the default: on B200 /
        # NCCL 2.28.9 / torch 2.11 a multi-branch capture containing the two collectives hung at replay.
        # A dist.PeerShardedMatcher (capturable = True) has no NCCL call at all: its five kernels exchange probes and
        # keys over NVLink peer memory and are captured like every other kernel of the step.
        self.matcher = matcher
        self.capture_collectives = (capture_collectives or getattr(matcher, "capturable", False)) and use_graph
        self._eager_match = matcher is not None and not (self.capture_collectives or getattr(matcher, "capturable", False))
        # The pipeline owns its scratch buffers: their addresses are baked into the captured graph, so they must live
        # exactly as long as the pipeline does (ops' shared grow-only cache may replace its buffers at any later call).
        b_, a_ = self.inp.face_levels[0].shape[0], sum(l.shape[2] * l.shape[3] for l in self.inp.face_levels)
        nc_ = self.inp.face_levels[0].shape[1] - 64
        self._ws_face = ops.alloc_workspace(device, ops.nms_workspace_bytes(b_, a_, nc_, self.det_max_candidates))
        ncp_ = self.inp.person_levels[0].shape[1] - 64
        ap_ = sum(l.shape[2] * l.shape[3] for l in self.inp.person_levels)
        self._ws_person = ops.alloc_workspace(device, ops.nms_workspace_bytes(self.inp.person_levels[0].shape[0], ap_, ncp_,
                                                                              self.det_max_candidates))
        with self._limits():
            self._ws_match = (ops.alloc_workspace(device, ops.match_workspace_bytes(self.inp.embeddings.shape[0], self.gallery.shape[0]))
                              if matcher is None else None)
        self._ws_crop = ops.alloc_workspace(device, ops.crop_workspace_bytes(self.inp.boxes.shape[0], 256, 192,
                                                                             self.inp.frames.dtype == torch.uint8))
        self._match_stream = torch.cuda.Stream(device, priority=-1) if self._eager_match else None
        self._side = [torch.cuda.Stream(device) for _ in range(3)]
        self._plan_stream = torch.cuda.Stream(device)
        with torch.cuda.stream(self._stream), self._limits():
            if self._eager_match:
                self.out["ids"], self.out["sims"] = matcher.match(self.inp.embeddings)
            self._enqueue()                       # warm-up: sizes workspaces, sets kernel attributes
            self._enqueue()
        self._stream.synchronize()
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with self._limits(), torch.cuda.graph(g, stream=self._stream):
                self._enqueue()
            self.graph = g

    def _limits(self):
        """CTA budgets of the two whole-machine kernels while this pipeline enqueues (or captures) its step."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            L = _lib.lib()
            prev_mode = L.spp_decode_nms_mode(1 if self.det_fused else 0)
            prev_c = L.spp_set_launch_limit(2, self.crop_free_ctas)
            prev_h = prev_m = None
            if self.match_sms > 0:
                sms = L.spp_device_sm_count()
                prev_h = L.spp_set_launch_limit(0, max(1, sms - self.match_sms))
                prev_m = L.spp_set_launch_limit(1, self.match_sms)
            try:
                yield
            finally:
                L.spp_decode_nms_mode(prev_mode)
                L.spp_set_launch_limit(2, prev_c)
                if prev_h is not None:
                    L.spp_set_launch_limit(0, prev_h)
                    L.spp_set_launch_limit(1, prev_m)
        return ctx()

    def _enqueue(self) -> None:
        """Enqueue one pass.  Four independent chains run on forked streams and join at the end (inside a
        capture this becomes a graph with parallel branches): the HBM-bound crop -> heatmap-decode chain on the
        pipeline stream, the latency-bound detection chains (face / person) and the match chain beside it, so
        the small kernels fill SM slots and hide behind the two bandwidth-bound ones."""
        i = self.inp
        main = torch.cuda.current_stream(self.device)
        if self.concurrent:
            fork = torch.cuda.Event()
            fork.record(main)
            for s in self._side:
                s.wait_event(fork)
        sides = self._side if self.concurrent else [main, main, main]
        n = 0
        flags = (ops.FLAG_SCALE_SCORE | ops.FLAG_BACKPROJECT) if self.mode == "softargmax" else 0
        kp = None
        # Enqueue (= graph launch) order matters for who gets SM slots first: the two detection chains go first (one small
        # CTA per image lands on its own SM), then the heatmap decode, whose persistent CTAs co-reside with them, then the
        # match chain — its GEMM needs whole SMs and must not sit in front of the detection kernels in a hardware queue.
        def det(which, stream):
            lv, key, ws = ((i.face_levels, "_face", self._ws_face) if which == 0 else (i.person_levels, "_person", self._ws_person))
            with torch.cuda.stream(stream):
                return ops.decode_nms(lv, conf_thres=self.conf, iou_thres=self.iou, out=self.out.get(key), workspace=ws,
                                      max_candidates=self.det_max_candidates)
        # The crop's plan kernel (source map, coordinate tables, band layout per box) only reads the boxes: when they are inputs
        # of the step it runs here, on the match chain's stream, beside the heatmap decode, and the crop itself is only the
        # stream kernel.
        crop_planned = self.concurrent and not self.select_on_device
        if crop_planned:
            self._plan_stream.wait_event(fork)
            with torch.cuda.stream(self._plan_stream):
                ops.crop_plan(i.boxes, i.frame_idx, i.frames.shape, i.frames.dtype == torch.uint8, self._ws_crop)
                plan_done = torch.cuda.Event()
                plan_done.record(self._plan_stream)
            n += 1
        late = self.det_after_heatmap if self.concurrent else 0
        face = det(0, sides[0]) if late < 2 else None
        person = det(1, sides[1]) if late < 1 else None
        if self.heatmap_first:
            kp = ops.heatmap_decode(i.heatmaps, i.flipped, i.perm, i.boxes, self.mode, 11, flags,
                                    out=(self.out["keypoints"], self.out["scores"], self.out["argmax"]) if "keypoints" in self.out else None)
            n += 1
        if late:
            done = torch.cuda.Event()
            done.record(main)
            if person is None:
                sides[1].wait_event(done)
                person = det(1, sides[1])
            if face is None:
                sides[0].wait_event(done)
                face = det(0, sides[0])
        # candidate scan + candidate decode + NMS kernel per head (the count memset is not a kernel); one fused kernel
        # per head after ops.set_decode_nms_mode("fused")
        n += 2 * (1 if _lib.lib().spp_decode_nms_mode(-1) == 1 else 3)
        if self.matcher is None:
            with torch.cuda.stream(sides[2]):
                ids, sims, keys = ops.match_top1(i.embeddings, self.gallery, self.threshold, self.id_offset, want_keys=True,
                                                 gallery_f32=self.gallery_f32, max_row_norm=self.max_row_norm,
                                                 workspace=self._ws_match,
                                                 out=(self.out["ids"], self.out["sims"], self.out["keys"]) if "keys" in self.out and self.out["keys"] is not None else None)
            n += 3          # normalise, tcgen05 GEMM + top-2, fp32 re-score
        elif self._eager_match:
            ids, sims, keys = self.out.get("ids"), self.out.get("sims"), None
            n += 4          # (eager, in step()) normalise, GEMM + top-2, re-score, key unpack
        else:
            with torch.cuda.stream(sides[2]):
                # peer matcher: push probes -> wait -> local top-1 -> push keys -> reduce (5 kernels, no NCCL);
                # NCCL matcher (capture_collectives): all_gather -> local top-1 -> all_reduce(MAX) -> unpack
                ids, sims = self.matcher.match(i.embeddings)
            keys = None
            n += getattr(self.matcher, "launches", 4)
        boxes, frame_idx = i.boxes, i.frame_idx
        if self.select_on_device:
            if self.concurrent:              # the selection needs both detection chains and the match
                for s in self._side:
                    e = torch.cuda.Event()
                    e.record(s)
                    main.wait_event(e)
            pf = min(self.sel_cap, ids.shape[0] // self._face_ids.shape[0])
            self._face_ids[:, :pf] = ids.view(self._face_ids.shape[0], -1)[:, :pf].to(torch.int32)
            sel_boxes, sel_ident, sel_rows, sel_count = ops.associate(face, self._face_ids, person, self.sel_cap)
            boxes, frame_idx = sel_boxes.view(-1, 4), self._sel_frame_idx
            self.out.update(sel_boxes=sel_boxes, sel_ident=sel_ident, sel_count=sel_count)
            n += 2
        if i.frames.dtype == torch.uint8:     # HF default for uint8 images: 1/255 rescale folded into mean / std
            mean, std = [m * 255.0 for m in (0.485, 0.456, 0.406)], [s * 255.0 for s in (0.229, 0.224, 0.225)]
            if crop_planned:
                main.wait_event(plan_done)
            pix = ops.crop_affine(i.frames, boxes, frame_idx, mean=mean, std=std, out=self.out.get("pixel_values"), workspace=self._ws_crop,
                                  planned=crop_planned)
        else:
            if crop_planned:
                main.wait_event(plan_done)
            pix = ops.crop_affine(i.frames, boxes, frame_idx, out=self.out.get("pixel_values"), workspace=self._ws_crop, planned=crop_planned)
        n += 1 if crop_planned else 2          # stream kernel (+ plan kernel when it was not run ahead)
        if kp is None:
            kp = ops.heatmap_decode(i.heatmaps, i.flipped, i.perm, boxes, self.mode, 11, flags,
                                    out=(self.out["keypoints"], self.out["scores"], self.out["argmax"]) if "keypoints" in self.out else None)
            n += 1
        if self.concurrent:
            for s in self._side:
                join = torch.cuda.Event()
                join.record(s)
                main.wait_event(join)
        self.launches_per_step = n
        self.out.update(_face=face, _person=person, face_dets=face.dets, face_count=face.count, person_dets=person.dets,
                        person_count=person.count, ids=ids, sims=sims, keys=keys, pixel_values=pix, keypoints=kp[0],
                        scores=kp[1], argmax=kp[2])

    def step(self) -> Dict[str, torch.Tensor]:
        """Enqueue one pass on the pipeline's stream (graph replay when captured)."""
        with torch.cuda.stream(self._stream):
            self._launch()
        return self.out

    def _launch(self) -> None:
        main = torch.cuda.current_stream(self.device)
        if self._eager_match:
            fork = torch.cuda.Event()
            fork.record(main)
        if self.graph is not None:
            self.graph.replay()
        else:
            with self._limits():
                self._enqueue()
        if self._eager_match:
            # gallery-sharded match: all_gather -> local top-1 -> all_reduce(MAX).  Enqueued AFTER the graph
            # launch so the host issues these eager calls while the device is busy with the graph; on the
            # device the chain runs beside the graph on its own stream.
            self._match_stream.wait_event(fork)
            with torch.cuda.stream(self._match_stream):
                ids, sims = self.matcher.match(self.inp.embeddings)
                self.out["ids"], self.out["sims"] = ids, sims
            join = torch.cuda.Event()
            join.record(self._match_stream)
            main.wait_event(join)

    @property
    def stream(self) -> torch.cuda.Stream:
        return self._stream

    # ---- host-fed pass --------------------------------------------------------------------------
    RESULT_KEYS = ("face_dets", "face_count", "person_dets", "person_count", "ids", "sims", "keypoints", "scores")

    def bind_host(self, host_inputs: StepInputs, copy_streams: int = 2) -> None:
        """Pin the host-side inputs and allocate pinned result buffers.  The H2D copies of a step are spread over
        ``copy_streams`` streams (largest tensors first, balanced by bytes): one stream already saturates a PCIe link on
        large copies, the second one hides the issue gaps between the dozen small tensors."""
        self._host_in = {k: (v if v.is_pinned() else v.contiguous().pin_memory()) for k, v in host_inputs.tensors().items()}
        self._host_out = {k: torch.empty(self.out[k].shape, dtype=self.out[k].dtype).pin_memory() for k in self.RESULT_KEYS}
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self._host_in.values())
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self._host_out.values())
        self.copy_streams = max(1, int(copy_streams))
        self._copy = [torch.cuda.Stream(self.device) for _ in range(self.copy_streams)]
        load = [0] * self.copy_streams
        self._copy_plan = [[] for _ in range(self.copy_streams)]
        for k, t in sorted(self._host_in.items(), key=lambda kv: -kv[1].numel() * kv[1].element_size()):
            j = load.index(min(load))
            self._copy_plan[j].append(k)
            load[j] += t.numel() * t.element_size()

    def run_host(self) -> Dict[str, torch.Tensor]:
        """H2D of every input from pinned memory, one pass, D2H of the compact results (async on the
        pipeline stream; call ``stream.synchronize()`` to wait)."""
        dst = self.inp.tensors()
        main = self._stream
        ready = torch.cuda.Event()
        ready.record(main)                       # the previous step's kernels have read the device inputs
        for cs, names in zip(self._copy, self._copy_plan):
            cs.wait_event(ready)
            with torch.cuda.stream(cs):
                for k in names:
                    dst[k].copy_(self._host_in[k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(cs)
            main.wait_event(done)
        with torch.cuda.stream(main):
            self._launch()
            for k, buf in self._host_out.items():
                buf.copy_(self.out[k], non_blocking=True)
        return self._host_out

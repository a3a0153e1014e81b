"""CPU-side checks: the C-ABI library loads and exports what include/*.h declares, the host logic
(key packing, shard bounds, shims' argument validation) and the world_size-2 gallery reduction over gloo."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import match as omatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(spp):
    spp.build()                                   # nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(spp._lib.LIB_PATH)
    declared = set()
    for header in ("spp.h", "spp_internal.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(spp_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), f"libspp.so does not export {name}"
    assert declared == set(spp._lib.SIGNATURES), "ctypes table and headers disagree"
    assert lib.spp_abi_version() == 2


def test_no_cpu_fallback(spp):
    with pytest.raises(RuntimeError, match="CUDA"):
        spp.non_max_suppression(torch.zeros(1, 5, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        spp.get_keypoints_from_heatmaps(torch.zeros(1, 17, 64, 48))
    with pytest.raises(RuntimeError, match="CUDA"):
        spp.crop_affine(torch.zeros(1, 3, 8, 8), torch.zeros(1, 4), torch.zeros(1, dtype=torch.int32))
    if not torch.cuda.is_available():
        L = spp._lib.lib()
        assert L.spp_device_sm_count() < 0 and b"no CUDA device" in L.spp_last_error()
    # workspace queries are pure host arithmetic
    assert spp._lib.lib().spp_nms_workspace_bytes(64, 19320, 1, 0) > 64 * 19320 * 8
    assert spp._lib.lib().spp_match_workspace_bytes(640, 10000, 512) > 640 * 512 * 6
    assert spp._lib.lib().spp_match_workspace_bytes(640, 10000, 256) == 0
    # crop: header + (out_w + out_h) table entries per crop (8 B for fp32 frames, 16 B for uint8) + one 64-byte descriptor per slab
    L = spp._lib.lib()
    assert L.spp_crop_workspace_bytes(0, 256, 192, 0) == 0
    assert L.spp_crop_workspace_bytes(640, 256, 192, 0) == 256 + 640 * (448 * 8 + 32 * 64)
    assert L.spp_crop_workspace_bytes(640, 256, 192, 1) == 256 + 640 * (448 * 16 + 32 * 64)
    assert spp.crop_workspace_bytes(640) == L.spp_crop_workspace_bytes(640, 256, 192, 0)
    # implementation policy and CTA budgets are process-wide host state: set, query, restore
    prev = L.spp_crop_policy(2)
    assert L.spp_crop_policy(-1) == 2 and L.spp_crop_policy(7) == -1
    L.spp_crop_policy(prev)
    assert L.spp_crop_policy(-1) == prev
    assert L.spp_set_launch_limit(2, 48) == 0 and L.spp_set_launch_limit(2, -1) == 48 and L.spp_set_launch_limit(2, 0) == 48
    assert L.spp_set_launch_limit(3, 1) == -1


def test_key_packing_round_trip_and_order(spp):
    d = spp.dist
    g = torch.Generator().manual_seed(0)
    sims = torch.cat([torch.randn(1000, generator=g), torch.tensor([0.0, -0.0, 1.0, -1.0, 0.5, 0.5])])
    ids = torch.randint(0, 2 ** 31 - 2, (sims.numel(),), generator=g)
    ids[-1], ids[-2] = 7, 9                                  # equal similarity: lower id must win
    keys = d.pack_keys(sims, ids)
    rid, rs = d.unpack_keys(keys)
    assert torch.equal(rid, ids) and torch.equal(rs.abs(), sims.abs())
    order = torch.argsort(keys, descending=True)
    assert (sims[order][:-1] >= sims[order][1:]).all()
    assert keys[-1] > keys[-2]                               # id 7 beats id 9 at equal similarity
    none = d.pack_keys(torch.tensor([0.3]), torch.tensor([-1]))
    assert none < keys.min() and d.unpack_keys(none)[0].item() == -1
    assert d.unpack_keys(d.pack_keys(torch.tensor([0.3]), torch.tensor([5])), threshold=0.4)[0].item() == -1
    assert [d.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, m, out):
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("person-recognition-for-pose-estimation_b200.dist")
    synth = importlib.import_module("person-recognition-for-pose-estimation_b200.synth")
    ms = synth.make_match_set(world * m, n, seed=3)
    gal = ms.gallery.to(torch.bfloat16).float()
    lo, hi = d.shard_bounds(n, world, rank)

    def local(p):      # oracle-based stand-in for the GPU kernel: top-1 of every probe in this shard
        ids, sims = omatch.match_top1(p, gal[lo:hi])
        return d.pack_keys(sims, ids + lo)

    matcher = d.ShardedGalleryMatcher(local, threshold=0.4)
    ids, sims = matcher.match(ms.embeddings[rank * m:(rank + 1) * m])
    ref_ids, ref_sims = omatch.match_top1(ms.embeddings[rank * m:(rank + 1) * m], gal, threshold=0.4)
    out[rank] = bool(torch.equal(ids, ref_ids) and torch.allclose(sims, ref_sims, rtol=0, atol=1e-6))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gallery_top1_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, 301, 12, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _peer_worker(rank, world, port, out):
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("person-recognition-for-pose-estimation_b200.dist")
    opened = []

    def alloc(nbytes):                       # stand-ins for spp_peer_alloc / spp_peer_open (those need a GPU)
        return 0x1000 * (rank + 1), bytes([rank]) * 64

    def open_(handle):
        opened.append(handle[0])
        return 0x1000 * (handle[0] + 1) + 1
    g = d.PeerGroup(24, alloc=alloc, open_=open_)
    want = [0x1000 * (r + 1) + (0 if r == rank else 1) for r in range(world)]
    out[rank] = bool(g.world == world and g.rank == rank and g.ptrs == want and sorted(opened) == [r for r in range(world) if r != rank]
                     and [g.struct.buffers[r] for r in range(world)] == want and g.struct.m_local == 24
                     and g.handles == [bytes([r]) * 64 for r in range(world)])
    g.close()
    dist.barrier()
    dist.destroy_process_group()


def test_peer_group_handle_exchange_gloo_world2():
    """Host side of the peer-memory exchange: every rank publishes its IPC handle, opens every other rank's, and the
    rank-ordered pointer table handed to the kernels (spp_peer_group) has its own buffer at its own rank."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_peer_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_peer_buffer_layout_is_pure_host_arithmetic(spp):
    L = spp._lib.lib()
    b = L.spp_peer_buffer_bytes(8, 640, 512)
    assert b >= 2 * 8 * 640 * 512 * 6 + 2 * 8 * 640 * 8 and b % 1024 == 0
    assert L.spp_peer_buffer_bytes(17, 640, 512) == 0 and L.spp_peer_buffer_bytes(8, 640, 256) == 0
    assert L.spp_sharded_match_workspace_bytes(8, 640, 125000, 512) > 0


def test_synthetic_inputs_are_deterministic(synth):
    a = synth.make_heatmaps(3, 17, seed=5)
    b = synth.make_heatmaps(3, 17, seed=5)
    assert torch.equal(a.heatmaps, b.heatmaps) and torch.equal(a.flipped, b.flipped)
    assert synth.num_anchors(736, 1280) == 19320 and synth.num_anchors(640, 640) == 8400
    hm = synth.make_head_maps(1, 64, 96, n_obj=2, seed=1)
    assert [tuple(l.shape) for l in hm.levels] == [(1, 65, 8, 12), (1, 65, 4, 6), (1, 65, 2, 3)]


def test_torch_ops_registered_with_fake_kernels(spp):
    """torch.ops.spp.* exist, infer shapes without a GPU (fake tensors), and have no CPU implementation."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("head_decode", "nms_decoded", "decode_nms", "l2_normalize", "match_top1", "crop_affine", "heatmap_decode"):
        assert hasattr(torch.ops.spp, name)
    with FakeTensorMode():
        hm = torch.empty(7, 17, 64, 48)
        kp, sc, am = torch.ops.spp.heatmap_decode(hm, None, None, None, "dark", 11, 0)
        assert kp.shape == (7, 17, 2) and sc.shape == (7, 17) and am.dtype == torch.int32
        levels = [torch.empty(2, 65, 32, 40), torch.empty(2, 65, 16, 20), torch.empty(2, 65, 8, 10)]
        assert torch.ops.spp.head_decode(levels, [8.0, 16.0, 32.0]).shape == (2, 5, 1680)
        dets, cnt, keys = torch.ops.spp.decode_nms(levels, [8.0, 16.0, 32.0], 0.001, 0.65, 300)
        assert dets.shape == (2, 300, 6) and cnt.shape == (2,) and keys.shape == (2, 300)
        ids, sims = torch.ops.spp.match_top1(torch.empty(5, 512), torch.empty(100, 512, dtype=torch.bfloat16), 0.4, 0)
        assert ids.shape == (5,) and ids.dtype == torch.int32 and sims.shape == (5,)
        pix = torch.ops.spp.crop_affine(torch.empty(1, 3, 64, 64), torch.empty(3, 4), torch.empty(3, dtype=torch.int32), 256, 192,
                                        [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
        assert pix.shape == (3, 3, 256, 192)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.spp.l2_normalize(torch.zeros(2, 512))          # CPU tensors: no kernel registered


def test_bench_roi_bytes_union_vs_sum():
    """bench.py's two source-byte figures for the crop roofline: every crop's own window and the union of the
    windows per frame (what has to cross HBM at least once)."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    boxes = torch.tensor([[100.0, 100.0, 60.0, 80.0], [100.0, 100.0, 60.0, 80.0], [700.0, 300.0, 60.0, 80.0]])
    same_frame = torch.tensor([0, 0, 0], dtype=torch.int32)
    other_frame = torch.tensor([0, 1, 2], dtype=torch.int32)
    total = bench.roi_bytes(boxes, 720, 1280)
    u_same = bench.roi_union_bytes(boxes, same_frame, 720, 1280)
    u_other = bench.roi_union_bytes(boxes, other_frame, 720, 1280)
    assert u_same < u_other                                   # two identical boxes in one frame are read once
    assert abs(u_same / u_other - 2.0 / 3.0) < 0.02
    assert 0.9 < u_other / total < 1.1                        # disjoint windows: union = sum (up to the rounding of the window edges)

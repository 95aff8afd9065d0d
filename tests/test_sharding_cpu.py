"""Host logic of the multi-GPU path, on CPU with the gloo backend (world_size 2): the view sharding used by
bench.py (rank r renders views [8r, 8r+8) of the shared cloud, no data-path collective) and the only collective
of the job, an all-gather of per-view PSNR.  The CUDA op itself is not called here (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pf3plat_b200.cameras import make_view_batch
from pf3plat_b200.sharding import (SharedCloudUploader, allreduce_scene_gradients, gather_metric, interleave_views,
                                   shard_views)
from pf3plat_b200.synthetic import make_scene


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        views = list(shard_views(8, rank, world))
        psnr_local = torch.tensor([float(v) for v in views])
        allp = gather_metric(psnr_local)
        sc = make_scene(100, len(views), 32, 32, first_view=views[0], total_views=8)
        # one scene's views split over the ranks: per-Gaussian gradient blocks are summed by ONE all-reduce
        g_means, g_sh = torch.full((100, 3), float(rank + 1)), torch.full((100, 25, 3), 10.0 * (rank + 1))
        allreduce_scene_gradients([g_means, None, g_sh])
        # shared-cloud upload: each rank copies its 1/N of the rows, one all-gather per array fills in the rest (P = 101 is
        # not a multiple of the world size: the padding rows stay outside the returned views)
        g = torch.Generator().manual_seed(5)
        host = {"means3D": torch.randn(101, 3, generator=g), "shs": torch.randn(101, 25, 3, generator=g)}
        up = SharedCloudUploader(host, torch.device("cpu"))
        got = up.upload()
        same = all(torch.equal(got[k], host[k]) for k in host) and up.bytes_per_step < sum(v.numel() * 4 for v in host.values())
        out.put((rank, views, allp.tolist(), sc.extrinsics[:, :2, 3].tolist(), float(g_means[7, 1]), float(g_sh[3, 2, 1]), same))
    finally:
        dist.destroy_process_group()


def test_view_sharding_and_psnr_gather_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, v0, g0, t0, m0, s0, u0), (r1, v1, g1, t1, m1, s1, u1) = res
    assert u0 and u1                                                # every rank ends up with the whole cloud
    assert m0 == m1 == 3.0 and s0 == s1 == 30.0                    # 1 + 2 and 10 + 20 on both ranks
    assert v0 == [0, 1, 2, 3] and v1 == [4, 5, 6, 7]               # contiguous, disjoint, complete
    assert g0 == g1 == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]    # every rank sees every view's metric, in view order
    full = make_scene(100, 8, 32, 32).extrinsics[:, :2, 3].tolist()
    assert t0 + t1 == full                                          # shards reproduce the unsharded camera path


def test_interleaved_views_cover_everything_once_and_match_the_strided_cameras():
    from pf3plat_b200.synthetic import make_cameras
    for n, world in [(8, 1), (64, 8), (10, 4), (3, 5)]:
        got = sorted(v for r in range(world) for v in interleave_views(n, r, world))
        assert got == list(range(n))
    full = make_cameras(64, 32, 32, total_views=64)[0]
    for r in (0, 3, 7):
        mine = interleave_views(64, r, 8)
        ext = make_cameras(len(mine), 32, 32, first_view=mine[0], total_views=64, view_stride=8)[0]
        assert torch.equal(ext, full[list(mine)])


def test_shard_views_covers_ragged_splits():
    for n, world in [(8, 1), (8, 3), (5, 4), (32, 8), (3, 8)]:
        got = [v for r in range(world) for v in shard_views(n, r, world)]
        assert got == list(range(n))
        sizes = [len(shard_views(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_view_batch_has_no_host_sync_inputs():
    sc = make_scene(10, 3, 16, 16)
    vb = make_view_batch(sc.extrinsics, sc.intrinsics, sc.near, sc.far)
    assert vb.viewmatrix.shape == (3, 4, 4) and vb.tanfov.shape == (3, 2)
    assert torch.allclose(vb.tanfov, torch.full((3, 2), 0.5 / 0.86), atol=1e-5)

"""Multi-rank host logic on CPU: world_size 2, gloo backend.  The local computation is the
oracle (cKDTree) so what is under test is the sharding, padding, all-gather / OR / MIN
assembly that the NCCL path uses unchanged."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anemoi_transform_b200 import distributed as atd


def test_shard_range_covers_everything_in_multiples():
    for n in (0, 1, 7, 64, 3120, 542_080):
        for ws in (1, 2, 3, 4, 8):
            for mult in (1, 4):
                spans = [atd.shard_range(n, r, ws, mult) for r in range(ws)]
                assert spans[0][0] == 0 and spans[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
                assert all((lo % mult == 0) for lo, _ in spans if lo < n)
    # 3120 fields over 8 ranks in multiples of 4: (u, v) / (q, t) pairs never straddle ranks
    assert all((hi - lo) % 4 == 0 for lo, hi in (atd.shard_range(3120, r, 8, 4) for r in range(8)))
    with pytest.raises(ValueError):
        atd.shard_range(10, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, ws: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from scipy.spatial import cKDTree

        from anemoi_transform_b200 import synthetic as syn
        from oracle import spatial as osp

        src = np.array(osp.latlon_to_xyz(*syn.regular_latlon(4.0))).T
        tgt = np.array(osp.latlon_to_xyz(*syn.octahedral(13))).T  # odd count: ragged shards
        tree = cKDTree(src)

        # sharded kNN: indices all-gathered
        idx = atd.sharded_query(tgt.shape[0], lambda lo, hi: torch.from_numpy(tree.query(tgt[lo:hi], k=3)[1]))
        # ball union: partial marks OR-ed
        lam = np.array(osp.latlon_to_xyz(*syn.rotated_lam(9, 11, 1.5))).T
        lo, hi = atd.shard_range(lam.shape[0], rank, ws)
        mark = torch.zeros(src.shape[0], dtype=torch.uint8)
        for sub in tree.query_ball_point(lam[lo:hi], 0.08):
            mark[sub] = 1
        mark = atd.all_reduce_or(mark)
        # resolution: MIN of per-shard minima
        lo, hi = atd.shard_range(src.shape[0], rank, ws)
        res = atd.all_reduce_min(float(tree.query(src[lo:hi], k=2)[0][:, 1].min()))
        # field sharding: no collective, each rank owns a slice
        mine = atd.shard_fields(list(range(22)), multiple=4)
        # fewer queries than ranks: the trailing rank owns an empty range and must still take part
        # in the collective (it used to raise before the all-gather and hang the others)
        one = atd.sharded_query(1, lambda lo, hi: torch.from_numpy(np.asarray(tree.query(tgt[lo:hi], k=3)[1]).reshape(hi - lo, 3)))
        # interleaved split (rank::ws) and its inverse
        full = torch.arange(23 * 2, dtype=torch.int64).reshape(23, 2)
        inter = atd.all_gather_strided(full[rank::ws].contiguous(), 23)
        inter1 = atd.all_gather_strided(torch.arange(7)[rank::ws].contiguous(), 7)
        # host trigonometry sharded over the ranks, slices all-gathered: bitwise the single-process xyz
        atd.LATLON_SHARD_MIN_POINTS = 0
        lat, lon = syn.octahedral(13)
        xyz = atd.latlon_to_xyz_device(lat, lon, _device="cpu")
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx.numpy(), mark=mark.numpy(), res=res, mine=np.array(mine), one=one.numpy(), xyz=np.stack([t.numpy() for t in xyz]), inter=inter.numpy(), inter1=inter1.numpy())
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    from scipy.spatial import cKDTree

    from anemoi_transform_b200 import synthetic as syn
    from oracle import spatial as osp

    ws = 2
    mp.spawn(_worker, args=(ws, _free_port(), str(tmp_path)), nprocs=ws, join=True)
    src = np.array(osp.latlon_to_xyz(*syn.regular_latlon(4.0))).T
    tgt = np.array(osp.latlon_to_xyz(*syn.octahedral(13))).T
    lam = np.array(osp.latlon_to_xyz(*syn.rotated_lam(9, 11, 1.5))).T
    tree = cKDTree(src)
    want_idx = tree.query(tgt, k=3)[1]
    want_mark = np.zeros(src.shape[0], dtype=np.uint8)
    for sub in tree.query_ball_point(lam, 0.08):
        want_mark[sub] = 1
    want_res = osp.resolution(src)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(ws)]
    for o in outs:  # every rank holds the full result
        assert np.array_equal(o["idx"], want_idx)
        assert np.array_equal(o["mark"], want_mark) and want_mark.sum() > 0
        assert float(o["res"]) == want_res
        assert np.array_equal(o["one"], want_idx[:1])
        assert np.array_equal(o["inter"], np.arange(46).reshape(23, 2)) and np.array_equal(o["inter1"], np.arange(7))
        assert np.array_equal(o["xyz"].view(np.uint64), np.array(osp.latlon_to_xyz(*syn.octahedral(13))).view(np.uint64))
    assert sorted(np.concatenate([o["mine"] for o in outs]).tolist()) == list(range(22))
    assert outs[0]["mine"].tolist() == list(range(12))  # 22 fields, multiples of 4 → 12 + 10

"""Multi-GPU sharding of the hot path — one process per GPU, `torch.distributed`.

What shards:

    regrid SpMM + epilogue   fields are independent columns → each rank regrids its own
                             contiguous slice of fields; the CSR matrix is replicated;
                             NO collective on the math path
    kNN / thinning / cutout  queries are independent → sources (and buckets) replicated,
                             queries split; NCCL all-gather of the int64 indices / uint8
                             classifications when every rank needs the full result
    global_on_lam_mask       LAM queries split; each rank marks a uint8[n_global] partial mask;
                             all-reduce(MAX) (= bitwise OR on 0/1 bytes), then compaction
    _resolution              self-queries split; all-reduce(MIN) of one float64

The collectives are a few MB at most (542,080 int64 indices = 4.3 MB; a 6.6 M-byte mask), so
over NVLink/NVSwitch they are latency-bound; nothing here is fused with a compute kernel
because no compute step consumes the gathered data on the device.

The functions take the local computation as a callable so the same code runs on NCCL (GPU)
and on gloo (CPU tensors, used by the world_size-2 tests).
"""

from __future__ import annotations

from typing import Callable

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int, multiple: int = 1) -> tuple[int, int]:
    """Contiguous [lo, hi) of `n` items for `rank`; shard sizes are multiples of `multiple`
    (e.g. 4 keeps (u, v) / (q, t) partner columns and float4 groups on one rank)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    units = -(-n // multiple)
    per = -(-units // world_size)
    lo = min(n, rank * per * multiple)
    hi = min(n, (rank + 1) * per * multiple)
    return lo, hi


def shard_fields(fields: list, multiple: int = 4) -> list:
    """This rank's slice of a list of fields (regrid shards by field, no collective)."""
    rank, ws = world()
    lo, hi = shard_range(len(fields), rank, ws, multiple)
    return fields[lo:hi]


def all_gather_rows(local, n_total: int):
    """Concatenate per-rank row blocks [hi-lo, …] (contiguous shards of `n_total` rows, as made
    by `shard_range`) into the full [n_total, …] tensor on every rank."""
    import torch

    dist = _dist()
    rank, ws = world()
    if ws == 1:
        return local
    per = -(-n_total // ws)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((ws * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return out[:n_total]


def all_reduce_or(mask):
    """Element-wise OR of 0/1 uint8 masks across ranks (NCCL has no bitwise OR: MAX on bytes)."""
    dist = _dist()
    if world()[1] > 1:
        dist.all_reduce(mask, op=dist.ReduceOp.MAX)
    return mask


def all_reduce_min(value: float, device=None) -> float:
    import torch

    dist = _dist()
    if world()[1] == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def sharded_query(n_queries: int, local_query: Callable[[int, int], "object"]):
    """Run `local_query(lo, hi)` on this rank's query range and all-gather the row blocks."""
    rank, ws = world()
    lo, hi = shard_range(n_queries, rank, ws)
    return all_gather_rows(local_query(lo, hi), n_queries)


# ---- GPU entry points ---------------------------------------------------------------------
def nearest_grid_points(source_latitudes, source_longitudes, target_latitudes, target_longitudes, max_distance=None, num_neighbours_to_return: int = 1):
    """`spatial.nearest_grid_points` with the target points sharded over the ranks; every
    rank returns the full index array (numpy int64)."""
    from . import spatial
    from .device import KnnIndex, to_device_f64

    index = KnnIndex(spatial.latlon_to_xyz(source_latitudes, source_longitudes))
    tx = spatial.latlon_to_xyz(np.asarray(target_latitudes), np.asarray(target_longitudes))
    k = int(num_neighbours_to_return)
    ub = float("inf") if max_distance is None else float(max_distance)

    def local(lo: int, hi: int):
        q = tuple(to_device_f64(a[lo:hi]) for a in tx)
        return index.query(q, k=k, distance_upper_bound=ub)[0]

    idx = sharded_query(tx[0].shape[0], local)
    idx = idx[:, 0] if k == 1 else idx
    return idx.cpu().numpy()


def global_on_lam_mask(lats, lons, global_lats, global_lons, distance_km=None):
    """`spatial.global_on_lam_mask` with the LAM points sharded over the ranks."""
    from . import spatial
    from .constants import R_earth_km
    from .device import KnnIndex, compact_mask

    rank, ws = world()
    global_index = KnnIndex(spatial.latlon_to_xyz(global_lats, global_lons))
    lam_xyz = spatial.latlon_to_xyz(np.asarray(lats), np.asarray(lons))
    if isinstance(distance_km, (int, float)):
        distance = distance_km / R_earth_km
    else:
        src = KnnIndex(lam_xyz) if distance_km == "lam" else global_index
        lo, hi = shard_range(src.n, rank, ws)
        distance = all_reduce_min(src.min_nn_distance(lo, hi - lo), device="cuda")
    lo, hi = shard_range(lam_xyz[0].shape[0], rank, ws)
    mark = global_index.ball_mark(tuple(a[lo:hi] for a in lam_xyz), distance)
    indices = compact_mask(all_reduce_or(mark)).cpu().numpy()
    return indices if indices.size else np.array(sorted(set()))


def cutout_mask(lats, lons, global_lats, global_lons, **kwargs):
    """`spatial.cutout_mask` with the cropped global points (kNN + triangle classification)
    sharded over the ranks, uint8 results all-gathered; every rank returns the full mask."""
    from . import spatial

    return spatial.cutout_mask(lats, lons, global_lats, global_lons, _sharded=True, **kwargs)


def thinning_mask(lats, lons, global_lats, global_lons, **kwargs):
    """`spatial.thinning_mask` with the cropped global points sharded over the ranks."""
    from . import spatial

    return spatial.thinning_mask(lats, lons, global_lats, global_lons, _sharded=True, **kwargs)

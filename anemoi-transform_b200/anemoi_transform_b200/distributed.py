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


def all_gather_strided(local, n_total: int):
    """Inverse of the interleaved split `x[rank::world_size]`: per-rank row blocks → the full
    [n_total, …] tensor on every rank.  Interleaving spreads queries whose cost depends on
    where they are (far from every source: a tree walk; near: one bucket ring) evenly over the
    ranks, which contiguous ranges do not — the cropped global points of config 5 are ordered by
    latitude and one contiguous eighth of them took 0.07 s where the whole set takes 0.12 s."""
    import torch

    dist = _dist()
    rank, ws = world()
    if ws == 1:
        return local
    per = -(-n_total // ws)
    tail = tuple(local.shape[1:])
    pad = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((ws * per,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    # out[r * per + j] is item j * ws + r
    return out.view((ws, per) + tail).transpose(0, 1).reshape((ws * per,) + tail)[:n_total]


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


# ---- fused query + all-gather over NVLink peer memory ----------------------------------------
class _DevicePointer:
    """A raw device allocation presented to torch through the CUDA array interface."""

    def __init__(self, ptr: int, shape: tuple[int, ...], typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


class PeerBuffers:
    """One device buffer per rank that every rank of the box can address (cudaIpc over
    NVLink / NVSwitch): `ptrs[r]` is rank r's buffer as seen from this process."""

    def __init__(self, nbytes: int):
        import ctypes

        import torch

        from ._cabi import call

        dist = _dist()
        self.rank, self.world = world()
        self.nbytes = int(nbytes)
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        call("at_peer_alloc", self.nbytes, ctypes.byref(own), handle)
        self._own = own.value
        handles: list = [None] * self.world
        dist.all_gather_object(handles, handle.raw)
        self.ptrs: list[int] = []
        self._opened: list[int] = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self._own)
                continue
            p = ctypes.c_void_p()
            call("at_peer_open", ctypes.create_string_buffer(h, 64), ctypes.byref(p))
            self.ptrs.append(p.value)
            self._opened.append(p.value)
        torch.cuda.synchronize()
        dist.barrier()  # every buffer is zeroed and mapped before anyone writes into a peer

    def close(self) -> None:
        import ctypes

        import torch

        from ._cabi import call

        if self._own is None:
            return
        torch.cuda.synchronize()
        if _dist().is_initialized():
            _dist().barrier()  # nobody is still storing into a buffer about to be freed
        for p in self._opened:
            call("at_peer_close", ctypes.c_void_p(p))
        call("at_peer_free", ctypes.c_void_p(self._own))
        self._own, self._opened, self.ptrs = None, [], []


class ShardedKnnQuery:
    """k nearest sources of a query set sharded over the ranks; every rank ends up with the
    full [n_queries, k] index array in HBM.

    On NCCL-capable boxes the all-gather is fused into the search (`at_knn_query_gather`): the
    query kernels store each index into every rank's gather buffer through NVLink peer mappings
    and a one-warp kernel exchanges arrival flags — no collective launch, no staging copy.  When
    peer mapping is unavailable the queries write straight into this rank's slice of a
    preallocated buffer and NCCL all-gathers it in place (no per-step allocation either)."""

    def __init__(self, index, qxyz, k: int = 1, distance_upper_bound: float = float("inf"), mode: str = "auto", exchange: str = "bulk", graph: bool = True):
        import torch

        from .device import to_device_f64

        self.index, self.k, self.ub = index, int(k), float(distance_upper_bound)
        self.rank, self.world = world()
        self.nq = int(qxyz[0].shape[0])
        # shards of an even number of queries: every rank's slice of the int64 gather buffer then
        # starts on a 16-byte boundary (the bulk exchange copies 16 bytes at a time)
        self.lo, self.hi = shard_range(self.nq, self.rank, self.world, 2)
        self.per = 2 * -(-(-(-self.nq // 2)) // self.world)  # the stride of shard_range(…, multiple=2)
        self.q = tuple(to_device_f64(a[self.lo : self.hi]) for a in qxyz)
        self._all_q = qxyz
        self.epoch = 0
        self.peers = None
        self.exchange = exchange
        self._use_graph, self._graphs = bool(graph), None
        rows = max(1, self.world * self.per)
        self.mode = "local" if self.world == 1 else mode
        if self.mode in ("auto", "peer"):
            try:
                self._slot_bytes = -(-rows * self.k * 8 // 256) * 256
                self.peers = PeerBuffers(2 * self._slot_bytes + 256)
                self.mode = "peer"
            except Exception as e:
                if mode == "peer":
                    raise
                import logging

                logging.getLogger(__name__).warning("peer mapping unavailable (%s): NCCL all-gather instead", e)
                self.mode = "nccl"
        if self.mode == "peer":
            import ctypes

            self._gather = [(ctypes.c_void_p * self.world)(*[p + s * self._slot_bytes for p in self.peers.ptrs]) for s in (0, 1)]
            self._flags = (ctypes.c_void_p * self.world)(*[p + 2 * self._slot_bytes for p in self.peers.ptrs])
            self._views = [torch.as_tensor(_DevicePointer(self.peers.ptrs[self.rank] + s * self._slot_bytes, (rows, self.k), "<i8"), device="cuda") for s in (0, 1)]
            self._error = torch.zeros((1,), dtype=torch.int32, device="cuda")
        else:
            self._out = torch.empty((rows, self.k), dtype=torch.int64, device="cuda")

    def step(self):
        """One sharded query → int64 [n_queries, k] CUDA tensor with every rank's answers (in
        peer mode a view that stays valid until the call after next)."""
        from ctypes import c_void_p

        from ._cabi import EXCHANGE_BULK, EXCHANGE_INLINE, call
        from .device import _ptr, stream_ptr

        n_local = self.hi - self.lo
        if self.mode == "peer":
            self.epoch += 1
            slot = self.epoch & 1

            def launch():  # epoch 0: the library counts the calls on the device (replayable)
                call("at_knn_query_gather", self.index._h, _ptr(self.q[0]), _ptr(self.q[1]), _ptr(self.q[2]), n_local, self.k, self.ub,
                     self._gather[slot], self._flags, self.world, self.rank, self.lo, None, None, 0, _ptr(self._error),
                     EXCHANGE_BULK if self.exchange == "bulk" else EXCHANGE_INLINE, stream_ptr())  # fmt: skip

            if not self._use_graph:
                launch()
            else:
                # A step is three short kernels; at 8 GPUs the host spends longer launching them
                # than the GPU running them.  Each slot's step is captured once (the first two
                # calls run eagerly and warm everything up) and replayed with one graph launch.
                import torch

                if self._graphs is None:
                    self._graphs = {}
                if self.epoch <= 2:
                    launch()
                elif slot not in self._graphs:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        launch()
                    self._graphs[slot] = g
                    g.replay()
                else:
                    self._graphs[slot].replay()
            return self._views[slot][: self.nq]
        mine = self._out[self.lo : self.lo + n_local]
        if n_local:
            call("at_knn_query", self.index._h, _ptr(self.q[0]), _ptr(self.q[1]), _ptr(self.q[2]), n_local, self.k, self.ub, c_void_p(mine.data_ptr()), None, None, stream_ptr())
        if self.mode == "nccl":
            _dist().all_gather_into_tensor(self._out, self._out[self.rank * self.per : (self.rank + 1) * self.per])
        return self._out[: self.nq]

    def timed_out(self) -> bool:
        return self.mode == "peer" and bool(int(self._error.item()))

    def check_against_single_gpu(self) -> bool:
        """The gathered result equals what this rank finds when it runs every query itself."""
        import torch

        got = self.step().clone()
        torch.cuda.synchronize()
        full, _, _ = self.index.query(self._all_q, k=self.k, distance_upper_bound=self.ub)
        return bool(torch.equal(got, full)) and not self.timed_out()

    def describe(self) -> str:
        return {
            "local": "single GPU",
            "peer": "queries split over ranks; indices go into every rank's gather buffer over NVLink peer mappings — "
            + ("one coalesced broadcast kernel after the search whose last CTA exchanges arrival flags" if self.exchange == "bulk" else "stored by the search kernels as each query finishes, then one flag-exchange kernel")
            + " (no NCCL call, no staging copy)",
            "nccl": "queries split over ranks, written in place into the gather buffer, in-place NCCL all-gather of the int64 indices",
        }[self.mode]

    def close(self) -> None:
        self._graphs = None
        if self.peers is not None:
            self._views = []
            self.peers.close()
            self.peers = None


#: below this many points every rank evaluates the whole array itself (not worth a collective)
LATLON_SHARD_MIN_POINTS = 4096


def latlon_to_xyz_device(lat, lon, _device: str = "cuda"):
    """`spatial.latlon_to_xyz` with the points sharded over the ranks → three float64 CUDA tensors
    holding every point on every rank.

    The trigonometry stays numpy on the host (bit-identical xyz is what makes the indices equal
    cKDTree's, spatial.py module docstring) but each rank evaluates only its slice — the same
    numpy calls on the same elements, verified slice-invariant once per process — and the
    slices are all-gathered on the device (3 x 8 bytes per point over NVLink).  In round 1
    every rank repeated the whole evaluation, which is why the sharded config-5 functions were
    slower than one GPU."""
    import torch

    from . import spatial
    from .device import to_device_f64

    dist = _dist()
    rank, ws = world()
    lat, lon = np.asarray(lat, dtype=np.float64).reshape(-1), np.asarray(lon, dtype=np.float64).reshape(-1)
    n = int(lat.size)
    if ws == 1 or n < LATLON_SHARD_MIN_POINTS or not spatial._threaded_trig_is_exact():
        return tuple(to_device_f64(a) for a in spatial.latlon_to_xyz(lat, lon))
    per = -(-n // ws)
    lo, hi = shard_range(n, rank, ws)
    local = torch.zeros((3, per), dtype=torch.float64, device=_device)
    if hi > lo:
        xyz = spatial.latlon_to_xyz(lat[lo:hi], lon[lo:hi])
        local[:, : hi - lo] = torch.from_numpy(np.stack(xyz)).to(_device)
    out = torch.empty((ws * 3, per), dtype=torch.float64, device=_device)
    dist.all_gather_into_tensor(out, local)
    full = out.view(ws, 3, per).permute(1, 0, 2).reshape(3, ws * per)
    return tuple(full[k, :n].contiguous() for k in range(3))


# ---- GPU entry points ---------------------------------------------------------------------
def nearest_grid_points(source_latitudes, source_longitudes, target_latitudes, target_longitudes, max_distance=None, num_neighbours_to_return: int = 1):
    """`spatial.nearest_grid_points` with the target points sharded over the ranks; every
    rank returns the full index array (numpy int64)."""
    from . import spatial
    from .device import KnnIndex

    index = KnnIndex(latlon_to_xyz_device(source_latitudes, source_longitudes))
    tx = latlon_to_xyz_device(target_latitudes, target_longitudes)
    k = int(num_neighbours_to_return)
    ub = float("inf") if max_distance is None else float(max_distance)

    query = ShardedKnnQuery(index, tx, k=k, distance_upper_bound=ub)
    try:
        idx = query.step().clone()
        if query.timed_out():
            raise RuntimeError("sharded nearest_grid_points: a rank did not deliver its indices")
    finally:
        query.close()
    idx = idx[:, 0] if k == 1 else idx
    return idx.cpu().numpy()


def global_on_lam_mask(lats, lons, global_lats, global_lons, distance_km=None):
    """`spatial.global_on_lam_mask` with the LAM points sharded over the ranks."""
    from . import spatial
    from .constants import R_earth_km
    from .device import KnnIndex, compact_mask

    rank, ws = world()
    global_index = KnnIndex(latlon_to_xyz_device(global_lats, global_lons))
    lam_xyz = latlon_to_xyz_device(lats, lons)
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

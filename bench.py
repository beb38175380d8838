#!/usr/bin/env python
"""Benchmark of the field-transform hot path on B200 — BASELINE.json's headline metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config 3 of BASELINE.json, the one the metric is quoted on): regrid 0.25°
(1440x721 = 1,038,240 points) → N320-shaped (542,080 points) with a 4-point bilinear CSR
matrix, 3120 float32 fields (10 vars x 13 levels x 24 steps).  One *step* = one pass of the
matrix over the 3120-field batch.  Synthetic data (no network): seeded random fields, a
locally built bilinear matrix, an N320-shaped reduced Gaussian grid (the real N320 `pl`
table is not available offline — same point count and density).

One JSON line on stdout (rank 0):
    value     fields/s with the batch resident in HBM (device-timed, max over ranks)
    e2e       fields/s through the C-ABI host pipeline (at_pipeline_regrid): pinned HOST
              buffers in, HOST buffers out, H2D + D2H inside the timed region
    roofline  algorithmic bytes per launch / measured launch time vs the measured HBM peak
    cpu_baseline  the oracle's C port of scipy's csr_matvec on all host cores (bounded sample)
    knn       config 2: N320 targets vs 0.25° sources, k=1 (queries sharded over ranks,
              NCCL all-gather of the indices)

Multi-GPU: one process per GPU (torchrun); fields shard with no collective on the math path,
every rank regrids its own 3120-field batch of an N x 3120-field job ("scaling": "weak").

--impl reference times the reference's CPU path (per-field csr_matvec, regrid.py:204-208,
309-310) on the host cores: the oracle's plain-C restatement of scipy's csr_matvec with
OpenMP over fields (the reference is pure Python + scipy; nothing to compile into
oracle/_ref), each step a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

N_FIELDS = 3120
WORKLOAD = "regrid 0.25deg (1440x721=1,038,240 pts) -> N320-shaped (542,080 pts), 4-nnz bilinear CSR, 3120 float32 fields"
# dram__bytes_read.sum + dram__bytes_write.sum of spmm_f32_kernel per launch on this workload,
# from the `ncu --set full` capture summarised in profiles/ (None until captured).
NCU_TRAFFIC_BYTES_PER_LAUNCH = 18_984_274_000  # profiles/r01_spmm_ncu.md: 12.237 GB read + 6.747 GB write
CPU_SAMPLE_FIELDS = 624  # 1/5 of the workload per CPU step


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs() -> tuple[float, str]:
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload():
    from anemoi_transform_b200 import synthetic as syn

    t_lat, t_lon = syn.n320_like()
    data, idx, ptr, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
    n_src_ref = int(np.unique(idx).size)
    alg_bytes = 4 * N_FIELDS * (n_src_ref + shape[0]) + 8 * data.size + 4 * (shape[0] + 1)
    return dict(data=data, idx=idx, ptr=ptr, shape=shape, n_src_ref=n_src_ref, alg_bytes=alg_bytes, t_lat=t_lat, t_lon=t_lon)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).

    nvidia-smi takes ~100 ms to produce its first row, longer than a 20-step timed region, so the
    sampler is started before the warm-up and `mark()` brackets the timed region: only rows that
    arrived between the two marks (plus the first one after) are reported."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int, period_ms: int = 10):
        self.rows, self.proc, self.gpu, self.period_ms = [], None, gpu_index, period_ms
        self.marks: list[int] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms), "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception as e:  # nvidia-smi missing: report that rather than fail the bench
            log("clock sampler unavailable:", e)
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_for_first_row(self, timeout_s: float = 5.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def mark(self):
        self.marks.append(len(self.rows))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(3 * self.period_ms / 1000.0)
        self.proc.terminate()
        lo, hi = (self.marks + [0, len(self.rows)])[:2] if len(self.marks) >= 2 else (0, len(self.rows))
        rows = self.rows[max(0, lo - 1) : hi + 1] or self.rows
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "period_ms": self.period_ms}


def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank (and allocate its pinned buffers) on the NUMA node its GPU hangs off, so
    H2D / D2H DMA does not cross the inter-socket link.  Best effort: sysfs may be hidden."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception as e:
        log("NUMA binding skipped:", e)
    return None


def cpu_baseline(w, steps: int = 3) -> dict:
    """The oracle's C port of the reference's per-field csr_matvec loop, all host threads."""
    from oracle import spmm as ospmm

    rng = np.random.default_rng(0)
    x = rng.standard_normal((CPU_SAMPLE_FIELDS, w["shape"][1]), dtype=np.float32)
    threads = ospmm.c_max_threads()
    ospmm.c_regrid_fields_f32(w["ptr"], w["idx"], w["data"], x[:threads])  # warm-up
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ospmm.c_regrid_fields_f32(w["ptr"], w["idx"], w["data"], x)
        times.append(time.perf_counter() - t0)
    # the reference exactly as it runs it: scipy, one field at a time, one thread
    from scipy.sparse import csr_array

    m = csr_array((w["data"], w["idx"], w["ptr"]), shape=w["shape"])
    m @ x[0]
    t0 = time.perf_counter()
    for f in range(16):
        m @ x[f]
    scipy_1t = 16 / (time.perf_counter() - t0)
    return {
        "value": CPU_SAMPLE_FIELDS / min(times),
        "unit": "fields/s",
        "cores": threads,
        "kind": "port",
        "sample": f"{CPU_SAMPLE_FIELDS} of the {N_FIELDS} fields per step (same matrix and grids), best of {steps}; C port of scipy csr_matvec, OpenMP over fields, host memory in and out",
        "scipy_single_thread_fields_per_s": scipy_1t,
        "host_cpus": os.cpu_count(),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import spmm as ospmm

    w = build_workload()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((CPU_SAMPLE_FIELDS, w["shape"][1]), dtype=np.float32)
    threads = ospmm.c_max_threads()
    for _ in range(args.warmup):
        ospmm.c_regrid_fields_f32(w["ptr"], w["idx"], w["data"], x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ospmm.c_regrid_fields_f32(w["ptr"], w["idx"], w["data"], x)
    dt = time.perf_counter() - t0
    value = CPU_SAMPLE_FIELDS * args.steps / dt
    sample = f"each step = {CPU_SAMPLE_FIELDS} of the {N_FIELDS} fields; plain-C port of scipy csr_matvec (the reference's `matrix @ data`, regrid.py:309-310), OpenMP over fields"
    print(
        json.dumps(
            {
                "impl": "reference",
                "metric": "regrid_fields_per_s",
                "value": value,
                "unit": "fields/s",
                "n_gpus": args.gpus,
                "steps": args.steps,
                "warmup": args.warmup,
                "ms_per_step": dt / args.steps * 1e3,
                "higher_is_better": True,
                "scaling": "weak",
                "vs_baseline": None,
                "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "sample_fields_per_step": CPU_SAMPLE_FIELDS},
                "cpu_baseline": {"value": value, "unit": "fields/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "fields/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
        ),
        flush=True,
    )


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from anemoi_transform_b200 import _cabi
    from anemoi_transform_b200.device import CsrMatrix, HostPipeline, KnnIndex

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    _cabi.load(check_device=True)  # fail loudly if the CUDA library or the device is missing
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = build_workload()
    n_tgt, n_src = w["shape"]
    csr = CsrMatrix(w["data"], w["idx"], w["ptr"], w["shape"])

    # ---- device-resident: value + roofline -------------------------------------------------
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    X = torch.empty((n_src, N_FIELDS), device=dev, dtype=torch.float32)
    for c0 in range(0, N_FIELDS, 260):  # fill in slabs: randn of the whole 13 GB would need a second copy
        X[:, c0 : c0 + 260] = torch.randn((n_src, min(260, N_FIELDS - c0)), device=dev, dtype=torch.float32, generator=gen) * 15.0 + 280.0
    Y = torch.empty((n_tgt, N_FIELDS), device=dev, dtype=torch.float32)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        csr.apply(X, out=Y, variant=args.variant)
    if rank == 0:
        sampler.wait_for_first_row()
    barrier()
    if rank == 0:
        sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        csr.apply(X, out=Y, variant=args.variant)
    ev1.record()
    barrier()
    if rank == 0:
        sampler.mark()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    value = N_FIELDS * world / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak_gbs()
    achieved = w["alg_bytes"] / (ms_per_step * 1e-3) / 1e9

    # parity spot check outside the timed region (rank 0): sampled columns bit-exact vs scipy
    parity = None
    if rank == 0:
        from scipy.sparse import csr_array

        m = csr_array((w["data"], w["idx"], w["ptr"]), shape=w["shape"])
        parity = True
        for col in (0, 1777, N_FIELDS - 1):
            ref = m @ X[:, col].cpu().numpy()
            parity = parity and bool(np.array_equal(ref.view(np.uint32), Y[:, col].cpu().numpy().view(np.uint32)))

    if args.value_only:
        if rank == 0:
            print(json.dumps({"metric": "regrid_fields_per_s", "value": value, "unit": "fields/s", "ms_per_step": ms_per_step, "partial": "--value-only", "roofline": {"achieved": achieved, "peak": peak, "frac": achieved / peak}, "parity_spot_check": parity}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- kNN (config 2) ---------------------------------------------------------------------
    from anemoi_transform_b200 import spatial
    from anemoi_transform_b200 import synthetic as syn

    s_xyz = spatial.latlon_to_xyz(*syn.regular_latlon(0.25))
    t_xyz = spatial.latlon_to_xyz(w["t_lat"], w["t_lon"])
    t_build0 = time.perf_counter()
    knn = KnnIndex(s_xyz)
    torch.cuda.synchronize()
    knn_build_ms = (time.perf_counter() - t_build0) * 1e3
    nq = t_xyz[0].shape[0]
    per = -(-nq // world)
    lo, hi = min(rank * per, nq), min((rank + 1) * per, nq)
    q = tuple(torch.from_numpy(a[lo:hi]).to(dev) for a in t_xyz)

    def knn_step():
        idx, _, _ = knn.query(q, k=1)
        if world > 1:
            pad = torch.full((per, 1), -1, dtype=torch.int64, device=dev)
            pad[: hi - lo] = idx
            out = torch.empty((world * per, 1), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(out, pad)
            return out
        return idx

    for _ in range(3):
        knn_step()
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    knn_iters = 20
    k0.record()
    for _ in range(knn_iters):
        knn_step()
    k1.record()
    barrier()
    knn_ms = max_over_ranks(k0.elapsed_time(k1)) / knn_iters
    knn_info = {
        "workload": "N320-shaped targets (542,080) vs 0.25deg sources (1,038,240), k=1, float64, exact",
        "queries_per_s": nq / (knn_ms * 1e-3),
        "ms_per_query_batch": knn_ms,
        "build_ms": knn_build_ms,
        "sharding": "queries split over ranks, NCCL all-gather of int64 indices" if world > 1 else "single GPU",
    }

    # ---- end to end through the C-ABI host pipeline ------------------------------------------
    del X, Y
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    # Pinned host buffers.  Every step moves all 3120 fields in and out over PCIe; when host
    # memory is short (many ranks on one box) the 3120 field pointers cycle over a smaller
    # pool of distinct pinned buffers — same bytes copied, smaller host footprint.
    import psutil

    per_field = 4 * (n_src + n_tgt)
    budget = 0.6 * psutil.virtual_memory().available / max(world, 1)
    pool = int(max(2 * args.chunk, min(N_FIELDS, budget // per_field)))
    host_in = torch.empty((pool, n_src), dtype=torch.float32).pin_memory()
    host_out = torch.empty((pool, n_tgt), dtype=torch.float32).pin_memory()
    rng = np.random.default_rng(99 + rank)
    hin = host_in.numpy()
    for f0 in range(0, pool, 64):
        hin[f0 : f0 + 64] = rng.standard_normal((min(64, pool - f0), n_src), dtype=np.float32)
    fields_in = [hin[f % pool] for f in range(N_FIELDS)]
    fields_out = [host_out.numpy()[f % pool] for f in range(N_FIELDS)]
    pipe = HostPipeline(csr, chunk_fields=args.chunk)
    pipe.regrid(fields_in[: 2 * args.chunk], fields_out[: 2 * args.chunk])  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.regrid(fields_in, fields_out)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e_parity = None
    if rank == 0:
        from scipy.sparse import csr_array

        m = csr_array((w["data"], w["idx"], w["ptr"]), shape=w["shape"])
        e2e_parity = all(bool(np.array_equal((m @ hin[f]).view(np.uint32), host_out.numpy()[f].view(np.uint32))) for f in (0, pool // 2, pool - 1))
    chunks = -(-N_FIELDS // args.chunk)
    e2e = {
        "value": N_FIELDS * world / e2e_s,
        "unit": "fields/s",
        "h2d_bytes_per_step": 4 * N_FIELDS * n_src,
        "d2h_bytes_per_step": 4 * N_FIELDS * n_tgt,
        "steps": e2e_steps,
        "ms_per_step": e2e_s * 1e3,
        "path": f"at_pipeline_regrid (C-ABI): pinned host fields -> chunked H2D ({args.chunk} fields) -> pack -> SpMM -> unpack -> D2H, 3 streams",
        "parity_spot_check": e2e_parity,
        "gpu_launches_per_step": 3 * chunks,
        "host_pool_fields": pool,
        "numa_node": numa_node,
    }
    pipe.close()

    if rank == 0:
        base = cpu_baseline(w) if world == 1 else None
        line = {
            "metric": "regrid_fields_per_s",
            "value": value,
            "unit": "fields/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "fields_per_gpu": N_FIELDS,
                "layout": "point-major [points x fields] float32 in HBM",
                "l2": "inputs larger than L2 (X 12.96 GB + Y 6.77 GB per step), no flush needed",
                "spmm_variant": args.variant,
                "parallelism": f"fields sharded over {world} GPU(s), no collective on the math path",
            },
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": args.steps,
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                "kernel": "spmm_f32_kernel (one launch per step)",
                "algorithmic_bytes_per_launch": w["alg_bytes"],
                "n_src_referenced": w["n_src_ref"],
                "peak_source": peak_src,
                "frac_of_8TBs_nominal": achieved / 8000.0,
            },
            "cpu_baseline": base,
            "knn": knn_info,
            "parity_spot_check": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reserve_stdout():
    """Keep the process's real stdout for the one JSON line: everything else that writes to
    fd 1 (NCCL prints its version banner there, libraries may print warnings) goes to stderr."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="at_spmm kernel shape (0 = default)")
    ap.add_argument("--chunk", type=int, default=64, help="fields per chunk of the host pipeline")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--value-only", action="store_true", help="device-resident leg only (for ncu captures); prints a partial line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

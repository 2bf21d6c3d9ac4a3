#!/usr/bin/env python
"""bench.py -- retrieved+blended query frames/sec (768-d, k=8) on N B200s, plus the HBM / tensor roofline
of the dominant kernel and the CPU restatement of the reference path timed on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C5|C3|C2|C1] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path (rvc/infer/pipeline.py:497-507: search k=8 -> gather -> 1/score^2
weights -> index_rate blend) over the whole synthetic query batch of the workload.  Default workload C5 =
BASELINE.json configs[4] (32M x 768 IVF-Flat index, nlist 90,509, 1M query frames): it fits one B200
(98.3 GB of list vectors), and it is the configuration the metric is quoted on at 1/2/4/8 GPUs, so the
same total work is sharded over N ranks (strong scaling).  Index lists are sharded by contiguous
centroid ranges; queries are sharded by home rank.

One JSON line is printed by rank 0 (see README / DESIGN.md for every key).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "codename-rvc-fork-3_b200"

METRIC = "retrieved+blended query frames/sec (768-d, k=8)"
UNIT = "frames/s"
INDEX_RATE = 0.75
K = 8


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]),
                    "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measure_fp8_peak(dev, seconds: float = 2.0):
    """MEASURED_PEAKS.json (driver-written) holds a bf16 cuBLAS figure only.  When the coarse step runs with e4m3 operands its
    roofline needs an fp8 denominator measured the same way: torch._scaled_mm (cuBLASLt) on float8_e4m3fn 8192^3, best of 10
    (burst) and back to back for `seconds` (sustained, power-capped clocks).  Returns {"burst", "sustained"} in TFLOP/s or None."""
    try:
        import torch
        n = 8192
        a = torch.randn((n, n), device=dev).to(torch.float8_e4m3fn)
        b = torch.randn((n, n), device=dev).to(torch.float8_e4m3fn).t()        # column-major B as cuBLASLt wants it
        one = torch.ones((), device=dev)
        f = lambda: torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)   # noqa: E731
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        return {"burst": flops / (best * 1e-3) / 1e12, "sustained": flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12}
    except Exception:
        return None


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node (sysfs local_cpulist of the PCI device) before
    any pinned host buffer is allocated, so that the H2D / D2H copies of the e2e leg do not cross the socket
    interconnect when 8 ranks transfer at once.  Best effort: returns a short description or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cpus local to GPU {local_rank}"
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up steps (nvidia-smi needs ~100 ms to deliver its first line) and polls every 20 ms; samples
    are kept if their arrival time falls inside the timed region, which the caller brackets with mark_begin() /
    mark_end().  If the region is shorter than the polling can resolve, the samples of the warm-up steps (same kernels,
    same load) are used as well and `window` says so."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, ln in rows:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 10:
                    continue
                try:
                    sm.append(float(f[2]))
                    mx.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        inside = [r for r in self.lines if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1 + 0.02]
        window = "timed region"
        sm, mx, reasons = parse(inside)
        if len(sm) < 2:
            window = "warm-up + timed region (timed region shorter than the 20 ms polling resolves)"
            sm, mx, reasons = parse([r for r in self.lines if self.t1 is None or r[0] <= self.t1 + 0.02])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path, all host threads
# ------------------------------------------------------------------------------------------------

def cpu_path(oracle, q, cent, list_off, vecs, ids_pos):
    """coarse (faiss BLAS path: sgemm + norms) -> single-list scan + top-8 (canonical order,
    OpenMP over queries) -> pipeline.py:500-506 verbatim (numpy).  Returns blended rows."""
    a, _ = oracle.coarse_assign_blas(q, cent)
    D, I, P = oracle.scan_topk(q, a, list_off, vecs, ids_pos, k=K, want_pos=True)
    with np.errstate(all="ignore"):
        out = oracle.tail_numpy_verbatim(D, np.where(P < 0, -1, P), vecs, INDEX_RATE, q)
    return out, a


def host_sample_by_construction(nlist, d, avg_len, nq, seed=1234):
    """A bounded sample of the workload built entirely on the host (no GPU code involved): all nlist
    centroids (the coarse cost per query is exact), and the lists the sampled queries probe, filled by
    the same construction as synth.fill_index_by_construction."""
    import torch
    g = torch.Generator().manual_seed(seed)
    nb = (nlist + 3) // 4
    centres = torch.randn((nb, d), generator=g) * 0.30
    cent = (centres[torch.arange(nlist) // 4] + torch.randn((nlist, d), generator=g) * 0.03).numpy()
    lists = torch.randint(0, nlist, (nq,), generator=g).numpy()
    q = (cent[lists] + (torch.randn((nq, d), generator=g) * 0.14).numpy()).astype(np.float32)
    import oracle
    a, _ = oracle.coarse_assign_blas(q, cent)                       # which lists will be probed (untimed setup)
    probed = np.unique(a)
    rng = np.random.default_rng(seed)
    sizes = np.zeros(nlist, np.int64)
    sizes[probed] = np.maximum(8, np.floor(avg_len * np.exp(0.35 * rng.standard_normal(len(probed)))).astype(np.int64))
    list_off = np.zeros(nlist + 1, np.int64)
    np.cumsum(sizes, out=list_off[1:])
    tot = int(list_off[-1])
    vecs = np.empty((tot, d), np.float32)
    rep = np.repeat(probed, sizes[probed])
    step = 1 << 18
    for s in range(0, tot, step):
        e = min(tot, s + step)
        vecs[s:e] = cent[rep[s:e]] + (torch.randn((e - s, d), generator=g) * 0.14).numpy()
    return q, cent.astype(np.float32), list_off, vecs, np.arange(tot, dtype=np.int64)


def run_reference_arm(args, cfg):
    """--impl reference: the reference's CPU path (oracle port; faiss cannot be installed) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    import torch
    threads, blas = use_all_host_threads()
    N, d, nq, nlist = cfg
    sample = min(nq, args.cpu_sample)
    t0 = time.time()
    q, cent, list_off, vecs, ids = host_sample_by_construction(nlist, d, N / nlist, sample)
    setup_s = time.time() - t0
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        cpu_path(oracle, q, cent, list_off, vecs, ids)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    val = sample / sec
    cores = threads
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "index_vectors": N, "d": d, "nlist": nlist, "query_frames": nq, "k": K,
                       "nprobe": 1, "index_rate": INDEX_RATE, "note": "each step = bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} of {nq} query frames per step against all {nlist} centroids and the "
                                       f"{int((np.diff(list_off) > 0).sum())} lists they probe (avg {N / nlist:.0f} rows); "
                                       f"oracle port of faiss 1.7.3 + numpy (faiss itself is not installable here); "
                                       f"omp_threads={oracle.num_threads()} torch_threads={torch.get_num_threads()} blas={blas} setup {setup_s:.0f}s"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def parity_sample(gpu_rows, cpu_rows):
    """The blended rows the GPU produced in the first timed step for the CPU-baseline sample, against the CPU restatement
    of the reference path on the same queries (faiss-like BLAS coarse, canonical scan, numpy tail verbatim).  Rows are
    expected to be bit-identical except where the BLAS-order and the canonical-order coarse distances pick different
    centroids on a near-tie (the documented ties of DESIGN.md).  Never raises: the bench line must survive it."""
    try:
        if gpu_rows is None or cpu_rows is None:
            return None
        m = min(len(gpu_rows), len(cpu_rows))
        g, c = np.asarray(gpu_rows[:m]), np.asarray(cpu_rows[:m])
        with np.errstate(all="ignore"):
            same = ((g == c) | (np.isnan(g) & np.isnan(c))).all(axis=1)
            other = ~same
            rel = 0.0
            if other.any():
                den = np.maximum(np.abs(c[other]), 1e-30)
                diff = np.abs(g[other] - c[other]) / den
                rel = float(np.nanmax(np.where(np.isfinite(diff), diff, 0.0)))
        return {"rows": int(m), "rows_bit_identical": int(same.sum()), "rows_differing": int(other.sum()),
                "max_rel_diff_in_differing_rows": rel,
                "against": "CPU port of the reference path on the same queries (BLAS-order coarse like faiss >= 20 queries; "
                           "differing rows = near-tie centroid choices between the BLAS and the canonical distance order)"}
    except Exception as ex:                                    # pragma: no cover
        return {"error": repr(ex)[:200]}


def build_roofline(phase_acc, coarse_on_tensor_cores, pk, *, nst, n_home, owned_q, nlist, d, scan_impl, workload, world,
                   coarse_e4m3=False, fp8_peak=None):
    """Per-phase rooflines from the library's CUDA-event phase times (ms) and the ALGORITHMIC work of SURVEY 8d.
    nst = vectors stored on this rank, n_home = queries whose coarse step runs here, owned_q = queries whose list lives
    here (uniform queries: nq / world).  Returns (dict by phase, name of the dominant phase)."""
    coarse_flops = 2.0 * n_home * nlist * d     # home queries x all centroids (replicated quantizer)
    scan_bytes = nst * (4 * d + 8) + owned_q * 4 * d + owned_q * K * 12      # B_s: every list probed at these sizes
    blend_bytes = owned_q * (K * 4 * d + 8 * d) + owned_q * K * 12           # B_g
    roof = {}
    if "coarse_main" in phase_acc:
        tf = coarse_flops / (phase_acc["coarse_main"] * 1e-3) / 1e12
        tc = coarse_on_tensor_cores
        # MEASURED_PEAKS.json holds two cuBLAS bf16 figures: the burst one (best of ten 0.7 ms GEMMs) for a kernel timed
        # alone, the sustained one (back to back for 4 s, power-capped clocks) for a kernel timed inside a long step.  A
        # launch that runs for tens of milliseconds inside back-to-back steps is in the sustained regime (the clocks
        # line shows sw_power_cap); microsecond launches (C1 / C2) are compared with the burst figure.  Both fractions
        # are reported.
        sustained = tc and phase_acc["coarse_main"] >= 20.0
        # e4m3 operands: MEASURED_PEAKS.json has no fp8 figure, so bench.py measures one the same way in the same run
        # (measure_fp8_peak: cuBLASLt fp8 GEMM, burst and sustained); only if that is impossible the denominator falls back
        # to 2 x the measured bf16 figure, and peak_source says so
        f8 = tc and coarse_e4m3
        if f8 and fp8_peak:
            pb, ps = fp8_peak["burst"], fp8_peak["sustained"]
            src = "measured in this run: cuBLASLt e4m3 GEMM 8192^3 (torch._scaled_mm), "
        else:
            fx = 2.0 if f8 else 1.0
            pb, ps = fx * pk["bf16_tflops"], fx * pk["bf16_tflops_sustained"]
            src = pk["source"] + " cuBLAS bf16" + (" x 2 for e4m3 operands (fp8 GEMM could not be measured)" if f8 else "") + ", "
        peak = (ps if sustained else pb) if tc else 72.5
        roof["coarse"] = {"bound": "tensor" if tc else "fp32-ffma", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                          "frac": tf / peak, "traffic": None,
                          "kernel": ("coarse_tc_kernel<2> (tcgen05 e4m3, fp32 accumulate)" if coarse_e4m3 else
                                     "coarse_tc_kernel<2> (tcgen05 bf16, fp32 accumulate)") if tc else "coarse_exact_kernel",
                          "ms": phase_acc["coarse_main"], "flops_per_launch": coarse_flops,
                          "peak_source": (src + ("sustained (the launch runs %.0f ms inside back-to-back steps)" % phase_acc["coarse_main"]
                                                 if sustained else "burst"))
                          if tc else "measured FFMA micro-peak (profiles/)"}
        if tc:
            roof["coarse"]["frac_vs_burst"] = tf / pb
            roof["coarse"]["frac_vs_sustained"] = tf / ps
    fused = "scan" in phase_acc and "blend" not in phase_acc      # the re-check kernel also gathered + blended
    if "scan" in phase_acc:
        sb = scan_bytes + (blend_bytes if fused else 0)
        gb = sb / (phase_acc["scan"] * 1e-3) / 1e9
        roof["scan"] = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / pk["hbm_gbs"],
                        "traffic": None, "kernel": ("scan_tc_kernel (tcgen05 tf32 + TMA) + scan_recheck_kernel (exact fp32"
                                                    + (" + gather/weights/blend)" if fused else ")")) if "scan_recheck" in phase_acc
                        else f"scan_topk_kernel_v{scan_impl or 4}", "ms": phase_acc["scan"], "bytes_per_launch": sb,
                        "bytes_note": "B_s + B_g (SURVEY 8d): scan and fused tail" if fused else "B_s (SURVEY 8d)",
                        "peak_source": pk["source"]}
    if "blend" in phase_acc:
        gb = blend_bytes / (phase_acc["blend"] * 1e-3) / 1e9
        roof["blend"] = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / pk["hbm_gbs"],
                         "traffic": None, "kernel": "gather_blend_kernel", "ms": phase_acc["blend"],
                         "bytes_per_launch": blend_bytes, "peak_source": pk["source"]}
    dominant = max(roof, key=lambda k: roof[k]["ms"]) if roof else None
    # the scan phase split into its two kernels (not candidates for `dominant`): the streaming tensor-core scan against
    # B_s (the north star's "list-scan HBM roofline") and the exact re-check + fused tail against B_g
    if fused and "scan_recheck" in phase_acc and phase_acc["scan_recheck"] > 0:
        t_tail = phase_acc["scan_recheck"] + phase_acc.get("scan_fallback", 0.0)
        t_stream = phase_acc["scan"] - t_tail
        if t_stream > 0:
            gb = scan_bytes / (t_stream * 1e-3) / 1e9
            roof["scan_stream"] = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / pk["hbm_gbs"],
                                   "traffic": None, "kernel": "scan_tc_kernel (tcgen05 tf32 + TMA), incl. pad fill / counters",
                                   "ms": t_stream, "bytes_per_launch": scan_bytes, "bytes_note": "B_s (SURVEY 8d)",
                                   "peak_source": pk["source"]}
        gb = blend_bytes / (t_tail * 1e-3) / 1e9
        roof["scan_tail"] = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / pk["hbm_gbs"],
                             "traffic": None, "kernel": "scan_recheck_kernel (exact fp32 re-check + gather/weights/blend)",
                             "ms": t_tail, "bytes_per_launch": blend_bytes, "bytes_note": "B_g (SURVEY 8d)",
                             "peak_source": pk["source"]}
    traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if dominant and world == 1 and os.path.exists(traffic_file):   # the capture is of the 1-GPU step
        try:
            tr = json.load(open(traffic_file))
            for k in roof:
                w = tr.get(workload, {})
                roof[k]["traffic"] = w.get(k, w.get(roof[k]["kernel"].split(" ")[0].split("<")[0]) if k in ("coarse", "scan") else None)
        except Exception:
            pass

    return roof, dominant


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would time the CPU legs on one core (round 1:
    the per-N reference arm ran 3.5 x slower under the launcher than alone).  Set every pool explicitly to the cores this
    process may use: OpenMP of the C restatement, the BLAS behind numpy (sgemm coarse step) and torch."""
    n = len(os.sched_getaffinity(0))
    import oracle
    import torch
    oracle.set_num_threads(n)
    torch.set_num_threads(n)
    blas = None
    try:
        import threadpoolctl
        global _BLAS_LIMIT                                     # keep the limiter object alive
        _BLAS_LIMIT = threadpoolctl.threadpool_limits(limits=n)
        blas = [(i.get("internal_api"), i.get("num_threads")) for i in threadpoolctl.threadpool_info()]
    except Exception:
        pass
    return n, blas


def make_row_fetcher(vecs, row_bounds, rank, world, dist):
    """Sharded setup helper: fetch(gpos) -> the stored rows at global list-ordered positions gpos (int64 device tensor),
    wherever they live: positions are routed to the owning rank with all_to_all, gathered there, routed back.  Collective
    (every rank calls it once per use)."""
    import torch

    def fetch(gpos):
        dev = gpos.device
        bounds = torch.as_tensor(row_bounds[1:-1], dtype=torch.int64, device=dev)
        owner = torch.bucketize(gpos, bounds, right=True)
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=world)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts)
        both = torch.stack([send_counts, recv_counts]).cpu()
        ss, rs = [int(x) for x in both[0]], [int(x) for x in both[1]]
        want = torch.empty(sum(rs), dtype=torch.int64, device=dev)
        dist.all_to_all_single(want, gpos[order].contiguous(), rs, ss)
        rows = vecs[want - int(row_bounds[rank])]
        back = torch.empty((gpos.numel(), vecs.shape[1]), dtype=vecs.dtype, device=dev)
        dist.all_to_all_single(back, rows, ss, rs)
        out = torch.empty_like(back)
        out[order] = back
        return out
    return fetch


def stage_probed_lists(index, probed, ranges, rank, world, nlist, d, dev, dist):
    """CPU legs: the rows of the lists in `probed` (sorted global list ids, known to every rank), from whichever shard
    stores them, assembled on rank 0's host as (sub_off (nlist + 1,), sub_vecs).  Other ranks return (None, None)."""
    import torch
    vecs_d, _, offs_d, _ = index.device_views()
    offs_h = offs_d.cpu().numpy()
    lo, hi = ranges[rank]
    mine = probed[(probed >= lo) & (probed < hi)]
    sz_mine = offs_h[mine + 1] - offs_h[mine]
    if len(mine):
        idx = torch.cat([torch.arange(int(offs_h[l]), int(offs_h[l + 1]), device=dev) for l in mine]) if sz_mine.sum() else \
            torch.empty(0, dtype=torch.int64, device=dev)
        rows = vecs_d[idx]
    else:
        rows = torch.empty((0, d), dtype=torch.float32, device=dev)
    if world == 1:
        sz = np.zeros(nlist, np.int64)
        sz[mine] = sz_mine
        sub_off = np.zeros(nlist + 1, np.int64)
        np.cumsum(sz, out=sub_off[1:])
        return sub_off, rows.cpu().numpy()
    # sizes of every probed list, in list order (lists are contiguous per rank, so rank order == list order)
    cnt = torch.tensor([len(mine), int(sz_mine.sum())], dtype=torch.int64, device=dev)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [(int(c[0]), int(c[1])) for c in cnts]
    szt = torch.as_tensor(sz_mine, dtype=torch.int64, device=dev)
    if rank == 0:
        parts_sz, parts_rows = [sz_mine], [rows.cpu().numpy()]
        for r in range(1, world):
            t = torch.empty(cnts[r][0], dtype=torch.int64, device=dev)
            if cnts[r][0]:
                dist.recv(t, src=r)
            parts_sz.append(t.cpu().numpy())
            got = torch.empty((cnts[r][1], d), dtype=torch.float32, device=dev)
            if cnts[r][1]:
                dist.recv(got, src=r)
            parts_rows.append(got.cpu().numpy())
            del got
        sz = np.zeros(nlist, np.int64)
        sz[probed] = np.concatenate(parts_sz)
        sub_off = np.zeros(nlist + 1, np.int64)
        np.cumsum(sz, out=sub_off[1:])
        return sub_off, np.concatenate(parts_rows)
    if cnts[rank][0]:
        dist.send(szt, dst=0)
    if cnts[rank][1]:
        dist.send(rows, dst=0)
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--coarse-mode", type=int, default=0)
    ap.add_argument("--coarse-f8-auto", type=int, default=-1, help="1 / 0 = let / do not let the auto coarse mode probe e4m3 operands (-1 = library default)")
    ap.add_argument("--scan-impl", type=int, default=0, help="0 = library default (auto by batch size), 5 = tcgen05 scan, 1-4 exact FFMA scans")
    ap.add_argument("--tc-resident", type=int, default=-1, help="coarse kernel: resident query k-blocks (0/8/9/10; -1 = library default)")
    ap.add_argument("--tc-cluster", type=int, default=-1, help="coarse kernel: 2 = CTA pairs (default), 1 = one CTA per SM")
    ap.add_argument("--fused-blend", type=int, default=-1, help="1 = re-check kernel also blends (default), 0 = separate gather_blend kernel")
    ap.add_argument("--queries", type=int, default=0, help="tuning runs only: override the workload's query-frame count (e.g. the per-rank share of an N-GPU run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-peer", action="store_true", help="N > 1: use the NCCL all-to-all transport instead of peer memory")
    ap.add_argument("--near-frac", type=float, default=0.8, help="fraction of queries that are a stored row + N(0, 0.05^2) (SURVEY 8d)")
    ap.add_argument("--query-skew", type=float, default=0.75, help="Zipf exponent of the extra skewed-query measurement (0 = skip it)")
    ap.add_argument("--ring", type=int, default=4, help="input buffers kept on the device (refilled between steps, outside the events)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    synth = importlib.import_module(PKG + ".synth")
    cfg = synth.CONFIGS[args.workload]
    if args.queries > 0:
        cfg = (cfg[0], cfg[1], args.queries, cfg[3])     # reported in config.query_frames
    if args.workload == "C4":
        import bench_build
        return bench_build.main(args, cfg)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return
    N, d, nq, nlist = cfg

    import torch
    import torch.distributed as dist
    rvcr = importlib.import_module(PKG)
    sharded = importlib.import_module(PKG + ".sharded")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # the launcher's own setting (e.g. INFO for the rank check) wins
        dist.init_process_group("nccl", device_id=dev)

    # ---- build this rank's shard (by construction, on the device) --------------------------------
    t_setup = time.time()
    sizes = synth.list_sizes(N, nlist)
    ranges = synth.shard_ranges(sizes, world)
    lo, hi = ranges[rank]
    offs_all = np.concatenate([[0], np.cumsum(sizes)])
    row_bounds = [int(offs_all[r[0]]) for r in ranges] + [int(offs_all[-1])]
    cent_all = synth.all_centroids(nlist, d, dev)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=local_rank)
    index.nprobe = 1
    synth.fill_index_by_construction(index, cent_all, sizes, N, list_range=(lo, hi))
    index.set_option("coarse_mode", args.coarse_mode)
    index.set_option("scan_impl", args.scan_impl)
    if args.coarse_f8_auto >= 0:
        index.set_option("coarse_f8_auto", args.coarse_f8_auto)
    if args.tc_resident >= 0:
        index.set_option("tc_resident", args.tc_resident)
    if args.fused_blend >= 0:
        index.set_option("fused_blend", args.fused_blend)
    if args.tc_cluster > 0:
        index.set_option("tc_cluster", args.tc_cluster)
    index.set_option("profile", 1)
    nq_lo, nq_hi = nq * rank // world, nq * (rank + 1) // world
    n_home = nq_hi - nq_lo
    vecs_view = index.device_views()[0]
    fetch = make_row_fetcher(vecs_view, row_bounds, rank, world, dist) if world > 1 else None

    def make_queries(zipf):
        """this rank's home queries: 80 % stored row + N(0, 0.05^2), 20 % fresh mixture samples"""
        if args.near_frac > 0:
            q, near = synth.device_queries(cent_all, nq, q_range=(nq_lo, nq_hi), zipf=zipf, near_frac=args.near_frac, sizes=sizes)
            synth.apply_near_rows(q, near, vecs_view, row_base=row_bounds[rank], fetch=fetch)
            return q
        return synth.device_queries(cent_all, nq, q_range=(nq_lo, nq_hi), zipf=zipf)

    q_home = make_queries(0.0)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    sh = None
    peer = None
    transport = None
    total_steps = args.warmup + args.steps
    R = max(1, min(args.ring, total_steps))
    SLOT_SCRATCH, SLOT_E2E = R, R + 1
    if world > 1:
        sh = sharded.ShardedIndex(sharded.CudaShardBackend(index))
        transport = "nccl all-to-all (dispatch + combine)"
        if not args.no_peer:
            try:                                                   # collective: raises on every rank or on none
                peer = sh.enable_peer((R + 2) * n_home + 16, d, max_batch_rows=-(-nq // world))
                transport = sh.transport_name()
            except Exception as ex:
                transport += f"; peer memory unavailable: {str(ex)[:160]}"

    def slot(i):
        """(n_home, d) device buffer number i; inside the peer buffer when that transport is on"""
        if peer is None:
            return torch.empty((n_home, d), dtype=torch.float32, device=dev)
        return peer.rows(n_home, i * n_home)

    def step(buf):
        if sh is None:
            index.retrieve_blend_(buf, INDEX_RATE)
        else:
            sh.retrieve_blend_(buf, INDEX_RATE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(src, nsteps, warm, keep_first=0):
        """`warm` untimed + `nsteps` timed steps over the ring, every step on a fresh (un-blended) copy of `src`.  The
        steps are enqueued back to back; each is bracketed by its own CUDA events and the ring slot it used is refilled
        right after its closing event, so the refill copies are outside every timed interval.  Returns (sum of the step
        times in ms, max over ranks; the first `keep_first` blended rows of the first timed step)."""
        for b in ring:
            b.copy_(src)
        barrier()
        for i in range(warm):
            step(ring[i % R])
            ring[i % R].copy_(src)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        kept = None
        l0 = index.stats()["kernels_launched"]
        barrier()
        sampler.mark_begin()
        for i in range(nsteps):
            b = ring[(warm + i) % R]
            ev[i][0].record()
            step(b)
            ev[i][1].record()
            if i == 0 and keep_first:
                kept = b[:keep_first].clone()
            b.copy_(src)
        barrier()
        sampler.mark_end()
        ms = float(sum(a.elapsed_time(b) for a, b in ev))
        nl = index.stats()["kernels_launched"] - l0            # librvcr kernels launched inside the timed steps (this rank)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([nl], device=dev, dtype=torch.int64)
            dist.all_reduce(lt)
            nl = int(lt.item())
        return ms, kept, nl

    ring = [slot(i) for i in range(R)]
    scratch = slot(SLOT_SCRATCH)
    sampler = ClockSampler(local_rank)
    sampler.start()
    scratch.copy_(q_home)
    step(scratch)                                             # first call: lazy setup (row norms, tensor maps, scratch)
    barrier()
    time.sleep(0.25)                                          # let nvidia-smi deliver its first samples
    keep = min(n_home, args.cpu_sample) if rank == 0 else 0
    ms_total, gpu_rows_dev, launches = timed_steps(q_home, args.steps, args.warmup - 1, keep_first=keep)
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = nq / (ms_step * 1e-3)
    gpu_rows = gpu_rows_dev.cpu().numpy() if gpu_rows_dev is not None else None
    result_check = gpu_rows[:4] if gpu_rows is not None else None

    # ---- per-phase device times for the roofline: the library keeps the CUDA events of its LAST call only, and reading
    # them needs a synchronize, so they cannot be collected inside the timed loop.  Each repetition therefore enqueues three
    # steps back to back -- the same power / clock regime as the timed region (a lone step after an idle gap runs ~3 %
    # faster: sw_power_cap has not pulled the clocks down yet) -- and reads the phases of the third.
    phase_acc = {}
    reps = max(1, min(args.steps, 3))
    for _ in range(reps):
        for b in (ring[0], ring[1 % R], scratch):
            b.copy_(q_home)
        step(ring[0])
        step(ring[1 % R])
        step(scratch)
        torch.cuda.synchronize()
        for k, v in index.phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v / reps
    st = index.stats()
    coarse_e4m3 = args.coarse_mode == 4 or index.get_option("coarse_f8_decision") == 1
    fp8_peak = measure_fp8_peak(dev) if coarse_e4m3 else None                 # every rank measures (keeps the ranks in step)
    roof, dominant = build_roofline(phase_acc, st["coarse_tc_queries"] > 0, peaks(), nst=int(sizes[lo:hi].sum()), n_home=n_home,
                                    owned_q=nq / world, nlist=nlist, d=d, scan_impl=args.scan_impl, workload=args.workload,
                                    world=world, coarse_e4m3=coarse_e4m3, fp8_peak=fp8_peak)

    # ---- the same steps under a skewed query mass (Zipf over lists): what real clips do to the shard balance ----
    skewed = None
    if args.query_skew > 0:
        try:
            q_skew = make_queries(args.query_skew)
            sk_steps = max(1, min(args.steps, 3))
            sampler = ClockSampler(local_rank)                     # not started: mark_begin / mark_end are no-ops on the record
            ms_sk, _, _ = timed_steps(q_skew, sk_steps, 1)
            a_sk = index.assign(q_skew[:min(n_home, 65536)])[0]
            own = torch.bucketize(a_sk, torch.as_tensor([r[0] for r in ranges[1:]], dtype=torch.int32, device=dev), right=True)
            cnt = torch.bincount(own.clamp_(min=0), minlength=world).to(torch.float64)
            if world > 1:
                dist.all_reduce(cnt)
            skewed = {"zipf_exponent": args.query_skew, "value": nq / (ms_sk / sk_steps * 1e-3), "unit": UNIT,
                      "ms_per_step": ms_sk / sk_steps, "steps": sk_steps,
                      "owner_load_max_over_mean": float(cnt.max() / cnt.mean()) if world > 1 else 1.0,
                      "note": "same index, query lists drawn with mass ~ (rank + 1)^-s over a fixed random ranking of the lists"}
            del q_skew
        except Exception as ex:                                    # pragma: no cover
            skewed = {"error": repr(ex)[:200]}

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region --------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n_home, d), dtype=torch.float32).pin_memory()
        src = q_home.cpu()
        e2e_steps = max(1, min(args.steps, 3))
        times = []
        for it in range(1 + e2e_steps):
            host.copy_(src)
            barrier()
            t0 = time.perf_counter()
            if sh is None:
                index.retrieve_blend_host_(host.numpy(), INDEX_RATE)     # H2D, search+blend, D2H, sync
            else:
                sh.retrieve_blend_host_(host, INDEX_RATE,                 # chunked H2D under the coarse step, D2H, sync
                                        device_rows=slot(SLOT_E2E) if peer is not None else None)
            barrier()
            if it > 0:
                times.append(time.perf_counter() - t0)
        sec = float(np.mean(times))
        if world > 1:
            t = torch.tensor([sec], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        # what the host link gives this rank while every rank copies at once (plain pinned copies of the same size, H2D then
        # D2H): the floor under the e2e step whatever the kernels do
        link = None
        try:
            dbuf = slot(SLOT_SCRATCH)
            barrier()
            t0 = time.perf_counter()
            dbuf.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            host.copy_(dbuf, non_blocking=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            tt = torch.tensor([t1 - t0, t2 - t1], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            gb = n_home * d * 4 / 1e9
            link = {"h2d_ms": float(tt[0]) * 1e3, "d2h_ms": float(tt[1]) * 1e3, "h2d_gbs_per_rank": gb / float(tt[0]),
                    "d2h_gbs_per_rank": gb / float(tt[1]), "note": "slowest rank, all ranks copying concurrently"}
        except Exception:
            pass
        e2e = {"value": nq / sec, "unit": UNIT, "h2d_bytes_per_step": int(nq) * d * 4, "d2h_bytes_per_step": int(nq) * d * 4,
               "ms_per_step": sec * 1e3, "host_link": link, "api": "IndexIVFFlat.retrieve_blend_host_ (rvcr_retrieve_blend_host), pinned host buffers"
               if sh is None else "ShardedIndex.retrieve_blend_host_, pinned host buffers per rank"}
        del host, src

    # ---- CPU legs on rank 0: the parity sample at every N, the timed CPU baseline at N = 1 only ----------
    cpu = None
    parity = None
    parity_str = None
    if not args.no_cpu_baseline:
        try:                                                   # the bench line must survive a failing baseline leg
            import oracle
            sample = min(n_home if world > 1 else nq, args.cpu_sample)
            cent_h = index.device_views()[3].cpu().numpy()
            probed_t = torch.zeros(sample + 1, dtype=torch.int64, device=dev)
            if rank == 0:
                threads, blas = use_all_host_threads()
                qs = q_home[:sample].cpu().numpy()
                a, _ = oracle.coarse_assign_blas(qs, cent_h)                 # untimed: which lists to stage on the host
                probed = np.unique(a[a >= 0])
                probed_t[0] = len(probed)
                probed_t[1:1 + len(probed)] = torch.as_tensor(probed, device=dev)
            if world > 1:
                dist.broadcast(probed_t, src=0)
            probed = probed_t[1:1 + int(probed_t[0])].cpu().numpy()
            sub_off, sub_vecs = stage_probed_lists(index, probed, ranges, rank, world, nlist, d, dev, dist)
            if rank == 0:
                sub_ids = np.arange(len(sub_vecs), dtype=np.int64)
                cpu_path(oracle, qs[:256], cent_h, sub_off, sub_vecs, sub_ids)      # warm-up (thread pools, page faults)
                t0 = time.perf_counter()
                passes = 0
                while True:                                                         # N = 1: about 10 s of CPU work, at most 8 passes
                    out_cpu, _ = cpu_path(oracle, qs, cent_h, sub_off, sub_vecs, sub_ids)
                    passes += 1
                    if world > 1 or time.perf_counter() - t0 >= 10.0 or passes >= 8:
                        break
                total_s = time.perf_counter() - t0
                sec = total_s / passes
                if world == 1:
                    cpu = {"value": sample / sec, "unit": UNIT, "cores": threads, "kind": "port",
                           "sample": f"first {sample} of {nq} query frames of this workload, all {nlist} centroids, the {len(probed)} lists "
                                     f"they probe staged on the host; {passes} passes, {total_s:.1f} s of CPU work; oracle port of faiss "
                                     f"1.7.3 + numpy (omp_threads={oracle.num_threads()}, torch_threads={torch.get_num_threads()}, blas={blas})"}
                parity = parity_sample(gpu_rows, out_cpu)
                if parity and "rows" in parity:
                    parity_str = (f"{parity['rows_bit_identical']}/{parity['rows']} blended rows of the first timed step bit-identical to the "
                                  f"CPU port of the reference path, {parity['rows_differing']} differing"
                                  + (f" (max rel diff {parity['max_rel_diff_in_differing_rows']:.2e}; near-tie centroid choices of the "
                                     f"BLAS-order coarse distances)" if parity["rows_differing"] else "")
                                  + (f"; lists gathered from their owner ranks ({world} GPUs)" if world > 1 else ""))
                del sub_vecs
        except Exception as ex:
            if rank == 0:
                cpu = {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                       "sample": "CPU leg failed: " + repr(ex)[:200]}
                parity_str = "parity sample failed: " + repr(ex)[:160]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "index_vectors": N, "d": d, "nlist": nlist, "query_frames": nq, "k": K,
                           "nprobe": 1, "index_rate": INDEX_RATE, "parallelism": f"lists sharded over {world} GPU(s)",
                           **({"transport": transport} if transport else {}),
                           "l2": "inputs larger than L2 (fresh 3 KB/query feature buffer per step; lists >> 126 MB)"
                           if nq * d * 4 > (126 << 20) else "inputs smaller than L2",
                           "coarse": (("tcgen05 e4m3 (chosen per centroid set by the library's probe: results identical to bf16)" if coarse_e4m3
                                       else "tcgen05 bf16") + " + exact fp32 re-check") if st["coarse_tc_queries"] else "exact FFMA",
                           "queries": f"{args.near_frac:.0%} stored row + N(0, 0.05^2), rest fresh mixture samples; index has 0.1 % exact "
                                      f"duplicate rows (SURVEY 8d)",
                           "timing": f"per-step CUDA events summed over {args.steps} back-to-back steps (max over ranks); ring of {R} "
                                     f"input buffers refilled between steps outside the events",
                           "parity_sample": parity_str,
                           "setup_s": round(setup_s, 1)},
                "roofline": dict(roof[dominant], dominant=dominant) if dominant else None,
                "roofline_all": roof, "phase_ms": phase_acc, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks, "host_affinity": numa, "skewed_queries": skewed,
                "coarse_stats": {"tc_queries": st["coarse_tc_queries"], "rechecked_candidates": st["coarse_candidates"],
                                 "exact_fallback_queries": st["coarse_fallback"]},
                "scan_stats": {"tc_candidates_rechecked": st.get("scan_tc_candidates", 0), "exact_fallback_queries": st.get("scan_tc_fallback", 0),
                               "queries": st["queries"]},
                "sanity": {"finite_first_rows": bool(np.isfinite(result_check).all()) if result_check is not None else None,
                           "parity_sample": parity}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

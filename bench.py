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
    cores = len(os.sched_getaffinity(0))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "index_vectors": N, "d": d, "nlist": nlist, "query_frames": nq, "k": K,
                       "nprobe": 1, "index_rate": INDEX_RATE, "note": "each step = bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} of {nq} query frames per step against all {nlist} centroids and the "
                                       f"{int((np.diff(list_off) > 0).sum())} lists they probe (avg {N / nlist:.0f} rows); "
                                       f"oracle port of faiss 1.7.3 + numpy (faiss itself is not installable here); "
                                       f"omp_threads={oracle.num_threads()} torch_threads={torch.get_num_threads()} setup {setup_s:.0f}s"},
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
                   coarse_e4m3=False):
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
        # coarse_mode 4 (experimental e4m3 operands): MEASURED_PEAKS.json has no fp8 figure; the tensor cores' e4m3 rate is
        # twice the bf16 rate, so the denominator is 2 x the measured bf16 figure, and peak_source says so
        fp8x = 2.0 if (tc and coarse_e4m3) else 1.0
        peak = fp8x * (pk["bf16_tflops_sustained"] if sustained else pk["bf16_tflops"]) if tc else 72.5
        roof["coarse"] = {"bound": "tensor" if tc else "fp32-ffma", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                          "frac": tf / peak, "traffic": None,
                          "kernel": ("coarse_tc_kernel<2> (tcgen05 e4m3, fp32 accumulate)" if coarse_e4m3 else
                                     "coarse_tc_kernel<2> (tcgen05 bf16, fp32 accumulate)") if tc else "coarse_exact_kernel",
                          "ms": phase_acc["coarse_main"], "flops_per_launch": coarse_flops,
                          "peak_source": (pk["source"] + (" cuBLAS bf16 sustained (the launch runs %.0f ms inside back-to-back steps)"
                                                          % phase_acc["coarse_main"] if sustained else " cuBLAS bf16 burst")
                                          + (" x 2 for e4m3 operands (no measured fp8 figure)" if coarse_e4m3 else ""))
                          if tc else "measured FFMA micro-peak (profiles/)"}
        if tc:
            roof["coarse"]["frac_vs_burst"] = tf / (fp8x * pk["bf16_tflops"])
            roof["coarse"]["frac_vs_sustained"] = tf / (fp8x * pk["bf16_tflops_sustained"])
    fused = "scan" in phase_acc and "blend" not in phase_acc      # the re-check kernel also gathered + blended
    if "scan" in phase_acc:
        sb = scan_bytes + (blend_bytes if fused else 0)
        gb = sb / (phase_acc["scan"] * 1e-3) / 1e9
        roof["scan"] = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / pk["hbm_gbs"],
                        "traffic": None, "kernel": ("scan_tc_kernel (tcgen05 tf32 + TMA) + scan_recheck_kernel (exact fp32"
                                                    + (" + gather/weights/blend)" if fused else ")")) if "scan_recheck" in phase_acc
                        else ("scan_tc_fused_kernel (tcgen05 tf32 + TMA; exact re-check + tail as a warp role)" if scan_impl == 6
                              else f"scan_topk_kernel_v{scan_impl or 4}"), "ms": phase_acc["scan"], "bytes_per_launch": sb,
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

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--coarse-mode", type=int, default=0)
    ap.add_argument("--coarse-f8-auto", type=int, default=0, help="1 = let the auto coarse mode probe e4m3 operands (experimental)")
    ap.add_argument("--scan-impl", type=int, default=0, help="0 = library default (auto by batch size), 5 = tcgen05 scan, 1-4 exact FFMA scans, 6 = experimental fused scan + re-check")
    ap.add_argument("--tc-resident", type=int, default=-1, help="coarse kernel: resident query k-blocks (0/8/9/10; -1 = library default)")
    ap.add_argument("--fused-blend", type=int, default=-1, help="1 = re-check kernel also blends (default), 0 = separate gather_blend kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-peer", action="store_true", help="N > 1: use the NCCL all-to-all transport instead of peer memory")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    synth = importlib.import_module(PKG + ".synth")
    cfg = synth.CONFIGS[args.workload]
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
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    # ---- build this rank's shard (by construction, on the device) --------------------------------
    t_setup = time.time()
    sizes = synth.list_sizes(N, nlist)
    ranges = synth.shard_ranges(sizes, world)
    lo, hi = ranges[rank]
    cent_all = synth.all_centroids(nlist, d, dev)
    index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=local_rank)
    index.nprobe = 1
    synth.fill_index_by_construction(index, cent_all, sizes, N, list_range=(lo, hi))
    index.set_option("coarse_mode", args.coarse_mode)
    index.set_option("scan_impl", args.scan_impl)
    if args.coarse_f8_auto:
        index.set_option("coarse_f8_auto", 1)
    if args.tc_resident >= 0:
        index.set_option("tc_resident", args.tc_resident)
    if args.fused_blend >= 0:
        index.set_option("fused_blend", args.fused_blend)
    index.set_option("profile", 1)
    nq_lo, nq_hi = nq * rank // world, nq * (rank + 1) // world
    q_home = synth.device_queries(cent_all, nq, q_range=(nq_lo, nq_hi))
    n_home = nq_hi - nq_lo
    del cent_all
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    sh = None
    peer = None
    transport = None
    total_steps = args.warmup + args.steps
    if world > 1:
        sh = sharded.ShardedIndex(sharded.CudaShardBackend(index))
        transport = "nccl all-to-all (dispatch + combine)"
        if not args.no_peer:
            try:                                                   # collective: raises on every rank or on none
                peer = sh.enable_peer((total_steps + 2) * n_home + 16, d)
                transport = "peer memory over NVLink (cudaIpc): owner pulls rows, blend epilogue pushes them home"
            except Exception as ex:
                transport += f"; peer memory unavailable: {str(ex)[:160]}"

    def fresh(slot):
        """a fresh (un-blended) copy of this rank's queries; inside the peer buffer when that transport is on"""
        if peer is None:
            return q_home.clone()
        return peer.rows(n_home, slot * n_home).copy_(q_home)

    def step(buf):
        if sh is None:
            index.retrieve_blend_(buf, INDEX_RATE)
        else:
            sh.retrieve_blend_(buf, INDEX_RATE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bufs = [fresh(i) for i in range(total_steps)]             # every step gets a fresh (un-blended) input
    sampler = ClockSampler(local_rank)
    sampler.start()
    step(bufs[0])                                             # first call: lazy setup (row norms, tensor maps, scratch)
    barrier()
    time.sleep(0.25)                                          # let nvidia-smi deliver its first samples
    for i in range(1, args.warmup):
        step(bufs[i])
    barrier()
    launches0 = index.stats()["kernels_launched"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for i in range(args.steps):
        step(bufs[args.warmup + i])
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = index.stats()["kernels_launched"] - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = nq / (ms_step * 1e-3)
    result_check = bufs[args.warmup][:4].cpu().numpy()
    gpu_rows = None                                           # blended rows of the first timed step, for the parity sample
    try:
        if world == 1 and not args.no_cpu_baseline:
            gpu_rows = bufs[args.warmup][:min(n_home, args.cpu_sample)].cpu().numpy()
    except Exception:
        gpu_rows = None
    del bufs

    # ---- per-phase device times (separate pass, same inputs) for the roofline --------------------
    phase_acc = {}
    reps = max(1, min(args.steps, 3))
    for _ in range(reps):
        b = fresh(total_steps)
        step(b)
        torch.cuda.synchronize()
        for k, v in index.phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v / reps
        del b
    st = index.stats()
    coarse_e4m3 = args.coarse_mode == 4 or (args.coarse_f8_auto and index.get_option("coarse_f8_decision") == 1)
    roof, dominant = build_roofline(phase_acc, st["coarse_tc_queries"] > 0, peaks(), nst=int(sizes[lo:hi].sum()), n_home=n_home,
                                    owned_q=nq / world, nlist=nlist, d=d, scan_impl=args.scan_impl, workload=args.workload,
                                    world=world, coarse_e4m3=coarse_e4m3)

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
                                        device_rows=peer.rows(n_home, (total_steps + 1) * n_home) if peer is not None else None)
            barrier()
            if it > 0:
                times.append(time.perf_counter() - t0)
        sec = float(np.mean(times))
        if world > 1:
            t = torch.tensor([sec], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        e2e = {"value": nq / sec, "unit": UNIT, "h2d_bytes_per_step": int(nq) * d * 4, "d2h_bytes_per_step": int(nq) * d * 4,
               "ms_per_step": sec * 1e3, "api": "IndexIVFFlat.retrieve_blend_host_ (rvcr_retrieve_blend_host), pinned host buffers"
               if sh is None else "ShardedIndex.retrieve_blend_host_, pinned host buffers per rank"}
        del host, src

    # ---- CPU baseline on rank 0 (N = 1 only): bounded sample of the same workload ------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:                                                   # the bench line must survive a failing baseline leg
            import oracle
            sample = min(nq, args.cpu_sample)
            qs = q_home[:sample].cpu().numpy()
            vecs_d, ids_d, offs_d, cent_d = index.device_views()
            cent_h = cent_d.cpu().numpy()
            a, _ = oracle.coarse_assign_blas(qs, cent_h)                 # untimed: which lists to stage on the host
            probed = np.unique(a[a >= 0])
            offs_h = offs_d.cpu().numpy()
            sz = np.zeros(nlist, np.int64)
            sz[probed] = offs_h[probed + 1] - offs_h[probed]
            sub_off = np.zeros(nlist + 1, np.int64)
            np.cumsum(sz, out=sub_off[1:])
            sub_vecs = np.empty((int(sub_off[-1]), d), np.float32)
            for l in probed:
                sub_vecs[sub_off[l]:sub_off[l + 1]] = vecs_d[offs_h[l]:offs_h[l + 1]].cpu().numpy()
            sub_ids = np.arange(len(sub_vecs), dtype=np.int64)
            cpu_path(oracle, qs[:256], cent_h, sub_off, sub_vecs, sub_ids)      # warm-up (thread pools, page faults)
            t0 = time.perf_counter()
            passes = 0
            while True:                                                         # about 10 s of CPU work, at most 8 passes
                out_cpu, _ = cpu_path(oracle, qs, cent_h, sub_off, sub_vecs, sub_ids)
                passes += 1
                if time.perf_counter() - t0 >= 10.0 or passes >= 8:
                    break
            total_s = time.perf_counter() - t0
            sec = total_s / passes
            import torch as _t
            cpu = {"value": sample / sec, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                   "sample": f"first {sample} of {nq} query frames of this workload, all {nlist} centroids, the {len(probed)} lists "
                             f"they probe staged on the host; {passes} passes, {total_s:.1f} s of CPU work; oracle port of faiss "
                             f"1.7.3 + numpy (omp_threads={oracle.num_threads()}, torch_threads={_t.get_num_threads()})"}
            parity = parity_sample(gpu_rows, out_cpu)
            del sub_vecs
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                   "sample": "CPU baseline leg failed: " + repr(ex)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "index_vectors": N, "d": d, "nlist": nlist, "query_frames": nq, "k": K,
                           "nprobe": 1, "index_rate": INDEX_RATE, "parallelism": f"lists sharded over {world} GPU(s)",
                           **({"transport": transport} if transport else {}),
                           "l2": "inputs larger than L2 (fresh 3 KB/query feature buffer per step; lists >> 126 MB)"
                           if nq * d * 4 > (126 << 20) else "inputs smaller than L2",
                           "coarse": (("tcgen05 e4m3 (experimental coarse_mode 4)" if coarse_e4m3 else "tcgen05 bf16")
                                      + " + exact fp32 re-check") if st["coarse_tc_queries"] else "exact FFMA",
                           "setup_s": round(setup_s, 1)},
                "roofline": dict(roof[dominant], dominant=dominant) if dominant else None,
                "roofline_all": roof, "phase_ms": phase_acc, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks, "host_affinity": numa,
                "coarse_stats": {"tc_queries": st["coarse_tc_queries"], "rechecked_candidates": st["coarse_candidates"],
                                 "exact_fallback_queries": st["coarse_fallback"]},
                "scan_stats": {"tc_candidates_rechecked": st.get("scan_tc_candidates", 0), "exact_fallback_queries": st.get("scan_tc_fallback", 0),
                               "queries": st["queries"]},
                "sanity": {"finite_first_rows": bool(np.isfinite(result_check).all()), "parity_sample": parity}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

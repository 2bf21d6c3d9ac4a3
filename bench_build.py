"""bench.py --workload C4: the index build of rvc/train/process/extract_index.py:59-71 at BASELINE config 4
(4 M synthetic 768-d training frames -> nlist = 32 000, single B200).

A "step" = one whole build pass with given centroids: `index.add` of every row (nearest-centroid assignment on the
tcgen05 coarse kernels + append) and the fold into the list layout (what `faiss.write_index` / the first search
consume).  `value` = vectors/s with the rows already resident in HBM, handed over in batches of 262 144 rows; `e2e` =
the reference's own loop shape, host numpy rows in `index.add(big_npy[i:i + 8192])` calls (extract_index.py:67-69), H2D
inside every call, and a D2H read of the list sizes at the end.  `train` (extract_index.py:65: 10 Lloyd iterations of
faiss Clustering on the same rows) is timed once and reported beside it.  `roofline` = the coarse kernel of one
262 144-row assignment (tensor bound; F_add = 2 N nlist d).  `cpu_baseline` / `--impl reference` = the CPU
port (sgemm coarse like faiss >= 20 rows, list append) on 65 536 rows, all host cores, stated as a sample.

The build path does not shard (one index, one list layout): under torch.distributed.run rank 0 alone runs it.
"""
from __future__ import annotations

import importlib
import json
import os
import time

import numpy as np

PKG = "codename-rvc-fork-3_b200"
METRIC = "index build: vectors assigned + added per second (768-d, nlist 32000)"
UNIT = "vectors/s"
BATCH_DEV = 262144
BATCH_REF = 8192          # extract_index.py:67


def _rows(cent, a, b, dev):
    """rows a..b of the synthetic training set: centroid[list] + N(0, 0.14^2), list uniform (device, seeded per range)."""
    import torch
    g = torch.Generator(device=dev).manual_seed(777 + a)
    lists = torch.randint(0, cent.shape[0], (b - a,), generator=g, device=dev)
    return cent[lists] + torch.randn((b - a, cent.shape[1]), generator=g, device=dev) * 0.14


def _cpu_port(oracle, xs, cent_h, nlist):
    a = oracle.coarse_assign_blas(xs, cent_h)[0]
    oracle.build_lists(xs, np.asarray(a, np.int32), nlist)


def main(args, cfg):
    import bench
    N, d, _, nlist = cfg
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import oracle
    sample = 65536
    if args.impl == "reference":
        threads, blas = bench.use_all_host_threads()
        g = torch.Generator().manual_seed(1234)
        cent_h = (torch.randn(((nlist + 3) // 4, d), generator=g) * 0.30)[torch.arange(nlist) // 4].numpy() + \
            (torch.randn((nlist, d), generator=g) * 0.03).numpy()
        xs = (cent_h[torch.randint(0, nlist, (sample,), generator=g).numpy()] +
              (torch.randn((sample, d), generator=g) * 0.14).numpy()).astype(np.float32)
        cent_h = cent_h.astype(np.float32)
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _cpu_port(oracle, xs, cent_h, nlist)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
        sec = float(np.mean(times))
        val = sample / sec
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "C4", "vectors": N, "d": d, "nlist": nlist, "note": "each step = bounded sample"},
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                           "sample": f"{sample} of {N} rows per step against all {nlist} centroids; oracle port of faiss 1.7.3 "
                                                     f"add (sgemm coarse + append); omp_threads={oracle.num_threads()} blas={blas}"},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return

    rvcr = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    t_setup = time.time()
    cent = synth.all_centroids(nlist, d, dev)
    cent_h = cent.cpu().numpy()
    x = torch.empty((N, d), dtype=torch.float32, device=dev)
    for c0 in range(0, N, 1 << 19):
        x[c0:c0 + (1 << 19)] = _rows(cent, c0, min(N, c0 + (1 << 19)), dev)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    def fresh_index():
        index = rvcr.index_factory(d, f"IVF{nlist},Flat", device=local_rank)    # extract_index.py:62
        index.nprobe = 1                                                          # :64
        index.set_centroids(cent_h)
        return index

    def build_pass(index, rows, batch):
        for i in range(0, N, batch):
            index.add(rows[i:i + batch])                                          # :67-69
        return index.list_sizes()                                                 # folds the staged rows into the lists

    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    total = args.warmup + args.steps
    times, launches = [], 0
    sizes = None
    for it in range(total):
        index = fresh_index()
        index.assign(x[:BATCH_DEV])                                               # centroid operands built outside the timing
        torch.cuda.synchronize()
        if it == args.warmup:
            time.sleep(0.25)
            sampler.mark_begin()
        l0 = index.stats()["kernels_launched"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sizes = build_pass(index, x, BATCH_DEV)
        e1.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(e0.elapsed_time(e1))
            launches += index.stats()["kernels_launched"] - l0
        if it + 1 < total:
            del index
            torch.cuda.empty_cache()
    sampler.mark_end()
    clocks = sampler.stop()
    ms_step = float(np.mean(times))
    value = N / (ms_step * 1e-3)
    assert int(sizes.sum()) == N == index.ntotal

    # ---- properties of the last build at full size + a sample against the oracle ----------------------------
    vecs, ids, offs, _ = index.device_views()
    seg = torch.repeat_interleave(torch.arange(nlist, device=dev), torch.as_tensor(sizes, device=dev))
    same = seg[1:] == seg[:-1]
    g = torch.Generator(device=dev).manual_seed(1)
    pos = torch.randint(0, N, (512,), generator=g, device=dev)
    a_o, _ = oracle.coarse_assign(vecs[pos].cpu().numpy(), cent_h)
    props = {"every_id_exactly_once": bool(torch.equal(torch.sort(ids).values, torch.arange(N, device=dev))),
             "ids_ascending_inside_every_list": bool((ids[1:][same] > ids[:-1][same]).all()),
             "stored_row_equals_added_row": bool(torch.equal(vecs[pos], x[ids[pos]])),
             "sample_lists_equal_oracle_argmin": bool(np.array_equal(seg[pos].cpu().numpy(), a_o)), "sample": 512}
    parity_str = ("512 sampled vectors in the list the CPU oracle's canonical argmin picks; every id stored once, ascending inside "
                  "every list, rows intact" if all(v for k, v in props.items() if k != "sample") else f"PROPERTY FAILED: {props}")
    del vecs, ids, offs, seg, same

    # ---- roofline: one 262 144-row assignment on the coarse kernels ---------------------------------------
    index.set_option("profile", 1)
    ph = {}
    for _ in range(3):
        index.assign(x[:BATCH_DEV])
        torch.cuda.synchronize()
        for k, v in index.phase_ms().items():
            ph[k] = ph.get(k, 0.0) + v / 3
    pk = bench.peaks()
    flops = 2.0 * BATCH_DEV * nlist * d
    tf = flops / (ph["coarse_main"] * 1e-3) / 1e12
    # the 262 144-row add calls are above the e4m3 probe's threshold: if the probe kept e4m3 operands the denominator is
    # the fp8 GEMM peak measured in this run, not the bf16 figure of MEASURED_PEAKS.json
    f8 = index.get_option("coarse_f8_decision") == 1
    fp8 = bench.measure_fp8_peak(dev) if f8 else None
    if f8 and fp8:
        pb, ps, src = fp8["burst"], fp8["sustained"], "measured in this run: cuBLASLt e4m3 GEMM 8192^3 (torch._scaled_mm),"
    elif f8:
        pb, ps = 2 * pk["bf16_tflops"], 2 * pk["bf16_tflops_sustained"]
        src = pk["source"] + " cuBLAS bf16 x 2 for e4m3 operands (fp8 GEMM could not be measured),"
    else:
        pb, ps, src = pk["bf16_tflops"], pk["bf16_tflops_sustained"], pk["source"] + " cuBLAS bf16,"
    roof = {"bound": "tensor", "achieved": tf, "peak": pb, "unit": "TFLOP/s", "frac": tf / pb,
            "traffic": None, "kernel": "coarse_tc_kernel<2> (tcgen05 %s, fp32 accumulate)" % ("e4m3" if f8 else "bf16"),
            "ms": ph["coarse_main"], "flops_per_launch": flops,
            "peak_source": src + " burst (a 262 144-row launch, ~10 ms)",
            "frac_vs_sustained": tf / ps, "dominant": "coarse"}
    del index
    torch.cuda.empty_cache()

    # ---- e2e: the reference's loop, host rows in 8192-row add calls ---------------------------------------
    e2e = None
    if not args.no_e2e:
        xh = x.cpu().numpy()
        ts = []
        for it in range(2):
            index = fresh_index()
            index.assign(x[:BATCH_REF])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sz = build_pass(index, xh, BATCH_REF)
            torch.cuda.synchronize()
            if it > 0:
                ts.append(time.perf_counter() - t0)
            assert int(sz.sum()) == N
            del index
        e2e = {"value": N / float(np.mean(ts)), "unit": UNIT, "h2d_bytes_per_step": N * d * 4, "d2h_bytes_per_step": nlist * 8,
               "ms_per_step": float(np.mean(ts)) * 1e3,
               "api": "IndexIVFFlat.add(host rows[i:i+8192]) loop + list_sizes() (extract_index.py:67-69), pageable numpy rows"}
        del xh

    # ---- index.train on the same rows (extract_index.py:65), once -----------------------------------------
    train = None
    try:
        xh = x[: N].cpu().numpy()
        tindex = rvcr.index_factory(d, f"IVF{nlist},Flat", device=local_rank)
        t0 = time.perf_counter()
        tindex.train(xh)
        torch.cuda.synchronize()
        t_train = time.perf_counter() - t0
        c = tindex.get_centroids()
        train = {"seconds": t_train, "niter": 10, "assign_pflop": 10 * 2.0 * N * nlist * d / 1e15,
                 "pflops_incl_h2d": 10 * 2.0 * N * nlist * d / 1e15 / t_train, "centroids_finite": bool(np.isfinite(c).all())}
        del tindex, xh
    except Exception as ex:                                        # pragma: no cover
        train = {"error": repr(ex)[:200]}

    cpu = None
    if not args.no_cpu_baseline:
        threads, blas = bench.use_all_host_threads()
        xs = x[:sample].cpu().numpy()
        _cpu_port(oracle, xs[:4096], cent_h, nlist)
        t0 = time.perf_counter()
        passes = 0
        while True:
            _cpu_port(oracle, xs, cent_h, nlist)
            passes += 1
            if time.perf_counter() - t0 >= 10.0 or passes >= 4:
                break
        sec = (time.perf_counter() - t0) / passes
        cpu = {"value": sample / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {sample} of {N} rows, all {nlist} centroids, {passes} passes; oracle port of faiss 1.7.3 add (sgemm coarse + "
                         f"append; omp_threads={oracle.num_threads()} blas={blas})"}

    print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic",
                      "config": {"workload": "C4", "vectors": N, "d": d, "nlist": nlist, "batch_rows": BATCH_DEV,
                                 "l2": "inputs larger than L2 (12.3 GB of rows per step)", "parity_sample": parity_str,
                                 "parallelism": "one GPU (the build does not shard)", "setup_s": round(setup_s, 1)},
                      "roofline": roof, "phase_ms": ph, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                      "train": train, "sanity": props}), flush=True)

#!/usr/bin/env python
"""bench.py -- MAS alignments/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A step = one ``maximum_path(neg_cent, mask)`` call over one batch of synthetic input
(SURVEY.md section 8d shapes).  Reported:

  value    alignments/s with neg_cent / mask resident in HBM, CUDA-event timed, whole job (all ranks)
  e2e      the same metric through the reference-facing C entry ``mas_maximum_path_c_host`` (the twin of
           core.pyx:38) with pinned HOST buffers: H2D of neg_cent and D2H of the int32 path inside the
           timed region
  roofline achieved algorithmic GB/s of the maximum_path kernel chain (forward DP + backtrack + write-out,
           chained with programmatic dependent launch so they overlap and are timed as one unit) against the
           measured HBM copy bandwidth; ``kernels_ms`` gives the per-kernel durations from a serialised pass
  cpu_baseline  the reference's own Cython (oracle/_ref, compiled from /root/reference) timed on this box's
           host cores on the same workload

Under torchrun every rank aligns its own shard of utterances (weak scaling: B per GPU fixed, like
``batch_size`` per replica in the reference's DDP, train.py:101); there is no collective in the timed region.
After timing, ranks all-gather their per-frame indices over NCCL only to verify them.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # name: (B per GPU, T_y, T_x)
    "c1": (1, 128, 32), "c2": (64, 1024, 192), "c3": (32, 1536, 256), "c4": (8, 4096, 512),
}
METRIC = "MAS alignments/sec"
UNIT = "alignments/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def busy_until(self, work, rows, timeout_s):
        """Run `work()` (enqueue + synchronize) over and over until `rows` samples have arrived."""
        import torch
        t0 = time.time()
        while self.proc is not None and len(self.rows) < rows and time.time() - t0 < timeout_s:
            work()
            torch.cuda.synchronize()

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def host_entry_groups(B, plane_bytes):
    """(first utterance, count) of the groups mas_maximum_path_c_host pipelines over PCIe (mas_api.cu): about
    12 MB each, the first and last a quarter of that.  Only used to count the bytes the entry copies."""
    groups = min(B, max(2, min(16, (plane_bytes + (6 << 20)) // (12 << 20))))
    per = -(-B // groups)
    q = per // 4 if per >= 4 else 0
    sizes, left = [], B

    def push(n):
        nonlocal left
        n = min(n, left)
        if n > 0:
            sizes.append(n)
            left -= n
    if q:
        push(q)
        push(per - q)
        while left > per:
            push(per)
        push(left - q)
        push(q)
    while left > 0:
        push(per)
    out, b0 = [], 0
    for n in sizes:
        out.append((b0, n))
        b0 += n
    return out


def workload_string(name, B, T_y, T_x, ragged, scaling="weak"):
    """`config.workload`, word for word the same in both arms (the driver compares the strings)."""
    per = "/GPU" if scaling == "weak" else " global, split over the GPUs"
    lens = ("variable lengths (SURVEY 8d: t_x~U[T_x/2,T_x], t_y~U[max(t_x,T_y/2),T_y], element 0 full)" if ragged
            else "full-length")
    return (f"{name}: B={B}{per} T_y<={T_y} T_x<={T_x} {lens} with [B,T_y,T_x] fp32 mask, "
            f"neg_cent~N(-400,20^2) fp32 -> fp32 path")


def make_lengths(rng, B, T_y, T_x, ragged):
    if not ragged:
        return np.full(B, T_y, np.int32), np.full(B, T_x, np.int32)
    t_xs = rng.integers((T_x + 1) // 2, T_x + 1, size=B)
    t_ys = np.array([rng.integers(max(tx, (T_y + 1) // 2), T_y + 1) for tx in t_xs])
    t_xs[0], t_ys[0] = T_x, T_y
    order = np.argsort(-t_ys, kind="stable")
    return t_ys[order].astype(np.int32), t_xs[order].astype(np.int32)


# ------------------------------------------------------------------------------------------- CPU reference
def cpu_reference_timing(B, T_y, T_x, t_ys, t_xs, budget_s=10.0, seed=1234):
    """The reference's CPU path on this box's host cores, same workload.  R1: as shipped (its Cython
    built with its setup.py's flags => serial) through the wrapper's marshalling; R2: core only;
    R4: the same Cython rebuilt with -O3 -fopenmp on all cores (steelman)."""
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the reference's wrapper (torch/numpy marshalling around the serial
    # Cython core) gets every host thread it can use, whatever launched us
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    from oracle import mas_oracle
    stock = mas_oracle.load_ref_core("stock")
    omp = mas_oracle.load_ref_core("omp")
    kind = "reference" if stock is not None else "port"
    core = stock if stock is not None else mas_oracle.maximum_path_c
    g = torch.Generator().manual_seed(seed)
    nc = torch.randn(B, T_y, T_x, generator=g) * 20 - 400
    mask = mas_oracle.attn_mask(torch.as_tensor(t_xs), torch.as_tensor(t_ys), T_x, T_y, torch.float32)

    def timeit(fn, budget):
        fn()
        ts = []
        t_end = time.perf_counter() + budget
        while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 200):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)), float(min(ts)), len(ts)

    res = {}
    ref_fn, _, _ = reference_maximum_path()
    med, best, n = timeit(lambda: ref_fn(nc, mask), budget_s * 0.5)
    res["wrapper"] = {"median_s": med, "min_s": best, "reps": n, "alignments_per_s": B / med}
    values = nc.numpy().astype(np.float32)
    ty32, tx32 = np.asarray(t_ys, np.int32), np.asarray(t_xs, np.int32)
    paths = np.zeros(values.shape, np.int32)

    def core_only(c):
        v = values.copy()
        paths.fill(0)
        t0 = time.perf_counter()
        c(paths, v, ty32, tx32)
        return time.perf_counter() - t0

    def time_core(c, budget):
        core_only(c)
        ts, t_end = [], time.perf_counter() + budget
        while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 200):
            ts.append(core_only(c))
        return float(np.median(ts)), float(min(ts)), len(ts)

    med, best, n = time_core(core, budget_s * 0.25)
    res["core_only"] = {"median_s": med, "min_s": best, "reps": n, "alignments_per_s": B / med}
    # (3) what training pays with the reference today: the wrapper fed CUDA tensors (blocking D2H of neg_cent and
    # of the mask sums, serial CPU loop, H2D of the path)
    try:
        if torch.cuda.is_available():
            nc_d, mask_d = nc.cuda(), mask.cuda()

            def wrapped_cuda():
                out = ref_fn(nc_d, mask_d)
                torch.cuda.synchronize()
                return out
            med, best, n = timeit(wrapped_cuda, budget_s * 0.15)
            res["wrapper_cuda_tensors"] = {"median_s": med, "min_s": best, "reps": n, "alignments_per_s": B / med}
            del nc_d, mask_d
    except Exception as ex:  # pragma: no cover
        res["wrapper_cuda_tensors"] = {"error": str(ex)}
    threads = len(os.sched_getaffinity(0))
    if omp is not None:
        os.environ.setdefault("OMP_NUM_THREADS", str(threads))
        med, best, n = time_core(omp, budget_s * 0.25)
        res["omp_steelman_core_only"] = {"median_s": med, "min_s": best, "reps": n, "alignments_per_s": B / med,
                                         "threads": threads}
    cpu_model = ""
    try:
        with open("/proc/cpuinfo") as f:
            cpu_model = next((l.split(":", 1)[1].strip() for l in f if l.startswith("model name")), "")
    except Exception:
        pass
    return kind, res, {"cpu_model": cpu_model, "cpu_count": os.cpu_count(), "affinity": threads}


print_line = lambda line: print(json.dumps(line), flush=True)


def reference_maximum_path():
    """The reference's own `monotonic_align.maximum_path` (its unmodified __init__.py over its own compiled core.pyx,
    installed into the git-ignored baseline/_ref/vits by tools/install_reference.py); when that tree is absent, the
    oracle's restatement of the same wrapper around the compiled core (or around the C port)."""
    pkg = os.path.join(ROOT, "baseline", "_ref", "vits", "monotonic_align")
    if os.path.exists(os.path.join(pkg, "__init__.py")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_monotonic_align_stock", os.path.join(pkg, "__init__.py"),
                                                      submodule_search_locations=[pkg])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_monotonic_align_stock"] = mod
        try:
            spec.loader.exec_module(mod)
            return mod.maximum_path, "reference", "baseline/_ref/vits/monotonic_align: the reference's unmodified __init__.py:7-20 + its core.pyx built with its setup.py's flags (no OpenMP => serial prange)"
        except Exception:
            pass
    from oracle import mas_oracle
    stock = mas_oracle.load_ref_core("stock")
    core = stock if stock is not None else mas_oracle.maximum_path_c
    return (lambda nc, mask: mas_oracle.maximum_path(nc, mask, core=core)), ("reference" if stock is not None else "port"), \
        "oracle.mas_oracle.maximum_path (restatement of __init__.py:7-20) around " + ("the reference's compiled core.pyx" if stock is not None else "the C port")


def reference_worker(job):
    """One CPU process of the reference arm: what one rank of the reference's mp.spawn (train.py:46) pays per step."""
    workload, full_length, steps, warm, threads, shard, nshards, strong = job
    import torch
    torch.set_num_threads(max(1, threads))           # BEFORE the timed loop (torchrun exports OMP_NUM_THREADS=1)
    B, T_y, T_x = WORKLOADS[workload]
    if strong:
        B = max(1, B // nshards)
    rng = np.random.default_rng(1234 + shard)
    t_ys, t_xs = make_lengths(rng, B, T_y, T_x, not full_length)
    fn, kind, path = reference_maximum_path()
    from oracle import mas_oracle
    g = torch.Generator().manual_seed(1234 + shard)
    nc = torch.randn(B, T_y, T_x, generator=g) * 20 - 400
    mask = mas_oracle.attn_mask(torch.as_tensor(t_xs), torch.as_tensor(t_ys), T_x, T_y, torch.float32)
    for _ in range(max(warm, 1)):
        fn(nc, mask)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(nc, mask)
    return time.perf_counter() - t0, B, kind, path, torch.get_num_threads()


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores.  Rank 0 alone runs
    and prints; with --gpus N > 1 it runs N worker processes side by side, one per would-be rank, each on its own shard
    and with its share of the host threads -- the reference under its own mp.spawn (train.py:46) runs one CPU alignment
    per rank -- and reports N x B / (slowest worker's time), so the whole-job figures of the two arms compare."""
    if rank != 0:
        return
    n = max(1, args.gpus)
    strong = args.scaling == "strong"
    B, T_y, T_x = WORKLOADS[args.workload]
    steps = min(args.steps, 50)
    warm = min(args.warmup, 5)
    threads_all = len(os.sched_getaffinity(0))
    threads = max(1, threads_all // n)
    jobs = [(args.workload, args.full_length, steps, warm, threads, i, n, strong) for i in range(n)]
    if n == 1:
        results = [reference_worker(jobs[0])]
    else:
        import multiprocessing as mp
        os.environ["OMP_NUM_THREADS"] = str(threads)
        with mp.get_context("spawn").Pool(n) as pool:
            results = pool.map(reference_worker, jobs)
    dt = max(r[0] for r in results)
    total_B = sum(r[1] for r in results)
    kind, path = results[0][2], results[0][3]
    value = total_B * steps / dt
    import torch
    rng = np.random.default_rng(1234)
    t_ys, t_xs = make_lengths(rng, results[0][1], T_y, T_x, not args.full_length)
    extra = host = None
    if n == 1:
        _, extra, host = cpu_reference_timing(results[0][1], T_y, T_x, t_ys, t_xs, budget_s=6.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, B, T_y, T_x, not args.full_length, args.scaling),
                   "arm": "CPU tensors in and out; " + path,
                   "processes": n, "torch_threads_per_process": results[0][4],
                   "worker_seconds": [r[0] for r in results]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n, "kind": kind,
                         "torch_threads": results[0][4] * n,
                         "sample": f"{steps} full batches of the workload in each of {n} process(es)", "variants": extra, "host": host},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_line(line)


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import vits_b200
    from vits_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    B_cfg, T_y, T_x = WORKLOADS[args.workload]
    strong = args.scaling == "strong"
    # weak: B utterances per GPU (batch_size is per replica in the reference's DDP, train.py:101); strong: the
    # configuration's batch is the GLOBAL batch, split over the ranks (c3, 32 -> 4 per GPU on 8 GPUs)
    B = max(1, B_cfg // world) if strong else B_cfg

    # rotating buffer sets so that consecutive steps never find their input or output in L2
    plane_bytes = B * T_y * T_x * 4
    nbuf = max(2, min(8, int(np.ceil(3 * 126e6 / (2 * plane_bytes)))))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ncs = [torch.randn(B, T_y, T_x, generator=g, device=dev) * 20 - 400 for _ in range(nbuf)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_variant(ragged, steps, warmup, sample_clocks):
        """One workload variant: parity gate on rank 0, graph capture, W warm-up + exactly K timed steps."""
        rng = np.random.default_rng(1234 + rank)
        t_ys, t_xs = make_lengths(rng, B, T_y, T_x, ragged)
        ty_d = torch.as_tensor(t_ys, device=dev)
        tx_d = torch.as_tensor(t_xs, device=dev)
        ym = torch.arange(T_y, device=dev)[None, :] < ty_d[:, None]
        xm = torch.arange(T_x, device=dev)[None, :] < tx_d[:, None]
        mask = (ym[:, :, None] & xm[:, None, :]).float()          # [B,T_y,T_x] like attn_mask.squeeze(1)
        outs = [None] * nbuf

        def step(i):
            outs[i % nbuf] = vits_b200.maximum_path(ncs[i % nbuf], mask)

        # correctness gate before timing: oracle on rank 0's first buffer (bit-exact)
        parity = None
        if rank == 0:
            from oracle import mas_oracle
            want = mas_oracle.maximum_path_numpy(ncs[0].cpu().numpy(), t_ys, t_xs)
            step(0)
            torch.cuda.synchronize()
            parity = bool(np.array_equal(outs[0].cpu().numpy().astype(np.int32), want))
            assert parity, "GPU path differs from the oracle -- refusing to report a number"

        # CUDA graphs remove the Python/ctypes enqueue cost from the device timeline.  One graph holds one
        # or more passes over all `nbuf` rotating buffer sets; single-step graphs cover the remainder so that
        # EXACTLY `steps` steps are timed.
        graphs = multi = None
        launches_per_step = None
        # passes over the buffer sets per graph: the ~10 us between two graph launches are harness cost, not the
        # path's; 20 steps per graph keep them under 2 % of the timed region
        passes = max(1, min(-(-20 // nbuf), steps // nbuf))
        if not args.no_graph:
            try:
                for i in range(nbuf):
                    step(i)
                torch.cuda.synchronize()
                graphs = []
                for i in range(nbuf):
                    gr = torch.cuda.CUDAGraph()
                    n0 = _lib.launch_count()
                    with torch.cuda.graph(gr):
                        step(i)
                    launches_per_step = _lib.launch_count() - n0
                    graphs.append(gr)
                multi = torch.cuda.CUDAGraph()
                with torch.cuda.graph(multi):
                    for i in range(nbuf * passes):
                        step(i)
            except Exception as e:  # pragma: no cover
                print(f"[bench] CUDA graph capture failed ({e}); timing eager launches", file=sys.stderr)
                graphs = multi = None

        def run_steps(n):
            if graphs is None:
                for i in range(n):
                    step(i)
                return
            per = nbuf * passes
            for _ in range(n // per):
                multi.replay()
            if n % per:
                tail_graph(n % per).replay()

        tails = {}

        def tail_graph(k):
            """k < steps_per_graph consecutive steps as one graph (captured once, before the timed region)."""
            if k not in tails:
                gk = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gk):
                    for i in range(k):
                        step(i)
                tails[k] = gk
            return tails[k]

        if graphs is not None:
            for n in (warmup, steps, nbuf):
                if n % (nbuf * passes):
                    tail_graph(n % (nbuf * passes))
            # every graph object the timed region replays has been replayed at least once before it (the first
            # replay of a graph pays its upload; r1's driver run timed exactly that: 46 instead of 40 us per step)
            if steps // (nbuf * passes):
                multi.replay()
            if steps % (nbuf * passes):
                tail_graph(steps % (nbuf * passes)).replay()
            torch.cuda.synchronize()

        # nvidia-smi needs a good fraction of a second to start (longer with several ranks starting one at once),
        # the K timed steps may last only milliseconds: rank 0 starts it first and keeps the GPU busy with the
        # same steps, untimed, until the first sample is in and again after the timed region until one more is,
        # so the samples bracket the timed region under its own load.
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        run_steps(warmup)
        if sampler:
            sampler.busy_until(lambda: run_steps(nbuf), len(sampler.rows) + 1, 6.0)
        barrier()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(steps)
        e1.record()
        n1 = _lib.launch_count()
        if sampler:
            sampler.busy_until(lambda: run_steps(nbuf), len(sampler.rows) + 1, 1.5)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = (n1 - n0) if graphs is None else launches_per_step * steps
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        alg_bytes = 4 * int(np.sum(t_ys.astype(np.int64) * t_xs)) + 4 * B * T_y * T_x
        return dict(t_ys=t_ys, t_xs=t_xs, ty_d=ty_d, tx_d=tx_d, mask=mask, ms_max=ms_max, step_ms=ms_max / steps,
                    value=world * B * steps / (ms_max * 1e-3), clocks=clocks, launches=int(launches), parity=parity,
                    graph=graphs is not None, alg_bytes=alg_bytes, steps_per_graph=nbuf * passes)

    # primary: the BASELINE.json configuration (variable lengths with masks) unless --full-length; the other
    # variant is timed with fewer steps and reported beside it
    primary_ragged = not args.full_length
    R = time_variant(primary_ragged, args.steps, args.warmup, True)
    O = time_variant(not primary_ragged, max(20, args.steps // 4), max(3, args.warmup // 2), False)
    t_ys, t_xs, ty_d, tx_d, mask = R["t_ys"], R["t_xs"], R["ty_d"], R["tx_d"], R["mask"]
    ms_max, value, clocks, launches, parity = R["ms_max"], R["value"], R["clocks"], R["launches"], R["parity"]
    graphs = True if R["graph"] else None

    # --- per-kernel durations: serialised pass (PDL off, eager), events between the three kernels ---
    kernels_ms = None
    breakdown = None
    if rank == 0:
        kernels_ms = per_kernel_ms(L, ncs, mask, B, T_y, T_x, dev)
        if not args.no_breakdown:
            try:
                breakdown = contraction_and_e2e(B, T_y, T_x, t_ys, t_xs, dev)
            except Exception as ex:  # pragma: no cover
                breakdown = {"error": str(ex)}

    # --- end to end through the host-buffer C entry (pinned host memory) ---
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 20))
        h_vals = [torch.empty(B, T_y, T_x, dtype=torch.float32).pin_memory() for _ in range(2)]
        for hv, nc in zip(h_vals, ncs):
            hv.copy_(nc)
        # zero-filled once, as the reference's caller does (np.zeros, __init__.py:15): the entry writes every
        # row below t_y_i in full and, like core.pyx, never touches the padded rows
        h_paths = [torch.zeros(B, T_y, T_x, dtype=torch.int32).pin_memory() for _ in range(2)]
        h_ty, h_tx = torch.as_tensor(t_ys), torch.as_tensor(t_xs)
        ty_np = np.clip(np.asarray(t_ys), 0, T_y)
        valid_row_bytes = int(4 * T_x * sum(int(ty_np[b0:b0 + n].max()) * n for b0, n in host_entry_groups(B, plane_bytes)))

        def host_step(i):
            rc = L.mas_maximum_path_c_host(h_paths[i % 2].data_ptr(), h_vals[i % 2].data_ptr(), h_ty.data_ptr(),
                                           h_tx.data_ptr(), B, T_y, T_x)
            assert rc == 0, rc
        for i in range(3):
            host_step(i)
        # three timed blocks of e2e_steps calls each, the median block reported (the host side of this leg -- copies
        # out of host memory, the entry's host threads -- shares the box's cores and memory with whatever else runs there:
        # single blocks of the same binary on the same box ranged 62-71 k alignments/s); every block is in `blocks_s`
        blocks = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                host_step(i)
            tb = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            blocks.append(float(tb.item()))
        dt = sorted(blocks)[1]
        td = torch.tensor([dt], device=dev, dtype=torch.float64)
        # the link itself, for scale: one plain pinned copy of a whole [B,T_y,T_x] plane each way
        link = None
        all_ranks_duplex = None
        if world > 1:
            # every rank's link busy in both directions at once: what the HOST (root complexes, memory) sustains
            # when all the ranks copy together -- the bound of this leg beyond a few GPUs per host
            da, db = (torch.empty(B, T_y, T_x, dtype=torch.float32, device=dev) for _ in range(2))
            sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

            def both_all(n):
                for _ in range(n):
                    with torch.cuda.stream(sa):
                        da.copy_(h_vals[0], non_blocking=True)
                    with torch.cuda.stream(sb):
                        h_vals[1].copy_(db, non_blocking=True)
                torch.cuda.synchronize()
            both_all(1)
            barrier()
            t1 = time.perf_counter()
            both_all(5)
            tl = torch.tensor([time.perf_counter() - t1], device=dev, dtype=torch.float64)
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
            per_plane_s = float(tl.item()) / 5
            all_ranks_duplex = {"aggregate_gbs": world * 2 * plane_bytes / per_plane_s / 1e9,
                                "bound": world * B / (per_plane_s * valid_row_bytes / plane_bytes),
                                "note": "all ranks copying one plane each way at the same time (pinned, max over ranks); "
                                        "bound = whole-job alignments/s if a step cost exactly those copies"}
            h_vals[1].copy_(ncs[1 % len(ncs)])
            del da, db
        tmin = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            d_tmp = torch.empty(B, T_y, T_x, dtype=torch.float32, device=dev)
            gbs = []
            for src, dst in ((h_vals[0], d_tmp), (d_tmp, h_vals[1])):
                dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                for _ in range(5):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                gbs.append(5 * plane_bytes / (time.perf_counter() - t1) / 1e9)
            d_tmp2 = torch.empty_like(d_tmp)
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

            def both(n):
                for _ in range(n):
                    with torch.cuda.stream(s1):
                        d_tmp.copy_(h_vals[0], non_blocking=True)
                    with torch.cuda.stream(s2):
                        h_vals[1].copy_(d_tmp2, non_blocking=True)
                torch.cuda.synchronize()
            both(1)
            t1 = time.perf_counter()
            both(5)
            duplex_s = (time.perf_counter() - t1) / 5
            h_vals[1].copy_(ncs[1 % len(ncs)])
            link = {"h2d_gbs": gbs[0], "d2h_gbs": gbs[1], "duplex_gbs": 2 * plane_bytes / duplex_s / 1e9,
                    "h2d_bound": B / (valid_row_bytes / (gbs[0] * 1e9)),
                    "duplex_bound": B / (duplex_s * valid_row_bytes / plane_bytes),
                    "note": "pinned cudaMemcpy of one whole plane each way, alone and then both at once; h2d_bound = "
                            "alignments/s if a step cost exactly its inbound copy (what the entry moves now: the path "
                            "comes back as the 4-byte-per-frame index and is materialised by host threads); duplex_bound "
                            "= the same if the dense path crossed the link too (round 1's pipeline, MAS_HOST_DENSE=1)"}
            del d_tmp, d_tmp2
        if rank == 0:
            from oracle import mas_oracle
            want = mas_oracle.maximum_path_numpy(h_vals[0].numpy(), t_ys, t_xs)
            assert np.array_equal(h_paths[0].numpy(), want), "e2e path differs from the oracle"
        e2e = {"value": world * B * e2e_steps / float(td.item()), "unit": UNIT,
               "h2d_bytes_per_step": valid_row_bytes + 8 * B, "d2h_bytes_per_step": 4 * B * T_y + 64,
               "copied": "in: leading rows of every utterance's neg_cent up to the longest of its group (the padded tail "
                         "is never needed, as in core.pyx:13-33); out: the int32 per-frame index [B,T_y] -- the dense int32 "
                         "path (one 1 per frame) is written into the caller's buffer by the entry's host threads, rows "
                         "below t_y in full (zeros and the one); host paths buffer zero-filled once by the caller",
               "pcie_link": link, "pcie_all_ranks_duplex": all_ranks_duplex,
               "blocks_s": blocks, "rank_seconds": {"min": float(tmin.item()), "max": float(td.item())},
               "steps": e2e_steps, "timer": "host wall clock around the synchronous C call, max over ranks; three blocks of "
                                            "`steps` calls, the median block is the value",
               "api": "mas_maximum_path_c_host (twin of core.pyx:38), pinned host buffers"}
        # the same leg through the repo's own PYTHON API, exactly as a user of the reference would call it with CPU
        # tensors: ordinary (pageable) torch tensors in, a new fp32 tensor out, mask given as the dense tensor
        py_steps = max(3, min(args.steps, 10))
        nc_cpu = [hv.clone() for hv in h_vals]          # pageable copies
        mask_cpu = mask.cpu()
        for i in range(2):
            out_cpu = vits_b200.maximum_path(nc_cpu[i % 2], mask_cpu)
        barrier()
        t0 = time.perf_counter()
        for i in range(py_steps):
            out_cpu = vits_b200.maximum_path(nc_cpu[i % 2], mask_cpu)
        dtp = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dtp, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert out_cpu.dtype == torch.float32 and np.array_equal(
                out_cpu.numpy().astype(np.int32), mas_oracle.maximum_path_numpy(nc_cpu[(py_steps - 1) % 2].numpy(), t_ys, t_xs))
        # ... and the C entry on PAGEABLE arrays, which is what the reference's own wrapper hands its native core
        # (np.zeros / astype(np.float32) results, __init__.py:14-15) and what the compiled binding receives: `values` is
        # staged into a pinned mirror by the entry's host threads
        pg_vals = [hv.clone() for hv in h_vals]
        pg_paths = [torch.zeros(B, T_y, T_x, dtype=torch.int32) for _ in range(2)]

        def pageable_step(i):
            rc = L.mas_maximum_path_c_host(pg_paths[i % 2].data_ptr(), pg_vals[i % 2].data_ptr(), h_ty.data_ptr(),
                                           h_tx.data_ptr(), B, T_y, T_x)
            assert rc == 0, rc
        for i in range(2):
            pageable_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(py_steps):
            pageable_step(i)
        dtg = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dtg, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert np.array_equal(pg_paths[(py_steps - 1) % 2].numpy(),
                                  mas_oracle.maximum_path_numpy(pg_vals[(py_steps - 1) % 2].numpy(), t_ys, t_xs))
        e2e["c_entry_pageable"] = {"value": world * B * py_steps / float(dtg.item()), "unit": UNIT, "steps": py_steps,
                                   "api": "mas_maximum_path_c_host on ordinary (pageable) host arrays, paths zero-filled once"}
        del pg_vals, pg_paths
        e2e["python_api"] = {"value": world * B * py_steps / float(dtp.item()), "unit": UNIT, "steps": py_steps,
                             "api": "vits_b200.maximum_path(neg_cent_cpu, mask_cpu): pageable CPU tensors in, new fp32 CPU "
                                    "tensor out (lengths from the mask on the host, mas_maximum_path_host inside)"}
        del nc_cpu, mask_cpu, out_cpu
        L.mas_host_release()

    # --- c3 strong scaling (BASELINE.json configs[2]: B = 32 global, batch-sharded over the GPUs), reported as an
    # extra key of every N > 1 line; `--scaling strong --workload c3` makes it the headline instead ---
    c3_strong = None
    if world > 1 and not (strong and args.workload == "c3"):
        c3_strong = strong_scaling_probe("c3", world, rank, dev, barrier)

    # --- multi-GPU verification (outside the timed region): all-gather the per-frame indices over NCCL ---
    verified = None
    if world > 1:
        idx = vits_b200.maximum_path_index(ncs[0], mask)
        gathered = torch.empty(world * B, T_y, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(gathered, idx)
        lens = torch.stack([ty_d, tx_d], 1).to(torch.int32)
        all_lens = torch.empty(world * B, 2, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_lens, lens)
        if rank == 0:
            gl = all_lens.cpu().numpy()
            ok = vits_b200.shard.check_index(gathered, gl[:, 0], gl[:, 1])
            verified = ok
            assert ok, "gathered paths violate the alignment invariants"

    if rank == 0:
        alg_bytes = R["alg_bytes"]
        peak, peak_src = measured_peak_gbs()
        step_ms = R["step_ms"]
        achieved = alg_bytes / (step_ms * 1e-3) / 1e9
        other = {"workload": "full-length" if primary_ragged else "ragged lengths", "value": O["value"], "unit": UNIT,
                 "ms_per_step": O["step_ms"], "steps": max(20, args.steps // 4),
                 "roofline_frac": O["alg_bytes"] / (O["step_ms"] * 1e-3) / 1e9 / peak, "parity_checked": O["parity"]}
        # DRAM bytes of one step's three kernels from `ncu --set full` (profiles/): only captured for c2, B = 64
        traffic, traffic_note = None, "not captured for this workload"
        if args.workload == "c2" and primary_ragged:
            traffic = 44.20e6 + 50.33e6
            traffic_note = ("c2 variable lengths, profiles/r02_final_chain_c2_variable_lengths_summary.txt (ncu --set full over one "
                            "call's three kernels): dram read+write inside the kernels' windows = mas_dp2 40.80 MB (30.6 MB "
                            "algorithmic: 32x64 TMA boxes overhang the band and t_x) + mas_writeout 1.40 MB + "
                            "mas_backtrack_stream 2.02 MB = 44.2 MB, plus the 50.33 MB dense path that ncu sees absorbed by the "
                            "126 MB L2 and that is written back after the window (counted here once, as it must reach HBM)")
        elif args.workload == "c2":
            traffic = 105.4e6
            traffic_note = ("full-length c2, profiles/r02_final_mas_dp2_summary.txt: mas_dp2 50.4 MB DRAM read + 1.6 MB of tagged "
                            "decision words (and 3.1 MB of tag clears), mas_writeout 50.3 MB written")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(args.workload, B_cfg, T_y, T_x, primary_ragged, args.scaling),
                       "arm": f"device-resident tensors, {B} utterances on each of {world} GPU(s)",
                       "other_variant": other,
                       "l2": f"inputs larger than L2: {nbuf} rotating (neg_cent, path) buffer sets = "
                             f"{2 * nbuf * plane_bytes / 1e6:.0f} MB, no flush kernel in the timed region",
                       "launch": (f"CUDA graph replay, {R['steps_per_graph']} consecutive steps per graph cycling over the {nbuf} rotating buffer sets" if graphs is not None else "eager ctypes launches"),
                       "parity_checked": parity, "multi_gpu_verified": verified},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "path_breakdown": breakdown,
            "c3_strong": c3_strong,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": alg_bytes,
                         "kernel": "maximum_path chain (mas_dp2 wavefront forward kernel + mas_writeout zero-fill + "
                                   "mas_backtrack_stream, concurrent on disjoint SMs through programmatic dependent launch, "
                                   "timed as one unit with CUDA events on the launching stream)",
                         "phase_timeline_us": kernels_ms},
        }
        if not args.no_cpu and world == 1:
            kind, res, host = cpu_reference_timing(B, T_y, T_x, t_ys, t_xs, budget_s=args.cpu_budget)
            if e2e is not None:   # what the host-buffer leg is compared with, level by level
                e2e["vs_cpu"] = {
                    "c_entry_vs_reference_wrapper": e2e["value"] / res["wrapper"]["alignments_per_s"],
                    "c_entry_vs_reference_core_only": e2e["value"] / res["core_only"]["alignments_per_s"],
                    "c_entry_vs_openmp_steelman_core_only": (e2e["value"] / res["omp_steelman_core_only"]["alignments_per_s"]
                                                             if "omp_steelman_core_only" in res else None),
                    "python_api_vs_reference_wrapper": e2e["python_api"]["value"] / res["wrapper"]["alignments_per_s"],
                    "note": "same-level pairs: C entry <-> core_only (core.pyx:38 on prepared arrays); Python API <-> wrapper "
                            "(__init__.py:7-20 on CPU tensors).  The OpenMP steelman is the reference's core rebuilt with "
                            "-fopenmp on every host thread; this leg is PCIe-bound (pcie_link.duplex_bound)"}
            line["cpu_baseline"] = {"value": res["wrapper"]["alignments_per_s"], "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": f"{res['wrapper']['reps']} full batches of the same workload "
                                              "(reference wrapper marshalling + compiled core.pyx, as shipped: serial)",
                                    "variants": res, "host": host}
        print_line(line)
    if world > 1:
        dist.destroy_process_group()


def strong_scaling_probe(workload, world, rank, dev, barrier, steps=40):
    """The workload's batch as a GLOBAL batch split over the ranks (c3, 32 -> 4 per GPU on 8 GPUs):
    every rank aligns B/world utterances; whole-job alignments/s = B / (slowest rank's time per step)."""
    import torch
    import torch.distributed as dist
    import vits_b200
    Bg, T_y, T_x = WORKLOADS[workload]
    B = max(1, Bg // world)
    rng = np.random.default_rng(4321 + rank)
    t_ys, t_xs = make_lengths(rng, B, T_y, T_x, True)
    ty_d, tx_d = torch.as_tensor(t_ys, device=dev), torch.as_tensor(t_xs, device=dev)
    ym = torch.arange(T_y, device=dev)[None, :] < ty_d[:, None]
    xm = torch.arange(T_x, device=dev)[None, :] < tx_d[:, None]
    mask = (ym[:, :, None] & xm[:, None, :]).float()
    nbuf = 8
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    ncs = [torch.randn(B, T_y, T_x, generator=g, device=dev) * 20 - 400 for _ in range(nbuf)]
    outs = [None] * nbuf

    def step(i):
        outs[i % nbuf] = vits_b200.maximum_path(ncs[i % nbuf], mask)
    for i in range(nbuf):
        step(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(steps):
            step(i)
    gr.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"workload": workload_string(workload, Bg, T_y, T_x, True, "strong"), "scaling": "strong", "n_gpus": world,
            "utterances_per_gpu": B, "ms_per_step": ms, "value": B * world / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "note": "device-resident inputs, one CUDA graph of all steps over 8 rotating buffers, max over ranks"}


def contraction_and_e2e(B, T_y, T_x, t_ys, t_xs, dev, C=192, reps=5):
    """neg_cent alone (ours vs the reference's fp32 torch expression, same
    inputs, HBM-resident) and z_p,m_p,logs_p -> path end to end on the device, CUDA-graph timed."""
    import torch
    import vits_b200
    from oracle import mas_oracle
    g = torch.Generator(device=dev).manual_seed(4321)
    nset = 3
    sets = [(torch.randn(B, C, T_y, generator=g, device=dev), torch.randn(B, C, T_x, generator=g, device=dev),
             torch.randn(B, C, T_x, generator=g, device=dev) * 0.3) for _ in range(nset)]
    y_len = torch.as_tensor(t_ys, device=dev)
    x_len = torch.as_tensor(t_xs, device=dev)

    def timeit(fn):
        for st in sets[:2]:
            fn(*st)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            keep = [fn(*st) for st in sets]
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        del keep
        return e0.elapsed_time(e1) / (reps * nset) * 1e3

    ref = mas_oracle.neg_cent_torch(*sets[0])
    ours = vits_b200.neg_cent(*sets[0])
    rel = float(((ours - ref).abs().amax() / ref.abs().amax()).item())
    t_ref = timeit(mas_oracle.neg_cent_torch)
    t_nc = timeit(vits_b200.neg_cent)
    t_e2e = timeit(lambda z, m, ls: vits_b200.maximum_path_from_lengths(vits_b200.neg_cent(z, m, ls), y_len, x_len))
    flops = 4.0 * B * T_y * T_x * C
    nbytes = 4.0 * (B * C * T_y + 2 * B * C * T_x + B * T_y * T_x)
    return {"neg_cent_us": t_nc, "neg_cent_torch_fp32_us": t_ref, "neg_cent_rel_err_of_max": rel,
            "neg_cent_algorithmic_gflop": flops / 1e9, "neg_cent_algorithmic_mb": nbytes / 1e6,
            "neg_cent_gbs": nbytes / t_nc / 1e3, "neg_cent_tflops": flops / t_nc / 1e6,
            "stats_to_path_us": t_e2e, "stats_to_path_alignments_per_s": B / t_e2e * 1e6,
            "note": "device-resident inputs, CUDA-graph replay over 3 rotating input sets; torch = the reference's "
                    "inline fp32 expression (SynthesizerTrn.py:223-232)"}


def per_kernel_ms(L, ncs, mask, B, T_y, T_x, dev, reps=5):
    """Phase timeline of one call from %globaltimer stamps written by the kernels themselves
    (mas_set_timeline): microseconds relative to the first forward CTA's start, median of `reps`."""
    import torch
    import vits_b200
    names = ["fwd_first_start", "fwd_last_dp_done", "fwd_last_end", "bt_first_start", "bt_last_end",
             "wo_first_start", "wo_last_zero_fill_done", "wo_last_end|wavefront:lengths_known",
             "bt_top_group_words_seen", "bt_all_tables_done", "bt_chain_over_groups_done", "bt_last_table_done",
             "bt_top_group_walked", "bt_last_table_words_complete", "bt16_walk_windows_loaded", "bt16_groups_rewalked"]
    tl = torch.zeros(16, dtype=torch.int64, device=dev)   # 16 slots (include/vits_mas.h: mas_set_timeline)
    out = {}
    try:
        for pdl in (1, 0):
            L.mas_set_tuning(0, 0, 0, pdl)
            runs = []
            for i in range(reps + 1):
                tl.zero_()
                tl[0] = tl[3] = tl[5] = -1
                torch.cuda.synchronize()
                L.mas_set_timeline(tl.data_ptr())
                vits_b200.maximum_path(ncs[i % len(ncs)], mask)
                torch.cuda.synchronize()
                L.mas_set_timeline(None)
                v = tl.cpu().numpy().astype(np.uint64)
                runs.append([(int(x) - int(v[0])) / 1e3 if int(x) not in (0, 2 ** 64 - 1) else None for x in v])
            runs = runs[1:]
            med = {}
            for k, n in enumerate(names):
                vals = [r[k] for r in runs if r[k] is not None]
                if vals:
                    med[n] = float(np.median(vals))
            out["pdl" if pdl else "serialised"] = med
    finally:
        L.mas_set_tuning(0, 0, 0, -1)
        L.mas_set_timeline(None)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS) + ["c5"],
                    help="c1-c4: the alignment path alone (SURVEY.md 8d shapes); c5: the reference's training step with "
                         "the drop-in (bench_c5.py)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's B per GPU; strong: the workload's B is the global batch, split over the GPUs")
    ap.add_argument("--c5-batch", type=int, default=64, help="utterances per GPU in the c5 training step")
    ap.add_argument("--full-length", action="store_true",
                    help="headline on the full-length variant (default: variable lengths, BASELINE.json configs[1])")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the neg_cent / stats->path timings")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL announces its version on
    # stdout when NCCL_DEBUG is set) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_line
    print_line = lambda line: print(json.dumps(line), file=json_out, flush=True)
    if args.workload == "c5":
        import bench_c5
        bench_c5.main(args, print_line)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

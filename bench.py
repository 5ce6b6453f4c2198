#!/usr/bin/env python
"""Benchmark of the hot path: SymEigsSolver + SparseSymMatProd on a synthetic sparse matrix.

    python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a CUDA through the C ABI)
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU (oracle port)

Workload (BASELINE.json `metric`): G_sym(n = 1e7, 20 nnz/row, seed 0), nev = 20, ncv = 60, tol 1e-10,
selection LargestAlge, default SimpleRandom(0) initial residual.  One "step" = one complete solve
(init() + compute()).  `value` = matrix operations (SpMV iterations) per second of device time with the
operator already resident in HBM; `e2e` = the same metric through the reference-facing call sequence with
HOST buffers (CSR upload + triangle expansion, solve, eigenvalues and eigenvectors copied back).
Strong scaling: n stays 1e7 for every N; rows are sharded across the N ranks (one process per GPU).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=float, default=1e7)
    ap.add_argument("--nnz-per-row", type=int, default=20)
    ap.add_argument("--nev", type=int, default=20)
    ap.add_argument("--ncv", type=int, default=60)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample-ops", type=int, default=12, help="matrix operations in the bounded CPU sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="do not run the other BASELINE configurations (C2, C3, C5) after the timed work")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        if shutil.which("nvidia-smi"):
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.th = threading.Thread(target=self._read, daemon=True)
                self.th.start()
            except Exception:
                self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def pinned_alloc():
    """Host arrays in pinned memory (torch is only the allocator here)."""
    import torch

    keep = []

    def alloc(count, dtype):
        t = torch.empty(int(count), dtype={np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)],
                        pin_memory=True)
        keep.append(t)
        return t.numpy()

    return alloc, keep


def host_threads():
    """Hardware threads this process may use.  Deliberately NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which would
    silently shrink the CPU arm at N > 1; the oracle takes its team size as an explicit num_threads() clause."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def golden_full_solves():
    """Full CPU solves of the BASELINE configurations, run once with tests/golden/make_baseline_golden.py and committed (the n = 1e7
    solve takes hours): operations, seconds, thread count and the host they were timed on."""
    out = {}
    for name in ("C2", "C4"):
        try:
            with open(os.path.join(ROOT, "tests", "golden", f"baseline_{name}.json")) as fh:
                g = json.load(fh)
            out[name] = {"nops": g["nops"], "niter": g["niter"], "nconv": g["nconv"], "seconds": g["seconds"], "threads": g["threads"],
                         "spmv_iters_per_sec": g["nops"] / g["seconds"], "host": "development container, 8 vCPU (not the GPU box)"}
        except (OSError, KeyError, ValueError):
            pass
    # ... and of the reference's own code (oracle/_ref: yixuan/spectra's headers over the Eigen stand-in, one thread, make_reference_golden.py)
    for name in ("C2",):
        try:
            with open(os.path.join(ROOT, "tests", "golden", f"reference_{name}.json")) as fh:
                g = json.load(fh)
            out[name + "_reference_own_code"] = {"nops": g["nops"], "niter": g["niter"], "nconv": g["nconv"], "seconds": g["solve_seconds"], "threads": 1,
                                                 "spmv_iters_per_sec": g["nops"] / g["solve_seconds"], "library": g["library"],
                                                 "host": "development container (not the GPU box)"}
        except (OSError, KeyError, ValueError):
            pass
    return out


def reference_own_code_sample(n, rp, ci, v, sample_ops):
    """The same bounded sample through THE REFERENCE'S OWN CODE: oracle/_ref/libspectra_ref.so is yixuan/spectra's headers compiled in the
    development container (`make -C oracle ref`; Eigen replaced by oracle/eigen_standin -- scalar loops, no SIMD, so real Eigen would be
    somewhat faster on the panel products).  Single-threaded, as the reference ships.  None when the library did not travel."""
    try:
        from oracle import ref as R

        if not os.path.exists(R._LIB_PATH) or n * 21 >= 2 ** 31:
            return None
        A = R.Compressed(n, np.asarray(rp, dtype=np.int32), ci, v, order="col")  # full symmetric CSR == the same matrix column-major; Lower is read
        nops, sec = R.lanczos_sample(A, sample_ops - 1)
        return {"value": nops / sec, "unit": "SpMV-iters/s", "cores": 1, "kind": "reference", "library": R.version(),
                "sample": f"Lanczos::init + first {nops - 2} steps of Lanczos::factorize_from over SparseSymMatProd<double> (the head of the n={n} solve), {sec:.1f} s"}
    except Exception as e:  # noqa: BLE001 -- an optional extra arm must not break the line
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args, n):
    """The reference algorithm (CPU restatement in oracle/) on bounded samples of the workload: once single-threaded -- what the
    reference, a single-threaded header library, does as shipped -- and once with the fastest OpenMP team of this host.  `value` is the
    team figure (the stronger baseline); both are in the line so that ratios at different N compare like with like."""
    from spectra_b200 import dist, synth
    rank, _, world = dist.env_rank()
    if rank != 0:
        return
    import oracle as O

    rp, ci, v = synth.csr(n, args.nnz_per_row, args.seed, True)
    A = O.Csr.adopt(n, rp, ci, v)
    avail = host_threads()
    cands = sorted({t for t in (1, 4, 8, 16, 32, 64, 128, avail) if t <= avail})
    xcal = np.ones(n)
    calib = {}
    for t in cands:
        A.set_threads(t)
        A.spmv(xcal)
        t0 = time.perf_counter()
        A.spmv(xcal)
        calib[t] = time.perf_counter() - t0
    threads = min(calib, key=calib.get)
    del xcal

    def sample(nthreads, reps, warm):
        times, ops = [], 0
        for it in range(warm + reps):
            r = O.sym_eigs(A, args.nev, args.ncv, O.LargestAlge, 1000, args.tol, threads=nthreads, op_limit=args.cpu_sample_ops, want_vectors=False)
            if it >= warm:
                times.append(r.seconds)
                ops = r.nops
        return ops, float(np.mean(times))

    ops, sec = sample(threads, args.steps, args.warmup)
    ops1, sec1 = sample(1, 1, 0)
    value = ops / sec
    ref_own = reference_own_code_sample(n, rp, ci, v, args.cpu_sample_ops)
    what = (f"init + first {ops - 2} Lanczos steps (op_limit={args.cpu_sample_ops}) of the n={n} solve, {threads} OpenMP threads "
            f"(fastest team of {cands} on the SpMV; {avail} hardware threads available, OMP_NUM_THREADS ignored)")
    line = {
        "impl": "reference", "metric": "spmv_iters_per_sec", "value": value, "unit": "SpMV-iters/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n),
        "cpu_baseline": {"value": value, "unit": "SpMV-iters/s", "cores": threads, "kind": "port", "sample": what,
                         "single_thread_value": ops1 / sec1, "single_thread_note": "the reference as shipped is single-threaded; same sample, 1 thread",
                         "reference_own_code": ref_own,
                         "value_is": "the fastest CPU arm measured here (the OpenMP port): the conservative denominator for GPU/CPU ratios",
                         "spmv_seconds_by_team": {str(k): round(val, 4) for k, val in calib.items()},
                         "full_solves_cached": golden_full_solves()},
        "e2e": {"value": value, "unit": "SpMV-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": f"SymEigsSolver<SparseSymMatProd<double>> G_sym n={n} nnz/row={args.nnz_per_row} nev={args.nev} ncv={args.ncv} tol={args.tol:g} "
                        f"selection=LargestAlge init=SimpleRandom(0)", "n": n, "nnz_per_row": args.nnz_per_row, "nev": args.nev, "ncv": args.ncv,
            "parallelism": f"row-sharded x{args.gpus}", "l2": "inputs larger than L2 (CSR + Krylov basis >> 126 MB); no flush needed"}


def run_other_configs(sb, synth):
    """BASELINE.json configs[1], [2], [4] on one GPU: SpMV-iters/s of one complete solve (device time of init() + compute())."""
    import scipy.sparse as sp

    out = {}

    def solve(make, sel, maxit=1000):
        best = None
        for _ in range(2):  # first solve warms up (allocation, clocks); the second is reported
            e = make()
            e.init()
            nconv = e.compute(sel, maxit)
            st = e.stats()
            best = {"nconv": int(nconv), "num_operations": int(e.num_operations()), "num_iterations": int(e.num_iterations()), "info": int(e.info()),
                    "ms": st["ms_total"], "spmv_iters_per_sec": e.num_operations() / (st["ms_total"] / 1e3), "eigenpairs_per_sec": nconv / (st["ms_total"] / 1e3),
                    "host_syncs": st["host_syncs"], "kernel_launches": st["kernel_launches"]}
            e.close()
        return best

    n2 = 1_000_000
    rp, ci, v = synth.csr(n2, 20, 0, True)
    op = sb.SparseGenMatProd.from_csr_slab(n2, 0, rp, ci, v)
    out["C2_sym_n1e6_k20_ncv60_LargestAlge"] = solve(lambda: sb.SymEigsSolver(op, 20, 60), sb.SortRule.LargestAlge)
    op.close()
    rp, ci, v = synth.csr(n2, 20, 1, False)
    op = sb.SparseGenMatProd.from_csr_slab(n2, 0, rp, ci, v)
    r = solve(lambda: sb.GenEigsSolver(op, 10, 30), sb.SortRule.LargestMagn, 40)
    r["note"] = "G_gen unplanted, maxit = 40 (the circular-law spectrum does not converge within the default 1000 restarts either); throughput of the Arnoldi loop"
    out["C3_gen_n1e6_k10_ncv30_LargestMagn_maxit40"] = r
    op.close()
    n5 = 200_000
    rp, ci, v = synth.band_csr(n5, 15, 0, 0.0)
    A = sp.csr_matrix((v, ci, rp), shape=(n5, n5))
    ops = sb.SparseSymShiftSolve(sp.tril(A).tocsc())
    out["C5_shift_invert_n2e5_band15_k10_ncv30_sigma0.5"] = solve(lambda: sb.SymEigsShiftSolver(ops, 10, 30, 0.5), sb.SortRule.LargestMagn)
    ops.close()
    # the same solver over a mesh-like pattern (27-point stencil on 30^3 points, half-bandwidth 931): the wide-band route of the shift-solve
    # (sequential block elimination with grid-wide block kernels); one solve includes nothing of set_shift, which is timed separately
    try:
        import itertools

        dims = (30, 30, 30)
        nm = int(np.prod(dims))
        idx = np.arange(nm).reshape(dims)
        rng = np.random.default_rng(5)
        rows, cols = [], []
        for o in (o for o in itertools.product((-1, 0, 1), repeat=3) if any(o)):
            src = tuple(slice(max(0, -d), dims[a] - max(0, d)) for a, d in enumerate(o))
            dst = tuple(slice(max(0, d), dims[a] - max(0, -d)) for a, d in enumerate(o))
            rows.append(idx[src].ravel())
            cols.append(idx[dst].ravel())
        r_, c_ = np.concatenate(rows), np.concatenate(cols)
        M = sp.csr_matrix((rng.uniform(-1, 1, r_.size), (r_, c_)), shape=(nm, nm))
        M = ((M + M.T) / 2 + sp.diags(rng.uniform(-1, 1, nm) + 0.3)).tocsc()
        opm = sb.SparseSymShiftSolve(sp.tril(M).tocsc())
        t0 = time.perf_counter()
        opm.set_shift(0.5)
        t_fac = time.perf_counter() - t0
        rm = solve(lambda: sb.SymEigsShiftSolver(opm, 10, 30, 0.5), sb.SortRule.LargestMagn)
        rm.update(set_shift_s=t_fac, layout=opm.layout(), status=opm.status())
        out["C5_mesh_shift_invert_27pt_30x30x30_k10_ncv30_sigma0.5"] = rm
        opm.close()
    except Exception as e:  # noqa: BLE001 -- an extra configuration must not break the line
        out["C5_mesh_shift_invert_27pt_30x30x30_k10_ncv30_sigma0.5"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def main():
    args = parse_args()
    n = int(args.n)
    if args.impl == "reference":
        run_reference(args, n)
        return

    import torch

    import spectra_b200 as sb
    from spectra_b200 import dist, synth

    rank, local_rank, world = dist.env_rank()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs {args.gpus} ranks (launch with torch.distributed.run); WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    sb.set_device(local_rank)
    dist.init_process_group("nccl" if world > 1 else None)
    comm = dist.make_comm()
    info = sb.device_info()
    peak, peak_src = measured_peak()

    # ---- synthetic input (host, pinned): this rank's row slab of G_sym ----
    row0, nrows = dist.slab_range(n, rank, world)
    alloc, _keep = pinned_alloc()
    t0 = time.time()
    rp, ci, v = synth.csr(n, args.nnz_per_row, args.seed, True, row0=row0, nrows=nrows, alloc=alloc)
    gen_s = time.time() - t0
    nnz_local = len(ci)
    nnz_total = int(dist.sum_over_ranks(nnz_local))

    def make_op():
        if world == 1:
            # the reference-facing call: Eigen-style compressed ColMajor matrix, lower triangle read and mirrored
            return sb.SparseSymMatProd((n, rp, ci, v, "col"), uplo="lower")
        return sb.SparseGenMatProd.from_csr_slab(n, row0, rp, ci, v, comm=comm)

    # =========================== value arm: operator resident in HBM ===========================
    op = make_op()
    comm_mode = "single GPU" if world == 1 else ("peer memory over NVLink (IPC windows, one-shot all-reduce, residual pushed by the correction pass)"
                                                  if op.peer_mode() else "NCCL all-gather + all-reduce")
    eigs = sb.SymEigsSolver(op, args.nev, args.ncv)
    sampler = None
    step_ms, step_wall, launches = [], [], 0
    nops = nconv = niter = 0
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            dist.barrier()
            torch.cuda.synchronize()
            sampler = ClockSampler(local_rank) if rank == 0 else None
        tw = time.perf_counter()
        eigs.init()
        nconv = eigs.compute(sb.SortRule.LargestAlge, 1000, args.tol)
        torch.cuda.synchronize()
        wall = time.perf_counter() - tw
        st = eigs.stats()
        if it >= args.warmup:
            step_ms.append(st["ms_total"])  # CUDA events on the solver stream around init()+compute()
            step_wall.append(wall * 1e3)
            launches += st["kernel_launches"]
        nops, niter = eigs.num_operations(), eigs.num_iterations()
    dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms_per_step = dist.max_over_ranks(float(np.mean(step_ms)))
    wall_per_step = dist.max_over_ranks(float(np.mean(step_wall)))
    value = nops / (ms_per_step / 1e3)
    info_ok = eigs.info() == sb.CompInfo.Successful
    last_stats = eigs.stats()

    # ---- accuracy of the returned eigenpairs (untimed): ||A x - lambda x|| / |lambda| ----
    evals = eigs.eigenvalues()
    X = eigs.eigenvectors(local=(world > 1))
    Xfull = eigs.eigenvectors() if world > 1 else X
    AX = op @ Xfull  # local rows
    num = np.array([dist.sum_over_ranks(float(np.sum((AX[:, c] - X[:, c] * evals[c]) ** 2))) for c in range(len(evals))])
    rel_res = np.sqrt(num) / np.abs(evals) if len(evals) else np.array([np.inf])
    del AX, Xfull

    # ---- per-kernel roofline (untimed extra solve with per-class CUDA events) ----
    sb.set_profiling(1)
    eigs.init()
    eigs.compute(sb.SortRule.LargestAlge, 1000, args.tol)
    ps = eigs.stats()
    sb.set_profiling(0)
    nl = nrows
    # one operator application = 12 nnz + 4 (n + 1) + 8 n (x) + 8 n (y); the fused step head adds the v_i write and the
    # v_{i-1} read (16 B/row); an application whose last kernel also carries the first panel pass (sell_step_dot_kernel) additionally
    # streams 8 nl i bytes of V (8(d) "GEMV-T": 8 n j + 8 n, the w values stay on chip)
    spmv_only_bytes = 12.0 * nnz_local + 4.0 * (nl + 1) + 8.0 * n + 8.0 * nl
    spmv_bytes = spmv_only_bytes + 16.0 * nl
    fused_bytes_total = 8.0 * nl * ps["fused_dot_cols"]
    op_bytes_total = spmv_bytes * ps["spmv_launches"] + fused_bytes_total
    panel_bytes_total = 8.0 * nl * (ps["panel_cols"] + 2 * ps["panel_launches"])
    # the plain operator kernels alone (no step head, no panel): what BASELINE's "SpMV at >= 70 % of the HBM roofline" is quoted on
    plain_ms = op.spmv_device_time(20)
    kern = {
        "panel_pass": {"launches": ps["panel_launches"], "ms_total": ps["ms_panel"], "avg_ms": ps["ms_panel"] / max(ps["panel_launches"], 1),
                       "algorithmic_bytes_per_launch": panel_bytes_total / max(ps["panel_launches"], 1),
                       "gbs": panel_bytes_total / max(ps["ms_panel"], 1e-9) / 1e6},
        "operator_step": {"launches": ps["spmv_launches"], "ms_total": ps["ms_spmv"], "avg_ms": ps["ms_spmv"] / max(ps["spmv_launches"], 1),
                          "algorithmic_bytes_per_launch": op_bytes_total / max(ps["spmv_launches"], 1), "gbs": op_bytes_total / max(ps["ms_spmv"], 1e-9) / 1e6,
                          "fused_first_panel_pass": {"launches": ps["fused_dot_launches"], "avg_cols": ps["fused_dot_cols"] / max(ps["fused_dot_launches"], 1)},
                          "layout": op.spmv_layout(),
                          "note": "all column-block kernels of one operator application inside the solver, incl. the fused step head and (sliced layout) "
                                  "the first panel pass V^T w streamed in the same kernel"},
        "spmv_plain": {"avg_ms": plain_ms, "algorithmic_bytes_per_launch": spmv_only_bytes, "gbs": spmv_only_bytes / plain_ms / 1e6,
                       "note": "y = A x alone (sb200_op_spmv_device, 20 launches back to back): SURVEY 8(d) bytes of the SpMV itself"},
        "restart_gemm": {"launches": ps["compress_launches"], "ms_total": ps["ms_compress"]},
        "small_dense": {"ms_total": ps["ms_small"]},
        "comm": {"ms_total": ps["ms_comm"]},
    }
    for kname in ("panel_pass", "operator_step", "spmv_plain"):
        kern[kname]["frac_of_hbm_peak"] = kern[kname]["gbs"] / peak
    dominant = "panel_pass" if ps["ms_panel"] >= ps["ms_spmv"] else "operator_step"
    # roofs of the gather phase measured live (csrc/microbench.cu): 1e8 uniformly random 8-byte gathers from one 40 MB operand slice, alone and
    # with the (index, value) stream a sliced SpMV reads -- two slices of that make one n = 1e7 operator application
    roofs = None
    if rank == 0 and world == 1:
        try:
            sl = min(n, 5_000_000)
            g_only = sb.bench_gather(sl, 100_000_000, 5)["ms"]
            g_stream = sb.bench_stream_gather(sl, 100_000_000, False, 5)["ms"]
            scale = nnz_total / 1e8
            roofs = {"slice_doubles": sl, "gathers": 100_000_000, "gather_only_ms": g_only, "stream_gather_ms": g_stream,
                     "spmv_floor_ms_gather_only": g_only * scale, "spmv_floor_ms_stream_gather": g_stream * scale,
                     "spmv_frac_ceiling_stream_gather": spmv_only_bytes / (g_stream * scale) / 1e6 / peak,
                     "note": "uniformly random columns: every gathered operand is its own 32 B sector; the L1TEX wavefront rate (1 per clock per SM) and the "
                             "L2 sector throughput bound the SpMV below the HBM roof"}
        except Exception as e:  # noqa: BLE001
            roofs = {"error": f"{type(e).__name__}: {e}"}
    # DRAM traffic per launch from the committed `ncu --set full` capture of this same workload and device layout (profiles/traffic.json;
    # dram__bytes_read.sum + dram__bytes_write.sum, summed over the column-block kernels of one operator application)
    traffic, traffic_src, traffic_ctx = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        ent = tj.get(dominant)
        if ent and int(ent["n"]) == n and int(ent.get("n_gpus", 1)) == world and ent.get("layout") == kern["operator_step"]["layout"]["format"]:
            traffic, traffic_src = float(ent["dram_bytes_per_launch"]), ent.get("source")
            traffic_ctx = {"captured_at_panel_width": ent.get("captured_at_panel_width"), "algorithmic_bytes_at_that_width": ent.get("algorithmic_bytes_at_that_width")}
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": kern[dominant]["gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": kern[dominant]["gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src, "traffic_context": traffic_ctx,
                "algorithmic_bytes_per_launch": kern[dominant]["algorithmic_bytes_per_launch"],
                "spmv_plain_frac": kern["spmv_plain"]["frac_of_hbm_peak"], "gather_roofs": roofs,
                "share_of_step": {"panel_pass": ps["ms_panel"] / ps["ms_total"], "operator_step": ps["ms_spmv"] / ps["ms_total"]}}

    # =========================== e2e arm: host buffers in, host results out ===========================
    e2e = None
    if not args.skip_e2e:
        del eigs
        op.close()
        e2e_wall = []
        h2d = rp.nbytes + ci.nbytes + v.nbytes + 8 * n  # CSR arrays + the default initial residual
        import torch as _t

        xbuf = _t.empty((args.nev, nrows if world > 1 else n), dtype=_t.float64, pin_memory=True).numpy().T  # F-ordered (rows x nev)
        d2h = 0
        for it in range(args.warmup + args.steps):
            dist.barrier()
            torch.cuda.synchronize()
            tw = time.perf_counter()
            op2 = make_op()
            e2 = sb.SymEigsSolver(op2, args.nev, args.ncv)
            e2.init()
            e2.compute(sb.SortRule.LargestAlge, 1000, args.tol)
            ev2 = e2.eigenvalues()
            xv = e2.eigenvectors(local=(world > 1), out=xbuf)
            torch.cuda.synchronize()
            dist.barrier()
            wall = time.perf_counter() - tw
            d2h = xv.nbytes + ev2.nbytes
            ops2 = e2.num_operations()
            if it >= args.warmup:
                e2e_wall.append(wall)
            del e2
            op2.close()
        e2e_s = dist.max_over_ranks(float(np.mean(e2e_wall)))
        e2e = {"value": ops2 / e2e_s, "unit": "SpMV-iters/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "s_per_step": e2e_s,
               "eigenpairs_per_sec": nconv / e2e_s}

    # =========================== CPU baseline: oracle port, 1 thread, bounded sample ===========================
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        import oracle as O

        A = O.Csr.adopt(n, rp, ci, v)
        r = O.sym_eigs(A, args.nev, args.ncv, O.LargestAlge, 1000, args.tol, threads=1, op_limit=args.cpu_sample_ops, want_vectors=False)
        cpu = {"value": r.nops / r.seconds, "unit": "SpMV-iters/s", "cores": 1, "kind": "port",
               "sample": f"init + first {r.nops - 2} Lanczos steps (op_limit={args.cpu_sample_ops}) of the same n={n} solve, {r.seconds:.1f} s; early steps have "
                         f"narrow panels, so this over-states the CPU's steady-state rate", "host_cores_available": host_threads(),
               "reference_own_code": reference_own_code_sample(n, rp, ci, v, args.cpu_sample_ops),
               "full_solves_cached": golden_full_solves()}

    # ---- the other BASELINE configurations on one GPU (C2, C3, C5): one solve each after a warm-up solve, device time; parity of
    # these configurations is the job of tests/ (test_sym_eigs_full_size_properties, test_gen_eigs_c3_unplanted_history, test_gpu_shift) ----
    configs = None
    if rank == 0 and world == 1 and not args.skip_configs:
        configs = {}
        try:
            configs.update(run_other_configs(sb, synth))
        except Exception as e:  # noqa: BLE001
            configs["error"] = f"{type(e).__name__}: {e}"

    if rank == 0:
        line = {
            "metric": "spmv_iters_per_sec", "value": value, "unit": "SpMV-iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n),
            "eigenpairs_per_sec": nconv / (ms_per_step / 1e3), "nconv": int(nconv), "num_operations": int(nops), "num_iterations": int(niter),
            "converged": bool(info_ok), "accuracy": {"max_rel_residual": float(np.max(rel_res)), "bound": 1e-10},
            "wall_ms_per_step": wall_per_step, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kern, "cpu_baseline": cpu,
            "clocks": clocks, "algo_counters": {k: last_stats[k] for k in ("lanczos_steps", "reorth_passes", "restarts", "expand_calls", "host_syncs", "fused_dot_launches")},
            "device": info, "nnz": nnz_total, "gen_seconds": gen_s, "configs": configs, "comm_mode": comm_mode,
        }
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

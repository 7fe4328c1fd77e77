#!/usr/bin/env python
"""bench.py -- the headline measurement: matrix+vector assembly of 3-D Poisson Q2 on a hex mesh.

A "step" is one pass of the hot path over one batch of synthetic input: one numeric (re)assembly
(cell integration + Dirichlet lifting + deterministic scatter + ghost-row reduction) of the whole
mesh.  N=1: BASELINE.json configs[1] (128^3 cells, one part).  N=2/4/8: parts (2,1,1)/(2,2,1)/(2,2,2),
128^3 OWNED cells per part (weak scaling; N=8 is configs[2], 256^3 cells), one part per GPU, NCCL
ghost-row reduction.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C] [--geometry cartesian|hex]

Prints ONE JSON line (rank 0).  `value` = assembled nnz/s with inputs resident in HBM; `e2e` = the same
through the C ABI with HOST buffers (pinned): per step the Dirichlet values and the source term go
host->device and the assembled CSR values + right-hand side come back device->host.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/) on the host cores
(the Julia reference cannot run in this image: no julia, no MPI).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARTS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
METRIC = "assembled nnz/s (3D Poisson Q2 hex, FP64 matrix+vector assembly)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region by a thread calling NVML directly (the timed region is
    tens of milliseconds: an `nvidia-smi -lms` child would not deliver a single sample in it)."""

    def __init__(self, index, period_s=0.002):
        import threading

        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.th = None
        try:
            import pynvml as N

            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            h = N.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for nm, bit in bits.items():
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception:
                        pass
                    time.sleep(period_s)

            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
        except Exception:
            self.th = None

    def stop(self):
        self._stop.set()
        if self.th is not None:
            self.th.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def algorithmic_bytes(nnz, nrows, nd, ncells, D, nnodes_geom):
    """values written once, rhs written once, cell_dof_ids read once, coordinates read once."""
    return 8 * nnz + 8 * nrows + 4 * nd * ncells + 8 * D * nnodes_geom


def host_parts(nthreads):
    """Factor the host thread count into a 3-D part grid (x fastest gets the largest factor)."""
    p, dims, d = nthreads, [1, 1, 1], 0
    f = 2
    while p > 1:
        while p % f:
            f += 1
        dims[d % 3] *= f
        p //= f
        d += 1
    return tuple(sorted(dims, reverse=True))


def cpu_reference_run(cells, steps, warmup, threads=None):
    """The reference algorithm on the host cores: oracle/assembly_ref.c (per-cell quadrature, one COO triplet per
    (cell,i,j), COO->CSR sort + duplicate sum), one OS thread per mesh part like one MPI rank per part in the
    reference's with_mpi mode; `cells`^3 owned cells per part, P = host threads parts, no ghost exchange timed."""
    from helpers import build_problem
    from oracle import c_oracle

    threads = threads or os.cpu_count() or 1
    parts = host_parts(threads)
    pr = build_problem(parts, tuple(p * cells for p in parts), 2, "boundary", lambda x: x[0] + x[1] + x[2], "sub")
    inputs = [c_oracle.part_inputs(pr, k) for k in range(len(pr.model.models))]
    times, nnz = [], 0
    for it in range(warmup + steps):
        t = time.perf_counter()
        res = c_oracle.assemble_poisson_q2(pr, source=1.0, threads=threads, keep=False, inputs=inputs)
        dt = time.perf_counter() - t
        nnz = sum(r[1] for r in res)
        if it >= warmup:
            times.append(dt)
    ncells = sum(len(i[2]) for i in inputs)
    return dict(ncells=ncells, nnz=nnz, seconds=float(np.mean(times)), threads=threads, parts=parts, cells=cells)


def cpu_sample_text(r):
    return (f"{r['parts']} parts x {r['cells']}^3 owned cells (same form/space/quadrature as the 128^3 workload), C restatement of the "
            f"reference algorithm (oracle/assembly_ref.c: per-cell quadrature, COO materialisation, COO->CSR), one thread per part, "
            f"{r['seconds']:.2f} s/step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.ref_cells, args.steps, min(args.warmup, 1))
    v = r["nnz"] / r["seconds"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Poisson Q2 hex, {r['parts']} parts x {r['cells']}^3 cells on {r['threads']} host threads "
                                   f"(bounded sample of the 128^3-cells-per-GPU workload; the Julia reference cannot run in this image)",
                       "cells_per_s": r["ncells"] / r["seconds"]},
            "cpu_baseline": {"value": v, "unit": "nnz/s", "cores": r["threads"], "kind": "port", "sample": cpu_sample_text(r)},
            "e2e": {"value": v, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--cells", type=int, default=128, help="owned cells per direction per part")
    ap.add_argument("--ref-cells", type=int, default=32, help="CPU baseline: owned cells per direction per part (one part per host thread)")
    ap.add_argument("--no-general", action="store_true", help="skip the extra measurement of the general (hex-node) route")
    ap.add_argument("--geometry", default="cartesian", choices=["cartesian", "hex"])
    ap.add_argument("--perturb", action="store_true", help="hex geometry with nodes moved by <= 0.1 h (general trilinear cells)")
    ap.add_argument("--strategy", default="sub", choices=["sub", "fully"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--spmv-steps", type=int, default=20)
    ap.add_argument("--cg-iters", type=int, default=20)
    ap.add_argument("--no-cg", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import graft_import

    g = graft_import.load()
    L = g.libgraft
    from helpers import build_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        backend = g.DistBackend()
    else:
        backend = g.DebugBackend(1)
    parts = PARTS[args.gpus]
    cells = tuple(p * args.cells for p in parts)
    t0 = time.time()
    u = lambda x: x[0] + x[1] + x[2]
    pr = build_problem(parts, cells, 2, "boundary", u, args.strategy, backend=backend)
    t_setup = time.time() - t0
    strategy = g.FullyAssembledRows() if args.strategy == "fully" else g.SubAssembledRows()
    kw = {}
    if args.perturb:
        args.geometry = "hex"
        prng = np.random.default_rng(0)
        kw["perturb"] = lambda m, xyz: xyz + prng.uniform(-0.1, 0.1, xyz.shape) * np.asarray(m.h)[None, :] * _interior_mask(m, m.vertex_multi_index())[:, None]
    assem = g.SparseMatrixAssembler(pr.U, pr.V, strategy, geometry=args.geometry, device=local_rank, **kw)
    dΩ = g.Measure(pr.trian, 4)
    form = g.Poisson(dΩ, source=1.0)
    lib, comm, ctx = assem.comm.lib, assem.comm.handle, assem.comm.ctxs[0]
    assem._set_form(form)
    t0 = time.time()
    assem._symbolic(form)
    t_symbolic_wall = time.time() - t0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, what=3):
        for _ in range(n):
            L.check(lib.graft_numeric(comm, what))
        L.check(lib.graft_sync(comm))

    # ---- device-resident numeric assembly ---------------------------------------------------------
    run_steps(args.warmup)
    st0 = assem.stats()[0]
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # EXACTLY K steps enqueued back to back on the library's compute stream, bracketed by timing marks recorded on that
    # stream (graft_mark / graft_elapsed = CUDA events); barrier + synchronize on both sides; max over ranks below
    import ctypes as C

    t0 = time.perf_counter()
    L.check(lib.graft_mark(comm, 0))
    for _ in range(args.steps):
        L.check(lib.graft_numeric(comm, 3))
    L.check(lib.graft_mark(comm, 1))
    el = C.c_double()
    L.check(lib.graft_elapsed(ctx, 0, 1, C.byref(el)))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    st1 = assem.stats()[0]
    dev_ms = float(el.value)
    tim = assem.timers()[0]   # phases of the last step
    # max over ranks of the device time
    if dist is not None:
        t = torch.tensor([dev_ms, wall * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
        cnt = torch.tensor([st1["nnz"], st1["ncells"], 0], device="cuda", dtype=torch.int64)
        # count only OWNED rows' nnz / owned cells for the aggregate
        m = A_own_nnz(assem, L, lib, ctx)
        cnt[0] = m
        cnt[1] = len(pr.model.cell_gids.indices[0].own_to_local)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tot_nnz, tot_cells = int(cnt[0]), int(cnt[1])
    else:
        wall_ms = wall * 1e3
        tot_nnz, tot_cells = st1["nnz"], st1["ncells"]
    ms_per_step = dev_ms / args.steps
    value = tot_nnz / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (per part)
    sp = pr.U.spaces[0]
    m_loc = pr.model.models[0]
    nnodes_geom = int(np.prod(m_loc.ncells_local + 1))
    B_num = algorithmic_bytes(st1["nnz"], st1["nrows"], sp.nd, st1["ncells"], 3, nnodes_geom)
    peak, peak_src = measured_peak()
    kern_ms = ms_per_step  # the whole numeric step (all its kernels)
    dominant = {("fused-affine", "cartesian"): "stream_groups_kernel", ("fused-affine", "hex"): "gemm_rows_kernel"}.get(
        (st1["path"], args.geometry), "integrate_small_kernel+gather_direct_kernel")
    achieved = B_num / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.cells == 128 and args.gpus == 1:
            traffic = sum(tj[k] for k in dominant.split("+")) if all(k in tj for k in dominant.split("+")) else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": dominant, "algorithmic_bytes_per_launch": B_num, "peak_source": peak_src,
                "note": "achieved = algorithmic bytes / device time of the whole numeric step (all kernels of the step, CUDA events on the library stream)",
                "phase_ms": {"integrate": float(tim[L.T_INTEGRATE]), "scatter": float(tim[L.T_SCATTER]), "exchange": float(tim[L.T_EXCHANGE])}}

    # ---- SpMV (mul!) device-resident ---------------------------------------------------------------
    spmv = None
    try:
        n_cols = st1["ncols"]; n_own = assem.rows[0].indices[0].own_length
        x = torch.ones(n_cols, dtype=torch.float64, device="cuda")
        y = torch.zeros(max(n_own, 1), dtype=torch.float64, device="cuda")
        xs, ys = L.ptr_array([x.data_ptr()]), L.ptr_array([y.data_ptr()])
        for _ in range(3):
            L.check(lib.graft_spmv_device(comm, 0, 0, 1.0, xs, 0.0, ys))
        L.check(lib.graft_sync(comm)); barrier()
        L.check(lib.graft_mark(comm, 2))
        for _ in range(args.spmv_steps):
            L.check(lib.graft_spmv_device(comm, 0, 0, 1.0, xs, 0.0, ys))
        L.check(lib.graft_mark(comm, 3))
        el2 = C.c_double()
        L.check(lib.graft_elapsed(ctx, 2, 3, C.byref(el2)))
        barrier()
        sp_ms = float(el2.value) / args.spmv_steps
        if dist is not None:
            t = torch.tensor([sp_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); sp_ms = float(t[0])
        B_spmv = 12 * st1["nnz"] + 4 * (st1["nrows"] + 1) + 8 * st1["nrows"] + 8 * st1["ncols"]
        spmv = {"ms": sp_ms, "GBps": B_spmv / (sp_ms * 1e-3) / 1e9, "frac_of_peak": B_spmv / (sp_ms * 1e-3) / 1e9 / peak,
                "algorithmic_bytes": B_spmv}
    except Exception as e:  # pragma: no cover
        spmv = {"error": str(e)}

    # ---- Jacobi-CG on top of mul!: a fixed number of iterations on the assembled system ----------------
    cg = None
    try:
        if not args.no_cg:
            n_own = assem.rows[0].indices[0].own_length
            bvec = np.empty(st1["nrows"], dtype=np.float64)
            L.check(lib.graft_vec_get(ctx, 0, L.ptr(bvec)))
            bs, xs0 = [np.ascontiguousarray(bvec[:n_own])], [np.zeros(n_own)]
            it, rr = L.c_i32(), L.c_dbl()
            L.check(lib.graft_cg(comm, 0, L.ptr_array(bs), L.ptr_array(xs0), 0.0, args.cg_iters, 1, C.byref(it), C.byref(rr)))
            cg_ms = float(assem.timers()[0][L.T_CG])
            if dist is not None:
                t = torch.tensor([cg_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); cg_ms = float(t[0])
            cg = {"iterations": int(it.value), "ms_per_iteration": cg_ms / max(int(it.value), 1), "relres": float(rr.value),
                  "note": "Jacobi-preconditioned CG, device resident: 1 mul! + 2 global reductions + 3 vector updates per iteration"}
    except Exception as e:  # pragma: no cover
        cg = {"error": str(e)[:200]}

    # ---- the other geometry routes of the same workload (reported beside the headline, N=1 only) ------------------
    routes = None
    if args.gpus == 1 and not args.no_general and args.geometry == "cartesian":
        routes = {}

        def time_route(name, **kw):
            try:
                a2 = g.SparseMatrixAssembler(pr.U, pr.V, strategy, geometry="hex", device=local_rank, **kw)
                a2._set_form(form); a2._symbolic(form)
                c2 = a2.comm.handle
                for _ in range(3):
                    L.check(lib.graft_numeric(c2, 3))
                L.check(lib.graft_sync(c2))
                ts = []
                for _ in range(5):
                    L.check(lib.graft_numeric(c2, 3)); L.check(lib.graft_sync(c2))
                    ts.append(a2.timers()[0][L.T_NUMERIC])
                ms = float(np.mean(ts))
                tm = a2.timers()[0]
                routes[name] = {"ms_per_step": ms, "phase_ms": {"integrate": float(tm[L.T_INTEGRATE]), "scatter": float(tm[L.T_SCATTER])}, "nnz_per_s": st1["nnz"] / (ms * 1e-3), "frac_of_hbm_peak": B_num / (ms * 1e-3) / 1e9 / peak,
                                "route": a2.stats()[0]["path"]}
                a2.close()
            except Exception as e:  # pragma: no cover
                routes[name] = {"error": str(e)[:200]}

        # the same mesh given as hex node coordinates: per-cell Jacobians / coefficients (affine cells, tensor-core route)
        time_route("hex_nodes_affine")
        # nodes moved by <= 0.1 h (seed 0): general trilinear cells, cell matrices integrated per cell (FP64) and gathered
        rng = np.random.default_rng(0)

        def perturb(m, xyz):
            d = rng.uniform(-0.1, 0.1, xyz.shape) * np.asarray(m.h)[None, :]
            return xyz + d * _interior_mask(m, m.vertex_multi_index())[:, None]

        time_route("hex_nodes_perturbed", perturb=perturb)

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        nnz, nrows = st1["nnz"], st1["nrows"]
        dv = torch.from_numpy(np.ascontiguousarray(pr.U.dirichlet_values[0])).pin_memory()
        src = torch.ones(1, dtype=torch.float64).pin_memory()
        vals = torch.empty(nnz, dtype=torch.float64).pin_memory()
        bh = torch.empty(nrows, dtype=torch.float64).pin_memory()
        vp, bp, dvp, srp = vals.numpy(), bh.numpy(), dv.numpy(), src.numpy()

        def e2e_step():
            if len(dvp):
                L.check(lib.graft_space_set_dirichlet_values(ctx, 0, len(dvp), L.ptr(dvp)))
            L.check(lib.graft_source_set(ctx, 0, L.SOURCE_CONST, L.ptr(srp), None))
            L.check(lib.graft_numeric(comm, 3))
            L.check(lib.graft_csr_get_values(ctx, 0, 0, L.ptr(vp)))
            L.check(lib.graft_vec_get(ctx, 0, L.ptr(bp)))

        e2e_steps = max(2, min(args.steps, 5))
        e2e_step(); barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t[0])
        e2e = {"value": tot_nnz * e2e_steps / dt, "unit": "nnz/s", "h2d_bytes_per_step": int(8 * len(dvp) + 8),
               "d2h_bytes_per_step": int(8 * nnz + 8 * nrows), "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "checksum": float(vp[: min(nnz, 1 << 20)].sum())}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.ref_cells, 2, 1)
        cpu = {"value": r["nnz"] / r["seconds"], "unit": "nnz/s", "cores": r["threads"], "kind": "port", "sample": cpu_sample_text(r),
               "cells_per_s": r["ncells"] / r["seconds"]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"3D Poisson Q2 hex, {'x'.join(str(c) for c in cells)} cells, parts {parts}, {args.strategy}-assembled rows, "
                                       f"geometry={args.geometry}, re-assembly (numeric phase) of matrix+vector",
                           "cells": tot_cells, "nnz": tot_nnz, "cells_per_s": tot_cells / (ms_per_step * 1e-3), "route": st1["path"],
                           "l2": "inputs+outputs (>9 GB per step) exceed the 126 MB L2; no explicit flush",
                           "symbolic_ms_device": float(tim[L.T_SYMBOLIC]), "symbolic_s_wall": t_symbolic_wall, "host_setup_s": t_setup,
                           "wall_ms_per_step": wall_ms / args.steps},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "spmv": spmv, "cg": cg, "other_routes": routes,
                "gpu_launches": int(st1["launches"] - st0["launches"]), "clocks": clocks}
        print(json.dumps(line))
    assem.close()
    if dist is not None:
        dist.destroy_process_group()


def _interior_mask(m, ijk):
    """1 for vertices strictly inside the GLOBAL domain (boundary vertices keep their place so that the domain is unchanged)."""
    gi = ijk + np.asarray(m.cmin)[None, :]
    return np.all((gi > 0) & (gi < np.asarray(m.ncells_global)[None, :]), axis=1).astype(np.float64)


def A_own_nnz(assem, L, lib, ctx):
    """nnz of the owned rows of this part (ghost rows are structural zeros owned elsewhere)."""
    import ctypes as C

    m, n, nnz = L.c_i64(), L.c_i64(), L.c_i64()
    L.check(lib.graft_csr_query(ctx, 0, 0, C.byref(m), C.byref(n), C.byref(nnz)))
    rowptr = np.empty(m.value + 1, dtype=np.int64)
    L.check(lib.graft_csr_get(ctx, 0, 0, L.ptr(rowptr), None, None))
    return int(rowptr[assem.rows[0].indices[0].own_length] - assem.index_base)


if __name__ == "__main__":
    main()

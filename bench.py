#!/usr/bin/env python
"""bench.py -- the headline measurement: matrix+vector assembly of 3-D Poisson Q2 on a hex mesh (BASELINE.json).

A "step" is one pass of the hot path over one batch of synthetic input: one numeric (re)assembly (cell integration +
Dirichlet lifting + deterministic scatter + ghost-row reduction) of the whole mesh.  N=1: BASELINE.json configs[1] (128^3
cells, one part).  N=2/4/8: parts (2,1,1)/(2,2,1)/(2,2,2), 128^3 OWNED cells per part (weak scaling; N=8 is configs[2],
256^3 cells), one part per GPU, NCCL ghost-row reduction.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--form poisson|elasticity|stokes]
                  [--cells C] [--geometry cartesian|hex|perturbed]

Prints ONE JSON line (rank 0).
  value          assembled nnz/s of the headline geometry, inputs resident in HBM.  For --geometry cartesian (default) this is
                 the BEST-CASE SPECIAL CASE: the mesh is a Cartesian descriptor, every cell matrix is the same and the numeric
                 step streams row templates (what Gridap itself exploits for a CartesianDiscreteModel).
  general_route  the SAME workload with the mesh given as node coordinates moved by <= 0.1 h (general trilinear hexes): per-cell
                 sum-factorised integration fused with the row reduction (sweep.cu) -- the north-star's kernels (1)+(3); its own
                 roofline entry.  (N=1, --form poisson.)
  e2e            the same metric through the C ABI with HOST buffers (pinned): per step the Dirichlet values and the source
                 term go host->device and the assembled CSR values + right-hand side come back device->host.
  invariants     size-independent checks of the assembled full-size system, outside the timed region.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/) on the host cores (the Julia reference cannot
run in this image: no julia, no MPI).  --form elasticity / stokes: BASELINE.json configs[3] / [4] at their per-GPU sizes.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARTS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
FORMS = {
    "poisson": dict(cells=128, metric="assembled nnz/s (3D Poisson Q2 hex, FP64 matrix+vector assembly)",
                    what="3D Poisson Q2 hex"),
    "elasticity": dict(cells=96, metric="assembled nnz/s (3D linear elasticity vector Q2 hex, 81x81 cell matrices, FP64 matrix+vector assembly)",
                       what="3D linear elasticity Q2^3 hex (lambda = mu = 1)"),
    "stokes": dict(cells=64, metric="assembled nnz/s (3D Stokes Taylor-Hood Q2/Q1 hex, block assembly, FP64 matrix+vector assembly)",
                   what="3D Stokes Taylor-Hood Q2^3/Q1 hex, 2x2 blocks"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


FP64_PEAK_TFLOPS = 37.0  # measured on this pool's B200 (profiles/r2_fp64_peak.jsonl: DFMA 36.9, DMMA 37.0 TFLOP/s)


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
    import graft_import

    g = graft_import.load()
    from oracle import c_oracle

    threads = threads or os.cpu_count() or 1
    parts = host_parts(threads)
    pr = g.build_problem(parts, tuple(p * cells for p in parts), 2, "boundary", lambda x: x[0] + x[1] + x[2], "sub")
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
    line = {"impl": "reference", "metric": FORMS["poisson"]["metric"], "value": v, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Poisson Q2 hex, {r['parts']} parts x {r['cells']}^3 cells on {r['threads']} host threads "
                                   f"(bounded sample of the 128^3-cells-per-GPU workload; the Julia reference cannot run in this image)",
                       "cells_per_s": r["ncells"] / r["seconds"], "same_config": False},
            "cpu_baseline": {"value": v, "unit": "nnz/s", "cores": r["threads"], "kind": "port", "sample": cpu_sample_text(r)},
            "e2e": {"value": v, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------------
class Workload:
    """One form + geometry on one assembler: problem inputs, blocks, algorithmic bytes."""

    def __init__(self, g, form, parts, cells, strategy, backend):
        self.g, self.form_name = g, form
        u3 = lambda x: x[0] + x[1] + x[2]
        if form == "poisson":
            self.pr = g.build_problem(parts, cells, 2, "boundary", u3, strategy, backend=backend)
            self.trials, self.tests, self.blocks = self.pr.U, self.pr.V, [(0, 0)]
            self.fields = [self.pr.U]
        elif form == "elasticity":
            self.pr = g.build_problem(parts, cells, 2, "boundary", lambda x: np.stack([x[0], x[1], x[2]]), strategy, ncomp=3, backend=backend)
            self.trials, self.tests, self.blocks = self.pr.U, self.pr.V, [(0, 0)]
            self.fields = [self.pr.U]
        else:
            self.pr = g.build_stokes_problem(parts, cells, strategy, backend=backend)
            self.trials, self.tests, self.blocks = [self.pr.U, self.pr.P], [self.pr.V, self.pr.Q], [(0, 0), (0, 1), (1, 0)]
            self.fields = [self.pr.U, self.pr.P]
        self.strategy = g.FullyAssembledRows() if strategy == "fully" else g.SubAssembledRows()

    def make_form(self):
        g = self.g
        dO = g.Measure(self.pr.trian, 4)
        if self.form_name == "poisson":
            return g.Poisson(dO, source=1.0)
        if self.form_name == "elasticity":
            return g.LinearElasticity(dO, 1.0, 1.0, source=1.0)
        return g.StokesTH(dO, nu=1.0, source=1.0)

    def assembler(self, geometry, device):
        g = self.g
        kw = {}
        if geometry == "perturbed":
            kw["perturb"] = g.vertex_perturbation(0.1, seed=0)
        a = g.SparseMatrixAssembler(self.trials, self.tests, self.strategy, geometry="cartesian" if geometry == "cartesian" else "hex",
                                    device=device, **kw)
        form = self.make_form()
        a._set_form(form)
        return a, form


def block_sizes(assem, L, blocks):
    import ctypes as C

    lib, ctx = assem.comm.lib, assem.comm.ctxs[0]
    out = {}
    for (bi, bj) in blocks:
        m, n, nnz = L.c_i64(), L.c_i64(), L.c_i64()
        L.check(lib.graft_csr_query(ctx, bi, bj, C.byref(m), C.byref(n), C.byref(nnz)))
        out[(bi, bj)] = (m.value, n.value, nnz.value)
    return out


def algorithmic_bytes(wl, assem, L, sizes):
    """values written once, rhs written once, cell_dof_ids read once, coordinates read once (per part)."""
    st = assem.stats()[0]
    m_loc = wl.pr.model.models[0]
    nnodes_geom = int(np.prod(np.asarray(m_loc.ncells_local) + 1))
    nnz = sum(s[2] for s in sizes.values())
    nrows = sum(sizes[(f, f)][0] if (f, f) in sizes else sizes[(f, 0)][0] for f in range(len(wl.fields)))
    nd = sum(U.spaces[0].nd for U in wl.fields)
    return 8 * nnz + 8 * nrows + 4 * nd * st["ncells"] + 8 * 3 * nnodes_geom, nnz, nrows


def own_nnz(assem, L, bi, bj):
    """nnz of the owned rows of this part (ghost rows are structural zeros owned elsewhere)."""
    import ctypes as C

    lib, ctx = assem.comm.lib, assem.comm.ctxs[0]
    m, n, nnz = L.c_i64(), L.c_i64(), L.c_i64()
    L.check(lib.graft_csr_query(ctx, bi, bj, C.byref(m), C.byref(n), C.byref(nnz)))
    rowptr = np.empty(m.value + 1, dtype=np.int64)
    L.check(lib.graft_csr_get(ctx, bi, bj, L.ptr(rowptr), None, None))
    return int(rowptr[assem.rows[bi].indices[0].own_length] - assem.index_base)


def time_steps(assem, L, steps, warmup, barrier, what=3, before_step=None):
    """EXACTLY `steps` numeric steps enqueued back to back on the library's compute stream, bracketed by timing marks recorded on
    that stream (graft_mark / graft_elapsed = CUDA events); barrier + synchronize on both sides.  Returns device ms (this rank).
    before_step: host call made before every step (e.g. an input setter that invalidates what the library caches between steps)."""
    import ctypes as C

    lib, comm, ctx = assem.comm.lib, assem.comm.handle, assem.comm.ctxs[0]
    for _ in range(warmup):
        if before_step:
            before_step()
        L.check(lib.graft_numeric(comm, what))
    L.check(lib.graft_sync(comm))
    barrier()
    L.check(lib.graft_mark(comm, 0))
    for _ in range(steps):
        if before_step:
            before_step()
        L.check(lib.graft_numeric(comm, what))
    L.check(lib.graft_mark(comm, 1))
    el = C.c_double()
    L.check(lib.graft_elapsed(ctx, 0, 1, C.byref(el)))
    barrier()
    return float(el.value)


def check_invariants(wl, assem, L, torch, dist, cells_global, geometry):
    """Size-independent properties of the assembled FULL-SIZE system (outside the timed region; device resident through
    graft_spmv_device).  Poisson / elasticity, block (0,0):
      nnz        global nnz of the owned rows == ncomp^2 (8N-9)^3  (Q2, Dirichlet on the whole boundary)
      symmetry   x.(A y) == y.(A x) for two seeded random vectors
      row_sums   number of owned rows with sum_j A_ij == 0 (to 1e-9 of the largest row sum) == (2N-5)^3: the rows whose stencil avoids the boundary (Poisson)
      rhs        sum_i (b - A u_h)_i == (1 - 1/(3N))^3 for u = x+y+z, f = 1 (Poisson, unperturbed coordinates)"""
    import ctypes as C

    lib, comm, ctx = assem.comm.lib, assem.comm.handle, assem.comm.ctxs[0]
    out = {}
    rows, cols = assem.rows[0].indices[0], assem.cols[0].indices[0]
    n_own, n_cols = rows.own_length, cols.local_length
    ncomp = wl.fields[0].spaces[0].ncomp

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def spmv(x):
        y = torch.zeros(max(n_own, 1), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()   # x / y are produced on torch's stream, the product runs on the library's stream
        L.check(lib.graft_spmv_device(comm, 0, 0, 1.0, L.ptr_array([x.data_ptr()]), 0.0, L.ptr_array([y.data_ptr()])))
        L.check(lib.graft_sync(comm))
        return y[:n_own]

    N = cells_global
    nnz_own = allsum(float(own_nnz(assem, L, 0, 0)))
    expect = float(ncomp * ncomp) * float(np.prod([8 * n - 9 for n in N]))
    out["nnz"] = {"got": nnz_own, "expected": expect, "ok": nnz_own == expect}
    gen = torch.Generator(device="cuda"); gen.manual_seed(1234 + int(os.environ.get("RANK", "0")))
    x = torch.zeros(n_cols, dtype=torch.float64, device="cuda"); y = torch.zeros(n_cols, dtype=torch.float64, device="cuda")
    x[:n_own] = torch.rand(n_own, generator=gen, device="cuda", dtype=torch.float64) - 0.5
    y[:n_own] = torch.rand(n_own, generator=gen, device="cuda", dtype=torch.float64) - 0.5
    # NB the column PRange lists own dofs first in the same order as the rows (own_to_global of rows == of cols)
    Ay, Ax = spmv(y), spmv(x)
    xAy, yAx = allsum(float(torch.dot(x[:n_own], Ay))), allsum(float(torch.dot(y[:n_own], Ax)))
    scale = allsum(float(torch.dot(Ax, Ax))) ** 0.5 * allsum(float(torch.dot(y[:n_own], y[:n_own]))) ** 0.5
    out["symmetry"] = {"rel_diff": abs(xAy - yAx) / max(scale, 1e-300), "ok": abs(xAy - yAx) <= 1e-10 * scale}
    if wl.form_name == "poisson":
        ones = torch.ones(n_cols, dtype=torch.float64, device="cuda")
        rs = spmv(ones)
        # the rows of this operator have |sum| either ~1e-16 |A_ii| (stencil inside the domain) or O(|A_ii|): a threshold of
        # 1e-9 max|sum| separates them
        thr = 1e-9 * float(rs.abs().max()) if n_own else 0.0
        if dist is not None:
            t = torch.tensor([thr], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); thr = float(t[0])
        cnt = allsum(float((rs.abs() <= thr).sum()))
        expect = float(np.prod([2 * n - 5 for n in N]))
        out["row_sums"] = {"zero_rows": cnt, "expected": expect, "ok": cnt == expect}
        if geometry != "perturbed":
            sp = wl.fields[0].spaces[0]
            fe = wl.fields[0].gids.indices[0]
            uh_free = sp.free_dof_coords.sum(axis=1)                       # u = x+y+z at the free dofs (FE-space local order)
            order = np.argsort(fe.l2g, kind="stable")
            lid = order[np.searchsorted(fe.l2g[order], cols.l2g)]        # column -> FE-space local dof
            xu = torch.from_numpy(np.ascontiguousarray(uh_free[lid])).cuda()
            Au = spmv(xu)
            bvec = np.empty(assem.brows[0].indices[0].local_length, dtype=np.float64)
            L.check(lib.graft_vec_get(ctx, 0, L.ptr(bvec)))
            tot = allsum(float(torch.from_numpy(bvec[:n_own]).cuda().sum() - Au.sum()))
            expect = float(np.prod([1.0 - 1.0 / (3.0 * n) for n in N]))
            out["rhs"] = {"sum_b_minus_Au": tot, "expected": expect, "ok": abs(tot - expect) <= 1e-9}
    out["ok"] = all(v["ok"] for v in out.values() if isinstance(v, dict))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--form", default="poisson", choices=list(FORMS))
    ap.add_argument("--cells", type=int, default=0, help="owned cells per direction per part (default: 128 / 96 / 64 by form)")
    ap.add_argument("--ref-cells", type=int, default=32, help="CPU baseline: owned cells per direction per part (one part per host thread)")
    ap.add_argument("--no-general", action="store_true", help="skip the measurement of the general (hex-node) routes")
    ap.add_argument("--geometry", default="cartesian", choices=["cartesian", "hex", "perturbed"])
    ap.add_argument("--perturb", action="store_true", help="same as --geometry perturbed")
    ap.add_argument("--strategy", default="sub", choices=["sub", "fully"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-invariants", action="store_true")
    ap.add_argument("--spmv-steps", type=int, default=20)
    ap.add_argument("--cg-iters", type=int, default=20)
    ap.add_argument("--no-cg", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.perturb:
        args.geometry = "perturbed"
    if not args.cells:
        args.cells = FORMS[args.form]["cells"]

    import ctypes as C

    import torch
    import graft_import

    g = graft_import.load()
    L = g.libgraft

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
    wl = Workload(g, args.form, parts, cells, args.strategy, backend)
    t_setup = time.time() - t0
    assem, form = wl.assembler(args.geometry, local_rank)
    lib, comm, ctx = assem.comm.lib, assem.comm.handle, assem.comm.ctxs[0]
    t0 = time.time()
    assem._symbolic(form)
    t_symbolic_wall = time.time() - t0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        if dist is None:
            return list(vals), [0] * len(vals)
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        m = torch.stack(allt)
        return [float(v) for v in m.max(dim=0).values], [int(i) for i in m.argmax(dim=0)]

    # ---- device-resident numeric assembly ---------------------------------------------------------
    for _ in range(args.warmup):
        L.check(lib.graft_numeric(comm, 3))
    L.check(lib.graft_sync(comm))
    st0 = assem.stats()[0]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.perf_counter()
    dev_ms = time_steps(assem, L, args.steps, 0, barrier)
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    st1 = assem.stats()[0]
    tim = assem.timers()[0]   # phases of the last step on this rank
    sizes = block_sizes(assem, L, wl.blocks)
    B_num, nnz_loc, nrows_loc = algorithmic_bytes(wl, assem, L, sizes)
    (dev_ms_max, wall_ms, t_int, t_sc, t_ex, t_sym), (arg_dev, _, a_int, a_sc, a_ex, a_sym) = allmax(
        [dev_ms, wall * 1e3, tim[L.T_INTEGRATE], tim[L.T_SCATTER], tim[L.T_EXCHANGE], tim[L.T_SYMBOLIC]])
    if dist is not None:
        cnt = torch.tensor([sum(own_nnz(assem, L, bi, bj) for (bi, bj) in wl.blocks), len(wl.pr.model.cell_gids.indices[0].own_to_local)],
                           device="cuda", dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tot_nnz, tot_cells = int(cnt[0]), int(cnt[1])
    else:
        tot_nnz, tot_cells = nnz_loc, st1["ncells"]
    ms_per_step = dev_ms_max / args.steps
    value = tot_nnz / (ms_per_step * 1e-3)

    peak, peak_src = measured_peak()
    route = st1["path"]
    dominant = {("fused-affine", "cartesian"): "stream_groups_kernel", ("fused-affine", "hex"): "gemm_rows_kernel",
                ("fused-sweep", "hex"): "sweep_q2_kernel", ("fused-sweep", "perturbed"): "sweep_q2_kernel",
                ("sumfact-gather", "perturbed"): "integrate_sumfact_q2_kernel+gather_direct_kernel"}.get(
        (route, args.geometry), "integrate_small_kernel+gather_direct_kernel" if args.form == "poisson" else "integrate_cells_mma_kernel+gather_direct_kernel")
    achieved = B_num / (ms_per_step * 1e-3) / 1e9

    def traffic_of(kernels):
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of the kernel from the committed ncu --set full capture (128^3, 1 GPU)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if args.cells == 128 and args.gpus == 1 and args.form == "poisson":
                return sum(tj[k] for k in kernels.split("+")) if all(k in tj for k in kernels.split("+")) else None
        except Exception:
            pass
        return None

    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_of(dominant),
                "kernel": dominant, "algorithmic_bytes_per_launch": B_num, "peak_source": peak_src,
                "note": "achieved = algorithmic bytes / device time of the whole numeric step (all kernels of the step, CUDA events on the library stream)",
                "phase_ms_max_over_ranks": {"integrate": t_int, "scatter": t_sc, "exchange": t_ex},
                "phase_argmax_rank": {"integrate": a_int, "scatter": a_sc, "exchange": a_ex, "step": arg_dev}}

    # ---- invariants of the assembled full-size system (outside the timed region) ---------------------
    invariants = None
    if not args.no_invariants and args.form in ("poisson", "elasticity"):
        try:
            invariants = check_invariants(wl, assem, L, torch, dist, cells, args.geometry)
        except Exception as e:  # pragma: no cover
            invariants = {"ok": False, "error": str(e)[:300]}

    # ---- SpMV (mul!) device-resident, block (0,0) --------------------------------------------------
    spmv = None
    try:
        m00, n00, nnz00 = sizes[(0, 0)]
        n_own = assem.rows[0].indices[0].own_length
        x = torch.ones(n00, dtype=torch.float64, device="cuda")
        y = torch.zeros(max(n_own, 1), dtype=torch.float64, device="cuda")
        xs, ys = L.ptr_array([x.data_ptr()]), L.ptr_array([y.data_ptr()])
        torch.cuda.synchronize()
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
        (sp_ms,), _ = allmax([float(el2.value) / args.spmv_steps])
        B_spmv = 12 * nnz00 + 4 * (m00 + 1) + 8 * m00 + 8 * n00
        spmv = {"ms": sp_ms, "GBps": B_spmv / (sp_ms * 1e-3) / 1e9, "frac_of_peak": B_spmv / (sp_ms * 1e-3) / 1e9 / peak,
                "algorithmic_bytes": B_spmv, "block": [0, 0]}
    except Exception as e:  # pragma: no cover
        spmv = {"error": str(e)[:200]}

    # ---- symbolic phase against what it has to write (the CSR pattern, once) -------------------------------
    B_sym = 4 * nnz_loc + 4 * (nrows_loc + 1)
    roofline_symbolic = {"bound": "hbm", "achieved": B_sym / (t_sym * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": B_sym / (t_sym * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes": B_sym, "ms": t_sym,
                         "note": "B_sym = 4 nnz + 4 (nrows + 1) per part; first symbolic phase of the process (device time, max over ranks; includes growing "
                                 "the memory pool: 15 GB of device allocations, which are several times slower inside this process -- torch context "
                                 "present -- than in a bare process: 178-266 ms first call, 174 ms repeated, scratch/run_symbolic.py).  Not bandwidth "
                                 "bound: the kernels take 132 ms, of which the per-row duplicate elimination (row_unique_kernel, two passes) is 80 %, "
                                 "see profiles/r2_launches_symbolic.csv"}

    # ---- Jacobi-CG on top of mul!: a fixed number of iterations on the assembled system ----------------
    cg = None
    try:
        if not args.no_cg and args.form != "stokes":
            n_own = assem.rows[0].indices[0].own_length
            bvec = np.empty(assem.brows[0].indices[0].local_length, dtype=np.float64)
            L.check(lib.graft_vec_get(ctx, 0, L.ptr(bvec)))
            bs, xs0 = [np.ascontiguousarray(bvec[:n_own])], [np.zeros(n_own)]
            it, rr = L.c_i32(), L.c_dbl()
            L.check(lib.graft_cg(comm, 0, L.ptr_array(bs), L.ptr_array(xs0), 0.0, args.cg_iters, 1, C.byref(it), C.byref(rr)))
            (cg_ms,), _ = allmax([float(assem.timers()[0][L.T_CG])])
            cg = {"iterations": int(it.value), "ms_per_iteration": cg_ms / max(int(it.value), 1), "relres": float(rr.value),
                  "note": "Jacobi-preconditioned CG, device resident: 1 mul! + 2 global reductions + 3 vector updates per iteration"}
    except Exception as e:  # pragma: no cover
        cg = {"error": str(e)[:200]}

    # ---- the general geometry routes of the same workload (N=1) --------------------------------------------------------
    routes, general = None, None
    if args.gpus == 1 and not args.no_general and args.geometry == "cartesian":
        routes = {}

        def time_route(name, geometry):
            try:
                a2, f2 = wl.assembler(geometry, local_rank)
                a2._symbolic(f2)
                # the fused sweep route keeps its per-cell geometry records (Jacobians, inverse maps, quadrature coefficients) between
                # steps while mesh, source and state are unchanged; the figure of record re-sets the source before every step,
                # which invalidates them: every step then starts from the node coordinates, like the reference's lazy cell arrays
                one = np.ones(wl.fields[0].spaces[0].ncomp)
                ctx2 = a2.comm.ctxs[0]
                redo = lambda: L.check(a2.comm.lib.graft_source_set(ctx2, 0, L.SOURCE_CONST, L.ptr(one), None))
                ms = time_steps(a2, L, 10, 3, barrier, before_step=redo) / 10
                tm = a2.timers()[0]
                ms_cached = time_steps(a2, L, 10, 3, barrier) / 10
                r = {"ms_per_step": ms, "ms_per_step_geometry_records_kept": ms_cached,
                     "phase_ms": {"integrate": float(tm[L.T_INTEGRATE]), "scatter": float(tm[L.T_SCATTER])},
                     "nnz_per_s": nnz_loc / (ms * 1e-3), "cells_per_s": st1["ncells"] / (ms * 1e-3), "frac_of_hbm_peak": B_num / (ms * 1e-3) / 1e9 / peak,
                     "route": a2.stats()[0]["path"], "symbolic_ms_device": float(tm[L.T_SYMBOLIC])}
                if not args.no_invariants and args.form in ("poisson", "elasticity"):
                    inv = check_invariants(wl, a2, L, torch, dist, cells, geometry)
                    r["invariants"] = "ok" if inv["ok"] else inv
                routes[name] = r
                a2.close()
            except Exception as e:  # pragma: no cover
                routes[name] = {"error": str(e)[:200]}

        # the same mesh given as hex node coordinates (affine cells), and with nodes moved by <= 0.1 h: general trilinear cells
        time_route("hex_nodes_affine", "hex")
        time_route("hex_nodes_perturbed", "perturbed")
        gp = routes.get("hex_nodes_perturbed", {})
        if "ms_per_step" in gp and args.form == "poisson":
            # algorithmic flops of the general-cell integration: 2 nq (D nd) nd = 118 098 per cell (dense B^T D B); the
            # sum-factorised kernels execute 2 x 11 664 + geometry per cell
            fl_alg, fl_exec = 118098.0 * st1["ncells"], (2 * 11664.0 + 27 * 150.0) * st1["ncells"]
            general = {"value": gp["nnz_per_s"], "unit": "nnz/s", "ms_per_step": gp["ms_per_step"], "route": gp["route"],
                       "ms_per_step_geometry_records_kept": gp.get("ms_per_step_geometry_records_kept"),
                       "kernel": {"fused-sweep": "sweep_geom_kernel+sweep_q2_kernel", "sumfact-gather": "integrate_sumfact_q2_kernel+gather_direct_kernel"}.get(gp["route"], "integrate_small_kernel+gather_direct_kernel"),
                       "config": "same workload, mesh as node coordinates moved by <= 0.1 h (general trilinear hexes, per-cell Jacobians at 27 points)",
                       "roofline": {"bound": "hbm", "achieved": B_num / (gp["ms_per_step"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": gp["frac_of_hbm_peak"],
                                    "traffic": traffic_of({"fused-sweep": "sweep_geom_kernel+sweep_q2_kernel", "sumfact-gather": "integrate_sumfact_q2_kernel+gather_direct_kernel"}.get(gp["route"], "integrate_small_kernel+gather_direct_kernel")),
                                    "algorithmic_bytes_per_launch": B_num},
                       "fp64": {"algorithmic_tflops": fl_alg / (gp["ms_per_step"] * 1e-3) / 1e12, "executed_tflops": fl_exec / (gp["ms_per_step"] * 1e-3) / 1e12,
                                "peak_tflops": FP64_PEAK_TFLOPS, "peak_source": "profiles/r2_fp64_peak.jsonl (DFMA/DMMA microbenchmark on this pool)"},
                       "invariants": gp.get("invariants")}

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        dvs = [torch.from_numpy(np.array(U.dirichlet_values[0], dtype=np.float64, copy=True)).pin_memory() for U in wl.fields]
        src = [torch.ones(U.spaces[0].ncomp, dtype=torch.float64).pin_memory() for U in wl.fields]
        vals = {b: torch.empty(sizes[b][2], dtype=torch.float64).pin_memory() for b in wl.blocks}
        bhs = [torch.empty(assem.brows[f].indices[0].local_length, dtype=torch.float64).pin_memory() for f in range(len(wl.fields))]

        def e2e_step():
            for f, U in enumerate(wl.fields):
                if len(dvs[f]):
                    L.check(lib.graft_space_set_dirichlet_values(ctx, f, len(dvs[f]), L.ptr(dvs[f].numpy())))
            L.check(lib.graft_source_set(ctx, 0, L.SOURCE_CONST, L.ptr(src[0].numpy()), None))
            L.check(lib.graft_numeric(comm, 3))
            for (bi, bj) in wl.blocks:
                L.check(lib.graft_csr_get_values(ctx, bi, bj, L.ptr(vals[(bi, bj)].numpy())))
            for f in range(len(wl.fields)):
                L.check(lib.graft_vec_get(ctx, f, L.ptr(bhs[f].numpy())))

        e2e_steps = max(2, min(args.steps, 5))
        e2e_step(); barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        (dt,), _ = allmax([time.perf_counter() - t0])
        v00 = vals[(0, 0)].numpy()
        e2e = {"value": tot_nnz * e2e_steps / dt, "unit": "nnz/s", "h2d_bytes_per_step": int(sum(8 * len(d) for d in dvs) + 8 * len(src[0])),
               "d2h_bytes_per_step": int(8 * nnz_loc + 8 * nrows_loc), "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "checksum": float(v00[: min(len(v00), 1 << 20)].sum())}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline and args.form == "poisson":
        r = cpu_reference_run(args.ref_cells, 2, 1)
        cpu = {"value": r["nnz"] / r["seconds"], "unit": "nnz/s", "cores": r["threads"], "kind": "port", "sample": cpu_sample_text(r),
               "cells_per_s": r["ncells"] / r["seconds"]}
    elif rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        cpu = {"value": None, "unit": "nnz/s", "cores": 0, "kind": "port",
               "sample": "the C restatement (oracle/assembly_ref.c) covers scalar Poisson Q2 only; run --form poisson for the CPU baseline"}

    if rank == 0:
        special = (" [best-case special case: Cartesian descriptor, all cell matrices equal, the step streams row templates; see "
                   "general_route for per-cell integration]") if args.geometry == "cartesian" else ""
        line = {"metric": FORMS[args.form]["metric"], "value": value, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"{FORMS[args.form]['what']}, {'x'.join(str(c) for c in cells)} cells, parts {parts}, {args.strategy}-assembled rows, "
                                       f"geometry={args.geometry}, re-assembly (numeric phase) of matrix+vector{special}",
                           "cells": tot_cells, "nnz": tot_nnz, "cells_per_s": tot_cells / (ms_per_step * 1e-3), "route": route,
                           "l2": "inputs+outputs (GBs per step) exceed the 126 MB L2; no explicit flush",
                           "symbolic_ms_device_max_over_ranks": t_sym, "symbolic_argmax_rank": a_sym, "symbolic_s_wall": t_symbolic_wall,
                           "host_setup_s": t_setup, "wall_ms_per_step": wall_ms / args.steps},
                "roofline": roofline, "roofline_symbolic": roofline_symbolic, "general_route": general, "invariants": ("ok" if invariants and invariants.get("ok") else invariants),
                "cpu_baseline": cpu, "e2e": e2e, "spmv": spmv, "cg": cg, "other_routes": routes,
                "gpu_launches": int(st1["launches"] - st0["launches"]), "clocks": clocks}
        print(json.dumps(line))
    assem.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Generate the golden fixtures of tests/golden/ from the CPU oracle (oracle/assembly_oracle.py).

The reference (Julia) cannot be imported or run in this image and its own tests hold no golden matrices
(SURVEY.md 8c), so these fixtures freeze the ORACLE's outputs on the reference's own test configurations: a
silent change of the oracle (or of the input producers of the host mirror) is then caught on CPU, and the GPU
parity tests compare the CUDA path against the same files.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import build_problem, neumann_cellvec_2d, oracle_assemble  # noqa: E402

CASES = {
    # reference test/PoissonTests.jl:14-45 (config 1): 4x4 cells on (0,4)^2, (2,2) parts, Q2, u=(x+y)^2, Neumann on tags 6, 8
    "poisson_config1_sub": dict(parts=(2, 2), cells=(4, 4), order=2, tags=[1, 2, 3, 5, 7], strategy="sub", domain=[0, 4, 0, 4], source=-4.0,
                                neumann=True),
    "poisson_config1_fully": dict(parts=(2, 2), cells=(4, 4), order=2, tags=[1, 2, 3, 5, 7], strategy="fully", domain=[0, 4, 0, 4], source=-4.0,
                                  neumann=True),
    # 3-D Q2 on (2,2,2) parts: the small sibling of config 3
    "poisson_3d_222_sub": dict(parts=(2, 2, 2), cells=(4, 4, 4), order=2, tags="boundary", strategy="sub", domain=None, source=1.0, neumann=False),
    # reference test/FESpacesTests.jl:61-70: Q1, l = int 1*v
    "poisson_q1_21_sub": dict(parts=(2, 1), cells=(5, 3), order=1, tags="boundary", strategy="sub", domain=None, source=1.0, neumann=False),
    # vector-valued block (config 4 in small): 2-D linear elasticity, Q2 x 2 components, lambda = 1.3, mu = 0.7
    "elasticity_2d_21_sub": dict(parts=(2, 1), cells=(3, 2), order=2, tags="boundary", strategy="sub", domain=None, source=0.5, neumann=False,
                                 form=("elasticity", 1.3, 0.7), ncomp=2),
}


# Round-2 features, frozen on the CPU side only (their CUDA parity tests compare with the live oracle: tests/test_gpu_*.py):
# general (perturbed) cells, the state-dependent p-Laplacian, a periodic model, the boundary facet term in 3-D.
CPU_CASES = {
    "poisson_3d_perturbed_211_sub": dict(parts=(2, 1, 1), cells=(4, 3, 3), order=2, tags="boundary", strategy="sub", domain=None, source=1.5,
                                         neumann=False, perturbed=(0.15, 7)),
    "plaplacian_2d_22_fully": dict(parts=(2, 2), cells=(4, 4), order=2, tags="boundary", strategy="fully", domain=[0, 4, 0, 4], source=0.7,
                                   neumann=False, form=("plaplacian",), state_seed=5),
    "poisson_periodic_2d_21_sub": dict(parts=(2, 1), cells=(6, 4), order=2, tags="boundary", strategy="sub", domain=None, source=1.0,
                                       neumann=False, isperiodic=(True, False)),
    "poisson_facets_3d_121_sub": dict(parts=(1, 2, 1), cells=(3, 4, 3), order=2, tags=[21, 23, 25], strategy="sub", domain=None, source=-1.0,
                                      neumann=False, facet_tags=[22, 24, 26], perturbed=(0.1, 3)),
}


def ufun(case):
    if case.get("ncomp", 1) > 1:
        D = len(case["cells"])
        return lambda x: np.stack([x[0] * (d + 1.0) - 0.5 * x[(d + 1) % D] for d in range(D)])
    if case["neumann"]:
        return lambda x: (x[0] + x[1]) ** 2
    return lambda x: sum(x[d] * (d + 1) for d in range(len(x)))


def build(case):
    pr = build_problem(case["parts"], case["cells"], case["order"], case["tags"], ufun(case), case["strategy"], domain=case["domain"],
                       ncomp=case.get("ncomp", 1), **({"isperiodic": case["isperiodic"]} if "isperiodic" in case else {}))
    extra = neumann_cellvec_2d(pr, lambda x, n: 2 * (x[0] + x[1]) * (n[0] + n[1])) if case["neumann"] else None
    return pr, extra


def run(case):
    from helpers import g, oracle_facet_cellvecs

    pr, extra = build(case)
    kw = {}
    if "perturbed" in case:
        pert = g.vertex_perturbation(case["perturbed"][0], seed=case["perturbed"][1])
        kw["perturb"] = lambda m, lids, X: pert(m, m.vertex_coordinates())[m.cell_vertex_ids()[lids - 1] - 1]
    if "state_seed" in case:
        glob = np.random.default_rng(case["state_seed"]).uniform(-0.5, 0.5, pr.U.gids.indices[0].n_global)
        kw["state"] = [glob[ids.l2g - 1] for ids in pr.U.gids.indices]
    if "facet_tags" in case:
        gfun = lambda x, n: np.cos(x[0]) * n[1] + x[2] * n[2] + 0.25
        extra = oracle_facet_cellvecs(pr, g.Boundary(pr.model, tags=case["facet_tags"]), gfun, 2 * case["order"], perturb=kw.get("perturb"))
    out, _ = oracle_assemble(pr, case.get("form", ("poisson",)), source=case["source"], extra_cellvec=extra, **kw)
    d = {}
    for k, p in enumerate(out):
        nown = len(p["rows"]["own_to_local"])
        d[f"p{k}_rows_l2g"] = p["rows"]["l2g"]; d[f"p{k}_rows_l2o"] = p["rows"]["l2o"]; d[f"p{k}_rows_nown"] = np.int64(nown)
        d[f"p{k}_cols_l2g"] = p["cols"]["l2g"]; d[f"p{k}_cols_l2o"] = p["cols"]["l2o"]
        d[f"p{k}_cols_nown"] = np.int64(len(p["cols"]["own_to_local"]))
        d[f"p{k}_rowptr"], d[f"p{k}_colind"], d[f"p{k}_vals"] = p["csr"]
        d[f"p{k}_b"] = p["b"]
    d["nparts"] = np.int64(len(out))
    return d


if __name__ == "__main__":
    for name, case in {**CASES, **CPU_CASES}.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run(case))
        print("wrote", name)

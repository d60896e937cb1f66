"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares; the
host-side tables / LHS / toy solver of the product package match the golden vectors of the live
reference; the JSON configuration loader mirrors the reference's attributes; the product fails
loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from tests.golden_util import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "surrdamh_b200.h")).read()
    return sorted(set(re.findall(r"SDB_API\s+[\w\s\*]+?\b(sdb_\w+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from surrdamh_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "run `python -m surrdamh_b200.build`"
    h = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(h, s), "header declares %s, library does not export it" % s
    assert sorted(_lib.EXPORTS) == syms          # the ctypes binding covers exactly the header
    assert _lib.lib().sdb_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the header structs: field order and natural alignment (sizes a C compiler gives)."""
    import subprocess
    import tempfile
    from surrdamh_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "surrdamh_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(sdb_model), sizeof(sdb_problem), sizeof(sdb_chain_state), sizeof(sdb_chain_args),
         offsetof(sdb_chain_args, state), offsetof(sdb_chain_args, seed), offsetof(sdb_chain_args, pre_cur));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as t:
        c = os.path.join(t, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(t, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [ctypes.sizeof(_lib.Model), ctypes.sizeof(_lib.Problem), ctypes.sizeof(_lib.ChainState), ctypes.sizeof(_lib.ChainArgs),
            _lib.ChainArgs.state.offset, _lib.ChainArgs.seed.offset, _lib.ChainArgs.pre_cur.offset]
    assert got == want


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call; compute without a device fails loudly."""
    from surrdamh_b200 import _lib as L
    h = L.lib()
    assert h.sdb_poly_eval(None, 10, 2, 1, None, None, None) == L.ERR_ARG
    assert b"sdb_poly_eval" in h.sdb_last_error()
    assert h.sdb_lu_solve(None, 5, None, 0, None, None, None) == L.ERR_ARG
    assert h.sdb_rbf_thin(None, 5, 2, 3, None, None, None) == L.ERR_ARG
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(L.SdbError):
            L.require_cuda()
        with pytest.raises(L.SdbError):
            L.dev(np.zeros(3))
        from surrdamh_b200 import surrogate_poly
        with pytest.raises(L.SdbError):
            surrogate_poly.Surrogate_apply(2, 1).apply([np.zeros((3, 1)), 1], np.zeros((4, 2)))
        with pytest.raises(L.SdbError):
            surrogate_poly.Surrogate_update(2, 1)


def test_tables_lhs_and_toy_solver_match_reference_golden():
    from surrdamh_b200 import tables
    from surrdamh_b200.lhs_normal import lhs_normal
    from surrdamh_b200.solvers import Solver_linela2exp_local, Solver_uninformative
    g = load("poly")
    for i, (d, m, deg) in enumerate(g["apply_cases"]):
        assert np.array_equal(tables.multi_indices(int(d), int(deg)), g[f"apply{i}_alpha"])
        assert np.array_equal(tables.hermite_table(int(deg)), g[f"apply{i}_herm"])
        assert tables.n_terms(int(d), int(deg)) == g[f"apply{i}_alpha"].shape[0]
    assert tables.multi_indices(3, 0).tolist() == [[0, 0, 0]]
    g = load("misc")
    lcases = [(4, 2, [5.0, 3.0], [[4, -2], [-2, 4]]), (7, 4, 0.0, 1.0), (16, 3, [1.0, 2.0, 3.0], [1.0, 0.5, 2.0])]
    for i, (n, d, mean, std) in enumerate(lcases):
        assert np.array_equal(lhs_normal(d, mean, std, n, 0), g[f"lhs{i}"])
    s = Solver_linela2exp_local()
    out = []
    for p in g["toy_params"]:
        s.set_parameters(p)
        out.append(s.get_observations())
    assert np.array_equal(np.array(out), g["toy_G"])
    u = Solver_uninformative(3, 2)
    u.set_parameters(np.ones(3))
    assert u.get_observations().tolist() == [0.0, 0.0]
    # degree rule of the polynomial update (surrogate_poly.py:71-75)
    assert tables.fit_degree(40, 2, 5, 1) == 5 and tables.fit_degree(20, 3, 3, 1) == 2 and tables.fit_degree(5, 4, 5, 1) == 1
    assert tables.fit_degree(3, 2, 5, 4) == 4                       # never decreases
    with pytest.raises(OverflowError):
        tables.fit_degree(10, 1, 5, 1)                               # the reference divides by log(1)


def test_configuration_mirrors_reference_attributes(tmp_path):
    from surrdamh_b200.configuration import Configuration
    from surrdamh_b200 import surrogate_poly, surrogate_rbf
    conf = {
        "saved_samples_name": "simple", "no_parameters": 2, "no_observations": 1,
        "problem_parameters": {"prior_mean": [5.0, 3.0], "prior_std": [[4, -2], [-2, 4]], "observations": [-1e-3], "noise_std": 2e-4},
        "solver_module_path": os.path.join(ROOT, "surrdamh_b200", "solvers.py"), "solver_module_name": "solvers",
        "solver_init_name": "Solver_linela2exp_local", "solver_parameters": {}, "no_solvers": 2,
        "samplers_list": [{"type": "MH", "max_samples": 100, "time_limit": 60, "proposal_std": 1.0, "surrogate_is_updated": True}],
        "surrogate_type": "poly", "surr_updater_parameters": {"max_degree": 5}}
    p = tmp_path / "c.json"
    p.write_text(json.dumps(conf))
    C = Configuration(4, str(p))
    assert (C.no_parameters, C.no_observations, C.no_samplers, C.no_full_solvers) == (2, 1, 4, 2)
    assert C.saved_samples_name == "simple" and C.initial_sample_type == "lhs" and C.use_surrogate and not C.debug
    assert C.surr_solver_init is surrogate_poly.Surrogate_apply and C.surr_updater_init is surrogate_poly.Surrogate_update
    assert C.surr_updater_parameters == {"no_parameters": 2, "no_observations": 1, "max_degree": 5}
    assert C.surr_solver_parameters == {"no_parameters": 2, "no_observations": 1}
    assert C.rank_surr_collector == 5 and C.solver_parent_rank == 4 and C.max_buffer_size == 1 << 30
    assert len(C.solver_parent_parameters) == 2 and C.solver_parent_parameters[0]["no_samplers"] == 4
    assert C.child_solver_init.__name__ == "Solver_linela2exp_local" and C.child_solver_parameters == {}
    assert (C.no_chains, C.refit_every, C.refit_mode, C.seed) == (4, 64, "broadcast", 0)       # additive keys, defaults
    conf2 = dict(conf, surrogate_type="rbf", surr_solver_parameters={"kernel_type": 1},
                 surr_updater_parameters={"kernel_type": 1, "no_keep": 500, "solver_type": "direct"}, no_chains=4096, refit_every=32)
    C2 = Configuration(4, conf=conf2)
    assert C2.surr_solver_init is surrogate_rbf.Surrogate_apply and C2.surr_updater_parameters["no_keep"] == 500
    assert C2.no_chains == 4096 and C2.refit_every == 32
    conf3 = {k: v for k, v in conf.items() if not k.startswith("surr")}
    assert Configuration(1, conf=conf3).use_surrogate is False


def test_shipped_example_configs_load():
    from surrdamh_b200.configuration import Configuration
    ex = os.path.join(ROOT, "examples")
    names = sorted(f for f in os.listdir(ex) if f.endswith(".json"))
    assert names
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        for f in names:
            C = Configuration(4, os.path.join("examples", f))
            assert C.list_alg and C.child_solver_init is not None
            for st in C.list_alg:
                assert st["type"] in ("MH", "DAMH") and "proposal_std" in st and "surrogate_is_updated" in st
    finally:
        os.chdir(cwd)


def test_host_solver_pool_matches_sequential_plugin_calls():
    """no_solvers > 1 with a host-evaluated G: the pool of solver processes (process_SOLVER / process_CHILD) returns,
    in order, exactly what one solver instance called sequentially returns."""
    from surrdamh_b200.solver_pool import HostSolverPool
    from surrdamh_b200.solvers import Solver_linela2exp_local
    rs = np.random.RandomState(3)
    P = rs.standard_normal((37, 2)) + np.array([5.0, 3.0])
    pool = HostSolverPool(3, os.path.join(ROOT, "surrdamh_b200", "solvers.py"), "solvers", "Solver_linela2exp_local", {})
    try:
        got = pool.evaluate(P)
        assert pool.evaluate(P[:1]).shape == (1, 1) and pool.no_calls == 38
    finally:
        pool.close()
    s = Solver_linela2exp_local()
    want = []
    for p in P:
        s.set_parameters(p)
        want.append([s.get_observations()])
    assert np.array_equal(got, np.array(want))

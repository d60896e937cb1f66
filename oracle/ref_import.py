"""Import the UNMODIFIED reference modules: from /root/reference in the build container, else from the
git-ignored copy oracle/_ref/ that `python -m oracle.make_ref` made of them (which travels to the GPU box).

TEST / BASELINE INFRASTRUCTURE.  Callers: oracle/make_golden.py, the "live reference" tests (which skip
themselves when neither tree is present), and bench.py's CPU legs (--impl reference, cpu_baseline), which time
the reference's own classes on the host cores.  Nothing under surrdamh_b200/ imports this.

Three shims are applied before import; the reference files are
never edited:
  1. np.math = math          -- surrogate_poly.py:147 uses np.math.factorial (gone in numpy 2)
  2. scipy minres(tol=) -> rtol=  -- surrogate_rbf.py:134 (scipy >= 1.14 renamed the kwarg)
  3. a stub `mpi4py` module  -- classes_SAMPLER.py:9 imports it only for a rank print
"""
import importlib
import math
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _has_reference(root):
    return os.path.isfile(os.path.join(root, "surrDAMH", "modules", "classes_SAMPLER.py"))


def _find_root():
    if os.environ.get("SURRDAMH_REFERENCE"):
        return os.environ["SURRDAMH_REFERENCE"]
    for root in ("/root/reference", os.path.join(_HERE, "_ref")):
        if _has_reference(root):
            return root
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return _has_reference(REF_ROOT)


_cache = {}


def load():
    """Return a namespace with the reference modules: cS, poly, rbf, lhs, solvers."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import numpy as np
    import scipy.sparse.linalg as splin

    if not hasattr(np, "math"):
        np.math = math  # shim 1

    if not getattr(splin.minres, "_sdb_shimmed", False):  # shim 2
        _orig = splin.minres

        def minres(A, b, *args, tol=None, **kw):
            if tol is not None and "rtol" not in kw:
                kw["rtol"] = tol
            return _orig(A, b, *args, **kw)

        minres._sdb_shimmed = True
        splin.minres = minres

    if "mpi4py" not in sys.modules:  # shim 3
        stub = types.ModuleType("mpi4py")

        class _Comm:
            def Get_rank(self):
                return 0

            def Get_size(self):
                return 1

        class _MPI:
            COMM_WORLD = _Comm()

        stub.MPI = _MPI
        sys.modules["mpi4py"] = stub

    pkg_dir = os.path.join(REF_ROOT, "surrDAMH")
    if pkg_dir not in sys.path:
        sys.path.insert(0, pkg_dir)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cS = importlib.import_module("modules.classes_SAMPLER")
        poly = importlib.import_module("modules.surrogate_poly")
        rbf = importlib.import_module("modules.surrogate_rbf")
        lhs = importlib.import_module("modules.lhs_normal")
    spec = importlib.util.spec_from_file_location(
        "ref_solver_examples", os.path.join(REF_ROOT, "examples", "solvers", "solver_examples.py"))
    solvers = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(solvers)
    ns = types.SimpleNamespace(cS=cS, poly=poly, rbf=rbf, lhs=lhs, solvers=solvers, root=REF_ROOT)
    _cache["ns"] = ns
    return ns


class Streams:
    """Explicit (z, u) streams shared by the two generator stubs below.

    z[step, :]   -- the d standard normals of iteration `step`
    u[step, 0:2] -- slot 0 feeds the stage-1 (or MH) test, slot 1 the stage-2 test.
    The proposal is drawn exactly once at the top of every iteration
    (classes_SAMPLER.py:150,182), so the normal stub advances the shared step
    counter and rewinds the uniform slot; the reference draws the second uniform
    only after a stage-1 pass (classes_SAMPLER.py:193-195), so uniforms are indexed
    (step, slot), never sequentially.
    """

    def __init__(self, z, u):
        self.z = z
        self.u = u
        self.step = -1
        self.slot = 0

    def normal_stub(self):
        return _StreamNormal(self)

    def uniform_stub(self):
        return _StreamUniform(self)


class _StreamUniform:
    """Stands in for Algorithm_PARENT's RandomState (classes_SAMPLER.py:39,123)."""

    def __init__(self, s):
        self.s = s

    def uniform(self, lo=0.0, hi=1.0):
        v = self.s.u[self.s.step, self.s.slot]
        self.s.slot += 1
        return v


class _StreamNormal:
    """Stands in for Proposal_GaussRandomWalk's RandomState (classes_SAMPLER.py:225,250).

    normal(mean, std) = mean + std*z, two roundings and no fma, which is what
    numpy's legacy generator computes (loc + scale*gauss).
    """

    def __init__(self, s):
        self.s = s

    def normal(self, mean, std):
        self.s.step += 1
        self.s.slot = 0
        return mean + std * self.s.z[self.s.step]


def inject_streams(alg, prop, streams):
    """Swap the name-mangled private generators of a reference Algorithm/Proposal pair."""
    alg._Algorithm_PARENT__generator = streams.uniform_stub()
    prop._Proposal_GaussRandomWalk__generator = streams.normal_stub()

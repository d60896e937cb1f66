"""TEST / BASELINE INFRASTRUCTURE -- populate oracle/_ref/ with the UNMODIFIED reference modules of the hot path.

    python -m oracle.make_ref            (build container only: needs /root/reference)

The reference is pure Python; "building" it is copying the files its own process scripts import for this path
 from where they lie under /root/reference into the git-ignored directory oracle/_ref/, which travels
to the GPU box with the repository snapshot like a built .so does.  Nothing is edited: the three environment shims (np.math,
scipy's tol -> rtol, a stub mpi4py) are applied by oracle/ref_import.py at import time.  Used by
  * bench.py --impl reference and bench.py's cpu_baseline leg ("kind": "reference"): the reference's own
    Algorithm_DAMH.run / Surrogate_apply.apply / Surrogate_update.update on the box's host cores;
  * tests/test_oracle_vs_reference.py on machines without /root/reference.
The product (surrdamh_b200/) never imports it.  A MANIFEST with the sha256 of every copied file is written next to them.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "surrDAMH/modules/classes_SAMPLER.py",       # Algorithm_MH / Algorithm_DAMH, Proposal_GaussRandomWalk, Problem_Gauss, Snapshot
    "surrDAMH/modules/surrogate_rbf.py",         # Surrogate_apply / Surrogate_update (RBF)
    "surrDAMH/modules/surrogate_poly.py",        # Surrogate_apply / Surrogate_update (polynomial)
    "surrDAMH/modules/lhs_normal.py",            # initial samples
    "examples/solvers/solver_examples.py",       # the toy forward model
]


def make(src_root="/root/reference", dest=DEST, quiet=False):
    if not os.path.isfile(os.path.join(src_root, FILES[0])):
        return None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src_root, "files": manifest}, f, indent=1)
    if not quiet:
        print("oracle/_ref: %d reference files copied from %s" % (len(FILES), src_root))
    return dest


if __name__ == "__main__":
    sys.exit(0 if make() else 1)

"""Build the CUDA library in-tree: surrdamh_b200/_C/libsurrdamh_b200.so (sm_100a only).

    python -m surrdamh_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the repository snapshot.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libsurrdamh_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "surrdamh_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and not _stale():
        return LIB
    if not os.path.isfile(nvcc):
        if os.path.isfile(LIB):
            return LIB
        raise RuntimeError("nvcc not found and %s is not built" % LIB)
    os.makedirs(OUT_DIR, exist_ok=True)
    import time
    t_start = time.time()
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        extra = os.environ.get("SDB_NVCC_EXTRA", "").split()
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % cmd[-3])
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    try:
        import json
        import time
        ver = subprocess.run([nvcc, "--version"], stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
        with open(os.path.join(OUT_DIR, "build_record.json"), "w") as f:
            json.dump({"built_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "seconds": round(time.time() - t_start, 1),
                       "forced": bool(force), "nvcc": ver, "flags": NVCC_FLAGS, "sources": [os.path.basename(s_) for s_ in sources()],
                       "library": os.path.basename(LIB)}, f, indent=1)
    except OSError:
        pass
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

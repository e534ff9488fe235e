"""Build libgpkl.so (hand-written sm_100a CUDA + the C ABI of include/gpkl.h) in-tree with nvcc.

    python gp-vae_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  Translation units compile in parallel and are cached
by mtime (objects under gp-vae_b200/build/); the .so is git-ignored but travels to the GPU box.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "gpkl", "libgpkl.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
FLAGS += os.environ.get("GPKL_EXTRA_NVCC_FLAGS", "").split()  # debug builds only, e.g. -DGPKL_PANEL_TRACE

# warp tier instantiations: (lanes per pair, rows per lane)
WARP_CFGS = [(8, 1), (16, 1), (32, 1), (16, 3), (32, 2)]


def jobs():
    """(source, defines, object) for every translation unit."""
    out = []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        name = os.path.basename(src)[:-3]
        if name == "gpkl_warp_inst":
            for lp, r in WARP_CFGS:
                for bwd in (0, 1):
                    out.append((src, ["-DGPKL_LP=%d" % lp, "-DGPKL_R=%d" % r, "-DGPKL_BWD=%d" % bwd],
                                os.path.join(OBJ, "%s_%d_%d_%d.o" % (name, lp, r, bwd))))
        else:
            out.append((src, [], os.path.join(OBJ, name + ".o")))
    return out


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + _deps())


def _compile(job, verbose):
    src, defs, obj = job
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + defs + ["-c", "-o", obj, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return job, r


def build_lib(force=False, verbose=False):
    js = jobs()
    srcs = sorted(set(j[0] for j in js)) + _deps()
    stamp0 = os.path.join(OBJ, "flags.txt")
    flags_ok = not os.path.isdir(OBJ) or (os.path.exists(stamp0) and open(stamp0).read() == " ".join(FLAGS))
    if not force and flags_ok and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(f) for f in srcs):
        return OUT  # library is newer than every source (also the case on the GPU box, where build/ does not travel)
    os.makedirs(OBJ, exist_ok=True)
    # objects built with other flags (debug builds via GPKL_EXTRA_NVCC_FLAGS) are stale
    stamp = os.path.join(OBJ, "flags.txt")
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(FLAGS):
        force = True
    todo = [j for j in js if force or _stale(j[2], j[0])]
    # longest first
    todo.sort(key=lambda j: ("warp_inst" not in j[0], j[1]), reverse=False)
    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1)) as ex:
        for job, r in ex.map(lambda j: _compile(j, verbose), todo):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s %s" % (job[0], " ".join(job[1])))
            if verbose:
                sys.stderr.write("== %s %s\n%s" % (os.path.basename(job[0]), " ".join(job[1]), r.stderr))
    with open(stamp, "w") as f:
        f.write(" ".join(FLAGS))
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + [j[2] for j in js],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libgpkl.so failed")
    return OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

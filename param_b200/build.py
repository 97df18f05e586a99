"""Build libparam_b200.so (sm_100a only) and the CPU oracle, in-tree.

    python -m param_b200.build            # build if sources are newer than the library
    python -m param_b200.build --force

nvcc cross-compiles without a GPU.  The library is plain CUDA runtime code behind a C ABI
(include/param_b200.h) — it does not link against torch.  It is linked against the *shared*
CUDA runtime so that, inside a torch process, it binds to the libcudart.so.12 torch already
loaded (one runtime instance, streams and the current device are shared).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "param_b200" / "csrc"
LIBDIR = ROOT / "param_b200" / "lib"
LIB = LIBDIR / "libparam_b200.so"
OBJDIR = ROOT / "build" / "obj"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "-cudart", "shared",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _headers():
    return sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "param_b200.h"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


class _BuildLock:
    """Exclusive advisory lock on build/.lock: N torchrun ranks on a fresh checkout must not compile into the
    same object files at once — one builds, the others wait and find the library up to date."""

    def __enter__(self):
        import fcntl
        (ROOT / "build").mkdir(parents=True, exist_ok=True)
        self._fh = open(ROOT / "build" / ".lock", "w")
        fcntl.flock(self._fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self._fh, fcntl.LOCK_UN)
        self._fh.close()


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = _sources()
    hdrs = _headers()
    if not force and not _stale(LIB, srcs + hdrs + [Path(__file__)]):
        return LIB
    with _BuildLock():
        # another process may have finished the build while this one waited for the lock
        if not force and not _stale(LIB, srcs + hdrs + [Path(__file__)]):
            return LIB
        return _build_cuda_locked(srcs, hdrs, force, verbose)


def _build_cuda_locked(srcs, hdrs, force: bool, verbose: bool) -> Path:
    if not Path(NVCC).exists():
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB}")
    OBJDIR.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        if force or _stale(obj, [src] + hdrs + [Path(__file__)]):
            cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    # link to a temporary name and rename: a process that dlopens the library never sees a half-written file
    tmp = LIB.with_suffix(".so.tmp%d" % os.getpid())
    cmd = [NVCC, "-shared", "-cudart", "shared", "-o", str(tmp), *map(str, objs),
           "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


def build_oracle(force: bool = False) -> Path:
    """Compile oracle/param_oracle.c (test infrastructure, CPU only)."""
    src = ROOT / "oracle" / "param_oracle.c"
    out = ROOT / "oracle" / "libparam_oracle.so"
    if not force and not _stale(out, [src]):
        return out
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-o", str(out), str(src), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed on oracle:\n{r.stdout}\n{r.stderr}")
    return out


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    force = "--force" in argv
    verbose = "-v" in argv or "--verbose" in argv
    lib = build_cuda(force=force, verbose=verbose)
    print(f"built {lib}")
    if (ROOT / "oracle" / "param_oracle.c").exists():
        print(f"built {build_oracle(force=force)}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())

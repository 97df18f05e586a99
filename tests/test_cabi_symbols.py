"""The C-ABI library loads on a CPU-only box and exports every symbol include/param_b200.h
declares (no compute calls here).  Also: the product package never touches oracle/."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "param_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from param_b200 import build, _cabi
    lib_path = build.build_cuda()
    assert lib_path.exists()
    declared = _declared()
    assert len(declared) >= 20
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared
    lib = ctypes.CDLL(str(lib_path))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in param_b200.h but not exported"


def test_abi_loads_and_reports_errors_without_gpu():
    from param_b200 import _cabi
    lib = _cabi.load()
    assert lib.pb200_abi_version() == 2
    assert b"invalid" in lib.pb200_error_string(-1)
    assert lib.pb200_launch_count() >= 0
    # argument validation happens before any CUDA call
    assert lib.pb200_embbag_fwd(None, 0, 4, None, 0, None, 0, 0, 0, None, 0, None, 4, 0, None) == -1


def test_product_package_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]|libparam_oracle", re.M)
    for py in (ROOT / "param_b200").rglob("*.py"):
        if py.name == "build.py":   # build() compiles the checker; it does not use it
            continue
        assert not pat.search(py.read_text()), f"{py} references oracle/"
    for src in (ROOT / "param_b200" / "csrc").glob("*"):
        assert "oracle" not in src.read_text().lower(), f"{src} references the oracle"


def test_kernels_refuse_cpu_tensors():
    import pytest
    import torch
    from param_b200 import ops
    from param_b200._cabi import PB200Error
    with pytest.raises(PB200Error):
        ops.embedding_bag_forward(torch.randn(4, 4), torch.zeros(2, dtype=torch.int64),
                                  torch.zeros(2, dtype=torch.int64))


def test_no_library_sort_left():
    """the backward's sort is hand-written (csrc/radix_sort.cu): no cub radix-sort kernel in the library"""
    import shutil
    import subprocess
    import pytest
    from param_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", str(build.build_cuda())], capture_output=True, text=True).stdout
    assert "DeviceRadixSort" not in out
    assert "radix_scatter_kernel" in out


def test_build_lock_is_exclusive_across_processes(tmp_path):
    """N ranks on a fresh checkout: one builds, the others wait (param_b200/build.py::_BuildLock)."""
    import subprocess
    import sys
    import time
    code = ("import sys, time; sys.path.insert(0, %r); from param_b200.build import _BuildLock\n"
            "with _BuildLock():\n"
            "    open(%r, 'a').write('in %%s %%.3f\\n' %% (sys.argv[1], time.time())); time.sleep(0.6)\n"
            "    open(%r, 'a').write('out %%s %%.3f\\n' %% (sys.argv[1], time.time()))\n")
    log = str(tmp_path / "lock.log")
    code = code % (str(ROOT), log, log)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(i)]) for i in range(3)]
    assert all(p.wait(timeout=60) == 0 for p in procs)
    events = [ln.split() for ln in open(log).read().splitlines()]
    # critical sections never interleave: the log is in/out pairs of the same process
    assert len(events) == 6
    for k in range(0, 6, 2):
        assert events[k][0] == "in" and events[k + 1][0] == "out" and events[k][1] == events[k + 1][1]


def test_sort_plan_geometry_covers_every_key_width():
    """The digit plan of the hand-written radix sort (csrc/sort_plan.cuh::sort_geometry) through its host-side query:
    digits of 8 or 10 bits, contiguous from bit 0, covering the key, with the fewest passes that do; tiles of whole
    bags (a multiple of 16, at most 2048) that cover the batch; keys / values at fixed offsets of the plan buffer."""
    import ctypes as C
    from param_b200 import _cabi
    lib = _cabi.load()

    def geom(n, T, B, rows):
        passes, tb, tpt = C.c_int32(), C.c_int32(), C.c_int32()
        bits, shifts = (C.c_int32 * 4)(), (C.c_int32 * 4)()
        ko, vo = C.c_int64(), C.c_int64()
        rc = lib.pb200_sort_plan_geometry(n, T, B, rows, C.byref(passes), bits, shifts, C.byref(tb), C.byref(tpt),
                                          C.byref(ko), C.byref(vo))
        assert rc == 0
        return passes.value, list(bits)[:passes.value], list(shifts)[:passes.value], tb.value, tpt.value, ko.value, vo.value

    for key_bits in range(1, 33):
        rows = (1 << key_bits) if key_bits < 32 else (1 << 32) - 1
        if key_bits > 1:
            rows -= 1 if key_bits < 32 else 0          # bits_for(2^k - 1) == k
        passes, bits, shifts, _, _, _, _ = geom(20 * 1000 * 4, 4, 1000, rows)
        assert all(b in (8, 10) for b in bits) and len(set(bits)) == 1
        assert shifts == [i * bits[0] for i in range(passes)]
        assert sum(bits) >= min(key_bits, 32), (key_bits, bits)
        # minimal: one pass fewer of the widest digit would not cover the key
        assert (passes - 1) * 10 < key_bits
    # the bench shape: 1 M rows = 20 key bits = two 10-bit passes; 10 M rows = three 8-bit passes; unknown = 4 x 8
    assert geom(256 * 65536 * 20, 256, 65536, 1_000_000)[:2] == (2, [10, 10])
    assert geom(25 * 65536 * 20, 25, 65536, 10_000_000)[:2] == (3, [8, 8, 8])
    assert geom(1000, 1, 50, 0)[:2] == (4, [8, 8, 8, 8])
    for n, T, B in ((256 * 65536 * 20, 256, 65536), (7 * 333 * 3, 7, 333), (5, 1, 5), (64 * 100 * 3000, 64, 100)):
        _, _, _, tile_bags, tiles, ko, vo = geom(n, T, B, 1000)
        assert tile_bags % 16 == 0 and 16 <= tile_bags <= 2048 and tiles * tile_bags >= B > (tiles - 1) * tile_bags
        assert ko == 0 and vo >= n * 4 and vo % 256 == 0
    assert lib.pb200_sort_plan_geometry(10, 0, 5, 100, None, None, None, None, None, None, None) == -1

"""Locate the UNMODIFIED reference tree and put it on sys.path the way its own runners expect.

The reference imports itself as `param_bench.train.comms.pt...` (train/comms/pt/comms.py:15-36), its
DLRM runner does script-directory imports (`import dlrm_data`, `from comms_utils import ...`,
train/comms/pt/dlrm.py:17-23), the compute driver imports `pytorch_emb` from its own directory, and
et_replay imports itself as `et_replay` (top-level package of the reference root).

Search order: $PARAM_REF (a checkout, or a directory that contains `param_bench/`), <repo>/baseline/_ref
(tools/make_baseline_ref.sh puts an unedited copy there; it travels to the GPU box), /root/reference.

et_replay imports two third-party modules that are not in this image and that this path never calls
(`pydot`, execution_trace.py:33; `intervaltree`, profiler_trace_analysis.py:28): empty
stand-ins are registered in sys.modules when the real ones are missing.  Nothing of the reference is edited.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path
from typing import Optional

_REPO = Path(__file__).resolve().parents[2]


def find_reference() -> Optional[Path]:
    """The directory that IS the reference checkout (contains train/ and et_replay/), or None."""
    cands = []
    env = os.environ.get("PARAM_REF")
    if env:
        cands += [Path(env), Path(env) / "param_bench"]
    cands += [_REPO / "baseline" / "_ref" / "param_bench", Path("/root/reference")]
    for c in cands:
        if (c / "train" / "comms" / "pt" / "comms.py").exists():
            return c
    return None


def _stub_missing_third_party() -> None:
    import torch  # noqa: F401  (torch.fx probes `import pydot` itself: it must see the truth, before the stand-in exists)
    for name in ("pydot", "intervaltree"):
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except ImportError:
            mod = types.ModuleType(name)
            mod.__doc__ = "empty stand-in registered by param_b200.integration.refpath (module absent from the image)"
            if name == "pydot":
                mod.Dot = type("Dot", (), {})
            if name == "intervaltree":
                mod.Interval = type("Interval", (), {})
                mod.IntervalTree = type("IntervalTree", (), {})
            sys.modules[name] = mod


def setup(require: bool = True) -> Optional[Path]:
    """Make `param_bench`, the runner script directories and `et_replay` importable.  Returns the
    reference root.  Idempotent."""
    ref = find_reference()
    if ref is None:
        if require:
            raise RuntimeError("reference tree not found: set PARAM_REF or run tools/make_baseline_ref.sh "
                               "(needs /root/reference) to create baseline/_ref")
        return None
    if ref.name == "param_bench":
        pkg_parent = ref.parent
    else:
        # a checkout under another name (e.g. /root/reference): expose it as `param_bench` through a
        # symlink in a scratch directory, as SURVEY section 8(c) does by hand
        pkg_parent = Path(tempfile.gettempdir()) / f"pb200_param_bench_{os.getuid()}"
        pkg_parent.mkdir(exist_ok=True)
        link = pkg_parent / "param_bench"
        if not link.exists():
            try:
                link.symlink_to(ref)
            except FileExistsError:
                pass
    paths = [str(pkg_parent), str(ref), str(ref / "train" / "comms" / "pt"), str(ref / "train" / "compute" / "pt")]
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["PYTHONPATH"] = os.pathsep.join(paths + [os.environ.get("PYTHONPATH", "")]).rstrip(os.pathsep)
    _stub_missing_third_party()
    return ref

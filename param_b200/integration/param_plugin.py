"""Drop the B200 kernels under the UNMODIFIED reference runners.

The reference tree is found by param_b200/integration/refpath.py ($PARAM_REF, baseline/_ref — an unedited copy
made by tools/make_baseline_ref.sh that travels to the GPU box —, /root/reference) and exposed as
`param_bench`, the reference's own import convention (train/comms/pt/comms.py:15-36).  On a box without the
reference use the stand-alone runners in param_b200/comms/pt/ instead.

    # config 3 — the reference's comms.py, b200 backend selected through its own plugin registry
    torchrun --nproc-per-node 8 -m -- param_b200.integration.param_plugin comms \
        --backend b200 --device cuda --collective all_to_all_single --b 1K --e 1G --z 1 --c 1
    #   (`-m --`: torchrun's own parser otherwise claims the runner's short options --e / --n / --b)
    #   (comms.py falls through to customized_backend[args.backend], comms.py:1506-1521)

    # config 4 — the reference's dlrm.py (backend class is hard-coded there, dlrm.py:1327-1333, so it
    # is patched in; the use_device_time shim fixes the reference's start-up crash)
    torchrun --nproc-per-node 8 -m -- param_b200.integration.param_plugin dlrm --mini-batch-size 8192 ...

    # config 1 — the reference's compute driver with the module swapped (nn.EmbeddingBag -> B200)
    python -m param_b200.integration.param_plugin emb --device gpu emb --dataset A

    # compute/python framework — the reference's run_benchmark.py with the B200 operator + its JSON config
    python -m param_b200.integration.param_plugin bench -c param_b200/compute/configs/b200_batched_embedding_bag.json \
        -d cuda -w 2 -i 5 -b --cuda-l2-cache on

    # config 5 — the reference's replay tools on a trace captured with tools/cfg5_capture.py
    torchrun --nproc-per-node 8 -m -- param_b200.integration.param_plugin comm_replay --trace-type et \
        --trace-path <dir> --backend b200
    torchrun --nproc-per-node 8 -m -- param_b200.integration.param_plugin et_replay --trace-path <dir> -m full \
        --replay-config param_b200/et/replay-config-b200-aten.json --backend b200
"""
from __future__ import annotations

import sys


def make_backend_class():
    """class B200ParamBackend(B200CommsMixin, PyTorchDistBackend) against the real reference."""
    from param_bench.train.comms.pt.pytorch_dist_backend import PyTorchDistBackend

    from ..comms.pt.backend import B200CommsMixin

    class B200ParamBackend(B200CommsMixin, PyTorchDistBackend):
        def __init__(self, bootstrap_info, commsParams):
            PyTorchDistBackend.__init__(self, bootstrap_info, commsParams)
            # the collectiveFunc / computeFunc tables were bound in the base __init__; rebind the
            # hot-path entries to the overrides (pytorch_backend_utils.py:161-180)
            self.collectiveFunc["all_to_all_single"] = self.all_to_all_single
            self.collectiveFunc["all_to_allv"] = self.all_to_allv
            self.collectiveFunc["all_to_all"] = self.all_to_all
            self.computeFunc["emb_lookup"] = self.emb_lookup

        # the runner passes its --backend value ("b200") down as the c10d backend name
        # (comms.py:1527-1532, :1455-1456): bootstrap and pass-through collectives run on NCCL
        def initialize_backend(self, master_ip, master_port, backend="nccl", eager_mode=False):
            return PyTorchDistBackend.initialize_backend(
                self, master_ip, master_port, backend="nccl" if backend == "b200" else backend,
                eager_mode=eager_mode)

        def initialize_groups(self, groupRanks=None, backend="nccl", force_new_group=False):
            return PyTorchDistBackend.initialize_groups(
                self, groupRanks, backend="nccl" if backend == "b200" else backend,
                force_new_group=force_new_group)

    return B200ParamBackend


def register(name: str = "b200"):
    """register_customized_backend(name, cls) — must run BEFORE the runner parses its arguments
    (choices are computed at parse time, comms_utils.py:1769-1771)."""
    from param_bench.train.comms.pt.pytorch_backend_utils import register_customized_backend

    cls = make_backend_class()
    register_customized_backend(name, cls)
    return cls


def run_comms(argv):
    register()
    from param_bench.train.comms.pt import comms

    sys.argv = ["comms.py"] + list(argv)
    comms.main()


def run_dlrm(argv):
    cls = register()
    import dlrm  # script-dir import, as the reference does (dlrm.py:17-23)

    orig = dlrm.commsDLRMBench.readArgs

    def read_args(self, parser, defaultModel="dlrm"):
        args = orig(self, parser, defaultModel)
        if not hasattr(args, "use_device_time"):
            args.use_device_time = False
        return args

    dlrm.commsDLRMBench.readArgs = read_args
    # PB200_PLUGIN_BACKEND=stock keeps the reference's own PyTorchDistBackend (c10d / NCCL): the comparator
    # run of the very same script
    import os
    if os.environ.get("PB200_PLUGIN_BACKEND", "b200") != "stock":
        dlrm.PyTorchDistBackend = cls
    sys.argv = ["dlrm.py"] + list(argv)
    dlrm.main()


def run_emb(argv):
    """reference driver.py / pytorch_emb.py with nn.EmbeddingBag replaced by the B200 module for
    --device gpu (the swap XlaEmbeddingBag makes for tpu, pytorch_emb.py:178-184)."""
    import pytorch_emb as ref_emb  # reference module (train/compute/pt on sys.path)

    from ..compute.pt.pytorch_emb import B200EmbeddingBag

    class _NN:
        def __getattr__(self, item):
            import torch.nn as nn
            return getattr(nn, item)

        @staticmethod
        def EmbeddingBag(features, embdim, mode="sum", **kw):
            return B200EmbeddingBag(features, embdim, mode=mode, **kw)

    ref_emb.nn = _NN()
    import driver  # reference driver

    sys.argv = ["driver.py"] + list(argv)
    driver.main()


def _register_et_backend_if_cuda(argv):
    """--backend b200 needs the class in et_replay's registry BEFORE its parser computes the choices
    (et_replay/comm/comms_utils.py:1446-1451)."""
    if "b200" in argv:
        from ..et.backend import register_et_backend
        register_et_backend("b200")


def run_comm_replay(argv):
    """the reference's et_replay/tools/comm_replay.py, unmodified; `--backend b200` routes every replayed
    all_to_all(v) to the peer-push kernel (comm_replay.py:1734-1761 picks customized_backend[...])"""
    _register_et_backend_if_cuda(argv)
    from et_replay.tools import comm_replay

    sys.argv = ["comm_replay.py"] + list(argv)
    comm_replay.main()


def run_et_replay(argv):
    """the reference's et_replay/tools/et_replay.py, unmodified: compute nodes are rebuilt from name +
    schema (the replay config's "import modules" pulls in param_b200.et / the aten override), comm nodes go to
    the backend chosen with --backend"""
    _register_et_backend_if_cuda(argv)
    from et_replay.tools import et_replay

    sys.argv = ["et_replay.py"] + list(argv)
    et_replay.main()


def run_trace_replay(argv):
    """the reference's train/comms/pt/commsTraceReplay.py, unmodified, on a basic / ET / kineto trace.  Its
    `"compute": "emb_lookup"` entries (commsTraceParser.py:137-147) call comms_utils.init_emb_lookup, which needs
    fbgemm_gpu (comms_utils.py:1966-1979: logs an error and returns without it): the B200 set-up of the same
    collectiveArgs fields (param_b200/comms/pt/emb_lookup.py) is put in its place, the replay loop is untouched."""
    cls = register()
    from param_bench.train.comms.pt import comms_utils as ref_utils, commsTraceReplay, pytorch_dist_backend

    from ..comms.pt.emb_lookup import init_emb_lookup

    if "b200" in argv:
        # initBackend constructs PyTorchDistBackend by name for --nw-stack pytorch-dist (commsTraceReplay.py:
        # 1311-1316, no plugin registry there): hand it the subclass
        pytorch_dist_backend.PyTorchDistBackend = cls

    ref_utils.init_emb_lookup = init_emb_lookup
    # same start-up defect as dlrm.py: commsParamsHolderBase reads args.use_device_time, which
    # only comms.py's parser defines (comms_utils.py:826)
    orig = commsTraceReplay.commsTraceReplayBench.readArgs

    def read_args(self, parser):
        args = orig(self, parser)
        if not hasattr(args, "use_device_time"):
            args.use_device_time = False
        return args

    commsTraceReplay.commsTraceReplayBench.readArgs = read_args
    sys.argv = ["commsTraceReplay.py"] + list(argv)
    commsTraceReplay.main()


def run_bench(argv):
    """the reference's train/compute/python/pytorch/run_benchmark.py with the B200 operator, input iterator and
    input-data generator registered (param_b200/compute/python_plugin.py) and _clear_cache taught sm_100"""
    from ..compute import python_plugin

    python_plugin.run_benchmark(argv)


def main():
    runners = {"comms": run_comms, "dlrm": run_dlrm, "emb": run_emb, "comm_replay": run_comm_replay,
               "et_replay": run_et_replay, "bench": run_bench, "trace_replay": run_trace_replay}
    if len(sys.argv) < 2 or sys.argv[1] not in runners:
        raise SystemExit("usage: param_plugin {comms|dlrm|emb|comm_replay|et_replay|bench|trace_replay} <runner args>")
    from . import refpath
    refpath.setup()
    # evidence for the logs: how many libparam_b200 kernels this process launched under the reference's runner
    import atexit

    def _report():
        try:
            from .. import _cabi
            if _cabi._lib is not None:
                print(f"[param_plugin {sys.argv[0]}] libparam_b200 kernels launched by this process: "
                      f"{_cabi.launch_count()}", flush=True)
        except Exception:  # noqa: BLE001
            pass

    atexit.register(_report)
    runners[sys.argv[1]](sys.argv[2:])


if __name__ == "__main__":
    main()

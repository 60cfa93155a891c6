"""Run the reference's own eval scripts UNMODIFIED on this backend.

The reference has no plugin interface: its scripts import `load_pretrained_model` from
`multi_modality_model.multi_modality_v1.model.builder` (eval/run_opus_ddp.py:8, eval_run_multichoice.py:8,
run_opus_online.py:7) and then only touch the duck-typed surface listed in SURVEY.md section 8b (`model.generate`,
`model.eval`, tokenizer attributes). `install()` registers this package under that module name *before* the script is
imported, so the script's import statement resolves to `opus_pllm_b200.builder` and nothing of the reference's model
code (peft / bitsandbytes / fair-esm / lightning imports) is executed. The reference's host-side helpers the scripts
use (`constants`, `conversation`, `mm_utils`, `utils`) are left alone: they are plain text / tokenizer code and are
imported from the user's reference checkout.

    python -m opus_pllm_b200.compat /path/to/multi_modality_v1/eval/run_opus_ddp.py --model-base-path ... (script args)

Two third-party modules the scripts import are stood in for when they are not installed:
  * `accelerate` (run_opus_ddp.py:15-16): `Accelerator().split_between_processes / .process_index / .is_main_process /
    .wait_for_everyone` and `accelerate.utils.gather_object`, on top of torch.distributed (dp.py)
  * `metrics_computing_opi.return_opi_metrics` (run_opus_ddp.py:18): metric scripts are out of scope (DESIGN.md
    section 7); the stand-in only reports that scoring was skipped. The real module is used when it imports.
ESM-2 weights: the reference fetches them through fair-esm's hub cache; this backend reads the same file
(`$TORCH_HOME/hub/checkpoints/esm2_t33_650M_UR50D.pt`) or `$OPUS_ESM_PATH`.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REF_BUILDER = "multi_modality_model.multi_modality_v1.model.builder"


def _accelerate_stub() -> tuple[types.ModuleType, types.ModuleType]:
    import torch.distributed as dist
    from . import dp

    class Accelerator:
        def __init__(self, *a, **k):
            self.process_index = int(os.environ.get("LOCAL_RANK", "0"))
            self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
            self._rank = int(os.environ.get("RANK", "0"))
            if self.num_processes > 1 and not dist.is_initialized():
                import torch
                torch.cuda.set_device(self.process_index)
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.process_index))

        @property
        def is_main_process(self):
            return self._rank == 0

        def wait_for_everyone(self):
            if self.num_processes > 1:
                dist.barrier()

        @contextlib.contextmanager
        def split_between_processes(self, inputs, apply_padding=False):
            yield dp.split_between_processes(inputs, self._rank, self.num_processes)

    def gather_object(obj):
        """accelerate.utils.gather_object: concatenation of every rank's list, in rank order"""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return obj
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, obj)
        return [x for p in parts for x in p]

    from importlib.machinery import ModuleSpec
    acc = types.ModuleType("accelerate")
    acc.Accelerator = Accelerator
    acc.__spec__ = ModuleSpec("accelerate", loader=None, is_package=True)
    acc.__path__ = []
    utils = types.ModuleType("accelerate.utils")
    utils.gather_object = gather_object
    utils.__spec__ = ModuleSpec("accelerate.utils", loader=None)
    acc.utils = utils
    return acc, utils


def install(force_stubs: bool = False) -> None:
    """Idempotent. After this, `from multi_modality_model.multi_modality_v1.model.builder import load_pretrained_model,
    return_cstp_path` yields this backend's functions."""
    from . import builder
    pkg = "multi_modality_model.multi_modality_v1.model"
    for name in ("multi_modality_model", "multi_modality_model.multi_modality_v1"):
        importlib.import_module(name)          # the user's reference checkout must be importable (host-side helpers)
    if not getattr(sys.modules.get(pkg), "__opus_b200__", False):
        m = types.ModuleType(pkg)
        m.__path__ = []                        # a package whose only submodule is the one registered below
        m.__opus_b200__ = True
        sys.modules[pkg] = m
    sys.modules[REF_BUILDER] = builder
    sys.modules[pkg].builder = builder
    try:
        importlib.import_module("transformers")   # probes `accelerate` while it imports: let it see the real state first
    except ImportError:
        pass
    try:
        if force_stubs:
            raise ImportError
        importlib.import_module("accelerate")
    except ImportError:
        acc, utils = _accelerate_stub()
        sys.modules["accelerate"], sys.modules["accelerate.utils"] = acc, utils
    try:
        if force_stubs:
            raise ImportError
        importlib.import_module("metrics_computing_opi")
    except Exception:
        met = types.ModuleType("metrics_computing_opi")

        def return_opi_metrics(results, input_path):
            print(f"[opus_pllm_b200.compat] metric scripts are not part of this backend: {len(results)} results for "
                  f"{input_path} were written but not scored")
        met.return_opi_metrics = return_opi_metrics
        sys.modules["metrics_computing_opi"] = met


def load_script(path: str, name: str | None = None) -> types.ModuleType:
    """Import a reference eval script by file path (as a module, `__main__` guard not triggered) after install()."""
    import importlib.util
    install()
    spec = importlib.util.spec_from_file_location(name or os.path.splitext(os.path.basename(path))[0], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(argv=None):
    import runpy
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m opus_pllm_b200.compat <reference eval script.py> [script arguments]")
    script = argv[0]
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    install()
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()

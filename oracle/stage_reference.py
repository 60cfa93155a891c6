"""TEST INFRASTRUCTURE. Stage the few UNMODIFIED reference files that the `-m gpu` tests execute under baseline/_ref/
(git-ignored, not gpurun-ignored: it travels to the GPU box next to the built .so and never enters the history):
the three eval scripts (multi_modality_v1/eval/run_opus_ddp.py, eval_run_multichoice.py, run_opus_online.py) and the
host-side helper modules they import (constants, conversation templates, mm_utils, utils). /root/reference itself does not
exist on the GPU box; no reference source is ever committed, and nothing of the reference's model code is staged.

Called by __graft_entry__.build() whenever /root/reference is present; a no-op otherwise."""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/multi_modality_model"
DST = os.path.join(ROOT, "baseline", "_ref", "multi_modality_model")
FILES = [
    "__init__.py",
    "multi_modality_v1/__init__.py",
    "multi_modality_v1/constants.py",
    "multi_modality_v1/conversation.py",
    "multi_modality_v1/mm_utils.py",
    "multi_modality_v1/utils.py",
    "multi_modality_v1/eval/__init__.py",
    "multi_modality_v1/eval/run_opus_ddp.py",
    "multi_modality_v1/eval/eval_run_multichoice.py",
    "multi_modality_v1/eval/run_opus_online.py",
]


def stage(force: bool = False) -> str | None:
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    shutil.rmtree(DST, ignore_errors=True)
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
    return DST


if __name__ == "__main__":
    print(stage(force=True))

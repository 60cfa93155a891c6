"""TEST INFRASTRUCTURE. Stage the UNMODIFIED reference Python tree under baseline/_ref/ (git-ignored, not
gpurun-ignored, so it travels to the GPU box next to the built .so) so that `-m gpu` tests can execute the reference's
own eval scripts (multi_modality_v1/eval/run_opus_ddp.py, eval_run_multichoice.py, run_opus_online.py) against this
backend. /root/reference itself does not exist on the GPU box and no reference source is ever committed.

Called by __graft_entry__.build() whenever /root/reference is present; a no-op otherwise."""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/multi_modality_model"
DST = os.path.join(ROOT, "baseline", "_ref", "multi_modality_model")


def stage(force: bool = False) -> str | None:
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST) and not force:
        newest_src = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(SRC) for f in fs)
        newest_dst = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(DST) for f in fs)
        if newest_dst >= newest_src:
            return DST
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.png", "*.jpg"))
    return DST


if __name__ == "__main__":
    print(stage(force=True))

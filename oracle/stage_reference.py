"""Stage the UNMODIFIED reference entry scripts under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun) so
that tests/test_gpu_integration.py can run them on top of the shadow packages:

    eval_incremental.py, learn_mapping.py, configs.py, util.py      (callers: must work unchanged)
    eval/language_eval.py, eval/util.py                             (the reference's OWN session loop: per-op mode)

Only possible in the build container (/root/reference does not exist on the GPU box).  Nothing staged here is product
code and nothing is copied into the tracked tree; `python -m oracle.stage_reference` is idempotent.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = ["eval_incremental.py", "learn_mapping.py", "configs.py", "util.py", "eval/__init__.py", "eval/language_eval.py",
         "eval/util.py"]


def stage(dst=None):
    dst = dst or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(REF):
        return None
    for rel in FILES:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), out)
    with open(os.path.join(dst, "README"), "w") as f:
        f.write("Unmodified files of /root/reference staged by oracle/stage_reference.py for the integration tests.\n")
    return dst


if __name__ == "__main__":
    print(stage() or "no /root/reference here", file=sys.stderr)

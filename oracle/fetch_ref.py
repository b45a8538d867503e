"""Brings the reference's OWN caller scripts next to the oracle as untracked files (``oracle/_ref/`` is git-ignored, not
gpurun-ignored: it travels to the GPU box like a built .so and never enters the history):

    python oracle/fetch_ref.py            # needs /root/reference (or CTP_REFERENCE_DIR); a no-op elsewhere

``tests/test_reference_callers.py`` executes ``oracle/_ref/test_pipelines.py`` UNCHANGED against a synthetic checkpoint
directory, which is how the drop-in claim for ``tests/test_pipelines.py:9-63`` is checked on the B200 box (where the reference
checkout does not exist).  Test infrastructure only: nothing in the product package reads this directory.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CTP_REFERENCE_DIR", "/root/reference")
FILES = ["tests/test_pipelines.py"]


def fetch() -> list:
    out = []
    if not os.path.isdir(REF):
        return out
    dst_dir = os.path.join(HERE, "_ref")
    os.makedirs(dst_dir, exist_ok=True)
    for rel in FILES:
        src = os.path.join(REF, rel)
        if os.path.exists(src):
            dst = os.path.join(dst_dir, os.path.basename(rel))
            shutil.copyfile(src, dst)
            out.append(dst)
    return out


if __name__ == "__main__":
    got = fetch()
    print("\n".join(got) if got else f"{REF} not present: nothing fetched", file=sys.stderr if not got else sys.stdout)

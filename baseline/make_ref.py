#!/usr/bin/env python
"""Recipe for baseline/_ref/ — a verbatim, git-ignored copy of the reference's Python package (`src/`), so that the
UNMODIFIED reference can run on the GPU box, where /root/reference does not exist:

  * `bench.py --impl reference` and the `cpu_baseline` leg time the reference's own GraphGPTPretrainBase (kind
    "reference") on the box's host cores;
  * tests/test_reference_loop_gpu.py drives the reference's own batch_training / ft_batch_training
    (src/utils/training_utils.py:7-95, 98-205) against GraphGPTEngine + the B200 model classes.

baseline/_ref/ is listed in .gitignore (reference sources never enter this repository's history) and NOT in
.gpurunignore (it travels with the snapshot, like the built .so).  __graft_entry__.build() runs this whenever
/root/reference is present.  Nothing in the product package imports baseline/.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref")


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: keeping whatever is in {DST}")
        return os.path.isdir(os.path.join(DST, "src"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, f), os.path.dirname(SRC))
            out = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), out)
            n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("verbatim copy of /root/reference/src/**/*.py (alibaba/graph-gpt @ 7aff6083) made by baseline/make_ref.py\n")
    if verbose:
        print(f"copied {n} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)

#!/usr/bin/env python
"""Recipe for oracle/_ref/: a runnable copy of the UNMODIFIED Python reference's hot path for the CPU arm of
bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is copying the four
modules of the path from where they lie under /root/reference — environment.py, draw_line.py, transforms.py and
visualize_voxel.py (imported by environment.py:9, never called on the path) — into oracle/_ref/, which is
git-ignored (no reference source enters the history) but not gpurun-ignored (it travels to the GPU box with the
snapshot, like the built .so files).  The 808 MB of dense tumour volumes do not travel: oracle/ref_runtime.py
regenerates the ones a run needs from the packed phantom table (ppo-radiotherapy_b200/data/phantom.npz) into a
temporary directory laid out as the reference expects (./data/lungs.npy, ./data/tumours/<x_y_z_r>.npy,
environment.py:28-29,90-95).

    python oracle/build_ref.py            # /root/reference -> oracle/_ref/
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("RT_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("environment.py", "draw_line.py", "transforms.py", "visualize_voxel.py")


def build(verbose: bool = True) -> bool:
    """Copy the reference modules to oracle/_ref/.  Returns False (and leaves any earlier copy alone) when the
    reference tree is absent, e.g. on the GPU box."""
    if not all(os.path.isfile(os.path.join(REF_SRC, m)) for m in MODULES):
        if verbose:
            print(f"build_ref: {REF_SRC} not present; keeping {OUT} as it is")
        return False
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        shutil.copyfile(os.path.join(REF_SRC, m), os.path.join(OUT, m))
    with open(os.path.join(OUT, "SOURCE.txt"), "w") as f:
        f.write(f"verbatim copies of {', '.join(MODULES)} from {REF_SRC} (oracle/build_ref.py); not part of the repository\n")
    if verbose:
        print(f"build_ref: copied {len(MODULES)} modules to {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)

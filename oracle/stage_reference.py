"""Stage the UNMODIFIED reference files of the update path where they can travel to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

``/root/reference`` exists only in the build container; ``gpurun`` and the driver snapshot ``/root/repo``.  This recipe
(called from ``__graft_entry__.build()``, or ``python -m oracle.stage_reference``) copies byte-for-byte

    /root/reference/ces/__init__.py, calibrate.py, utils.py   ->   baseline/_ref/ces/

``baseline/_ref/`` is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so
``bench.py --impl reference`` can time the real ``sampling.eks_update_aldi`` (ces/calibrate.py:451-490) and
``sampling.run`` (:270-416) on the GPU box's host cores.  ``oracle/reference_loader.py`` looks at ``/root/reference``
first and here second.  A pip install of the reference is not possible (no setup.py / pyproject, and
``ces/calibrate.py`` raises TabError on import under Python 3) -- the loader tab-expands it in memory.
"""
import filecmp
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = os.environ.get("CES_REFERENCE_ROOT", "/root/reference")
TARGET = os.path.join(ROOT, "baseline", "_ref")
FILES = ("ces/__init__.py", "ces/calibrate.py", "ces/utils.py")


def stage(verbose=False):
    """Copy the files if the reference is present; returns the staged directory or None."""
    if not os.path.isfile(os.path.join(SOURCE, "ces", "calibrate.py")):
        return TARGET if os.path.isfile(os.path.join(TARGET, "ces", "calibrate.py")) else None
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(TARGET, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged %s" % rel)
    with open(os.path.join(TARGET, "README"), "w") as fh:
        fh.write("Unmodified copies of agarbuno/ces files (%s) staged by oracle/stage_reference.py for\n"
                 "bench.py --impl reference.  Git-ignored; not part of this repository.\n" % ", ".join(FILES))
    return TARGET


if __name__ == "__main__":
    print(stage(verbose=True))

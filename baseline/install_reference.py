"""Puts the UNMODIFIED reference modules under baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot) so that
`bench.py --impl reference` and the in-line cpu_baseline drive the reference's own code (kind "reference") instead of the
oracle's port.

The reference has no setup.py / pyproject.toml: `pip install --no-index --no-build-isolation --target baseline/_ref /root/reference`
ends with "Neither 'setup.py' nor 'pyproject.toml' found" (recorded in DESIGN.md §5), so the install is a byte-for-byte copy of its
top-level modules. Nothing under baseline/_ref/ is committed or imported by the product package.

    python baseline/install_reference.py [/root/reference]
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
MODULES = ("GNAN.py", "models.py", "trainer.py", "pre_process_datasets.py", "batched_pyg_main.py", "datasets.py", "main.py")


def install(src="/root/reference"):
    if not os.path.exists(os.path.join(src, "GNAN.py")):
        return None
    os.makedirs(DST, exist_ok=True)
    for m in MODULES:
        s, d = os.path.join(src, m), os.path.join(DST, m)
        if os.path.exists(s) and not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    return DST


if __name__ == "__main__":
    print(install(*sys.argv[1:2]))

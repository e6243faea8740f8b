#!/usr/bin/env python
"""Install the UNMODIFIED reference under ``baseline/_ref`` (git-ignored, travels to the GPU box with the
working tree) so that the CPU-baseline leg of ``bench.py`` and the drop-in tests can run the reference's own
code where ``/root/reference`` does not exist.

The reference is a script tree without ``setup.py`` / ``pyproject.toml`` — ``pip install`` has nothing to build —
so "installing" it is copying its Python sources and the one data file its scenes read.  Nothing under
``baseline/_ref`` is tracked, edited, or imported by the product package.

    python baseline/install_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
KEEP = (".py", ".jld2")


def install(src: str = "/root/reference") -> int:
    if not os.path.isdir(src):
        return 0
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("result_images", "extras", ".git", "__pycache__")]
        for f in files:
            if not f.endswith(KEEP):
                continue
            rel = os.path.relpath(os.path.join(root, f), src)
            out = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), out)
            n += 1
    return n


if __name__ == "__main__":
    print(f"installed {install(*(sys.argv[1:2]))} reference files into {DEST}")

"""Stage the unmodified reference package under oracle/_ref/ (git-ignored).  Build-container only.

    python oracle/make_ref.py            # copies /root/reference/edm2 -> oracle/_ref/edm2 (Python sources only)

The reference has no native code to compile: its `edm2/` package IS the runnable artefact.  The copy is
never committed (oracle/_ref/ is in .gitignore) and is not listed in .gpurunignore, so it travels to the GPU box like
a built .so.  See oracle/ref_shim.py for who may use it.
"""
import os
import shutil
import sys

SRC = os.environ.get("ONIRIS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def main():
    pkg = os.path.join(SRC, "edm2")
    if not os.path.isdir(pkg):
        print(f"{pkg} not found: nothing staged", file=sys.stderr)
        return 1
    out = os.path.join(DST, "edm2")
    if os.path.isdir(out):
        shutil.rmtree(out)
    shutil.copytree(pkg, out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb", "*.png", "*.jpg", "*.mp4", "*.pt"))
    n = sum(len(files) for _, _, files in os.walk(out))
    print(f"staged {n} files under {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

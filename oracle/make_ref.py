"""Recipe for oracle/_ref: the UNMODIFIED reference, importable on the GPU box.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference (quicophy/mdopt v1.1.1) is pure Python: there is nothing to compile.  This script copies the modules of
the decode path from /root/reference as they lie -- ``mdopt/`` (the library) and ``examples/decoding/decoding.py`` (the
decoder) -- into ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored, so it travels with the snapshot like a built .so).
Third-party modules the reference imports but the image lacks come from the stand-ins in ``oracle/refshim``.
``bench.py --impl reference`` and the ``cpu_baseline`` leg import it from there (``kind: "reference"``); when the
directory is absent they fall back to the numpy port (``kind: "port"``).  Nothing under ``mdopt_b200/`` touches it.

    python oracle/make_ref.py            # run in the build container (build() does)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def make(force: bool = False) -> bool:
    if not os.path.isdir(os.path.join(SRC, "mdopt")):
        return os.path.isdir(os.path.join(DST, "mdopt"))   # GPU box: use what travelled
    if os.path.isdir(DST) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "examples", "decoding"))
    shutil.copytree(os.path.join(SRC, "mdopt"), os.path.join(DST, "mdopt"),
                    ignore=shutil.ignore_patterns("__pycache__"))
    shutil.copy(os.path.join(SRC, "examples", "decoding", "decoding.py"), os.path.join(DST, "examples", "decoding"))
    for pkg in ("examples", os.path.join("examples", "decoding")):
        init = os.path.join(SRC, pkg, "__init__.py")
        if os.path.exists(init):
            shutil.copy(init, os.path.join(DST, pkg))
        else:
            open(os.path.join(DST, pkg, "__init__.py"), "a").close()
    return True


def import_reference():
    """(decode_css, qecstruct stand-in) of the unmodified reference in oracle/_ref, or None when it is absent."""
    if not os.path.isdir(os.path.join(DST, "mdopt")):
        return None
    for p in (DST, os.path.join(HERE, "refshim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import qecstruct as qec  # refshim
    from examples.decoding.decoding import decode_css

    return decode_css, qec


if __name__ == "__main__":
    print("oracle/_ref ready" if make(force=True) else "no /root/reference here and no oracle/_ref")

"""Build the CPU oracles (TEST INFRASTRUCTURE ONLY) into oracle/_build/.

    python oracle/build.py            # g++/gcc only, a few seconds

The reference-compiled checkers (oracle/_ref/) are built by oracle/build_ref.py.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")

TARGETS = [
    # (output, compiler, sources, flags)
    ("libfrustum_oracle.so", "g++", ["frustum_oracle.cpp"], ["-std=c++17", "-O2", "-ffp-contract=off"]),
    ("libops_oracle.so", "gcc", ["ops_oracle.c"], ["-std=c11", "-O2"]),
]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    for name, cc, srcs, flags in TARGETS:
        out = os.path.join(OUT, name)
        srcs = [os.path.join(HERE, s) for s in srcs]
        if not force and not _stale(out, srcs):
            continue
        cmd = [cc, "-shared", "-fPIC", *flags, "-o", out + ".tmp", *srcs, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        os.replace(out + ".tmp", out)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("oracle built in", OUT)

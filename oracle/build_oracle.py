"""Builds oracle/liborc.so (the CPU checker).  Test infrastructure: called by tests/conftest.py,
__graft_entry__.build() and bench.py's CPU-baseline legs - never by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liborc.so")


def build_oracle(force: bool = False) -> str:
    deps = [os.path.join(HERE, f) for f in ("oracle.c", "oracle.h", "Makefile")]
    stale = not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)
    if force or stale:
        r = subprocess.run(["make", "-C", HERE, "-B", "liborc.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            if os.path.exists(LIB) and not force:
                return LIB  # e.g. no compiler on the box: keep the prebuilt checker
            raise RuntimeError("oracle build failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))

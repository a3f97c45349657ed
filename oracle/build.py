"""Build the C oracle (oracle/walrus_oracle.c) into oracle/_build/.  TEST INFRASTRUCTURE ONLY.

Two shared objects: liboracle.so (double, the checker and the CPU baseline) and liboracle_ld.so
(long double, extended-precision yardstick).  Compiled with -march=native, so a stamp of the CPU flags
forces a rebuild when the snapshot lands on a different host (the GPU box).
The reference (/root/reference) is Python + Numba with no C sources, so there is no `oracle/_ref` build.
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "walrus_oracle.c")


def _cpu_stamp():
    flags = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    with open(SRC, "rb") as fh:
        src = fh.read()
    return hashlib.sha256(flags.encode() + src).hexdigest()


def ensure(verbose=False):
    """Compile if missing or stale; return (path_double, path_long_double)."""
    os.makedirs(OUT, exist_ok=True)
    stamp_file = os.path.join(OUT, "stamp")
    stamp = _cpu_stamp()
    so = os.path.join(OUT, "liboracle.so")
    so_ld = os.path.join(OUT, "liboracle_ld.so")
    fresh = False
    if os.path.exists(stamp_file) and os.path.exists(so) and os.path.exists(so_ld):
        with open(stamp_file) as fh:
            fresh = fh.read().strip() == stamp
    if not fresh:
        base = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=gnu11", SRC, "-lm"]
        for extra, target in (([], so), (["-DORACLE_LONG_DOUBLE"], so_ld)):
            cmd = base + extra + ["-o", target]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        with open(stamp_file, "w") as fh:
            fh.write(stamp)
    return so, so_ld


if __name__ == "__main__":
    print(ensure(verbose=True))

"""Development aid: run the CPU checkers (oracle/vgi_oracle.c and the reference-shader library of oracle/glsl_shim) under
UndefinedBehaviorSanitizer or AddressSanitizer. Builds instrumented copies into a temporary directory and runs
tests/test_ref_shaders.py, tests/test_oracle_kat.py and tests/test_dump.py against them.

    python tools/sanitize_oracle.py ubsan        # needs /root/reference for the shader library (skipped otherwise)
    python tools/sanitize_oracle.py asan         # re-executes itself with libasan preloaded

Round 1: both clean (no report) over the whole suite."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "glsl_shim"))
GCC = "/usr/bin/gcc"


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "ubsan"
    flag = {"ubsan": "-fsanitize=undefined", "asan": "-fsanitize=address"}[mode]
    if mode == "asan" and "libasan" not in os.environ.get("LD_PRELOAD", ""):
        lib = subprocess.check_output([GCC, "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, LD_PRELOAD=lib, ASAN_OPTIONS="detect_leaks=0:halt_on_error=0")
        sys.exit(subprocess.call([sys.executable] + sys.argv, env=env))
    import build_ref as B
    from oracle import pyoracle
    tmp = tempfile.mkdtemp(prefix="vgi_sanitize_")
    olib = os.path.join(tmp, "liboracle.so")
    subprocess.check_call([GCC, "-O1", "-g", "-std=gnu11", "-fPIC", "-fopenmp", "-ffp-contract=off", flag, "-fno-omit-frame-pointer",
                           "-shared", "-o", olib, os.path.join(ROOT, "oracle", "vgi_oracle.c"), "-lm"])
    pyoracle._LIB_PATH = olib
    pyoracle.build = lambda force=False: olib
    if B.reference_available():
        B.CXXFLAGS = [f for f in B.CXXFLAGS if f != "-O3"] + ["-O1", "-g", flag, "-fno-omit-frame-pointer"]
        rlib = os.path.join(tmp, "librefshaders.so")
        objs = []
        for u in B.UNITS:
            obj = os.path.join(tmp, u.replace(".", "_") + ".o")
            r = B.compile_unit(B.unit_source(u), obj)
            if r.returncode:
                raise SystemExit(r.stderr[-4000:])
            objs.append(obj)
        subprocess.check_call([B.GXX, "-shared", "-fopenmp", flag, "-o", rlib] + objs)
        B.OUT_SO = rlib
        B.build = lambda force=False, verbose=False: rlib
    import pytest
    tests = [os.path.join(ROOT, "tests", t) for t in ("test_ref_shaders.py", "test_oracle_kat.py", "test_dump.py")]
    sys.exit(pytest.main(tests + ["-q", "-p", "no:cacheprovider"]))


if __name__ == "__main__":
    main()

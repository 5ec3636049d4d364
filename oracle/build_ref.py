#!/usr/bin/env python
"""Build the UNMODIFIED-IN-BEHAVIOUR reference (jjgoings/McMurchie-Davidson) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

The reference is Cython + Python and does not build as-is on this image (Cython 3.3, SciPy 1.18,
NumPy 2.3).  This script copies /root/reference to a scratch dir under /tmp, applies four mechanical
compatibility patches that do not change the arithmetic, runs the reference's own
`setup.py build_ext --inplace`, and installs the resulting package (python files, basis data, built
.so files) into oracle/_ref/ — which is git-ignored (never part of history) but NOT gpurun-ignored,
so it travels to the GPU box where it is timed as the CPU baseline.

Patches (all on the scratch copy, the read-only reference is never touched):
  1. cython/basis.pxi:29     `long(view)` -> `int(view)`           (Cython 3: `long` is not a builtin)
  2. cython/*.pyx            `xrange` -> `range`                    (py3)
  3. cython/basis.pxi:4, cython/onee.pyx:8, mmd/integrals/reference.py:2
                             scipy `factorial2(-1)` now returns 0 (old scipy: 1) which turns every
                             s/p norm into inf/NaN -> shim `fact2(n) = 1 if n <= 0 else factorial2(n)`
  4. bitstring stub          mmd/postscf.py:7 imports `bitstring` at module top (CI code only)
  5. mmd/slater.py:45        `np.int` -> `int` (only CI code; harmless)

Usage: python oracle/build_ref.py [--force]
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
SCRATCH = "/tmp/mmd_refbuild"

FACT2_SHIM = (
    "from scipy.special import factorial2 as _scipy_fact2\n"
    "def fact2(n):\n"
    "    # (-1)!! = 1 (old SciPy semantics the reference relies on)\n"
    "    return 1.0 if n <= 0 else float(_scipy_fact2(int(n), exact=True))\n"
)

BITSTRING_STUB = '''"""Minimal stand-in for the `bitstring` package (absent from this image).
Only what mmd/postscf.py + mmd/slater.py touch: BitArray(bin=...).uint / .bin"""
class BitArray(object):
    def __init__(self, bin=None, uint=None, length=None):
        if bin is not None:
            self.bin = str(bin)
        else:
            self.bin = format(int(uint), "0%db" % int(length))
    @property
    def uint(self):
        return int(self.bin, 2)
'''


def _sub(path, pattern, repl, count=0, must=True):
    with open(path) as f:
        src = f.read()
    new, n = re.subn(pattern, repl, src, count=count)
    if must and n == 0:
        raise RuntimeError("patch did not apply: %s :: %s" % (path, pattern))
    with open(path, "w") as f:
        f.write(new)
    return n


def available():
    return os.path.exists(os.path.join(OUT, "mmd", "integrals", "__init__.py")) and any(
        fn.startswith("twoe") and fn.endswith(".so")
        for fn in os.listdir(os.path.join(OUT, "mmd", "integrals"))
    )


def build(force=False):
    if available() and not force:
        return OUT
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s not present (GPU box?) and oracle/_ref not prebuilt" % REF)
    if os.path.exists(SCRATCH):
        shutil.rmtree(SCRATCH)
    shutil.copytree(REF, SCRATCH, ignore=shutil.ignore_patterns("backup", "huckel", "examples", "*.pyc"))
    subprocess.check_call(["chmod", "-R", "u+w", SCRATCH])
    cy = os.path.join(SCRATCH, "cython")
    # 1
    _sub(os.path.join(cy, "basis.pxi"), r"return long\(view\)", "return int(view)")
    # 2
    for fn in os.listdir(cy):
        if fn.endswith(".pyx") or fn.endswith(".pxi"):
            _sub(os.path.join(cy, fn), r"\bxrange\b", "range", must=False)
    _sub(os.path.join(SCRATCH, "mmd", "utils", "spectrum.py"), r"\bxrange\b", "range", must=False)
    # 3
    _sub(os.path.join(cy, "basis.pxi"), r"from scipy\.special import factorial2 as fact2 *\n", FACT2_SHIM)
    _sub(os.path.join(cy, "onee.pyx"), r"from scipy\.special import factorial2 as fact2 *\n", FACT2_SHIM)
    # grad.pyx may import fact2 as well
    _sub(os.path.join(cy, "grad.pyx"), r"from scipy\.special import factorial2 as fact2 *\n", FACT2_SHIM, must=False)
    _sub(os.path.join(SCRATCH, "mmd", "integrals", "reference.py"),
         r"from scipy\.(special|misc) import factorial2 as fact2 *\n", FACT2_SHIM)
    # 4
    with open(os.path.join(SCRATCH, "bitstring.py"), "w") as f:
        f.write(BITSTRING_STUB)
    # 5
    _sub(os.path.join(SCRATCH, "mmd", "slater.py"), r"np\.int\b", "int", must=False)

    env = dict(os.environ)
    env["CFLAGS"] = env.get("CFLAGS", "") + " -O2 -w"
    log = os.path.join(SCRATCH, "build.log")
    with open(log, "w") as lf:
        rc = subprocess.call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=SCRATCH,
                             stdout=lf, stderr=subprocess.STDOUT, env=env)
    if rc != 0:
        sys.stderr.write(open(log).read()[-4000:])
        raise RuntimeError("reference build failed, see %s" % log)

    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    shutil.copytree(os.path.join(SCRATCH, "mmd"), os.path.join(OUT, "mmd"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.c"))
    shutil.copy(os.path.join(SCRATCH, "bitstring.py"), os.path.join(OUT, "bitstring.py"))
    # the reference's own tests, kept as a smoke suite for the oracle build (run from here only)
    shutil.copytree(os.path.join(SCRATCH, "tests"), os.path.join(OUT, "tests"),
                    ignore=shutil.ignore_patterns("__pycache__"))
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("reference oracle installed at", p)

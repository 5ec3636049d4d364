#!/usr/bin/env python
"""Convert Gaussian-94 (.gbs, EMSL Basis Set Exchange) basis-set files into the compact JSON the
drop-in package ships (mcmurchie-davidson_b200/mmd/basis/<name>.json).

Basis-set tables are public data (EMSL BSE), not reference source code; the reference tree is not
present on the GPU box, so the sets are re-shipped in this neutral format.  `SP` shells are split
into an S and a P shell sharing exponents, as the reference's parser does (mmd/molecule.py:172-181).

Usage: python tools/convert_basis.py /root/reference/mmd/basis  mcmurchie-davidson_b200/mmd/basis
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mcmurchie-davidson_b200"))
from mmd._b200.basisio import parse_g94  # noqa: E402


def main(src, dst):
    os.makedirs(dst, exist_ok=True)
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".gbs"):
            continue
        data = parse_g94(os.path.join(src, fn))
        out = {"name": fn[:-4], "format": "mmd-b200-basis-1",
               "elements": {str(z): [[mom, [[e, c] for e, c in prims]] for mom, prims in shells]
                            for z, shells in sorted(data.items())}}
        with open(os.path.join(dst, fn[:-4] + ".json"), "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print(fn, "->", len(data), "elements")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

"""Deterministic synthetic geometries for the benchmark configurations.
Standalone (imports nothing from the package) so the golden-vector generator can load it by path.
All coordinates in Angstrom, written with 9 decimals; charge 0, multiplicity 1."""
import math

WATER = (("O", 0.0, -0.075791844, 0.0), ("H", 0.866811829, 0.601435779, 0.0), ("H", -0.866811829, 0.601435779, 0.0))


def _fmt(rows):
    return "\n" + "\n".join(["0 1"] + ["%s %.9f %.9f %.9f" % r for r in rows]) + "\n"


def water_cluster(nx, ny, nz, spacing=3.1):
    """(H2O)_n on a simple-cubic grid, loop order ix (outer) -> iy -> iz (inner)."""
    rows = []
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                for s, x, y, z in WATER:
                    rows.append((s, x + spacing * ix, y + spacing * iy, z + spacing * iz))
    return _fmt(rows)


def water():
    return water_cluster(1, 1, 1)


def benzene(r_c=1.397, r_h=2.481):
    rows = [("C", r_c * math.cos(math.pi / 3 * k), r_c * math.sin(math.pi / 3 * k), 0.0) for k in range(6)]
    rows += [("H", r_h * math.cos(math.pi / 3 * k), r_h * math.sin(math.pi / 3 * k), 0.0) for k in range(6)]
    return _fmt(rows)


def alkane(n, r_cc=1.54, r_ch=1.09):
    """All-trans C_n H_(2n+2) zig-zag in the xz plane, tetrahedral angle."""
    th = math.radians(109.4712206)
    s, c = math.sin(th / 2), math.cos(th / 2)
    carbons = [(i * r_cc * s, 0.0, (i % 2) * r_cc * c) for i in range(n)]
    rows = [("C",) + p for p in carbons]
    for i, (x, y, z) in enumerate(carbons):
        sg = -1.0 if i % 2 == 0 else 1.0
        rows.append(("H", x, +r_ch * s, z + sg * r_ch * c))
        rows.append(("H", x, -r_ch * s, z + sg * r_ch * c))
    x, y, z = carbons[0]
    rows.append(("H", x - r_ch * s, 0.0, z + r_ch * c))
    x, y, z = carbons[-1]
    sg = 1.0 if (n - 1) % 2 == 0 else -1.0
    rows.append(("H", x + r_ch * s, 0.0, z + sg * r_ch * c))
    return _fmt(rows)


def methane():
    return ("\n0 1\nC 0.000000 0.000000 0.000000\nH 0.626425042 -0.626425042 -0.626425042\n"
            "H 0.626425042 0.626425042 0.626425042\nH -0.626425042 0.626425042 -0.626425042\n"
            "H -0.626425042 -0.626425042 0.626425042\n")


CONFIGS = {
    "h2o_sto3g": (water, "sto-3g"),
    "h2o_ccpvdz": (water, "cc-pvdz"),
    "benzene_631gss": (benzene, "6-31gss"),
    "w8_ccpvdz": (lambda: water_cluster(2, 2, 2), "cc-pvdz"),
    "c20h42_631gs": (lambda: alkane(20), "6-31gs"),
    "w32_ccpvdz": (lambda: water_cluster(4, 4, 2), "cc-pvdz"),
}


def config(name):
    gen, basis = CONFIGS[name]
    return gen(), basis

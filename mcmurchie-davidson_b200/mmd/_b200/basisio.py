"""Basis-set I/O for the drop-in package: JSON tables shipped with the package, plus a
Gaussian-94 reader so user-supplied .gbs files keep working (reference: mmd/molecule.py:135-189).

Returned structure (same as the reference's `basis_data`):
    {atomic_number: [(momentum_letter, [(exponent, coefficient), ...]), ...]}
"""
import json
import os

SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
           "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr "
           "Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()


def atomic_number(sym):
    return SYMBOLS.index(str(sym))


def _num(tok):
    return float(tok.replace("D", "E").replace("d", "e"))


def parse_g94(path):
    """Reader for the EMSL Gaussian-94 format: blocks separated by '****', first line of a block
    'Sym 0', then shells 'L nprim scale' followed by nprim rows of exponent + coefficient(s)."""
    with open(path) as f:
        blocks = f.read().split("****")
    table = {}
    for block in blocks[1:]:
        rows = [ln.split() for ln in block.splitlines() if ln.strip() and not ln.lstrip().startswith("!")]
        if not rows:
            continue
        z = atomic_number(rows[0][0])
        shells = []
        k = 1
        while k < len(rows):
            mom, nprim = rows[k][0].upper(), int(rows[k][1])
            prim_rows = rows[k + 1:k + 1 + nprim]
            k += 1 + nprim
            if mom == "SP":
                shells.append(("S", [(_num(r[0]), _num(r[1])) for r in prim_rows]))
                shells.append(("P", [(_num(r[0]), _num(r[2])) for r in prim_rows]))
            else:
                shells.append((mom, [(_num(r[0]), _num(r[1])) for r in prim_rows]))
        table[z] = shells
    return table


def load_json(path):
    with open(path) as f:
        raw = json.load(f)
    return {int(z): [(mom, [(float(e), float(c)) for e, c in prims]) for mom, prims in shells]
            for z, shells in raw["elements"].items()}


def load_basis(name, search_dir):
    """`name` as the reference spells it ('sto-3g', '6-31gss', 'cc-pvdz', ...) or a path."""
    name = str(name)
    if os.path.isfile(name):
        return load_json(name) if name.endswith(".json") else parse_g94(name)
    stem = os.path.join(search_dir, name.lower())
    if os.path.isfile(stem + ".json"):
        return load_json(stem + ".json")
    if os.path.isfile(stem + ".gbs"):
        return parse_g94(stem + ".gbs")
    raise FileNotFoundError("basis set '%s' not found in %s" % (name, search_dir))

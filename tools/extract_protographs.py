"""One-off data extraction: dump every protograph table the reference defines
(ldpc/py/ldpc.py:59-660: five custom tables, IEEE 802.16e, IEEE 802.11n) into a
compact sparse JSON, by IMPORTING the reference module and calling
`code.assign_proto` -- no reference source text is copied.  Run in the build
container (needs /root/reference and `make -C oracle`):

    python tools/extract_protographs.py

Output: sparc_ldpc_b200/data/protographs.json
  { key: {"rows": Mp, "cols": Np, "edges": [[row, col, shift], ...]} }
  key = "<standard>|<rate>|<ptype or ->|<z or *>"   ('*' = any expansion factor)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

_, _, _, ldpc = rh.load_reference()

combos = []
for std, rates in (
    ("2_7_12_good", ["1/2"]),
    ("2_7_12_good_dc6", ["1/2", "0.45"]),
    ("2_5_12_good_threshold08", ["0.45", "3/8"]),
    ("2_7_12_bad", ["1/2"]),
):
    for r in rates:
        combos.append((std, r, None, None, 4))
for r in ("1/2", "5/6"):
    combos.append(("802.16", r, None, None, 4))
for r in ("2/3", "3/4"):
    for pt in ("A", "B"):
        combos.append(("802.16", r, pt, None, 4))
for z in (27, 54, 81):
    for r in ("1/2", "2/3", "3/4", "5/6"):
        combos.append(("802.11n", r, None, z, z))

out = {}
for std, rate, ptype, zkey, z in combos:
    c = ldpc.code(std, rate, z, ptype or "A")
    p = c.proto
    edges = [[int(r), int(col), int(p[r, col])] for r in range(p.shape[0]) for col in range(p.shape[1]) if p[r, col] != -1]
    key = "|".join([std, rate, ptype or "-", str(zkey) if zkey else "*"])
    out[key] = {"rows": int(p.shape[0]), "cols": int(p.shape[1]), "edges": edges}

dst = os.path.join(ROOT, "sparc_ldpc_b200", "data", "protographs.json")
with open(dst, "w") as f:
    f.write("{\n")
    items = list(out.items())
    for i, (k, v) in enumerate(items):
        f.write(' "%s": %s%s\n' % (k, json.dumps(v, separators=(",", ":")), "," if i + 1 < len(items) else ""))
    f.write("}\n")
print("wrote", dst, len(out), "protographs")

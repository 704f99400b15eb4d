"""Pack the gtinv coupling-coefficient tables into pypolymlp_b200/data/gtinv.pack.

    python tools/pack_gtinv.py [/root/reference/src/pypolymlp/cxx/src/polymlp]

Input: the reference's data files polymlp_gtinv_data_v{V}_order{O}.bin (little-endian tables of
l-combinations, m-combinations and coefficients; they are DATA, produced offline by the reference's
polyinv generator).  Output container: "PMGT", int32 n_tables, then per table
int32 version, int32 order, int32 n_bytes, raw table bytes.  Only the tables that exist are packed
(v2 orders 3-6 are absent from the reference checkout).
"""
import glob
import os
import re
import struct
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/pypolymlp/cxx/src/polymlp"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pypolymlp_b200", "data", "gtinv.pack")
tables = []
for path in sorted(glob.glob(os.path.join(src, "polymlp_gtinv_data_v*_order*.bin"))):
    m = re.search(r"_v(\d+)_order(\d+)\.bin$", path)
    with open(path, "rb") as f:
        tables.append((int(m.group(1)), int(m.group(2)), f.read()))
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "wb") as f:
    f.write(b"PMGT" + struct.pack("<i", len(tables)))
    for v, o, raw in tables:
        f.write(struct.pack("<iii", v, o, len(raw)) + raw)
print(f"packed {len(tables)} tables -> {os.path.normpath(out)} ({os.path.getsize(out)} bytes)")

"""Join an ncu report's SASS source page with nvdisasm's line info: samples, instructions and shared-memory
wavefronts per source line (and per line range) of one kernel.
usage: python tools/ncu_hotspots.py report.ncu-rep object.o kernel_substring [top]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: h.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared",
                                "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global", "stall_long_sb", "stall_short_sb",
                                "stall_barrier", "stall_mio", "stall_lg", "stall_wait", "stall_math", "stall_not_selected",
                                "stall_no_inst", "stall_selected", "stall_dispatch", "stall_branch_resolving")}
agg = {}
base = None
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = a
    off = a - base
    key = line_of.get(off, ((None, 0), ""))[0] or ("?", 0)
    d = agg.setdefault(key, {})
    for n, c in col.items():
        if n in ("Address", "Source"):
            continue
        try:
            d[n] = d.get(n, 0.0) + float(r[c].replace(",", "") or 0)
        except ValueError:
            pass
tot = {n: sum(d.get(n, 0) for d in agg.values()) for n in col if n not in ("Address", "Source")}
print("totals:", {k: int(v) for k, v in tot.items()})
src_cache = {}


def src(f, l):
    for root in ("sparc_ldpc_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.isfile(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:80] if 0 < l <= len(src_cache[p]) else ""
    return ""


print("%-18s %6s %6s %6s %6s %7s | stalls: long short bar mio lg wait math notsel noinst" % ("line", "samp%", "inst%", "wf%", "ideal%", "tag%"))
for key, d in sorted(agg.items(), key=lambda kv: -kv[1].get("# Samples", 0))[:top]:
    s = d.get("# Samples", 0) or 1
    print("%-18s %6.2f %6.2f %6.2f %6.2f %7.2f | %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f %4.0f | %s" % (
        "%s:%d" % key, 100 * d.get("# Samples", 0) / tot["# Samples"], 100 * d.get("Instructions Executed", 0) / tot["Instructions Executed"],
        100 * d.get("L1 Wavefronts Shared", 0) / max(tot["L1 Wavefronts Shared"], 1),
        100 * d.get("L1 Wavefronts Shared Ideal", 0) / max(tot["L1 Wavefronts Shared"], 1),
        100 * d.get("L1 Tag Requests Global", 0) / max(tot["L1 Tag Requests Global"], 1),
        *[100 * d.get(k, 0) / s for k in ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_mio", "stall_lg", "stall_wait",
                                          "stall_math", "stall_not_selected", "stall_no_inst")], src(*key)))

"""Per-source-line shares of a kernel's executed instructions / stall samples from an ncu report.

  python scripts/ncu_lines.py <report.ncu-rep> <cubin> <mangled kernel name> [top N] [--outer FILE]

The report's SASS page (ncu --page source --csv) gives counters per instruction address; nvdisasm -gi of the cubin maps
every address to its inline chain.  Two tables: by innermost (file, line), and by the line of the chain that lies in
--outer FILE (default qk_group.cuh: the kernel-level statement the instruction was inlined into)."""
import collections
import csv
import os
import re
import subprocess
import sys

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 30
outer_file = sys.argv[sys.argv.index("--outer") + 1] if "--outer" in sys.argv else "qk_group.cuh"

dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(kern + ":")][0]
addr2 = {}
chain = []
pending = []
for l in dis[start + 1:]:
    if re.match(r"^_Z\w+:", l):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if pending:
            chain = pending
            pending = []
        addr2[int(m.group(1), 16)] = (chain, m.group(2))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
inner = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
outer = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ix["Address"]], 16)
    if base is None:
        base = a
    ch, _ = addr2.get(a - base, ([("?", 0)], ""))
    w = float(r[ix["Instructions Executed"]] or 0)
    t = float(r[ix["Thread Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    k = ch[0]
    o = next((c for c in reversed(ch) if c[0] == outer_file), ch[-1])
    for agg, key in ((inner, k), (outer, o)):
        agg[key][0] += w
        agg[key][1] += t
        agg[key][2] += s
tw = sum(v[0] for v in inner.values())
ts = sum(v[2] for v in inner.values())
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "quokkasim_b200", "csrc")
src = {}


def text(f, l):
    if f not in src:
        p = os.path.join(srcdir, f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""


print("total warp instructions %.4e, samples %d" % (tw, ts))
for name, agg, col in (("by kernel-level line (%s), most instructions first" % outer_file, outer, 0),
                       ("by innermost line, most instructions first", inner, 0),
                       ("by innermost line, most stall samples first", inner, 2)):
    print("==== " + name)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][col])[:top]:
        print("%-16s %4d  inst %5.2f%%  samp %5.2f%%  lanes %4.1f | %s"
              % (k[0], k[1], 100 * v[0] / tw, 100 * v[2] / max(ts, 1), v[1] / v[0] if v[0] else 0, text(*k)))

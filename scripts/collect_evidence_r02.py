"""gpurun_out/ev2 (scratch, scripts/gpu_evidence_r02.sh) -> profiles/r02_* (tracked).

Bench lines, GPU test log, launch list, DRAM traffic per full-size step (keyed by the sha256 of the library it was
captured from: bench.py only quotes it while the build is the same), ncu --set full summaries with stall shares and
per-source-line shares, and the SASS evidence of the bulk asynchronous copies."""
import csv
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "ev2")
PR = os.path.join(ROOT, "profiles")
TAG = "r02"
LIB = os.path.join(ROOT, "quokkasim_b200", "libquokka_b200.so")


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


lib_sha = sha(LIB)
sys.path.insert(0, ROOT)
from quokkasim_b200 import build as qbuild  # noqa: E402

lib_id = qbuild.built_id(LIB)
for f in sorted(os.listdir(EV)):
    if (f.startswith("bench_") or f.startswith("sweep_")) and f.endswith(".json") and os.path.getsize(os.path.join(EV, f)):
        shutil.copy(os.path.join(EV, f), os.path.join(PR, "%s_%s" % (TAG, f)))
for f, g in (("pytest_gpu.log", "pytest_gpu.log"), ("launches_default.csv", "launches_default.csv"),
             ("lockstep_sweep.txt", "lockstep_sweep.txt"), ("gpu.txt", "gpu.txt")):
    if os.path.exists(os.path.join(EV, f)):
        shutil.copy(os.path.join(EV, f), os.path.join(PR, "%s_%s" % (TAG, g)))

# one table of the sweep
rows = []
for f in sorted(os.listdir(EV)):
    if f.startswith("sweep_c5_") and f.endswith(".json"):
        try:
            d = json.loads(open(os.path.join(EV, f)).readline())
            rows.append((d["config"]["replications_total"], d["config"]["logging"], d["value"], d["e2e"]["value"],
                         d["placement"]["state"], d["placement"]["threads_per_block"], d["placement"]["step_chunks"]))
        except Exception:
            pass
if rows:
    with open(os.path.join(PR, TAG + "_sweep_c5.txt"), "w") as out:
        out.write("C5 haulage (8 Sources, 64 Processes, 16 shared stocks, 8 Sinks), 1 GPU: replications -> events/s\n")
        out.write("%12s %6s %12s %12s  placement\n" % ("replications", "log", "device", "end-to-end"))
        for r in sorted(rows):
            out.write("%12d %6s %12.4e %12.4e  %s, %d threads per block, %d chunks\n" % r)

# DRAM traffic of one full-size step
shape = {"discrete_queue": (1048576, 10000), "haulage": (262144, 10000)}
traffic = {}
for wl in shape:
    for lg in ("ring", "off"):
        f = os.path.join(EV, "traffic_%s_%s.csv" % (wl, lg))
        if not os.path.exists(f):
            continue
        rws = [r for r in csv.reader(open(f, errors="replace")) if len(r) > 10]
        if len(rws) < 2:
            continue
        h = rws[0]
        m = {}
        for r in rws[1:]:
            try:
                m[r[h.index("Metric Name")]] = float(r[h.index("Metric Value")].replace(",", ""))
            except ValueError:
                pass
        if "dram__bytes_read.sum" not in m:
            continue
        kname = rws[1][h.index("Kernel Name")]
        traffic.setdefault(wl, {})[lg] = {
            "replications": shape[wl][0], "events": shape[wl][1], "lib_sha256": lib_sha, "lib_build_id": lib_id, "kernel": kname,
            "dram_bytes_per_step": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
            "kernel_ns_under_ncu": m.get("gpu__time_duration.sum"),
            "warps_active_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": m.get("smsp__inst_executed.sum"),
            "active_lanes_per_instruction": m.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "l2_sector_hit_pct": m.get("lts__t_sector_hit_rate.pct"),
            "registers_per_thread": m.get("launch__registers_per_thread"), "block_size": m.get("launch__block_size"),
            "source": "ncu --metrics ... -k regex:qk_step_kernel -s 2 -c 1, QK_SLICE=0 (one plain launch = one step), "
                      "scripts/gpu_evidence_r02.sh",
        }
json.dump(traffic, open(os.path.join(PR, TAG + "_traffic.json"), "w"), indent=1)

# ncu --set full summaries
tmp = "/tmp/qk_cub_ev"
shutil.rmtree(tmp, ignore_errors=True)
os.makedirs(tmp)
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
for f in sorted(os.listdir(EV)):
    if not (f.startswith("prof_") and f.endswith(".ncu-rep")):
        continue
    tag = f[5:-8]
    rep = os.path.join(EV, f)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], stdout=subprocess.PIPE,
                         text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rws = list(csv.reader(raw.splitlines()))
    hdr, vals = rws[0], rws[2]
    kname = vals[hdr.index("Kernel Name")]
    stalls = []
    for h, v in zip(hdr, vals):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                stalls.append((float(v), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    txt = "%s\ncommand: see scripts/gpu_evidence_r02.sh (full %s); library sha256 %s\n\n%s" % (kname, tag, lib_sha, out)
    txt += "\nwarp stall samples (share of all samples)\n"
    for v, h in sorted(stalls, reverse=True)[:10]:
        txt += "  %-28s %5.1f %%\n" % (h, 100 * v / tot)
    mk = re.search(r"qk_step_kernel<(?:\(bool\))?(\d), (?:\(int\))?(-?\d+), (?:\(int\))?(\d)>", kname)
    if mk:
        sm = int(mk.group(2))
        mangled = "_Z14qk_step_kernelILb%sELi%sELi%sEEv7KParams" % (mk.group(1), ("n%d" % -sm) if sm < 0 else str(sm),
                                                                     mk.group(3))
        cubin = os.path.join(tmp, "qk_step_l%sf%s.sm_100a.cubin" % (mk.group(1), mk.group(3)))
        if os.path.exists(cubin):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, cubin, mangled, "36",
                                "--outer", "qk_kernels.cuh"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            txt += "\nper-source-line shares (scripts/ncu_lines.py)\n" + r.stdout
    open(os.path.join(PR, "%s_ncu_%s.txt" % (TAG, tag)), "w").write(txt)

# SASS evidence: bulk asynchronous copies (TMA engine) in the step kernels
sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
cur = None
counts = {}
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        counts[cur] = {"UTMALDG": 0, "UTMASTG": 0, "UBLKCP": 0, "SYNCS": 0, "LDGSTS": 0, "REDUX": 0, "REDG": 0}
        continue
    if cur:
        for k in counts[cur]:
            if re.search(r"\b%s\b" % k, l) or (k + ".") in l:
                counts[cur][k] += 1
with open(os.path.join(PR, TAG + "_sass_bulk_copies.txt"), "w") as out:
    out.write("cuobjdump -sass quokkasim_b200/libquokka_b200.so (sha256 %s): instruction counts per kernel\n" % lib_sha)
    out.write("UTMALDG / UTMASTG = cp.async.bulk.tensor.2d (TMA tile of a state plane: every slot x the block's replications) ; "
              "UBLKCP = cp.async.bulk (TMA engine, one row / the model tables) ; SYNCS = mbarrier operations ; "
              "LDGSTS = cp.async (per-thread) ; REDUX = warp reduction ; REDG = fire-and-forget atomics (statistics planes)\n\n")
    for k in sorted(counts):
        if "qk_step_kernel" in k or "qk_group_kernel" in k or "qk_phased_kernel" in k:
            c = counts[k]
            out.write("%-60s UTMALDG %2d  UTMASTG %2d  UBLKCP %3d  SYNCS %3d  LDGSTS %3d  REDUX %3d  REDG %3d\n" % (
                k, c["UTMALDG"], c["UTMASTG"], c["UBLKCP"], c["SYNCS"], c["LDGSTS"], c["REDUX"], c["REDG"]))
print("profiles/ updated:", sorted(f for f in os.listdir(PR) if f.startswith(TAG)))

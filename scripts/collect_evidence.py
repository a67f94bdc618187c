"""gpurun_out/ev (scratch) -> profiles/ (tracked): bench lines, launch list, DRAM traffic, ncu summaries, hot lines."""
import csv, json, os, re, shutil, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "ev")
PR = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def cp(src, dst):
    if os.path.exists(os.path.join(EV, src)):
        shutil.copy(os.path.join(EV, src), os.path.join(PR, dst))


cp("bench_full_ring.json", TAG + "_bench_full_ring.json")
cp("bench_full_off.json", TAG + "_bench_full_off.json")
cp("bench_reference.json", TAG + "_bench_reference.json")
cp("launches_default.csv", TAG + "_launches_default.csv")
cp("pytest_gpu.log", TAG + "_pytest_gpu.log")

# DRAM traffic of the full-size step launch
traffic = {"discrete_queue": {}}
for lg in ("ring", "off"):
    f = os.path.join(EV, "traffic_full_%s.csv" % lg)
    if not os.path.exists(f):
        continue
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]
    m = {r[h.index("Metric Name")]: float(r[h.index("Metric Value")].replace(",", "")) for r in rows[1:]}
    traffic["discrete_queue"][lg] = {
        "replications": 1048576, "events": 10000,
        "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
        "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
        "kernel_ns_under_ncu": m["gpu__time_duration.sum"],
        "warps_active_pct": m["sm__warps_active.avg.pct_of_peak_sustained_active"],
        "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "warp_instructions": m["smsp__inst_executed.sum"],
        "active_lanes_per_instruction": m["smsp__thread_inst_executed_per_inst_executed.ratio"],
        "registers_per_thread": m["launch__registers_per_thread"], "block_size": m["launch__block_size"],
        "blocks_per_sm_limit_registers": m["launch__occupancy_limit_registers"],
        "blocks_per_sm_limit_shared_mem": m["launch__occupancy_limit_shared_mem"],
        "source": "ncu --metrics ... -k regex:qk_step_kernel -s 2 -c 1 python bench.py --steps 1 --warmup 0 "
                  "--no-cpu-baseline --e2e-steps 1 --log %s (scripts/gpu_evidence.sh)" % lg,
    }
json.dump(traffic, open(os.path.join(PR, TAG + "_traffic.json"), "w"), indent=1)

# ncu --set full summaries + per-region / per-line shares
names = {"off": "stats_only", "ring": "logging"}
lib = os.path.join(ROOT, "quokkasim_b200", "libquokka_b200.so")
tmp = "/tmp/qk_cub"
shutil.rmtree(tmp, ignore_errors=True)
os.makedirs(tmp)
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
dis = {}
for f in os.listdir(tmp):
    if f.endswith(".cubin") and "qk_step" in f:
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
        dis[f] = out.split("\n")
src = open(os.path.join(ROOT, "quokkasim_b200", "csrc", "qk_kernels.cuh")).read().split("\n")
marks = []
for i, l in enumerate(src):
    m = re.match(r"__device__ __forceinline__ \S+ (qk_\w+)\(", l) or re.match(r"__global__ void .* (qk_step_kernel)\(", l)
    if m:
        marks.append((i + 1, m.group(1)))
    if "for (int it = 0; it < QK_SCAN_ITERS" in l:
        marks.append((i + 1, "LOOP.scan"))
    if "/* pop the scheduler queue" in l:
        marks.append((i + 1, "LOOP.select"))
    if "if (!QK_ANY_SYNC(!done || rem != 0)) break" in l:
        marks.append((i + 1, "LOOP.refill+update"))
    if "/* step_until(t) leaves the clock" in l:
        marks.append((i + 1, "epilogue"))


def region(f, l):
    if f != "qk_kernels.cuh":
        return f
    r = "?"
    for ln, nm in marks:
        if ln <= l:
            r = nm
    return r


for lg, nm in names.items():
    rep = os.path.join(EV, "prof_small_%s.ncu-rep" % lg)
    if not os.path.exists(rep):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], stdout=subprocess.PIPE,
                         text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, vals = rows[0], rows[2]
    stalls = []
    for h, v in zip(hdr, vals):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                stalls.append((float(v), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    out += "\nwarp stall samples (share of all samples)\n"
    for v, h in sorted(stalls, reverse=True)[:10]:
        out += "  %-28s %5.1f %%\n" % (h, 100 * v / tot)
    open(os.path.join(PR, "%s_ncu_step_kernel_%s.txt" % (TAG, nm)), "w").write(out)
    # source page
    sp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    srows = list(csv.reader(sp.splitlines()))
    kname = srows[0][1]
    mk = re.search(r"qk_step_kernel<\(bool\)(\d), \(int\)(-?\d+), \(int\)(\d)>", kname)
    mangled = "_Z14qk_step_kernelILb%sELi%sELi%sEEv7KParams" % (mk.group(1), mk.group(2), mk.group(3))
    lines = None
    for f, ls in dis.items():
        if any(l.startswith(mangled + ":") for l in ls):
            lines = ls
    if lines is None:
        continue
    start = [i for i, l in enumerate(lines) if l.startswith(mangled + ":")][0]
    addr2line, cur = {}, None
    for l in lines[start:]:
        if l.startswith("//-----") and "text." in l and mangled not in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur is not None:
            addr2line[int(m.group(1), 16)] = cur
    h2 = srows[1]
    ix = {h: i for i, h in enumerate(h2)}
    agg_l = collections.defaultdict(lambda: [0, 0, 0])
    agg_r = collections.defaultdict(lambda: [0, 0, 0])
    base = None
    for r in srows[2:]:
        a = int(r[ix["Address"]], 16)
        if base is None:
            base = a
        f, l = addr2line.get(a - base, ("?", 0))
        w = float(r[ix["Instructions Executed"]] or 0)
        t = float(r[ix["Thread Instructions Executed"]] or 0)
        s = float(r[ix["# Samples"]] or 0)
        for agg, key in ((agg_l, (f, l)), (agg_r, region(f, l))):
            agg[key][0] += w
            agg[key][1] += t
            agg[key][2] += s
    tw = sum(v[0] for v in agg_r.values())
    tt = sum(v[1] for v in agg_r.values())
    ts = sum(v[2] for v in agg_r.values()) or 1
    txt = "%s\nwarp instructions %.3e, thread instructions %.3e, samples %d\n\nby code region\n" % (kname, tw, tt, ts)
    for k, v in sorted(agg_r.items(), key=lambda x: -x[1][0])[:28]:
        txt += "  %-28s inst %5.2f %%  thread-inst %5.2f %%  samples %5.2f %%  active lanes %4.1f\n" % (
            k, 100 * v[0] / tw, 100 * v[1] / tt, 100 * v[2] / ts, v[1] / v[0] if v[0] else 0)
    txt += "\nhottest source lines (by stall samples)\n"
    for (f, l), v in sorted(agg_l.items(), key=lambda x: -x[1][2])[:40]:
        text = src[l - 1].strip()[:90] if f == "qk_kernels.cuh" and 0 < l <= len(src) else ""
        txt += "  %-16s %4d  inst %5.2f %%  samples %5.2f %%  lanes %4.1f | %s\n" % (
            f, l, 100 * v[0] / tw, 100 * v[2] / ts, v[1] / v[0] if v[0] else 0, text)
    open(os.path.join(PR, "%s_hot_lines_%s.txt" % (TAG, nm)), "w").write(txt)
print("profiles/ updated:", sorted(f for f in os.listdir(PR) if f.startswith(TAG)))

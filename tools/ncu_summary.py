"""Turn ncu outputs under gpurun_out/ into the text summaries committed under profiles/.
usage: python tools/ncu_summary.py <tag>     (reads gpurun_out/launches.csv, prof_*.ncu-rep)"""
import collections, csv, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = "gpurun_out"
os.makedirs("profiles", exist_ok=True)

def launch_list():
    p = os.path.join(src, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if l.startswith('"')]
    rd = csv.reader(lines); hdr = next(rd)
    ik, iv, iu, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.defaultdict(lambda: [0, 0.0]); n = 0
    for r in rd:
        v = float(r[iv].replace(",", "")); u = r[iu]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[re.sub(r"\(.*", "", r[ik])][0] += 1; agg[re.sub(r"\(.*", "", r[ik])][1] += v; n += 1
    tot = sum(v[1] for v in agg.values())
    with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {n} launches, {tot:.1f} us total\n")
        f.write("# share   total_us   launches  kernel\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{us / tot * 100:6.2f}%  {us:10.1f}  {c:5d}  {k}\n")

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio"]

def full(name):
    p = os.path.join(src, name + ".ncu-rep")
    if not os.path.exists(p):
        return
    out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = [(w, [i for i, h in enumerate(hdr) if h.endswith(w)]) for w in WANT]
    with open(f"profiles/{tag}_{name}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({name}.ncu-rep, {len(rows)} launches)\n")
        for r in rows:
            f.write("----\n")
            for w, ii in idx:
                if ii:
                    f.write(f"{w:82s} {r[ii[0]][:60]} {units[ii[0]]}\n")

launch_list()
for n in sorted(os.listdir(src)):
    if n.endswith(".ncu-rep"):
        full(n[:-8])
print(os.listdir("profiles"))

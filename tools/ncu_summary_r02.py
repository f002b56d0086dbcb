"""Turn the CSVs written by tools/gpu_ncu_r02.sh (gpurun_out/<tag>_launches.csv, <tag>_full_*.csv) into the text summaries and
the traffic table committed under profiles/.

    python tools/ncu_summary_r02.py [tag]          ->  profiles/<tag>_launches_summary.txt, profiles/<tag>_full_<name>.txt,
                                                      profiles/<tag>_traffic.json (DRAM bytes per launch, per kernel family)"""
import collections, csv, glob, json, os, re, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = "gpurun_out"
os.makedirs("profiles", exist_ok=True)

FAMILY = [("gemm_nt_kernel", "ns_gemm_nt"), ("gemm_tn_kernel", "ns_gemm_tn"), ("attn_fwd_db", "ns_attention_fwd"),
          ("attn_bwd_fused", "ns_attention_bwd_ws"), ("ln_fwd", "ns_layernorm_fwd"), ("ln_bwd", "ns_layernorm_bwd"),
          ("aug_btc", "ns_aug_pass"), ("ce_", "ns_cross_entropy"), ("lora_da_kernel", "ns_lora_da"), ("dropout_bits", "ns_dropout_bits"), ("lora_bwd_b_kernel", "ns_lora_bwd_b"),
          ("cross_absorbed", "ns_cross_attention_absorbed")]


def short(name):
    return re.sub(r"\(.*", "", name).replace("ns::", "")


def launch_list():
    p = os.path.join(src, f"{tag}_launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if l.startswith('"')]
    rd = csv.reader(lines); hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0]); n = 0
    for r in rd:
        v = float(r[iv].replace(",", "")); u = r[iu]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = short(r[ik]); agg[k][0] += 1; agg[k][1] += v; n += 1
    tot = sum(v[1] for v in agg.values())
    with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none over one EAGER training step of bench.py (lora_dropout 0.05):\n"
                f"# {n} launches (a window of about one step), {tot:.1f} us total -- cold-cache and serialised, so compare SHARES with kernel_shares of the bench line\n")
        f.write("# share   total_us   launches  kernel\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{us / tot * 100:6.2f}%  {us:10.1f}  {c:5d}  {k}\n")


WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio"]


def full(path, traffic):
    name = os.path.basename(path)[len(tag) + 6:-4]
    rd = list(csv.reader(open(path)))
    rd = [r for r in rd if r]
    if len(rd) < 3:
        return
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = [(w, [i for i, h in enumerate(hdr) if h == w or h.endswith(w)]) for w in WANT]
    ik = hdr.index("Kernel Name")
    col = lambda r, w: next((r[i] for i, h in enumerate(hdr) if h == w), None)
    with open(f"profiles/{tag}_full_{name}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, one launch per kernel at the bench shapes (tools/gpu_ncu_r02.sh; {len(rows)} launches)\n")
        for r in rows:
            f.write("----\n")
            for w, ii in idx:
                if ii:
                    f.write(f"{w:84s} {short(r[ii[0]])[:110] if w == 'Kernel Name' else r[ii[0]][:40]} {units[ii[0]]}\n")
            try:
                rdb = float(col(r, "dram__bytes_read.sum").replace(",", "")); wrb = float(col(r, "dram__bytes_write.sum").replace(",", ""))
                mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                ur = units[hdr.index("dram__bytes_read.sum")]; uw = units[hdr.index("dram__bytes_write.sum")]
                tot = rdb * mult.get(ur, 1.0) + wrb * mult.get(uw, 1.0)
                for pat, fam in FAMILY:
                    if pat in r[ik]:
                        t = traffic.setdefault(fam, {"launches": 0, "dram_bytes": 0.0, "kernels": []})
                        t["launches"] += 1; t["dram_bytes"] += tot; t["kernels"].append({"kernel": short(r[ik])[:90], "grid": col(r, "Grid Size"), "dram_bytes": tot})
                        break
            except Exception:
                pass


launch_list()
traffic = {}
for p in sorted(glob.glob(os.path.join(src, f"{tag}_full_*.csv"))):
    full(p, traffic if "decode" not in os.path.basename(p) else {})     # the traffic table is about the training-step shapes
for fam, t in traffic.items():
    t["dram_bytes_per_launch_avg"] = t["dram_bytes"] / max(t["launches"], 1)
if traffic:
    json.dump({"source": f"ncu --set full, tools/gpu_ncu_r02.sh (tag {tag}): dram__bytes_read.sum + dram__bytes_write.sum per launch", **traffic},
              open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(sorted(os.listdir("profiles")))

"""Turns ncu reports (.ncu-rep) into the summaries committed under profiles/: one CSV row per captured launch with the
metrics DESIGN.md quotes, and the DRAM traffic per launch of the two edge kernel families for bench.py's roofline block.

usage: python tools/ncu_summary.py --out profiles/r2_g_kernels_ncu_summary_R4096.csv --traffic profiles/r2_edge_traffic.json \
           --replicas 4096 label1=path1.ncu-rep label2=path2.ncu-rep ...   (or the raw-page CSV exports of the reports)"""
import argparse
import csv
import io
import json
import subprocess

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
MULT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

ap = argparse.ArgumentParser()
ap.add_argument("--out", required=True)
ap.add_argument("--traffic", default=None)
ap.add_argument("--replicas", type=int, default=4096)
ap.add_argument("--note", default="")
ap.add_argument("reports", nargs="+")
a = ap.parse_args()

rows_out, traffic = [], {"edge_fwd": [], "edge_bwd": []}
for spec in a.reports:
    label, _, path = spec.partition("=")
    if path.endswith(".csv"):      # already exported with `ncu -i report --page raw --csv` (e.g. on the GPU box)
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    for r in body:
        rows_out.append({"capture": label, **{w: r[i] for w, i in idx}})
        if "dram__bytes_read.sum" in hdr and label.startswith("full"):
            rd, wr, kn = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
            if r[rd] and r[wr]:
                t = float(r[rd]) * MULT[units[rd]] + float(r[wr]) * MULT[units[wr]]
                if "edge_fwd" in r[kn]:
                    traffic["edge_fwd"].append(t)
                elif "edge_bwd" in r[kn]:
                    traffic["edge_bwd"].append(t)
with open(a.out, "w") as f:
    f.write('"# ' + a.note.replace('"', "'") + '"\n')
    w = csv.writer(f)
    cols = ["capture"] + WANT
    w.writerow(cols)
    for o in rows_out:
        w.writerow([o.get(c, "") for c in cols])
if a.traffic:
    j = {"replicas": a.replicas, "atoms": 50,
         "source": f"{a.out} (ncu --set full captures: dram__bytes_read.sum + dram__bytes_write.sum per launch)"}
    for k, v in traffic.items():
        if v:
            j[k] = sum(v) / len(v)
    j["per_launch"] = traffic
    json.dump(j, open(a.traffic, "w"), indent=1)
print(f"{len(rows_out)} launches -> {a.out}")

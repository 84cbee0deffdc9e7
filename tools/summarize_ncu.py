"""Summarise an ncu report: python tools/summarize_ncu.py report.ncu-rep out_prefix [kernel-launch-index]
Writes <out_prefix>.md (key metrics + stall breakdown) and prints a JSON dict with dram bytes per launch."""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"
md = [f"# ncu summary: {rep}", ""]
summary = []
for li, val in enumerate(rows[2:]):
    d = dict(zip(hdr, val)); u = dict(zip(hdr, units))
    name = d.get("Kernel Name", "?")
    md += [f"## launch {li}: `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in d: md.append(f"| {k} | {d[k]} | {u[k]} |")
    st = sorted(((float(v), k[len(STALL):].replace('_per_issue_active.ratio', '')) for k, v in d.items() if k.startswith(STALL) and v not in ("", "n/a")), reverse=True)
    md += ["", "warp-issue stall reasons (average warps stalled per issue-active cycle):", ""] + [f"* {n}: {v:.3f}" for v, n in st[:8]] + [""]
    def num(k):
        try: return float(d[k])
        except Exception: return None
    def to_bytes(k):
        v = num(k)
        if v is None: return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[k], 1)
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    summary.append({"kernel": name, "time_ms": num("gpu__time_duration.sum"), "time_unit": u.get("gpu__time_duration.sum"),
                    "dram_bytes_per_launch": None if rd is None else rd + wr, "inst": num("smsp__inst_executed.sum"),
                    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")})
open(out + ".md", "w").write("\n".join(md))
print(json.dumps(summary))

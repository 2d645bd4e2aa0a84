"""Summarises an .ncu-rep (read here with `ncu -i`, no GPU needed) into profiles/<name>.md:
key raw metrics of each captured kernel and the hottest SASS lines with their stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of `{rep}`", "", note, ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines += [f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}", "",
                  "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u[k]} |")
        st = sorted(((float(v), k[len(STALLS):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                     if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        lines += ["", "warp stall reasons (warps stalled per issue-active cycle): " +
                  ", ".join(f"{n} {v:.2f}" for v, n in st[:8]), ""]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    h = None
    data = []
    for r in srows:
        if r and r[0] == "Address":
            h = r
            continue
        if h and len(r) == len(h):
            data.append(r)
    if h:
        col = {n: i for i, n in enumerate(h)}
        tot = sum(int(r[col["# Samples"]]) for r in data) or 1
        top = sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:25]
        lines += ["## hottest SASS instructions (share of warp-stall samples)", "", "| % samples | executed | SASS | top stall |", "|---|---|---|---|"]
        stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        for r in top:
            sv = sorted(((int(r[col[n]]), n) for n in stall_cols), reverse=True)[0]
            lines.append(f"| {100 * int(r[col['# Samples']]) / tot:.2f} | {r[col['Instructions Executed']]} | `{r[col['Source']].strip()}` | {sv[1]} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()

"""Summarises the CSV exports of tools/ncu_export.py (made on the GPU box) into profiles/<name>.md."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    base, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(open(base + "_raw.csv")))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary ({base}_raw.csv / _source.csv, exported on the GPU box by tools/ncu_export.py)", "", note, ""]
    for vals in rows[2:]:
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
        lines += [f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}", "", "| metric | value | unit |", "|---|---|---|"]
        lines += [f"| {k} | {d[k]} | {u[k]} |" for k in KEYS if k in d]
        st = sorted(((float(v), k[len(STALLS):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                     if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        lines += ["", "warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:9]), ""]
    try:
        srows = list(csv.reader(open(base + "_source.csv")))
    except OSError:
        srows = []
    if srows:
        h = srows[0]
        col = {n: i for i, n in enumerate(h)}
        tot = int(srows[1][col["total_samples"]])
        lines += [f"## hottest SASS instructions (share of {tot} warp-stall samples over {srows[1][col['total_lines']]} SASS lines)", "",
                  "| % samples | executed | SASS | top stall |", "|---|---|---|---|"]
        stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        for r in srows[1:31]:
            sv = sorted(((int(r[col[n]] or 0), n) for n in stall_cols), reverse=True)[0]
            lines.append(f"| {100 * int(r[col['# Samples']]) / tot:.2f} | {r[col['Instructions Executed']]} | `{r[col['Source']].strip()}` | {sv[1]} |")
        agg = {}
        for r in srows[1:]:
            for n in stall_cols:
                agg[n] = agg.get(n, 0) + int(r[col[n]] or 0)
        top = sum(int(r[col["# Samples"]]) for r in srows[1:])
        lines += ["", f"stall mix of the {len(srows) - 1} hottest lines ({100 * top / tot:.0f} % of all samples): " +
                  ", ".join(f"{n[6:]} {100 * v / max(1, sum(agg.values())):.0f} %" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8])]
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: writes the text summary and the per-launch DRAM traffic JSON
that bench.py's roofline.traffic reads.   usage: ncu_full_summary.py REPORT.ncu-rep OUT.txt OUT_traffic.json "header" """
import csv, io, json, re, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sectors_op_red.sum",
           "lts__t_sectors_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"]
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def main(rep, out_txt, out_json, header):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    traffic, lines = {}, [f"# {header}", "# per-launch values; one mapping iteration = 240,000 samples per launch"]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"--- {name}")
        for m in METRICS:
            if m not in hdr:
                continue
            i = hdr.index(m)
            lines.append(f"   {m:90s} {r[i]:>14s} {units[i]}")
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("unnamed>::", "").strip()
        b = sum(float(r[hdr.index(m)]) * UNIT.get(units[hdr.index(m)], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[short] = int(b)
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_json, "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])

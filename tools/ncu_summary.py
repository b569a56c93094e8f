#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of one kernel into profiles/:

    python tools/ncu_summary.py gpurun_out/r1s_score_tiled.ncu-rep profiles/r1s_k_score_tiled_ncu_full.md [traffic.json]

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv` / `--page source --csv`.  Writes a
markdown table of the metrics the design discussion uses, the share of executed instructions and
stall samples per code region, and (optionally) the per-launch DRAM traffic that bench.py reports as
`roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, md = sys.argv[1], sys.argv[2]
    traffic_json = sys.argv[3] if len(sys.argv) > 3 else None
    rows = ncu_csv(rep, "raw")
    hdr, units, launches = rows[0], rows[1], rows[2:]
    name = launches[0][hdr.index("Kernel Name")].split("(")[0]
    lines = ["# ncu --set full: %s" % name, "", "Source: `%s` (%d launch(es) captured; values of the first)." % (rep, len(launches)), "",
             "| metric | unit | value |", "|---|---|---|"]
    first = launches[0]
    vals = {}
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            vals[m] = (first[i], units[i])
            lines.append("| %s | %s | %s |" % (m, units[i], first[i]))
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(first[i])
            except ValueError:
                continue
            if v >= 0.3:
                lines.append("| %s | warps / issue | %.3f |" % (h.replace("smsp__average_warps_issue_stalled_", "stall: ").replace("_per_issue_active.ratio", ""), v))

    def to_bytes(v, u):
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    if traffic_json and "dram__bytes_read.sum" in vals:
        rd = to_bytes(*vals["dram__bytes_read.sum"])
        wr = to_bytes(*vals["dram__bytes_write.sum"])
        json.dump({"kernel": name, "source": rep.split("/")[-1], "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "traffic_bytes_per_launch": rd + wr}, open(traffic_json, "w"))
        lines += ["", "DRAM traffic per launch: %.3f MB read + %.3f MB written." % (rd / 1e6, wr / 1e6)]

    # code regions: instructions executed / stall samples per 2 KiB of SASS
    try:
        src = ncu_csv(rep, "source")
        h2 = src[1]
        ia, ie, ism = h2.index("Address"), h2.index("Instructions Executed"), h2.index("# Samples")
        data = [(int(r[ia], 16), int(r[ie]), int(r[ism])) for r in src[2:] if len(r) > ie]
        base = data[0][0]
        tot_i = float(sum(d[1] for d in data)) or 1.0
        tot_s = float(sum(d[2] for d in data)) or 1.0
        reg = {}
        for a, e, sm in data:
            k = (a - base) // 0x800
            reg.setdefault(k, [0, 0])
            reg[k][0] += e
            reg[k][1] += sm
        lines += ["", "Executed instructions and stall samples per 2 KiB SASS region (>= 1 % of either):", "",
                  "| SASS offset | instructions % | samples % |", "|---|---|---|"]
        for k in sorted(reg):
            pi, ps = 100 * reg[k][0] / tot_i, 100 * reg[k][1] / tot_s
            if pi >= 1 or ps >= 1:
                lines.append("| 0x%05x | %.1f | %.1f |" % (k * 0x800, pi, ps))
    except Exception as ex:                                  # source page needs -lineinfo + --import-source
        lines += ["", "(no source page: %s)" % ex]
    open(md, "w").write("\n".join(lines) + "\n")
    print("wrote", md)


if __name__ == "__main__":
    main()

"""Summarise an .ncu-rep (ncu --set full capture) into profiles/: per-kernel key metrics + DRAM bytes per launch.

usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_ncu_full_tc5_summary.txt [profiles/r01_ncu_dram_traffic.json]
"""
import csv, io, json, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines, traffic = [], {}
    for r in data:
        name = re.sub(r"<.*$", "", r[col["Kernel Name"]].split("(")[0].replace("void ", "")).strip()
        lines.append("== %s" % name)
        for k in KEYS:
            if k in col:
                lines.append("   %-78s %s %s" % (k, r[col[k]], units[col[k]]))
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)
        traffic.setdefault(name, []).append(b)
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3:
        try:
            cur = json.load(open(sys.argv[3]))
        except FileNotFoundError:
            cur = {}
        for k, v in traffic.items():
            cur[k] = sum(v) / len(v)
        json.dump(cur, open(sys.argv[3], "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()

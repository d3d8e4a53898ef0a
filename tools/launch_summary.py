"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...) per kernel: launches, total ms, share.

usage: python tools/launch_summary.py gpurun_out/x.csv profiles/rNN_ncu_launches_summary.txt "<the command that was profiled>"
"""
import csv, re, sys


def main():
    src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.reader(lines))
    col = {h: i for i, h in enumerate(rows[0])}
    agg = {}
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]])[:60]
        v = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[col["Metric Unit"]], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    txt = [cmd, "%-60s %5s %12s %7s %12s" % ("kernel", "n", "total ms", "share", "ms/launch")]
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        txt.append("%-60s %5d %12.3f %6.1f%% %12.4f" % (name, n, ms, 100 * ms / tot, ms / n))
    open(out, "w").write("\n".join(txt) + "\n")
    print("\n".join(txt[:14]))


if __name__ == "__main__":
    main()

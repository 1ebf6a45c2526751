"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python
bench.py --steps 1 --warmup 3 --no-extras --no-graph --sync-wgrad --serial-branches`) per kernel for ONE steady-state training
step: the launches from the last-but-K `ingest` kernel to the next one (K chosen so that the step is the timed one, not the
per-kernel event pass that follows it).

usage: python tools/ncu_by_kernel.py launches.csv [out.txt] [--step -3] [--csv step_launches.csv]"""
import collections
import csv
import sys


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "")
    cut = name.find("(")
    return (name if cut < 0 else name[:cut])[:60]


def main():
    path = sys.argv[1]
    out_path = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    step = -3
    if "--step" in sys.argv:
        step = int(sys.argv[sys.argv.index("--step") + 1])
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us))
    starts = [i for i, (n, _) in enumerate(rows) if "ingest" in n]
    a = starts[step]
    b = starts[step + 1] if step + 1 < 0 else len(rows)
    sel = rows[a:b]
    tot = collections.OrderedDict()
    for n, us in sel:
        k = short(n)
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += us
    total = sum(us for _, us in sel)
    lines_out = ["# one steady-state training step: %d consecutive launches, serialized total %.2f ms (ncu --metrics gpu__time_duration.sum "
                 "--clock-control none; eager single-stream issue: --no-graph --sync-wgrad --serial-branches)" % (len(sel), total / 1e3)]
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines_out.append("%-62s %4d %9.3f ms %6.1f%%  avg %7.1f us" % (k, n, us / 1e3, 100.0 * us / total, us / n))
    if "--csv" in sys.argv:  # compact per-launch list of the selected step
        with open(sys.argv[sys.argv.index("--csv") + 1], "w") as f:
            f.write("index,kernel,duration_ns\n")
            for i, (n, us) in enumerate(sel):
                f.write('%d,"%s",%d\n' % (i, short(n), int(round(us * 1e3))))
    txt = "\n".join(lines_out)
    print(txt)
    if out_path:
        open(out_path, "w").write(txt + "\n")


if __name__ == "__main__":
    main()

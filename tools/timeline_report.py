"""Reads the chrome trace written by `python bench.py --timeline FILE` (CUPTI kernel records of two replayed training steps)
and reports, for the LAST step in it: wall span, GPU-busy time (union of kernel intervals), idle time, the time during
which >= 2 kernels ran concurrently, per-stream busy time, the idle gaps by the kernel that follows them, and per-kernel
totals.  Diagnostic only: CUPTI records perturb the schedule slightly; nothing here is a bench value.

usage: python tools/timeline_report.py gpurun_out/timeline.json.trace.json [out.txt]"""
import collections
import json
import sys


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "")
    return name.split("(")[0][:48]


def main():
    tr = json.load(open(sys.argv[1]))
    ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    # split into steps at the ingest kernel (first kernel of a step)
    starts = [i for i, e in enumerate(ev) if "ingest" in e["name"]]
    if len(starts) >= 2:
        ev = ev[starts[-1]:]
    out = []
    t0 = ev[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in ev)
    span = t1 - t0
    # sweep
    pts = []
    for e in ev:
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    busy = multi = 0.0
    depth, last = 0, t0
    for t, d in pts:
        if depth >= 1:
            busy += t - last
        if depth >= 2:
            multi += t - last
        depth += d
        last = t
    out.append("last step: %d kernels, span %.3f ms, busy (>= 1 kernel) %.3f ms, idle %.3f ms, >= 2 kernels concurrently %.3f ms, sum of durations %.3f ms"
               % (len(ev), span / 1e3, busy / 1e3, (span - busy) / 1e3, multi / 1e3, sum(e["dur"] for e in ev) / 1e3))
    by_stream = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        s = e.get("args", {}).get("stream")
        by_stream[s][0] += 1
        by_stream[s][1] += e["dur"]
    out.append("per stream (kernels, sum of durations ms):")
    for s, (n, d) in sorted(by_stream.items(), key=lambda kv: -kv[1][1]):
        out.append("  stream %-6s %5d  %8.3f" % (s, n, d / 1e3))
    # idle gaps: intervals with depth 0, attributed to the kernel that ends the gap
    gaps = collections.defaultdict(lambda: [0, 0.0])
    end = t0
    hist = collections.Counter()
    for e in ev:
        if e["ts"] > end:
            g = e["ts"] - end
            k = short(e["name"])
            gaps[k][0] += 1
            gaps[k][1] += g
            hist[min(int(g), 20)] += 1
        end = max(end, e["ts"] + e["dur"])
    out.append("idle gaps (nothing running) by the kernel that follows: count, total us")
    for k, (n, d) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
        out.append("  %-50s %4d %9.1f" % (k, n, d))
    out.append("gap histogram (us -> count): " + " ".join("%d:%d" % kv for kv in sorted(hist.items())))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        k = short(e["name"])
        tot[k][0] += 1
        tot[k][1] += e["dur"]
    out.append("per kernel (launches, total ms, avg us) inside the replayed graph (durations include slow-down from sharing the GPU):")
    for k, (n, d) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
        out.append("  %-50s %4d %8.3f %8.1f" % (k, n, d / 1e3, d / n))
    # main-chain view: the stream with the largest busy time
    main = max(by_stream.items(), key=lambda kv: kv[1][1])[0]
    me = [e for e in ev if e.get("args", {}).get("stream") == main]
    g_in = 0.0
    for a, b in zip(me, me[1:]):
        g_in += max(0.0, b["ts"] - (a["ts"] + a["dur"]))
    out.append("main stream %s: %d kernels, busy %.3f ms, gaps between its consecutive kernels %.3f ms" % (main, len(me), sum(e["dur"] for e in me) / 1e3, g_in / 1e3))
    txt = "\n".join(out)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")


if __name__ == "__main__":
    main()

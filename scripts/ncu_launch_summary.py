"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py --no-graph into the launch
list of ONE training step (the launches between two consecutive sgd_kernel launches) and a
per-kernel summary:  python scripts/ncu_launch_summary.py <ncu.csv> <out_list.csv> <out_summary.txt>"""
import collections
import csv
import re
import sys


def main(src, out_list, out_summary, which=-1):
    txt = open(src, newline="").read()
    lines = txt[txt.index('"ID","Process ID"'):].splitlines()
    rd = csv.reader(lines)
    hdr = next(rd)
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    rows = []
    for r in rd:
        if len(r) != len(hdr):
            continue
        try:
            t = float(r[v].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[u], 1e-3)
        rows.append((r[k], t * scale))
    sgd = [i for i, (n, _) in enumerate(rows) if "sgd_kernel" in n]
    assert len(sgd) >= 2, "need two optimizer launches to delimit a step"
    lo, hi = sgd[which - 1] + 1, sgd[which] + 1      # (previous sgd, this sgd]
    step = rows[lo:hi]
    with open(out_list, "w") as f:
        f.write("idx,kernel,duration_us\n")
        for i, (n, t) in enumerate(step):
            f.write('%d,"%s",%.3f\n' % (i, n.replace('"', "'"), t))
    groups = collections.OrderedDict()
    for n, t in step:
        short = re.sub(r"^void ", "", n)
        short = re.sub(r"^sib::", "", short)
        short = short.split("(")[0][:60]
        g = groups.setdefault(short, [0, 0.0])
        g[0] += 1
        g[1] += t
    total = sum(t for _, t in step)
    with open(out_summary, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline\n")
        f.write("# ONE training step (launches after one sgd_kernel up to and including the next): %d launches, "
                "%.3f ms (cold-cache, serialised per-launch times: compare SHARES)\n" % (len(step), total / 1e3))
        for n, (c, t) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
            f.write("%-62s n=%4d  %8.3f ms  %4.1f%%\n" % (n, c, t / 1e3, 100 * t / total))
    print(open(out_summary).read())


if __name__ == "__main__":
    main(*sys.argv[1:4])

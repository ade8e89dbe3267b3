"""Reduce an `ncu --csv` log of `bench.py --steps 1 --warmup 3 --no-graph` (metrics: gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum, sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed)
to ONE training step (the launches between the last two optimizer kernels) and write
  <out_json>    per-family DRAM traffic / time / tensor-pipe activity (read by bench.py: roofline.traffic,
                roofline_bn.traffic)
  <out_summary> per-kernel launch summary of that step
      python scripts/ncu_step_traffic.py gpurun_out/r2/ncu_step.csv profiles/r02_step_dram_traffic.json profiles/r02_ncu_launch_summary.txt
"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def main(src, out_json, out_summary):
    txt = open(src, newline="").read()
    lines = txt[txt.index('"ID","Process ID"'):].splitlines()
    rd = csv.reader(lines)
    hdr = next(rd)
    i_id, i_k, i_m, i_v, i_u = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = collections.OrderedDict()
    for r in rd:
        if len(r) != len(hdr):
            continue
        try:
            val = float(r[i_v].replace(",", "")) * UNIT.get(r[i_u], 1.0)
        except ValueError:
            continue
        rec = launches.setdefault(r[i_id], {"name": r[i_k]})
        rec[r[i_m]] = val
    rows = list(launches.values())
    opt = [i for i, r in enumerate(rows) if "sgd_kernel" in r["name"]]
    assert len(opt) >= 2, "need two optimizer launches to delimit a step"
    step = rows[opt[-2] + 1:opt[-1] + 1]

    def fam(name):
        if re.search(r"igemm|wgrad_|halo3x3", name):
            return "conv"
        if re.search(r"bn_(finalize_apply|bwd_apply|bwd_reduce|apply|stats|act_maxpool)", name):
            return "bn"
        return "other"

    fams = collections.defaultdict(lambda: collections.defaultdict(float))
    groups = collections.OrderedDict()
    for r in step:
        t = r.get("gpu__time_duration.sum", 0.0)
        rd_b, wr_b = r.get("dram__bytes_read.sum", 0.0), r.get("dram__bytes_write.sum", 0.0)
        tp = r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        f = fams[fam(r["name"])]
        f["launches"] += 1
        f["time_us"] += t
        f["dram_bytes_read"] += rd_b
        f["dram_bytes_write"] += wr_b
        f["tensor_pct_x_time"] += tp * t
        short = re.sub(r"^void ", "", r["name"])
        short = re.sub(r"^sib::", "", short).split("(")[0][:60]
        g = groups.setdefault(short, [0, 0.0, 0.0])
        g[0] += 1
        g[1] += t
        g[2] += rd_b + wr_b
    out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none "
                     "python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-baseline "
                     "(launches between the last two optimizer kernels = one step; cold-cache, serialised)",
           "launches_per_step": len(step)}
    for name, f in fams.items():
        out[name] = {"launches": int(f["launches"]), "time_us_cold_serialised": f["time_us"],
                     "dram_bytes_read": f["dram_bytes_read"], "dram_bytes_write": f["dram_bytes_write"],
                     "dram_bytes_per_step": f["dram_bytes_read"] + f["dram_bytes_write"],
                     "tensor_pipe_active_pct_time_weighted": f["tensor_pct_x_time"] / f["time_us"] if f["time_us"] else 0.0}
    json.dump(out, open(out_json, "w"), indent=1)
    total = sum(g[1] for g in groups.values())
    with open(out_summary, "w") as fsum:
        fsum.write("# %s\n" % out["source"])
        fsum.write("# ONE training step: %d launches, %.3f ms summed kernel time (compare SHARES)\n" % (len(step), total / 1e3))
        for n, (c, t, b) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
            fsum.write("%-62s n=%4d  %8.3f ms  %4.1f%%  dram %8.1f MB\n" % (n, c, t / 1e3, 100 * t / total, b / 1e6))
    print(open(out_summary).read())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])

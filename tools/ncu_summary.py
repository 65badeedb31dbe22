#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) or an ncu launch-list CSV into a small text table for profiles/.

  python tools/ncu_summary.py rep   gpurun_out/prof.ncu-rep  > profiles/rN_xxx.txt
  python tools/ncu_summary.py list  gpurun_out/launches.csv  > profiles/rN_launches.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__waves_per_multiprocessor", "waves"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
        ("smsp__inst_executed.sum", "warp_inst")]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    gi = hdr.index("Grid Size")
    print("# %s (ncu --set full --clock-control none; per launch, cold cache, serialised)" % path)
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("fgnn::<unnamed>::", "")
        parts = ["%-34s grid=%-14s" % (name[:34], r[gi])]
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                try:
                    v = "%.4g" % float(v)
                except ValueError:
                    pass
                parts.append("%s=%s%s" % (short, v, units[i].replace("byte", "B").replace("register/thread", "")
                                          .replace("inst", "").replace("%", "%")))
        print("  ".join(parts))


def launch_list(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("fgnn::<unnamed>::", "")
        v = float(row["Metric Value"])
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg.setdefault((name, row["Grid Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache serialised: compare SHARES)" % path)
    for (k, g), a in agg.items():
        print("%-44s grid=%-14s n=%4d avg=%9.2f us share=%.3f" % (k[:44], g, a[0], a[1] / a[0], a[1] / tot))
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))


def traffic(path, pattern="gather_bulk"):
    """JSON for profiles/gather_traffic.json: mean DRAM read+write bytes per launch of the kernels matching
    `pattern` in an `ncu --set full` report (bench.py reports it as roofline.traffic)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    rd, wr, us, name = [], [], [], None
    for r in rows[2:]:
        if pattern not in r[ki]:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("fgnn::<unnamed>::", "")
        rd.append(float(r[ri]) * scale[units[ri]])
        wr.append(float(r[wi]) * scale[units[wi]])
        us.append(float(r[ti]) * tscale.get(units[ti], 1.0))
    n = max(1, len(rd))
    print(json.dumps({"kernel": name, "launches": len(rd), "dram_read_bytes_per_launch": sum(rd) / n,
                      "dram_write_bytes_per_launch": sum(wr) / n, "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n,
                      "ncu_time_us_per_launch": sum(us) / n, "report": path,
                      "how": "ncu --set full --clock-control none, bench.py timed region (cold cache, serialised replay)"}))


def stalls(path):
    """Warp-state breakdown per kernel (smsp__average_warps_issue_stalled_*_per_issue_active: warps stalled in that
    state per issued instruction) plus the DRAM sector figures: which latency the kernel sits on."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
    cols = [(i, h[len(pre):-len(suf)]) for i, h in enumerate(hdr) if h.startswith(pre) and h.endswith(suf)]
    extra = [("dram__sectors_read.sum", "dram_rd_sectors"), ("dram__sectors_write.sum", "dram_wr_sectors"),
             ("gpu__time_duration.sum", "time"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
             ("lts__t_sectors_op_atom.sum", "l2_atom_sectors"), ("lts__t_sectors_op_red.sum", "l2_red_sectors")]
    print("# %s: warps stalled per issued instruction, by state (top 5), ncu --set full" % path)
    seen = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("fgnn::<unnamed>::", "")
        key = (name, r[gi])
        if key in seen:
            continue
        seen[key] = 1
        vals = []
        for i, st in cols:
            try:
                vals.append((float(r[i]), st))
            except ValueError:
                pass
        vals.sort(reverse=True)
        tot = sum(v for v, _ in vals) or 1.0
        parts = ["%-30s grid=%-14s" % (name[:30], r[gi])]
        parts.append(" ".join("%s=%.2f(%.0f%%)" % (st, v, 100 * v / tot) for v, st in vals[:5]))
        for k, short in extra:
            if k in hdr:
                try:
                    parts.append("%s=%.4g%s" % (short, float(r[hdr.index(k)]), rows[1][hdr.index(k)]))
                except ValueError:
                    pass
        print("  ".join(parts))


if __name__ == "__main__":
    {"rep": rep, "list": launch_list, "traffic": traffic, "stalls": stalls}[sys.argv[1]](*sys.argv[2:])

#!/usr/bin/env python
"""Summarise ncu --csv logs into the small files kept under profiles/.

  launches <launches.csv> <out.csv>          per-kernel-name count / total / mean gpu__time_duration (one forward)
  traffic  <conv_traffic.csv> <out.json>     DRAM and L2->SM bytes summed over the conv launches of one forward
  kernels  <kernels.csv> <out.csv>           ONE ROW PER KERNEL of the last forward: launches, time, DRAM GB/s and % of
                                             peak, tensor-pipe % (time-weighted), L2->SM bytes, LTS / SM throughput %
"""
import csv
import json
import sys


def rows(path):
    """One dict per launch from either ncu CSV shape: `--csv --log-file` (long form, one line per launch x metric) or
    `ncu -i report.ncu-rep --page raw --csv` (wide form: a header line, a units line, one line per launch — what a
    `--set full` capture exports)."""
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    head = next(csv.reader(lines[:1]))
    if "Metric Name" not in head:                       # wide form
        table = list(csv.reader(lines))
        units = dict(zip(table[0], table[1]))
        rs = [dict(zip(table[0], t)) for t in table[2:] if len(t) == len(table[0])]
    else:
        r = list(csv.DictReader(lines))
        out, units = {}, {}
        for x in r:
            d = out.setdefault(x["ID"], {"Kernel Name": x["Kernel Name"]})
            d[x["Metric Name"]] = x["Metric Value"]
            units[x["Metric Name"]] = x["Metric Unit"]
        rs = [out[k] for k in sorted(out, key=int)]
    # keep the LAST forward only: a forward starts at its conv0 launch and ends at the main head_kernel
    starts = [i for i, x in enumerate(rs) if "conv0" in x["Kernel Name"]]
    if starts:
        rs = rs[starts[-1]:]
    return rs, units


def num(s):
    return float(s.replace(",", ""))


def to_us(v, unit):
    return num(v) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[unit]


def to_bytes(v, unit):
    return num(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def launches(src, dst):
    rs, units = rows(src)
    agg = {}
    for r in rs:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += to_us(r["gpu__time_duration.sum"], units["gpu__time_duration.sum"])
    total = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "mean_us", "share"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, "%.3f" % t, "%.3f" % (t / n), "%.4f" % (t / total)])
        w.writerow(["TOTAL", sum(a[0] for a in agg.values()), "%.3f" % total, "", "1.0"])
    print(open(dst).read())


def traffic(src, dst):
    rs, units = rows(src)
    rs = [r for r in rs if "conv0" not in r["Kernel Name"] and "conv" in r["Kernel Name"]]    # the tcgen05 conv family
    d = {
        "source": "ncu (%s), batch 128, one forward = %d conv launches (conv_igemm_persistent / conv_halo / pool_conv; "
                  "conv0 excluded)" % (src, len(rs)),
        "conv_launches_per_step": len(rs),
        "sum_gpu_time_us": sum(to_us(r["gpu__time_duration.sum"], units["gpu__time_duration.sum"]) for r in rs),
        "dram_bytes_read": sum(to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) for r in rs),
        "dram_bytes_write": sum(to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"]) for r in rs),
    }
    d["dram_bytes_total"] = d["dram_bytes_read"] + d["dram_bytes_write"]
    k = "l1tex__m_xbar2l1tex_read_bytes.sum"
    if k in units:
        d["l2_to_sm_bytes"] = sum(to_bytes(r[k], units[k]) for r in rs)
    json.dump(d, open(dst, "w"), indent=1)
    print(json.dumps(d, indent=1))


def kernels(src, dst):
    rs, units = rows(src)
    T, DR, DW = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
    DP = "dram__throughput.avg.pct_of_peak_sustained_elapsed"
    TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    XB = "l1tex__m_xbar2l1tex_read_bytes.sum"
    LT = "lts__throughput.avg.pct_of_peak_sustained_elapsed"
    SM = "sm__throughput.avg.pct_of_peak_sustained_elapsed"
    agg = {}
    for r in rs:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, {"n": 0, "us": 0.0, "dram": 0.0, "xbar": 0.0, "w": {DP: 0.0, TP: 0.0, LT: 0.0, SM: 0.0}})
        t = to_us(r[T], units[T])
        a["n"] += 1
        a["us"] += t
        a["dram"] += to_bytes(r[DR], units[DR]) + to_bytes(r[DW], units[DW])
        if XB in r:
            a["xbar"] += to_bytes(r[XB], units[XB])
        for m in a["w"]:
            v = r.get(m)
            if v is None and m == DP:               # the full set names it gpu__dram_throughput
                v = r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
            if v is not None:
                a["w"][m] += num(v) * t             # time-weighted mean over the kernel's launches
    total = sum(a["us"] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share", "dram_GB", "dram_GBps", "dram_pct_of_peak", "tensor_pipe_pct_active",
                    "l2_to_sm_GB", "l2_to_sm_TBps", "lts_throughput_pct", "sm_throughput_pct"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            t = a["us"]
            w.writerow([k, a["n"], "%.1f" % t, "%.4f" % (t / total), "%.3f" % (a["dram"] / 1e9), "%.0f" % (a["dram"] / t / 1e3),
                        "%.1f" % (a["w"][DP] / t), "%.1f" % (a["w"][TP] / t), "%.3f" % (a["xbar"] / 1e9),
                        "%.2f" % (a["xbar"] / t / 1e6), "%.1f" % (a["w"][LT] / t), "%.1f" % (a["w"][SM] / t)])
        w.writerow(["TOTAL", sum(a["n"] for a in agg.values()), "%.1f" % total, "1.0",
                    "%.3f" % (sum(a["dram"] for a in agg.values()) / 1e9), "", "", "",
                    "%.3f" % (sum(a["xbar"] for a in agg.values()) / 1e9), "", "", ""])
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic, "kernels": kernels}[sys.argv[1]](sys.argv[2], sys.argv[3])

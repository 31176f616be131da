#!/usr/bin/env python
"""Summarise ncu --csv logs into the small files kept under profiles/.

  launches <launches.csv> <out.csv>          per-kernel-name count / total / mean gpu__time_duration (one forward)
  traffic  <conv_traffic.csv> <out.json>     DRAM and L2->SM bytes summed over the conv launches of one forward
"""
import csv
import json
import sys


def rows(path):
    """ncu --csv --log-file output is long-form (one line per launch x metric): pivot it to one dict per launch."""
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
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
    rs = [r for r in rs if "conv0" not in r["Kernel Name"]]
    d = {
        "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... -k regex:conv_ (%s), "
                  "batch 128 bf16, one forward = %d conv launches (conv0 excluded)" % (src, len(rs)),
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


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])

"""Per-kernel table from an ncu launch list (`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--clock-control none --csv`).  usage: python profiles/summarize_launches.py launches.csv [first_launch] [n_launches]
(default: the last captured train step)"""
import collections
import csv
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ii, ki, mi, ui, vi = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ii], {"name": r[ki], "us": 0.0, "bytes": 0.0})
        v = float(r[vi].replace(",", "")) * UNIT[r[ui]]
        if r[mi].startswith("gpu__time_duration"):
            d["us"] = v
        elif r[mi].startswith("dram__bytes"):
            d["bytes"] += v
    data = list(launches.values())
    if len(sys.argv) > 3:
        data = data[int(sys.argv[2]):int(sys.argv[2]) + int(sys.argv[3])]
    else:
        # default: the LAST train step = from the last weight-preparation launch (tcs_prep_kernel / tc_prep_kernel) on
        prep = [i for i, d in enumerate(data) if "prep_kernel" in d["name"] and "embed" not in d["name"]]
        data = data[prep[-1]:] if prep else data[len(data) // 2:]
    agg = collections.OrderedDict()
    for d in data:
        k = d["name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:52]
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["us"]
        a[2] += d["bytes"]
    tot = sum(a[1] for a in agg.values())
    totb = sum(a[2] for a in agg.values())
    for k, (c, us, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print("%-54s n=%4d %9.1f us %7.1f us/launch %5.1f%%  %7.2f GB  %5.0f GB/s" % (k, c, us, us / c, 100 * us / tot, b / 1e9, b / us / 1e3 if us else 0))
    print("total %.3f ms over %d launches, %.2f GB DRAM traffic (%.0f GB/s)" % (tot / 1e3, len(data), totb / 1e9, totb / tot / 1e3))


if __name__ == "__main__":
    main()

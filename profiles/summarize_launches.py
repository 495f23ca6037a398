"""Per-kernel table from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
usage: python profiles/summarize_launches.py launches.csv [first_launch] [n_launches]   (default: the last third)"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
    if len(sys.argv) > 3:
        data = data[int(sys.argv[2]):int(sys.argv[2]) + int(sys.argv[3])]
    else:
        data = data[-(len(data) // 3):]
    agg = collections.OrderedDict()
    for k, v in data:
        k = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:58]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print("%-60s n=%4d %10.1f us %6.1f us/launch %5.1f%%" % (k, c, v / 1e3, v / 1e3 / c, 100 * v / tot))
    print("total %.3f ms over %d launches" % (tot / 1e6, len(data)))


if __name__ == "__main__":
    main()

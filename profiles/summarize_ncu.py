"""Turns an .ncu-rep (ncu --set full) into a small JSON summary kept under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_ncu_<kernel>.json [launch_index]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    row = data[idx]
    d = {"report": rep, "kernel": row[hdr.index("Kernel Name")], "launches_in_report": len(data)}
    for i, h in enumerate(hdr):
        if h in KEYS:
            try:
                d[h] = {"value": float(row[i].replace(",", "")), "unit": units[i]}
            except ValueError:
                d[h] = {"value": row[i], "unit": units[i]}

    def mb(k):
        v = d[k]
        f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[v["unit"]]
        return v["value"] * f
    d["dram_bytes_per_launch"] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    # stall mix from the source page
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    if idx < len(blocks) and blocks[idx]:
        h2, dat = blocks[idx][0], blocks[idx][1:]
        cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h2[i]: sum(int(r[i] or 0) for r in dat if len(r) > i) for i in cols}
        s = sum(tot.values()) or 1
        d["stall_mix_pct"] = {k: round(100.0 * v / s, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]}
    with open(out, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d, indent=1)[:1500])


if __name__ == "__main__":
    main()

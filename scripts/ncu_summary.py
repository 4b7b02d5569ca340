"""Per-kernel summary of an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log:
launches, mean device time, mean DRAM bytes read / written per launch, share of the summed time.

    python scripts/ncu_summary.py gpurun_out/x.csv [> profiles/x_summary.csv]
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("lsb::", "")


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rd:
        if len(r) < len(hdr):
            continue
        key = r[col["ID"]]
        d = launches.setdefault(key, {"name": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        v = r[col["Metric Value"]].replace(",", "")
        try:
            val = float(v)
        except ValueError:
            continue
        unit = r[col["Metric Unit"]]
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = val
    agg = OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "grid": d["grid"], "block": d["block"]})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(a["us"] for a in agg.values()) or 1.0
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "mean_us", "share_of_time", "mean_dram_read_MB", "mean_dram_write_MB", "grid", "block"])
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        w.writerow([k, n, f"{a['us'] / n:.2f}", f"{a['us'] / total:.4f}", f"{a['rd'] / n / 1e6:.3f}", f"{a['wr'] / n / 1e6:.3f}", a["grid"], a["block"]])


if __name__ == "__main__":
    main(sys.argv[1])

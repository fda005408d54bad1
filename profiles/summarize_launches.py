"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def main(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    unit = "ns"
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("asw::<unnamed>::", "").replace("void ", "")
        agg.setdefault(name, []).append(float(row["Metric Value"].replace(",", "")))
        unit = row["Metric Unit"]
    tot = sum(sum(v) for v in agg.values())
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    print(f"{'kernel':58s} {'n':>5s} {'avg_us':>10s} {'total_us':>10s} {'share':>6s}")
    for k, v in agg.items():
        print(f"{k[:58]:58s} {len(v):5d} {sum(v) / len(v) * scale:10.1f} {sum(v) * scale:10.1f} {sum(v) / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1])

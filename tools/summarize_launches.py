"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections
import csv
import sys


def main(path, only_wb=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        if only_wb and not k.startswith("wb::"):
            k = "(torch: synthetic data generation / glue)"
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms of kernel time (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:45s} launches={v[0]:5d} total_ms={v[1]:10.3f} share={v[1] / tot:6.1%}")


if __name__ == "__main__":
    main(sys.argv[1], only_wb="--wb" in sys.argv)

"""Stall samples per SASS opcode (ncu report, --page source): where do warps wait?"""
import collections
import csv
import subprocess
import sys


def main(path, top=18):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    hdr = rows[1]
    si = hdr.index("Source"); ni = hdr.index("Warp Stall Sampling (All Samples)"); ei = hdr.index("Instructions Executed")
    agg = collections.defaultdict(lambda: [0.0, 0.0])
    for r in rows[2:]:
        if len(r) <= max(si, ni, ei):
            continue
        txt = r[si].strip()
        if txt.startswith("@"):
            txt = txt.split(" ", 1)[1].strip()
        op = txt.split(" ")[0].split(".")[0]
        if op in ("LDG", "LDS", "STS", "STG", "LD", "ST"):
            op = txt.split(" ")[0]
        try:
            agg[op][0] += float(r[ni]); agg[op][1] += float(r[ei])
        except ValueError:
            pass
    tot = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
    print(f"# {path}: {tot:.0f} stall samples, {ti:.3g} warp instructions")
    for op, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{op:22s} samples {v[0] / tot:6.1%}   instructions {v[1] / ti:6.1%}")


if __name__ == "__main__":
    main(sys.argv[1])

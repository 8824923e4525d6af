"""Groups consecutive SASS lines with (nearly) equal execution counts: a poor man's per-loop profile."""
import collections
import csv
import subprocess
import sys


def main(path, min_share=0.01):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    hdr = rows[1]
    si = hdr.index("Source"); ei = hdr.index("Instructions Executed"); wi = hdr.index("Warp Stall Sampling (All Samples)")
    blocks = []
    cur = None
    for r in rows[2:]:
        try:
            n = float(r[ei]); w = float(r[wi])
        except (ValueError, IndexError):
            continue
        txt = r[si].strip()
        if txt.startswith("@"):
            txt = txt.split(" ", 1)[1].strip()
        op = txt.split(" ")[0].split(".")[0]
        if cur and abs(cur["n"] - n) < 0.03 * max(cur["n"], 1) + 1000:
            cur["ops"][op] += 1; cur["lines"] += 1; cur["tot"] += n; cur["stall"] += w
        else:
            cur = {"n": n, "ops": collections.Counter({op: 1}), "lines": 1, "tot": n, "stall": w}
            blocks.append(cur)
    tot = sum(b["tot"] for b in blocks); st = sum(b["stall"] for b in blocks)
    print(f"# {path}: {tot:.3g} warp instructions")
    for b in blocks:
        if b["tot"] / tot > min_share or b["stall"] / st > min_share:
            print(f"count {b['n'] / 1e6:8.2f}M lines {b['lines']:4d} inst {b['tot'] / tot:6.1%} stall {b['stall'] / st:6.1%}  ",
                  dict(b["ops"].most_common(6)))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.01)

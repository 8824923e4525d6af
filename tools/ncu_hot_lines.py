"""Top source lines by warp-stall samples from an ncu report captured with --import-source on."""
import csv
import subprocess
import sys


def main(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda"],
                         capture_output=True, text=True).stdout.splitlines()
    files = []
    cur = None
    rows = []
    hdr = None
    for line in out:
        if line.startswith('"File Name"'):
            cur = next(csv.reader([line]))[1]
            hdr = None
            continue
        r = next(csv.reader([line]))
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            d["file"] = cur
            rows.append(d)
    key = None
    for k in rows[0].keys():
        if "Samples" in k and "All" in k:
            key = k
    if key is None:
        key = "# Samples"

    def f(x):
        try:
            return float(x)
        except Exception:
            return 0.0
    tot = sum(f(r.get(key, 0)) for r in rows)
    print(f"# {path}: {tot:.0f} samples ({key})")
    for r in sorted(rows, key=lambda r: -f(r.get(key, 0)))[:top]:
        print(f"{f(r[key]) / tot:6.1%}  {r['file'].split('/')[-1]}:{r['Line No']:>4}  {r['Source'].strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

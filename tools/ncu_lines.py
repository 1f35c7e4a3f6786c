"""Per-source-line summary of an ncu report (stall samples and executed instructions), from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.  usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = curfile = hdr = None
    data = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            curfile = r[1]
        elif r[0] == "Function Name":
            cur = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif cur and hdr and len(r) > 7 and r[2] == "-":
            try:
                samp, inst = int(r[4]), int(r[7])
            except ValueError:
                continue
            e = data.setdefault(cur, {}).setdefault((curfile.split("/")[-1], r[0], r[1].strip()[:100]), [0, 0])
            e[0] += samp
            e[1] += inst
    for k, d in data.items():
        tot = sum(v[0] for v in d.values()) or 1
        toti = sum(v[1] for v in d.values()) or 1
        print(f"===== {k[:90]}  samples {tot}  warp-instructions {toti}")
        for key, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{100 * v[0] / tot:5.1f}% samp {100 * v[1] / toti:5.1f}% inst  {key[0]}:{key[1]}  {key[2]}")


if __name__ == "__main__":
    main()

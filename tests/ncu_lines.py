"""ncu_lines.py -- per-source-line instruction / stall-sample totals of one kernel from an ncu report (diagnostic).
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv ; python tests/ncu_lines.py x.csv [top N] [blocks]"""
import csv
import sys


def load(path):
    out = {}          # (file, line) -> [inst, samples, src]
    cur = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No") or r[0] == "":
            continue
        try:
            line = int(r[0])
            inst = int(r[7]); samp = int(r[6])
        except (ValueError, IndexError):
            continue
        e = out.setdefault((cur, line), [0, 0, r[1]])
        e[0] += inst; e[1] += samp
    return out


if __name__ == "__main__":
    d = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    blocks = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    ti = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values())
    print("total warp-instructions %d (%.0f per block), samples %d" % (ti, ti / blocks, ts))
    for (f, l), v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-14s %4d  inst %5.2f%%  samples %5.2f%%  %s" % (f, l, 100.0 * v[0] / ti, 100.0 * v[1] / ts, v[2][:100]))

"""ncu_sections.py -- section-level split of k_analysis from an ncu source-page csv (diagnostic).
usage: python tests/ncu_sections.py x.csv blocks   (section boundaries are read from the marker comments in analysis.cuh)"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import load  # noqa: E402

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "perceptual-audio-codec_b200", "csrc", "analysis.cuh")
MARKS = [("E1 power spectrum", r"// 1\. power spectrum"), ("E2 peaks, maskers, scans", r"// 2\. findpeaks"),
         ("E3 carries, scan values, loud list", r"// 3\. cross-warp"), ("E4 per-line sums + loud skirts", r"// 4\. upper skirts"),
         ("kernel prologue", r"^k_analysis\("), ("A load + dequantise PCM", r"A\. load one"), ("C MDCT + overall scale", r"C\. SineWindow"),
         ("B raw FFT, M/S decision, split", r"B\. raw FFTs"), ("F2 Hann taps", r"F2_M, F2_S ="), ("E curve loop", r"E\. six masked"),
         ("F SMR, band maxima, stores", r"F\. SMR candidates"), ("(mono kernel)", r"^k_calc_smrs\(")]


def main():
    d = load(sys.argv[1])
    blocks = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lines = open(SRC).read().splitlines()
    starts = []
    for name, pat in MARKS:
        for i, l in enumerate(lines, 1):
            if re.search(pat, l):
                starts.append((i, name))
                break
    starts.sort()
    ti = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values())
    agg = {}
    for (f, l), v in d.items():
        if f != "analysis.cuh":
            key = f + " (inlined helpers)"
        else:
            key = "before sections"
            for s0, name in starts:
                if l >= s0:
                    key = name
        e = agg.setdefault(key, [0, 0]); e[0] += v[0]; e[1] += v[1]
    print("total %.0f warp-instructions per block" % (ti / blocks))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-38s samples %5.1f%%  inst %5.1f%% (%6.0f per block)" % (k, 100.0 * v[1] / ts, 100.0 * v[0] / ti, v[0] / blocks))


if __name__ == "__main__":
    main()

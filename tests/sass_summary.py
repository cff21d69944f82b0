"""sass_summary.py -- opcode census of libpacb200.so per kernel (diagnostic; writes the evidence file under profiles/).
usage: python tests/sass_summary.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "perceptual-audio-codec_b200", "libpacb200.so")
WATCH = ["MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "MUFU", "REDUX", "SHFL", "LDGSTS", "LDS", "STS", "LDG", "STG", "ATOMS", "BAR", "DFMA", "DADD", "DMUL",
         "FFMA", "HMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "HGMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and kern:
            op = m.group(1)
            counts[kern]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[kern][w] += 1
    print("# SASS opcode census of perceptual-audio-codec_b200/libpacb200.so (cuobjdump -sass, sm_100a), static instruction counts per kernel")
    print("# tensor-core / TMA mnemonics (UTC*MMA, LDTM/STTM, UTMALDG/UTMASTG/UBLKCP, HMMA, HGMMA) are listed so that their absence is visible:")
    print("# nothing on this path is a contraction (FFT butterflies, a masker x line spreading sum, integer bit work), see DESIGN.md section 3")
    tot = collections.Counter()
    for k, c in counts.items():
        tot.update(c)
        print("%-70s total %6d  %s" % (k[:70], c["_total"], "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))
    print("ALL KERNELS: total %d  %s" % (tot["_total"], "  ".join("%s %d" % (w, tot[w]) for w in WATCH)))


if __name__ == "__main__":
    sys.exit(main())

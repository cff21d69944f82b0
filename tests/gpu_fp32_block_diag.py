"""compare fp32 and fp64 threshold curves of individual corpus blocks"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "perceptual-audio-codec_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pacb200, pacb200_batch as pbat
from conftest import corpus_files
files = corpus_files()
e32 = _pacb200.Engine(0, "fp32"); e64 = _pacb200.Engine(0, "fp64")
def frac(pcm):
    pcm = np.asarray(pcm, dtype=np.float64); return np.sign(pcm) * 2.0 * np.abs(pcm) / 65535.0
for name, b in (("harmonic_test2", 23), ("piano1", 220), ("percussion_test3", 280), ("pop_test2", 748)):
    pcm = pbat.read_wav(files[name])[1]
    seg = np.zeros((2048, 2), np.int16); lo = (b - 1) * 1024
    src = pcm[max(lo, 0):lo + 2048]; seg[max(0, -lo):max(0, -lo) + len(src)] = src
    x = frac(seg).T[None].copy()
    r32, r64 = e32.analysis(x), e64.analysis(x)
    print(name, b, "lrms", hex(int(r32["lrms"][0])), hex(int(r64["lrms"][0])))
    for c, cn in enumerate(("L", "R", "M", "S", "M'", "S'")):
        d = r32["bthr"][0][c] - r64["bthr"][0][c]
        i = int(np.argmax(np.abs(d)))
        big = np.nonzero(np.abs(d) > 1e-3)[0]
        print("   curve %-2s max |d| %.3e dB at line %d (thr %.2f); lines with |d|>1e-3: %d %s" % (cn, np.abs(d).max(), i, r64["bthr"][0][c][i], len(big), (big[0], big[-1]) if len(big) else ""))
    ds = r32["smr"][0] - r64["smr"][0]
    print("   smr diff max", np.abs(ds).max(), "at", np.unravel_index(np.argmax(np.abs(ds)), ds.shape))
    np.savez(os.path.join(ROOT, "gpurun_out", "blk_%s_%d.npz" % (name, b)), x=x, b32=r32["bthr"][0], b64=r64["bthr"][0], s32=r32["smr"][0], s64=r64["smr"][0])

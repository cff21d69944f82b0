"""gpu_cmp.py kind seed n -- field-by-field diff of GPU fp64 trace vs oracle trace for a seeded synthetic stream."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "perceptual-audio-codec_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import _pacb200, oracle as orc
from test_gpu_parity import synth_pcm
kind, seed, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
pcm = synth_pcm(seed, n, kind)
e = _pacb200.Engine(0, "fp64")
(got,), tr = e.encode_batch(pcm[None], trace=True)
O = orc.get()
want, otr, _ = O.encode_stream(pcm, trace=True)
print("equal", got == want, len(got), len(want))
nb = len(otr["lrms"])
for f in ("lrms", "oscale", "ba", "sf", "tableID", "nbytes", "extraBits", "bitDeposit"):
    a, b = tr[f][0][:nb], otr[f]
    bad = np.nonzero((a != b).reshape(nb, -1).any(1))[0]
    print("%-10s mismatching blocks: %s" % (f, list(bad[:10])))
    if len(bad):
        i = bad[0]
        print("   gpu   ", a[i].tolist() if hasattr(a[i], "tolist") else a[i])
        print("   oracle", b[i].tolist() if hasattr(b[i], "tolist") else b[i])
ds = np.abs(tr["smr"][0][:nb] - otr["smr"])
print("smr max diff per block", ds.reshape(nb, -1).max(1)[:8])
i = int(np.argmax(ds.reshape(nb, -1).max(1)))
print("worst block", i, "\n gpu   ", np.round(tr["smr"][0][i], 4).tolist(), "\n oracle", np.round(otr["smr"][i], 4).tolist())
dl = np.abs(tr["lines"][0][:nb] - otr["lines"])
print("lines max diff", dl.max(), "signbit mismatches", int(np.sum(np.signbit(tr["lines"][0][:nb]) != np.signbit(otr["lines"]))))
w = np.argwhere(np.signbit(tr["lines"][0][:nb]) != np.signbit(otr["lines"]))[:5]
for idx in w:
    print("  ", idx.tolist(), tr["lines"][0][tuple(idx)], otr["lines"][tuple(idx)])

# ---- analysis-level diff on the worst block
NL44 = [5, 4, 5, 5, 5, 5, 7, 7, 7, 9, 10, 11, 13, 15, 17, 21, 26, 32, 42, 51, 61, 83, 116, 163, 304]
nL = np.array(NL44, np.int32)
blk = np.zeros((2048, 2), np.int16)
lo, hi = (i - 1) * 1024, (i + 1) * 1024
s0, s1 = max(lo, 0), min(hi, len(pcm))
blk[s0 - lo:s1 - lo] = pcm[s0:s1]
x = np.sign(blk.astype(float)) * 2.0 * np.abs(blk.astype(float)) / 65535.0
r = e.analysis(x.T[None].copy())
lr = O.lrms(x[:, 0], x[:, 1], nL)
d0, d1 = O.sine_window(x[:, 0]), O.sine_window(x[:, 1])
X = [O.mdct(d0, 1024, 1024), O.mdct(d1, 1024, 1024)]
sc = [O.scale_factor(np.max(np.abs(X[c])), 4) for c in range(2)]
Xs = [X[c] * (1 << sc[c]) for c in range(2)]
smr, lines, bthr = O.stereo_smr(d0, d1, Xs[0], Xs[1], sc, 44100, nL, lr)
for c in range(6):
    d = np.abs(r["bthr"][0][c] - bthr[c])
    print("curve", c, "max diff %.3e at line %d" % (d.max(), d.argmax()), " gpu %.4f oracle %.4f" % (r["bthr"][0][c][d.argmax()], bthr[c][d.argmax()]))
print("clipped samples in block:", int(np.sum(blk == -32768)), "M nonzero:", int(np.sum(blk[:, 0] + blk[:, 1] != 0)))
import model_analysis as ma
T = ma.Tables()
ml = ma.analysis(T, x[:, 0].copy(), x[:, 1].copy(), NL44)
for c in range(6):
    d = np.abs(ml[3][c] - bthr[c])
    print("model curve", c, "max diff vs oracle %.3e" % d.max())
print("analysis-api smr vs oracle-api smr max diff %.3e" % np.max(np.abs(r["smr"][0] - smr)))
print("analysis-api smr vs gpu trace smr  max diff %.3e" % np.max(np.abs(r["smr"][0] - tr["smr"][0][i])))
print("oracle-api smr vs oracle trace smr max diff %.3e" % np.max(np.abs(smr - otr["smr"][i])))
print("oscale api", r["oscale"][0], "trace gpu", tr["oscale"][0][i], "trace oracle", otr["oscale"][i], "sc", sc)
print("lrms api %x trace gpu %x oracle %x" % (r["lrms"][0], tr["lrms"][0][i], otr["lrms"][i]))
print("lines api vs gpu trace %.3e ; api vs oracle trace %.3e" % (np.max(np.abs(r["lines"][0] - tr["lines"][0][i])), np.max(np.abs(r["lines"][0] - otr["lines"][i]))))

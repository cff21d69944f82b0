"""gpu_diag.py -- verbose first-light diagnostics on the GPU box (not a pytest file): prints where the CUDA path
and the oracle disagree instead of stopping at the first assert."""
import os
import sys
import time
import traceback

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "perceptual-audio-codec_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import _pacb200  # noqa: E402
import oracle as orc  # noqa: E402

NL44 = [5, 4, 5, 5, 5, 5, 7, 7, 7, 9, 10, 11, 13, 15, 17, 21, 26, 32, 42, 51, 61, 83, 116, 163, 304]


def section(t):
    print("\n==== " + t, flush=True)


def stage_check(eng, tag):
    st = np.load(os.path.join(REPO, "tests", "golden", "stages.npz"))
    keys = [str(k) for k in st["index"]]
    data = []
    for k in keys:
        pcm = st[k + ".pcm"].astype(np.float64)
        x = np.sign(pcm) * 2.0 * np.abs(pcm) / 65535.0
        data.append(x.T.copy())
    r = eng.analysis(np.stack(data))
    for i, k in enumerate(keys):
        ref_m, ref_b, ref_s, ref_l = st[k + ".mdct"], st[k + ".bthr"], st[k + ".smr"], st[k + ".lines"]
        lr = sum(int(v) << b for b, v in enumerate(st[k + ".lrms"]))
        em = np.max(np.abs(r["mdct"][i] - ref_m)) / max(np.max(np.abs(ref_m)), 1e-300)
        eb = np.max(np.abs(r["bthr"][i] - ref_b))
        es = np.max(np.abs(r["smr"][i] - ref_s))
        es_rel = np.max(np.abs(r["smr"][i] - ref_s) / np.maximum(np.abs(ref_s), 1.0))
        el = np.max(np.abs(r["lines"][i] - ref_l)) / max(np.max(np.abs(ref_l)), 1e-300)
        print("%s %-18s lrms %s oscale %s  mdct rel %.2e  bthr abs %.2e  smr abs %.2e rel %.2e  lines rel %.2e"
              % (tag, k, "ok" if int(r["lrms"][i]) == lr else "DIFF(%x vs %x)" % (int(r["lrms"][i]), lr),
                 "ok" if list(r["oscale"][i]) == list(st[k + ".oscale"]) else "DIFF", em, eb, es, es_rel, el), flush=True)


def main():
    os.system("nvidia-smi -L; nproc; free -g | head -2")
    section("context")
    e64 = _pacb200.Engine(0, "fp64")
    e32 = _pacb200.Engine(0, "fp32")
    print("version", _pacb200.lib().pac_version().decode(), "bands", list(e64.nLines))
    O = orc.get()

    section("L2 entry points")
    try:
        x = 0.5 * np.sin(0.01 * np.arange(2048.) ** 1.1)
        xs = O.sine_window(x)
        for e, t in ((e64, "fp64"), (e32, "fp32")):
            X = e.mdct(xs)
            R = O.mdct(xs, 1024, 1024)
            print(t, "mdct rel err", np.max(np.abs(X - R)) / np.max(np.abs(R)))
            y = e.imdct(R)
            Ry = O.imdct(R, 1024, 1024)
            print(t, "imdct rel err", np.max(np.abs(y - Ry)) / np.max(np.abs(Ry)))
        print("mdct N=8", e64.mdct(np.arange(8.)), O.mdct(np.arange(8.), 4, 4))
        print("imdct N=8", e64.imdct(O.mdct(np.arange(8.), 4, 4)))
        w = e64.window(0, np.ones(8)); print("sine", np.max(np.abs(w - O.sine_window(np.ones(8)))))
        w = e64.window(2, np.ones(2048)); print("kbd", np.max(np.abs(w - O.kbd_window(np.ones(2048)))))
        rng = np.random.default_rng(5)
        xs2 = np.concatenate([rng.uniform(-1, 1, 200), 10.0 ** rng.uniform(-7, 0, 200), [0.0, -0.0, 1.0, -1.0]])
        for ba in (2, 5, 16):
            ok = np.array_equal(e64.vmantissa(xs2, 3, 4, ba), O.vmantissa(xs2, 3, 4, ba))
            ok2 = [int(v) for v in e64.scale_factor(xs2, 4, ba)] == [O.scale_factor(v, 4, ba) for v in xs2]
            print("ba", ba, "vmantissa", ok, "scale_factor", ok2)
        smr = rng.uniform(-30, 40, 25)
        for extra, mask in ((0, 0), (137, 0x1555555), (-50, 0x1ffffff)):
            b1, d1 = e64.bitalloc(2116.48, extra, 16, smr, mask)
            b2, d2 = O.bitalloc(2116.48, extra, 16, 25, NL44, smr, [(mask >> b) & 1 for b in range(25)])
            print("bitalloc", extra, hex(mask), list(b1[0]) == list(b2), int(d1[0]) == d2)
    except Exception:
        traceback.print_exc()

    section("analysis stage vs reference dumps")
    for e, t in ((e64, "fp64"), (e32, "fp32")):
        try:
            stage_check(e, t)
        except Exception:
            traceback.print_exc()

    section("whole-stream encode vs oracle (piano_test2, castanets)")
    for name in ("piano_test2", "castanets"):
        try:
            rate, pcm = orc.read_wav(os.path.join(REPO, "tests", "golden", name + ".wav"))
            t0 = time.time()
            (got,), tr = e64.encode_batch(pcm[None], trace=True)
            t1 = time.time()
            want, otr, ofs = O.encode_stream(pcm, orc.default_params(rate), trace=True)
            gold = open(os.path.join(REPO, "tests", "golden", name + ".wak"), "rb").read()
            print(name, "gpu %.2fs  bytes %d vs oracle %d  equal=%s  equal_golden=%s final=%s oracle_final=%s"
                  % (t1 - t0, len(got), len(want), got == want, got == gold, list(e64.last_final_state[0]), ofs), flush=True)
            nb = len(otr["lrms"])
            for fld in ("lrms", "oscale", "ba", "sf", "tableID", "nbytes", "extraBits", "bitDeposit"):
                a, b = tr[fld][0][:nb], otr[fld]
                bad = np.nonzero((a != b).reshape(nb, -1).any(1))[0]
                print("   %-10s first mismatching block: %s (of %d mismatching)" % (fld, bad[0] if len(bad) else None, len(bad)))
            es = np.max(np.abs(tr["smr"][0][:nb] - otr["smr"]))
            el = np.max(np.abs(tr["lines"][0][:nb] - otr["lines"]))
            print("   smr max abs diff %.3e   lines max abs diff %.3e" % (es, el))
            if got != want:
                n = min(len(got), len(want))
                d = [i for i in range(n) if got[i] != want[i]][:8]
                print("   first differing bytes", d)
            # decode
            dec = e64.decode_batch([want])[0]
            wpcm = O.decode_stream(want)[0]
            print("   decode equal=%s shapes %s %s maxdiff %s" % (np.array_equal(dec[0], wpcm), dec[0].shape, wpcm.shape,
                  np.max(np.abs(dec[0].astype(int) - wpcm.astype(int))) if dec[0].shape == wpcm.shape else "n/a"), flush=True)
            # fp32
            (g32,), tr32 = e32.encode_batch(pcm[None], trace=True)
            mm = {f: float(np.mean(tr32[f][0][:nb] != otr[f])) for f in ("lrms", "oscale", "ba", "sf", "tableID")}
            print("   fp32: bytes %d  mismatch rates %s" % (len(g32), mm))
            d32 = e32.decode_batch([want])[0]
            print("   fp32 decode: max |pcm diff| =", np.max(np.abs(d32[0].astype(int) - wpcm.astype(int))))
        except Exception:
            traceback.print_exc()

    section("per-block API")
    try:
        rate, pcm = orc.read_wav(os.path.join(REPO, "tests", "golden", "piano_test2.wav"))
        x = np.sign(pcm.astype(np.float64)) * 2.0 * np.abs(pcm.astype(np.float64)) / 65535.0
        blk = np.stack([x[1024 * 9:1024 * 11].T, x[1024 * 30:1024 * 32].T])
        states = [[0, 0], [100, 2000]]
        r = e64.encode_blocks(blk, states)
        print("encode_blocks ok: states", states, "chunk sizes", [[len(c) for c in cc] for cc in r["chunks"]])
        u = e64.unpack_blocks(r["chunks"])
        print("unpack == encode fields:", all(np.array_equal(u[k], r[k]) for k in ("sf", "ba", "mant", "oscale", "lrms", "tableID")))
        y = e64.decode_blocks(u["sf"], u["ba"], u["mant"], u["oscale"], u["lrms"])
        print("decode_blocks out", y.shape, float(np.max(np.abs(y))))
    except Exception:
        traceback.print_exc()

    section("throughput probe (fp32, synthetic noise+tones, device-resident)")
    try:
        import torch
        S, sec = 256, 10
        n = sec * 44100
        g = torch.Generator(device="cuda").manual_seed(1)
        t = torch.arange(n, device="cuda", dtype=torch.float32) / 44100.0
        sig = torch.zeros(S, n, 2, device="cuda")
        for k in range(4):
            f = 50.0 * (320.0 ** torch.rand(S, 1, 1, device="cuda", generator=g))
            amp = 10 ** (-(6 + 24 * torch.rand(S, 1, 2, device="cuda", generator=g)) / 20)
            sig += amp * torch.sin(2 * np.pi * f * t.view(1, n, 1))
        sig += 10 ** (-40 / 20) * torch.randn(S, n, 2, device="cuda", generator=g)
        pcm = (sig.clamp(-1, 1) * 32767).round().to(torch.int16).contiguous()
        cap = e32.encode_bound(n)
        out = torch.empty(S, cap, dtype=torch.uint8, device="cuda")
        for e, tag in ((e32, "fp32"), (e64, "fp64")):
            e.encode_batch(pcm, out=out, cap=cap)
            torch.cuda.synchronize()
            t0 = time.time()
            _, ob = e.encode_batch(pcm, out=out, cap=cap)
            torch.cuda.synchronize()
            dt = time.time() - t0
            print("%s: %d streams x %ds in %.3fs -> %.0f audio-s/s  (avg %.0f B/block)" % (tag, S, sec, dt, S * sec / dt, ob.sum() / (S * e.num_blocks(n))))
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()

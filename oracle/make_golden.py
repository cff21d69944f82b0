#!/usr/bin/env python
"""
make_golden.py -- generate tests/golden/* from the REAL reference, run here.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference);
the fixtures it writes are committed so that the GPU box, where the reference
tree does not exist, can still check the oracle and the CUDA path against the
reference's own outputs.

What it does
  1. imports the mechanically patched reference (oracle/ref_py3.py);
  2. encodes + decodes every inputs/*.wav with it (one process per file) and
     proves the 12 committed goldens (`coded/<n>.wak`, `outputs/<n>.wav`,
     SURVEY.md Appendix D) are reproduced byte for byte;
  3. writes tests/golden/manifest.json: sha256 + size of each input, of the
     reference .wak and of the reference decoded .wav, and the end-of-stream
     reservoir state (bitDeposit, extraBits);
  4. copies the small committed whole-file fixtures (piano_test2, castanets);
  5. dumps per-stage values for a few blocks of several files
     (tests/golden/stages.npz) -- the per-kernel parity tap points;
  6. writes tests/golden/kats.json -- known-answer vectors taken by calling the
     reference's own functions on the inputs of its `__main__` self-tests.

Usage:  python oracle/make_golden.py [--jobs 8] [--only stages|kats|files]
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import shutil
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, HERE)
import ref_py3  # noqa: E402

REF_INPUTS = os.path.join(ref_py3.REF_ROOT, "inputs")
REF_CODED = os.path.join(ref_py3.REF_ROOT, "coded")
REF_OUTPUTS = os.path.join(ref_py3.REF_ROOT, "outputs")

# the pairs HEAD reproduces
GOLDEN_NAMES = ["harmonic_test2", "harmonic_test4", "percussion_test1", "percussion_test2",
                "percussion_test3", "piano_test2", "piano_test3", "pop_test2", "rock",
                "rock_test2", "speech_test2", "speech_test3"]
COMMITTED_FILES = ["piano_test2", "castanets"]

# (file, first block index, count) for the stage dumps
STAGE_BLOCKS = [("castanets", 60, 3), ("rock", 200, 3), ("speech_test2", 300, 3),
                ("harmonic_test2", 100, 3), ("piano_test2", 0, 2), ("piano_test2", 688, 2),
                ("pop_test2", 500, 2)]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def _one_file(name):
    tmp = tempfile.mkdtemp(prefix="pac_gold_")
    try:
        with ref_py3.Ref() as R:
            wav = os.path.join(REF_INPUTS, name + ".wav")
            pac = os.path.join(tmp, name + ".wak")
            out = os.path.join(tmp, name + ".wav")
            sys.stdout = open(os.devnull, "w")  # bitalloc.py:178 prints '*'
            dep, extra = R.encode_file(wav, pac)
            R.decode_file(pac, out)
            sys.stdout = sys.__stdout__
            win = open(wav, "rb").read()
            pb = open(pac, "rb").read()
            ob = open(out, "rb").read()
            rec = {"name": name, "wav_sha256": sha(win), "wav_bytes": len(win),
                   "pac_sha256": sha(pb), "pac_bytes": len(pb),
                   "out_sha256": sha(ob), "out_bytes": len(ob),
                   "bitDeposit_end": int(dep), "extraBits_end": int(extra),
                   "committed_golden": name in GOLDEN_NAMES}
            if name in GOLDEN_NAMES:
                gp = open(os.path.join(REF_CODED, name + ".wak"), "rb").read()
                go = open(os.path.join(REF_OUTPUTS, name + ".wav"), "rb").read()
                rec["matches_committed_pac"] = (gp == pb)
                rec["matches_committed_out"] = (go == ob)
            if name in COMMITTED_FILES:
                shutil.copyfile(wav, os.path.join(GOLD, name + ".wav"))
                with open(os.path.join(GOLD, name + ".wak"), "wb") as f:
                    f.write(pb)
                with open(os.path.join(GOLD, name + ".out.wav"), "wb") as f:
                    f.write(ob)
            return rec
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_files(jobs):
    names = sorted(f[:-4] for f in os.listdir(REF_INPUTS) if f.endswith(".wav"))
    with mp.Pool(jobs) as pool:
        recs = pool.map(_one_file, names, chunksize=1)
    bad = [r["name"] for r in recs if r.get("committed_golden")
           and not (r["matches_committed_pac"] and r["matches_committed_out"])]
    if bad:
        raise SystemExit("patched reference does NOT reproduce committed goldens: %s" % bad)
    man = {"generator": "oracle/make_golden.py (patched reference, Python %s, NumPy %s)"
           % (sys.version.split()[0], np.__version__),
           "params": {"nMDCTLines": 1024, "nScaleBits": 4, "nMantSizeBits": 4,
                      "targetBitsPerSample": 2.27, "nTableIDBits": 4},
           "files": {r["name"]: r for r in recs}}
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    print("files: %d encoded+decoded; %d committed goldens reproduced byte-for-byte"
          % (len(recs), sum(1 for r in recs if r.get("committed_golden"))))


def make_kbd():
    """KBDWindow as the codec's window: the REFERENCE's own encode and decode with the name `SineWindow` in codec.py's
    namespace rebound to window.KBDWindow (codec.py:59-60 and :239-240 are its only two uses on the path) -- no source edited, the
    reference's KBDWindow (window.py:56-78: alpha = 4, returns a copy) does the windowing.  Pins the oracle's and the engine's
    PAC_WINDOW_KBD option on a 40-block clip of piano_test2 (tests/golden/kbd_piano.*)."""
    tmp = tempfile.mkdtemp(prefix="pac_gold_kbd_")
    try:
        with ref_py3.Ref() as R:
            import struct
            src = open(os.path.join(REF_INPUTS, "piano_test2.wav"), "rb").read()
            p = src.find(b"data")
            n = 40 * 1024 + 100
            body = src[p + 8:p + 8 + n * 4]
            hdr = bytearray(src[:p + 8])
            hdr[p + 4:p + 8] = struct.pack("<L", len(body))
            hdr[4:8] = struct.pack("<L", len(hdr) - 8 + len(body))
            wav = os.path.join(tmp, "kbd_piano.wav")
            open(wav, "wb").write(bytes(hdr) + body)
            assert R.codec.SineWindow is R.window.SineWindow
            R.codec.SineWindow = R.window.KBDWindow
            pac = os.path.join(tmp, "kbd_piano.wak")
            out = os.path.join(tmp, "kbd_piano.out.wav")
            sys.stdout = open(os.devnull, "w")
            dep, extra = R.encode_file(wav, pac)
            R.decode_file(pac, out)
            sys.stdout = sys.__stdout__
            for f in ("kbd_piano.wav", "kbd_piano.wak", "kbd_piano.out.wav"):
                shutil.copyfile(os.path.join(tmp, f), os.path.join(GOLD, f))
            with open(os.path.join(GOLD, "kbd_piano.json"), "w") as f:
                json.dump({"bitDeposit_end": int(dep), "extraBits_end": int(extra), "pac_bytes": os.path.getsize(pac),
                           "generator": "oracle/make_golden.py --only kbd: patched reference with codec.SineWindow = window.KBDWindow"}, f)
            print("kbd: %d bytes coded, final state (%d, %d)" % (os.path.getsize(pac), dep, extra))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_stages():
    """Per-stage taps, captured by wrapping the reference's own functions."""
    want = {}
    for name, b0, n in STAGE_BLOCKS:
        want.setdefault(name, set()).update(range(b0, b0 + n))
    dump = {}
    index = []
    tmp = tempfile.mkdtemp(prefix="pac_gold_")
    try:
        for name in sorted(want):
            with ref_py3.Ref() as R:
                cap = {}
                state = {"iblk": -1, "on": False}
                o_bthr = R.psychoac.calcBTHR
                o_gsm = R.codec.getStereoMaskThreshold
                o_edc = R.codec.EncodeDualChannel
                o_enc = R.codec.Encode
                o_ba = R.codec.BitAlloc

                def w_bthr(data, MDCTdata, MDCTscale, sampleRate, sfBands, noDrop):
                    r = o_bthr(data, MDCTdata, MDCTscale, sampleRate, sfBands, noDrop)
                    if state["on"]:
                        cap.setdefault("bthr", []).append(np.array(r, dtype=np.float64))
                    return r

                def w_gsm(data, MDCTdata, MDCTscale, sampleRate, sfBands, LRMS, cp):
                    if state["on"]:
                        cap["mdct"] = np.array(MDCTdata, dtype=np.float64)
                        cap["oscale"] = np.array(MDCTscale, dtype=np.int32)
                    smr, lines = o_gsm(data, MDCTdata, MDCTscale, sampleRate, sfBands, LRMS, cp)
                    if state["on"]:
                        cap["smr"] = np.array(smr, dtype=np.float64)
                        cap["lines"] = np.array(lines, dtype=np.float64)
                    return smr, lines

                def w_ba(bitBudget, extraBits, maxMantBits, nBands, nLines, SMR, LRMS):
                    if state["on"]:
                        cap.setdefault("ba_extra_in", []).append(int(extraBits))
                        cap["bitBudget"] = float(bitBudget)
                    r = o_ba(bitBudget, extraBits, maxMantBits, nBands, nLines, SMR, LRMS)
                    if state["on"]:
                        cap.setdefault("ba_diff", []).append(int(r[1]))
                    return r

                def w_edc(data, cp, LRMS, huffman):
                    r = o_edc(data, cp, LRMS, huffman)
                    if state["on"]:
                        sf, ba, mant, osc = r
                        cap["sf"] = np.array(sf, dtype=np.int32)
                        cap["ba"] = np.array(ba, dtype=np.int32)
                        for ch in range(2):
                            m = np.zeros(1024, dtype=np.int64)
                            m[:len(mant[ch])] = mant[ch]
                            cap["mant%d" % ch] = m
                            cap["nmant%d" % ch] = len(mant[ch])
                    return r

                def w_enc(data, cp, huffman):
                    if state["on"]:
                        cap["extraBits_in"] = int(cp.extraBits)
                        cap["bitDeposit_in"] = int(huffman.bitDeposit)
                    r = o_enc(data, cp, huffman)
                    if state["on"]:
                        cap["tableID"] = np.array(r[4], dtype=np.int32)
                        cap["lrms"] = np.array(r[6], dtype=np.int32)
                        cap["extraBits_out"] = int(cp.extraBits)
                        cap["bitDeposit_out"] = int(huffman.bitDeposit)
                    return r

                R.psychoac.calcBTHR = w_bthr
                R.codec.getStereoMaskThreshold = w_gsm
                R.codec.EncodeDualChannel = w_edc
                R.codec.Encode = w_enc
                R.codec.BitAlloc = w_ba

                wav = os.path.join(REF_INPUTS, name + ".wav")
                pac = os.path.join(tmp, name + ".wak")
                raw = open(wav, "rb").read()
                # locate PCM payload the same way pcmfile.py:32-57 does
                pos = raw.find(b"data", 12)
                nsamp = struct.unpack("<L", raw[pos + 4:pos + 8])[0] // 4
                body = raw[pos + 8:pos + 8 + nsamp * 4]
                body = body[:(len(body) // 4) * 4]
                pcm = np.frombuffer(body, dtype="<i2").reshape(-1, 2)

                # drive the reference's own loop so the reservoir state is genuine
                huffman = R.new_huffman()
                R.pacfile.huffman = huffman
                inFile = R.pcmfile.PCMFile(wav)
                outFile = R.pacfile.PACFile(pac)
                cp = inFile.OpenForReading()
                cp.nMDCTLines = 1024
                cp.nScaleBits = 4
                cp.nMantSizeBits = 4
                cp.targetBitsPerSample = 2.27
                cp.nTableIDBits = 4
                cp.nSamplesPerBlock = 1024
                outFile.OpenForWriting(cp)
                last = max(want[name])
                iblk = 0
                sys.stdout = open(os.devnull, "w")
                while iblk <= last:
                    data = inFile.ReadDataBlock(cp)
                    if not data:
                        break
                    state["on"] = iblk in want[name]
                    cap.clear()
                    p0 = outFile.fp.tell()
                    outFile.WriteDataBlock(data, cp, huffman)
                    if state["on"]:
                        outFile.fp.flush()
                        with open(pac, "rb") as f:
                            f.seek(p0)
                            chunk = f.read()
                        key = "%s.%d" % (name, iblk)
                        win = np.zeros((2048, 2), dtype=np.int16)
                        lo, hi = (iblk - 1) * 1024, (iblk + 1) * 1024
                        s0, s1 = max(lo, 0), min(hi, len(pcm))
                        if s1 > s0:
                            win[s0 - lo:s1 - lo] = pcm[s0:s1]
                        dump[key + ".pcm"] = win
                        dump[key + ".bthr"] = np.stack(cap["bthr"])  # L,R,M,S,M',S'
                        for k in ("mdct", "oscale", "smr", "lines", "sf", "ba", "mant0", "mant1",
                                  "tableID", "lrms"):
                            dump[key + "." + k] = cap[k]
                        dump[key + ".state"] = np.array(
                            [cap["extraBits_in"], cap["bitDeposit_in"], cap["extraBits_out"],
                             cap["bitDeposit_out"], cap["nmant0"], cap["nmant1"],
                             cap["ba_extra_in"][0], cap["ba_extra_in"][1],
                             cap["ba_diff"][0], cap["ba_diff"][1]], dtype=np.int64)
                        dump[key + ".chunk"] = np.frombuffer(chunk, dtype=np.uint8)
                        index.append(key)
                    iblk += 1
                sys.stdout = sys.__stdout__
                inFile.fp.close()
                outFile.fp.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    dump["index"] = np.array(index)
    np.savez_compressed(os.path.join(GOLD, "stages.npz"), **dump)
    print("stages: %d blocks dumped" % len(index))


def make_kats():
    """Known-answer vectors: the reference's own functions on the inputs of its
    `__main__` self tests (bitpack.py:183-190, quantize.py:37,383-451,
    psychoac.py:696-715) plus small MDCT / window / band-layout vectors."""
    k = {}
    with ref_py3.Ref() as R:
        # bitpack.py:183-190
        bp = R.bitpack.PackedBits()
        bp.Size(2)
        vals, widths = (3, 5, 11, 3, 1), (4, 3, 5, 3, 1)
        for v, w in zip(vals, widths):
            bp.WriteBits(v, w)
        packed = bp.GetPackedData()
        bp.ResetPointers()
        back = [int(bp.ReadBits(w)) for w in widths]
        k["bitpack"] = {"values": vals, "widths": widths, "bytes_hex": packed.hex(), "readback": back}
        # quantize.py:37
        x = np.array(R.quantize.chartInputs, dtype=np.float64)
        q = R.quantize
        k["quantize"] = {
            "inputs": x.tolist(),
            "vQuantizeUniform8": [int(v) for v in q.vQuantizeUniform(x, 8)],
            "vQuantizeUniform12": [int(v) for v in q.vQuantizeUniform(x, 12)],
            "vDequantizeUniform8": q.vDequantizeUniform(q.vQuantizeUniform(x, 8), 8).tolist(),
            "ScaleFactor_3_5": [int(q.ScaleFactor(v, 3, 5)) for v in x],
            "vMantissa_s0_3_5": [int(v) for v in q.vMantissa(x, 0, 3, 5)],
            "vDequantize_s0_3_5": q.vDequantize(0, q.vMantissa(x, 0, 3, 5), 3, 5).tolist(),
        }
        # a denser sweep of the block-floating-point quantiser, incl. the coder's own (4, ba) settings
        rng = np.random.default_rng(1234)
        xs = np.concatenate([rng.uniform(-1, 1, 64), 10.0 ** rng.uniform(-7, 0, 64) * rng.choice([-1, 1], 64),
                             [0.0, -0.0, 1.0, -1.0, 0.999999999, 1e-12]])
        sweep = []
        for ba in (2, 3, 5, 8, 12, 16):
            sfs = [int(q.ScaleFactor(v, 4, ba)) for v in xs]
            sfb = int(q.ScaleFactor(np.max(np.abs(xs)), 4, ba))
            mant = [int(v) for v in q.vMantissa(xs, sfb, 4, ba)]
            deq = q.vDequantize(sfb, np.array(mant, dtype=np.int64), 4, ba).tolist()
            mant3 = [int(v) for v in q.vMantissa(xs * 2.0 ** -9, 9, 4, ba)]
            deq3 = q.vDequantize(9, np.array(mant3, dtype=np.int64), 4, ba).tolist()
            sweep.append({"ba": ba, "ScaleFactor": sfs, "blockScale": sfb, "vMantissa": mant,
                          "vDequantize": deq, "vMantissa_sf9": mant3, "vDequantize_sf9": deq3})
        k["bfp_sweep"] = {"inputs": xs.tolist(), "cases": sweep}
        # mdct.py:49-88 small vectors
        k["mdct"] = {"MDCT_arange8_4_4": R.mdct.MDCT(np.arange(8.), 4, 4).tolist(),
                     "IMDCT_of_that": R.mdct.IMDCT(R.mdct.MDCT(np.arange(8.), 4, 4), 4, 4).tolist()}
        x2048 = np.sin(0.01 * np.arange(2048.) ** 1.1) * 0.5
        k["mdct"]["x2048_formula"] = "0.5*sin(0.01*arange(2048)**1.1)"
        xw = R.window.SineWindow(x2048.copy())
        X = R.mdct.MDCT(xw, 1024, 1024)
        k["mdct"]["MDCT_sine_x2048"] = X.tolist()
        k["mdct"]["IMDCT_MDCT_sine_x2048"] = R.mdct.IMDCT(X, 1024, 1024).tolist()
        # window.py
        k["window"] = {"KBDWindow_ones8": R.window.KBDWindow(np.ones(8)).tolist(),
                       "SineWindow_ones8": R.window.SineWindow(np.ones(8)).tolist(),
                       "HanningWindow_ones8": R.window.HanningWindow(np.ones(8)).tolist(),
                       "KBDWindow_ones2048_head": R.window.KBDWindow(np.ones(2048))[:16].tolist()}
        # psychoac.py:124-156 band layouts
        P = R.psychoac
        k["bands"] = {str(fs): [int(v) for v in P.AssignMDCTLinesFromFreqLimits(1024, fs)]
                      for fs in (44100, 48000, 22050, 32000)}
        k["bands"]["512@44100"] = [int(v) for v in P.AssignMDCTLinesFromFreqLimits(512, 44100)]
        # psychoac.py:15-64 scalar helpers
        fr = [10.0, 50.0, 100.0, 440.0, 1000.0, 3300.0, 8000.0, 16000.0, 22000.0]
        k["psy_scalar"] = {"f": fr, "Bark": [float(P.Bark(f)) for f in fr],
                           "Thresh": [float(P.Thresh(f)) for f in fr],
                           "SPL": [float(P.SPL(v)) for v in (1.0, 1e-3, 1e-13, 0.0)],
                           "Intensity": [float(P.Intensity(v)) for v in (96.0, 0.0, -30.0)]}
        # psychoac.py:696-713 six-tone signal, mono CalcSMRs
        FS, N = 48000, 1024
        n = np.arange(N)
        amps = (0.60, 0.11, 0.10, 0.08, 0.05, 0.03)
        freqs = (420., 530., 640., 840., 4200., 8400.)
        x = sum(a * np.cos(2 * np.pi * f * n / FS) for a, f in zip(amps, freqs))
        sfb = P.ScaleFactorBands(P.AssignMDCTLinesFromFreqLimits(N // 2, FS))
        mdct = R.mdct.MDCT(R.window.SineWindow(x.copy()), N // 2, N // 2)[:N // 2] * 16.0
        smr = P.CalcSMRs(x.copy(), mdct, 4, FS, sfb)
        thr = P.getMaskedThreshold(x.copy(), mdct, 4, FS, sfb)
        k["calcsmrs_sixtone"] = {"FS": FS, "N": N, "amps": amps, "freqs": freqs, "scale": 4,
                                 "SMR": np.asarray(smr).tolist(),
                                 "maskedThreshold": np.asarray(thr).tolist()}
        # bitalloc.py:129-184 on that SMR vector, plus reservoir cases
        ba_cases = []
        nLines = sfb.nLines
        for budget, extra, lr in ((1000.5, 0, 0), (2116.48, 137, 1), (300.0, -50, 0), (6000.0, 0, 1)):
            LRMS = np.full(25, lr, dtype=int)
            LRMS[::3] = 1 - lr
            sys.stdout = open(os.devnull, "w")
            bits, diff = R.bitalloc.BitAlloc(budget, extra, 16, 25, nLines, np.asarray(smr), LRMS)
            sys.stdout = sys.__stdout__
            ba_cases.append({"bitBudget": budget, "extraBits": extra, "LRMS": LRMS.tolist(),
                             "nLines": [int(v) for v in nLines], "SMR": np.asarray(smr).tolist(),
                             "bits": [int(b) for b in bits], "bitDifference": int(diff)})
        k["bitalloc"] = ba_cases
        # the allocators HEAD does not call (bitalloc.py:22-125), on the same SMR vector and on the band peak SPLs
        alt = []
        B = R.bitalloc
        peak = float(np.max(smr) + 20.0)
        for budget in (1000, 2116.48, 300.0, 6000.0, 25.0, 523 * 3):
            alt.append({"mode": "uniform", "bitBudget": budget, "nLines": [int(v) for v in nLines],
                        "bits": [int(b) for b in B.BitAllocUniform(budget, 16, 25, nLines)]})
        # BitAllocConstSNR / ConstMNR only terminate when the greedy loop lands on exactly zero remaining bits (otherwise
        # `while remaining_bits > 0` spins for ever, bitalloc.py:74,109).  The C restatement detects that state, so it is used
        # here to PICK budgets for which the reference returns; the recorded answers are the reference's own.
        import oracle as orc
        O = orc.get()
        for mode, fn, level in (("constmnr", B.BitAllocConstMNR, np.asarray(smr, dtype=float)), ("constsnr", B.BitAllocConstSNR, peak)):
            ok = []
            for budget in list(range(40, 4000, 37)) + [512 * 16, 512 * 20]:
                try:
                    O.bitalloc_alt(mode, budget, 16, 25, nLines, level)
                    ok.append(budget)
                except RuntimeError:
                    pass
            assert len(ok) >= 4, (mode, ok)
            for budget in [ok[i] for i in sorted(set(np.linspace(0, len(ok) - 1, 8).astype(int).tolist()))]:
                got = fn(budget, 16, 25, nLines, level.copy() if hasattr(level, "copy") else level)
                alt.append({"mode": mode, "bitBudget": budget, "nLines": [int(v) for v in nLines],
                            "level": np.broadcast_to(np.asarray(level, dtype=float), (25,)).tolist(), "bits": [int(b) for b in got]})
        k["bitalloc_alt"] = alt
        # Huffman tables: structure facts (Huffman.py:138-153, huffmanTables.pickle)
        h = R.new_huffman()
        k["huffman_tables"] = {str(i): {"nsym": len(t.encodingTable), "maxkey": max(t.encodingTable),
                                        "esc": t.encodingTable[-1],
                                        "code0": t.encodingTable.get(0),
                                        "maxlen": max(len(c) for c in t.encodingTable.values())}
                               for i, t in h.huffmanTables.items()}
        # Huffman.py:353-371 reservoir arithmetic
        seq = []
        for dep in (0, 5, 10, 11, 99, 100, 101, 250, 12345, -1, -77):
            h.bitDeposit = dep
            w = h.withdrawBits()
            seq.append([dep, int(w), int(h.bitDeposit)])
        k["withdraw"] = seq
        # Huffman trainer (Huffman.py:156-250) on seeded geometric codes, two countFreq calls, then a second trainer in the same
        # process (the class-level statistics dict keeps counting, :30-34)
        def trainer_codes(seed, n, p):
            g = np.random.default_rng(seed)
            c = g.geometric(p, n) - 1
            c[g.integers(0, n, n // 50)] = g.integers(0, 30000, n // 50)       # rare large codes -> escapes
            return c
        H = R.Huffman
        cwd = os.getcwd()
        os.chdir(R.dir)
        try:
            c1 = trainer_codes(20161018, 40000, 0.03)
            t1 = H.HuffmanTrainer(11)
            t1.countFreq([int(v) for v in c1[:25000]])
            t1.countFreq([int(v) for v in c1[25000:]])
            t1.constructHuffmanTable()
            tab1 = dict(t1.huffmanCodeTable)
            c2 = trainer_codes(7, 20000, 0.2)
            t2 = H.HuffmanTrainer(12)
            t2.countFreq([int(v) for v in c2])
            t2.constructHuffmanTable()
            tab2 = dict(t2.huffmanCodeTable)
            with open("huffmanTables.pickle", "rb") as fh:
                import pickle
                stored = pickle.load(fh)
            assert stored[11].encodingTable == tab1 and stored[12].encodingTable == tab2 and len(stored) == 12
        finally:
            os.chdir(cwd)
        k["huffman_trainer"] = {"gen": "rng=default_rng(seed); c=geometric(p,n)-1; c[rng.integers(0,n,n//50)]=rng.integers(0,30000,n//50)",
                                "first": {"seed": 20161018, "n": 40000, "p": 0.03, "split": 25000,
                                          "table": {str(a): b for a, b in tab1.items()}},
                                "second": {"seed": 7, "n": 20000, "p": 0.2, "table": {str(a): b for a, b in tab2.items()}}}
        # (getMatchScore, Huffman.py:50-62, is Python-2-only as written -- np.array(dict.values()) -- and HEAD never calls it, so it
        # is not exercised here; with the class-level probability dict it compares the dict with itself and returns 3.0.)
    with open(os.path.join(GOLD, "kats.json"), "w") as f:
        json.dump(k, f, indent=0)
    print("kats written")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if not ref_py3.available():
        raise SystemExit("reference tree not present; fixtures can only be generated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    if a.only in ("", "kats"):
        make_kats()
    if a.only in ("", "stages"):
        make_stages()
    if a.only in ("", "kbd"):
        make_kbd()
    if a.only in ("", "files"):
        make_files(a.jobs)


if __name__ == "__main__":
    main()

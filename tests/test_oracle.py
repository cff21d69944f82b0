"""
Pins the CPU oracle (oracle/pac_oracle.c) against the reference:
  * the reference's own self-test vectors (tests/golden/kats.json, made by calling the reference's functions),
  * per-stage dumps of the reference running on real audio (tests/golden/stages.npz),
  * whole-file goldens: tests/golden/manifest.json holds sha256 of what the reference writes for every
    inputs/*.wav (12 of them equal to the reference's committed coded/*.wak + outputs/*.wav).
CPU only.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import corpus_files

NL44 = [5, 4, 5, 5, 5, 5, 7, 7, 7, 9, 10, 11, 13, 15, 17, 21, 26, 32, 42, 51, 61, 83, 116, 163, 304]


def sha(b):
    return hashlib.sha256(b).hexdigest()


# ---------------------------------------------------------------- KATs

def test_band_layouts(oracle, kats):
    for fs in ("44100", "48000", "22050", "32000"):
        assert list(oracle.band_layout(1024, int(fs))) == kats["bands"][fs]
    assert list(oracle.band_layout(512, 44100)) == kats["bands"]["512@44100"]
    assert kats["bands"]["44100"] == NL44


def test_quantizer_kats(oracle, kats):
    q = kats["quantize"]
    x = np.array(q["inputs"])
    assert list(oracle.vquantize_uniform(x, 8)) == q["vQuantizeUniform8"]
    assert list(oracle.vquantize_uniform(x, 12)) == q["vQuantizeUniform12"]
    assert [oracle.scale_factor(v, 3, 5) for v in x] == q["ScaleFactor_3_5"]
    assert list(oracle.vmantissa(x, 0, 3, 5)) == q["vMantissa_s0_3_5"]
    got = oracle.vdequantize(0, q["vMantissa_s0_3_5"], 3, 5)
    np.testing.assert_array_equal(got, np.array(q["vDequantize_s0_3_5"]))
    assert np.signbit(got[3]) and got[3] == 0          # -0.0 survives (Q22)
    np.testing.assert_array_equal(oracle.vdequantize_uniform(q["vQuantizeUniform8"], 8),
                                  np.array(q["vDequantizeUniform8"]))


def test_bfp_sweep(oracle, kats):
    s = kats["bfp_sweep"]
    x = np.array(s["inputs"])
    for c in s["cases"]:
        ba = c["ba"]
        assert [oracle.scale_factor(v, 4, ba) for v in x] == c["ScaleFactor"]
        sfb = oracle.scale_factor(np.max(np.abs(x)), 4, ba)
        assert sfb == c["blockScale"]
        assert list(oracle.vmantissa(x, sfb, 4, ba)) == c["vMantissa"]
        np.testing.assert_array_equal(oracle.vdequantize(sfb, c["vMantissa"], 4, ba), np.array(c["vDequantize"]))
        assert list(oracle.vmantissa(x * 2.0 ** -9, 9, 4, ba)) == c["vMantissa_sf9"]
        np.testing.assert_array_equal(oracle.vdequantize(9, c["vMantissa_sf9"], 4, ba), np.array(c["vDequantize_sf9"]))


def test_mdct_kats(oracle, kats):
    m = kats["mdct"]
    np.testing.assert_allclose(oracle.mdct(np.arange(8.), 4, 4), m["MDCT_arange8_4_4"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(oracle.imdct(m["MDCT_arange8_4_4"], 4, 4), m["IMDCT_of_that"], rtol=0, atol=1e-13)
    x = 0.5 * np.sin(0.01 * np.arange(2048.) ** 1.1)
    X = oracle.mdct(oracle.sine_window(x), 1024, 1024)
    ref = np.array(m["MDCT_sine_x2048"])
    assert np.max(np.abs(X - ref)) <= 1e-14 * np.max(np.abs(ref)) + 1e-17
    xr = oracle.imdct(ref, 1024, 1024)
    ref2 = np.array(m["IMDCT_MDCT_sine_x2048"])
    assert np.max(np.abs(xr - ref2)) <= 1e-13 * np.max(np.abs(ref2))


def test_window_kats(oracle, kats):
    w = kats["window"]
    np.testing.assert_allclose(oracle.sine_window(np.ones(8)), w["SineWindow_ones8"], rtol=0, atol=1e-16)
    np.testing.assert_allclose(oracle.hann_window(np.ones(8)), w["HanningWindow_ones8"], rtol=0, atol=1e-16)
    np.testing.assert_allclose(oracle.kbd_window(np.ones(8)), w["KBDWindow_ones8"], rtol=1e-12, atol=1e-16)
    np.testing.assert_allclose(oracle.kbd_window(np.ones(2048))[:16], w["KBDWindow_ones2048_head"], rtol=1e-11, atol=1e-18)


def test_psy_scalars(oracle, kats):
    p = kats["psy_scalar"]
    L = oracle.lib
    np.testing.assert_allclose([L.orc_bark(f) for f in p["f"]], p["Bark"], rtol=1e-15)
    np.testing.assert_allclose([L.orc_thresh(f) for f in p["f"]], p["Thresh"], rtol=1e-14)
    np.testing.assert_allclose([L.orc_spl(v) for v in (1.0, 1e-3, 1e-13, 0.0)], p["SPL"], rtol=1e-15)
    np.testing.assert_allclose([L.orc_intensity(v) for v in (96.0, 0.0, -30.0)], p["Intensity"], rtol=1e-15)


def test_calcsmrs_sixtone(oracle, kats):
    """psychoac.py:696-713 test signal through the mono CalcSMRs path (SURVEY Appendix D vector)."""
    k = kats["calcsmrs_sixtone"]
    FS, N = k["FS"], k["N"]
    n = np.arange(N)
    x = sum(a * np.cos(2 * np.pi * f * n / FS) for a, f in zip(k["amps"], k["freqs"]))
    nLines = oracle.band_layout(N // 2, FS)
    mdct = oracle.mdct(oracle.sine_window(x), N // 2, N // 2) * 16.0
    smr, thr = oracle.calc_smrs(x, mdct, 4, FS, nLines)
    np.testing.assert_allclose(thr, k["maskedThreshold"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(smr, k["SMR"], rtol=0, atol=1e-9)
    assert [round(v, 4) for v in smr[:4]] == [-1.7574, 13.2098, 0.6205, 13.295]


def test_bitalloc_kats(oracle, kats):
    for c in kats["bitalloc"]:
        bits, diff = oracle.bitalloc(c["bitBudget"], c["extraBits"], 16, 25, c["nLines"], c["SMR"], c["LRMS"])
        assert list(bits) == c["bits"]
        assert diff == c["bitDifference"]


def test_bitalloc_alt_kats(oracle, kats):
    """bitalloc.py:22-125 (BitAllocUniform / ConstSNR / ConstMNR): the reference's own answers, and the non-terminating state."""
    for c in kats["bitalloc_alt"]:
        bits = oracle.bitalloc_alt(c["mode"], c["bitBudget"], 16, 25, c["nLines"], c.get("level"))
        assert list(bits) == c["bits"], (c["mode"], c["bitBudget"])
    c = kats["bitalloc_alt"][-1]
    with pytest.raises(RuntimeError):                       # 0.48 bits can never be placed: the reference spins for ever
        oracle.bitalloc_alt("constmnr", 2116.48, 16, 25, c["nLines"], c["level"])


def _trainer_codes(seed, n, p):
    g = np.random.default_rng(seed)
    c = g.geometric(p, n) - 1
    c[g.integers(0, n, n // 50)] = g.integers(0, 30000, n // 50)
    return c


def test_huffman_trainer_kats(kats):
    """Huffman.py:156-250: the restated trainer reproduces the tables the reference built here, including the second trainer of
    a process (class-level statistics / queue / table carry over)."""
    import oracle as orc
    h = kats["huffman_trainer"]
    f = h["first"]
    c1 = _trainer_codes(f["seed"], f["n"], f["p"])
    t1, st = orc.train_huffman([c1[:f["split"]], c1[f["split"]:]])
    assert {str(a): b for a, b in t1.items()} == f["table"]
    g = h["second"]
    t2, st = orc.train_huffman([_trainer_codes(g["seed"], g["n"], g["p"])], state=st)
    assert {str(a): b for a, b in t2.items()} == g["table"]
    # prefix-free, escape present
    codes = sorted(t1.values())
    assert all(not codes[i + 1].startswith(codes[i]) for i in range(len(codes) - 1)) and -1 in t1


def test_huffman_table_facts(oracle, kats):
    for tid, f in kats["huffman_tables"].items():
        t = oracle.tables[int(tid)]
        assert len(t) == f["nsym"] and max(t) == f["maxkey"] and t[-1] == f["esc"] and t.get(0) == f["code0"]
    assert oracle.tables[5][-1] == "01000101"


# ---------------------------------------------------------------- per-stage dumps of the reference

def _keys(stages):
    return [str(k) for k in stages["index"]]


def test_stage_mdct_and_scale(oracle, stages):
    for key in _keys(stages):
        pcm = stages[key + ".pcm"].astype(np.float64)
        ref = stages[key + ".mdct"]
        osc = stages[key + ".oscale"]
        for ch in range(2):
            x = np.sign(pcm[:, ch]) * 2.0 * np.abs(pcm[:, ch]) / 65535.0
            X = oracle.mdct(oracle.sine_window(x), 1024, 1024)
            s = oracle.scale_factor(np.max(np.abs(X)), 4)
            assert s == osc[ch], key
            X *= (1 << s)
            assert np.max(np.abs(X - ref[ch])) <= 1e-13 * max(np.max(np.abs(ref[ch])), 1e-300), key


def test_stage_lrms_smr_lines(oracle, stages):
    nLines = np.array(NL44, dtype=np.int32)
    for key in _keys(stages):
        pcm = stages[key + ".pcm"].astype(np.float64)
        x = np.sign(pcm) * 2.0 * np.abs(pcm) / 65535.0
        lrms = oracle.lrms(x[:, 0], x[:, 1], nLines)
        assert list(lrms) == list(stages[key + ".lrms"]), key
        d0, d1 = oracle.sine_window(x[:, 0]), oracle.sine_window(x[:, 1])
        smr, lines, bthr = oracle.stereo_smr(d0, d1, stages[key + ".mdct"][0], stages[key + ".mdct"][1],
                                             stages[key + ".oscale"], 44100, nLines, lrms)
        np.testing.assert_allclose(bthr, stages[key + ".bthr"], rtol=0, atol=1e-9, err_msg=key)
        np.testing.assert_allclose(smr, stages[key + ".smr"], rtol=0, atol=1e-9, err_msg=key)
        np.testing.assert_array_equal(lines, stages[key + ".lines"], err_msg=key)


def test_stage_alloc_quant_chunk(oracle, stages, gold_dir):
    """ba / sf / mantissas / tableID / reservoir / chunk bytes of the dumped blocks, via the stream trace."""
    import oracle as omod
    files = corpus_files()
    by_file = {}
    for key in _keys(stages):
        name, blk = key.rsplit(".", 1)
        by_file.setdefault(name, []).append(int(blk))
    done = 0
    for name, blks in by_file.items():
        if name not in files:
            continue
        rate, pcm = omod.read_wav(files[name])
        last = max(blks)
        pcm = pcm[:(last + 1) * 1024]         # blocks 0..last see exactly the same samples
        data, tr, _ = oracle.encode_stream(pcm, omod.default_params(rate), trace=True)
        off = 76 + np.concatenate([[0], np.cumsum((tr["nbytes"] + 4).reshape(-1))])
        for b in blks:
            key = "%s.%d" % (name, b)
            np.testing.assert_array_equal(tr["ba"][b], stages[key + ".ba"], err_msg=key)
            np.testing.assert_array_equal(tr["sf"][b], stages[key + ".sf"], err_msg=key)
            np.testing.assert_array_equal(tr["tableID"][b], stages[key + ".tableID"], err_msg=key)
            np.testing.assert_array_equal(tr["oscale"][b], stages[key + ".oscale"], err_msg=key)
            st = stages[key + ".state"]
            assert tr["extraBits"][b] == st[2] and tr["bitDeposit"][b] == st[3], key
            for ch in range(2):
                nm = int(st[4 + ch])
                packed = tr["mant"][b, ch][np.repeat(tr["ba"][b, ch] > 0, NL44)]
                np.testing.assert_array_equal(packed, stages[key + ".mant%d" % ch][:nm], err_msg=key)
            chunk = data[off[2 * b]:off[2 * b + 2]]
            assert chunk == stages[key + ".chunk"].tobytes(), key
            done += 1
    assert done >= 5


# ---------------------------------------------------------------- whole files

@pytest.mark.parametrize("name", ["piano_test2", "castanets"])
def test_committed_wholefile_goldens(oracle, gold_dir, manifest, name):
    enc, _, fs = oracle.encode_wav(os.path.join(gold_dir, name + ".wav"))
    gold = open(os.path.join(gold_dir, name + ".wak"), "rb").read()
    assert enc == gold
    rec = manifest["files"][name]
    assert sha(gold) == rec["pac_sha256"] and len(gold) == rec["pac_bytes"]
    assert fs == (rec["bitDeposit_end"], rec["extraBits_end"])
    dec = oracle.decode_to_wav_bytes(gold)
    assert dec == open(os.path.join(gold_dir, name + ".out.wav"), "rb").read()
    assert sha(dec) == rec["out_sha256"]


def test_castanets_end_state(manifest):
    # SURVEY.md Appendix D end-state checks
    assert (manifest["files"]["castanets"]["bitDeposit_end"], manifest["files"]["castanets"]["extraBits_end"]) == (11755, 404306)
    assert (manifest["files"]["piano_test2"]["bitDeposit_end"], manifest["files"]["piano_test2"]["extraBits_end"]) == (6175, 3803)
    assert sum(1 for r in manifest["files"].values() if r.get("committed_golden") and r["matches_committed_pac"]
               and r["matches_committed_out"]) == 12


@pytest.mark.slow
def test_full_corpus_against_reference_hashes(oracle, manifest):
    """All inputs/*.wav (config 2): oracle bytes == reference bytes (sha256 from the manifest)."""
    files = corpus_files()
    names = [n for n in manifest["files"] if n in files]
    if len(names) <= 2:
        pytest.skip("tests/golden/_corpus not populated")
    from concurrent.futures import ThreadPoolExecutor

    def one(n):
        rec = manifest["files"][n]
        assert sha(open(files[n], "rb").read()) == rec["wav_sha256"]
        enc, _, fs = oracle.encode_wav(files[n])          # ctypes drops the GIL: files run in parallel
        assert len(enc) == rec["pac_bytes"] and sha(enc) == rec["pac_sha256"], n
        assert fs == (rec["bitDeposit_end"], rec["extraBits_end"]), n
        assert sha(oracle.decode_to_wav_bytes(enc)) == rec["out_sha256"], n
        return n

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        assert sorted(ex.map(one, names)) == sorted(names)


# ---------------------------------------------------------------- edge cases

def test_edge_streams(oracle):
    import oracle as omod
    # empty stream: ceil(0/1024)+1 = 1 block of silence (Q20), header numSamples += 1024 (Q16: 0 % 1024 == 0)
    enc, tr, _ = oracle.encode_stream(np.zeros((0, 2), np.int16), trace=True)
    assert enc[:4] == b"PAC " and len(tr["lrms"]) == 1
    assert int.from_bytes(enc[10:14], "little") == 1024
    pcm, rate, ns = oracle.decode_stream(enc)
    assert pcm.shape == (1024, 2) and not pcm.any()
    # ragged length + full-scale + -32768 (Q21)
    rng = np.random.default_rng(7)
    x = rng.integers(-32768, 32767, size=(3000, 2), dtype=np.int16)
    x[5] = (-32768, 32767)
    enc, tr, _ = oracle.encode_stream(x, trace=True)
    assert len(tr["lrms"]) == 4 and int.from_bytes(enc[10:14], "little") == 3000
    pcm, rate, ns = oracle.decode_stream(enc)
    assert pcm.shape == (4 * 1024, 2)
    # L == R  =>  every band with energy goes M/S (Q11) and S codes carry +0 sign bits
    y = (8000 * np.sin(0.05 * np.arange(4096))).astype(np.int16)
    enc, tr, _ = oracle.encode_stream(np.stack([y, y], 1), trace=True)
    assert tr["lrms"][1] == (1 << 25) - 1
    assert not tr["mant"][1:4, 1].any()
    # threaded batch == serial
    xs = np.stack([x, x[::-1].copy()])
    outs = oracle.encode_batch(xs, nthreads=2)
    assert outs[0] == enc_ser(oracle, x) and outs[1] == enc_ser(oracle, x[::-1].copy())


def enc_ser(oracle, x):
    return oracle.encode_stream(x)[0]


def test_kbd_window_option_matches_reference_golden(oracle, gold_dir):
    """PAC_WINDOW_KBD: tests/golden/kbd_piano.* were produced by the reference's own encode/decode with the name
    SineWindow in codec.py rebound to window.KBDWindow (oracle/make_golden.py:make_kbd).  KBDWindow returns a copy, so the
    psychoacoustic model sees the un-windowed block; the C oracle restates that and must reproduce the files byte for byte."""
    import json
    import oracle as omod
    rate, pcm = omod.read_wav(os.path.join(gold_dir, "kbd_piano.wav"))
    meta = json.load(open(os.path.join(gold_dir, "kbd_piano.json")))
    enc, _, fs = oracle.encode_stream(pcm, omod.default_params(rate, window=1))
    want = open(os.path.join(gold_dir, "kbd_piano.wak"), "rb").read()
    assert len(enc) == meta["pac_bytes"] and enc == want
    assert fs == (meta["bitDeposit_end"], meta["extraBits_end"])
    dec, sr, ns = oracle.decode_stream(want, window=1)
    assert omod.wav_bytes(dec, sr, ns) == open(os.path.join(gold_dir, "kbd_piano.out.wav"), "rb").read()
    # and the option really changes the stream (the sine-window encode of the same clip differs)
    assert oracle.encode_stream(pcm, omod.default_params(rate))[0] != want

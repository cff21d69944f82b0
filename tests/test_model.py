"""The restructured analysis algebra (tests/model_analysis.py == what csrc/analysis.cuh computes) reproduces the
reference's own per-stage dumps.  CPU only."""
import numpy as np

import model_analysis as ma

NL44 = [5, 4, 5, 5, 5, 5, 7, 7, 7, 9, 10, 11, 13, 15, 17, 21, 26, 32, 42, 51, 61, 83, 116, 163, 304]


def test_pieces_against_numpy():
    rng = np.random.default_rng(3)
    x = rng.standard_normal(2048)
    np.testing.assert_allclose(ma.rfft_packed(x), np.fft.fft(x)[:1025], rtol=0, atol=2e-12)
    n = np.arange(2048)
    h = 0.5 * (1 - np.cos(2.0 * (n + 0.5) * np.pi / 2048))
    F = np.fft.fft(x)[:1025]
    np.testing.assert_allclose(ma.hann_taps(F, 2048), np.fft.fft(h * x)[:1025], rtol=0, atol=2e-12)
    np.testing.assert_allclose(ma.hann_taps(ma.hann_taps(F, 2048), 2048), np.fft.fft(h * h * x)[:1025], rtol=0, atol=2e-12)
    k = np.arange(1024)
    ref = (2. / 2048) * np.real(np.fft.fft(x * np.exp(-1j * np.pi * n / 2048))[:1024]
                                * np.exp(1j * (-2. * np.pi / 2048) * 512.5 * (k + 0.5)))
    np.testing.assert_allclose(ma.mdct_fold_fft(x), ref, rtol=0, atol=1e-13)


def test_model_matches_reference_stage_dumps(stages):
    T = ma.Tables()
    for key in [str(k) for k in stages["index"]]:
        pcm = stages[key + ".pcm"].astype(np.float64)
        x = np.sign(pcm) * 2.0 * np.abs(pcm) / 65535.0
        lrms, osc, X, bthr, smr, lines = ma.analysis(T, x[:, 0].copy(), x[:, 1].copy(), NL44)
        assert list(lrms) == list(stages[key + ".lrms"]), key
        assert list(osc) == list(stages[key + ".oscale"]), key
        ref = stages[key + ".mdct"]
        assert np.max(np.abs(X - ref)) <= 1e-12 * max(np.max(np.abs(ref)), 1e-300), key
        np.testing.assert_allclose(bthr, stages[key + ".bthr"], rtol=0, atol=1e-8, err_msg=key)
        np.testing.assert_allclose(smr, stages[key + ".smr"], rtol=0, atol=1e-8, err_msg=key)
        rl = stages[key + ".lines"]
        assert np.max(np.abs(lines - rl)) <= 1e-12 * max(np.max(np.abs(rl)), 1e-300), key


def test_scan_based_curve_matches_direct_evaluation(stages):
    """fp32 fast path (csrc/analysis.cuh:masked_curve_fast): scans over the bins for the fixed-slope skirts, prefix-sum plateau,
    pairwise upper skirts only for maskers above 40 dB.  Against the direct masker x line evaluation
    (psychoac.py:431-456) in float64: tolerance 1e-4 dB = the 1e-5 * max(|SMR|, 10 dB) budget of north_star.
    The loud skirts are culled per 64-line half-chunk as in the kernel (30 bits below the half-chunk's smallest partial threshold)."""
    T = ma.Tables()
    G = ma.Geometry(T)
    worst, stats = 0.0, {}
    for key in [str(k) for k in stages["index"]][::3]:
        pcm = stages[key + ".pcm"].astype(np.float64)
        x = np.sign(pcm) * 2.0 * np.abs(pcm) / 65535.0
        for sig, drop in ((x[:, 0], 15.0), (0.5 * (x[:, 0] - x[:, 1]), 0.0)):
            F = np.fft.fft(sig * T.sine * T.hann)[:T.M + 1]
            ref = ma.curve(T, F, drop)
            got = ma.curve_v3(T, G, F, drop, dtype=np.float32, stats=stats)
            worst = max(worst, float(np.max(np.abs(got.astype(np.float64) - ref))))
    print("worst |dB error|", worst, stats)
    assert worst <= 1e-4


def test_flat_pack_positions_equal_sequential_layout():
    """k_pack places every field of a chunk from ONE running token-bit prefix P (csrc/pack.cuh): header of band b at S[b] + P(lo_b),
    its sign run at S[b] + hdr + P(lo_b), the token of line i at S[b] + hdr + signs_b + P(i), the LRMS bits after everything.
    Checked here against the layout written band after band as PACFile.WriteDataBlock does (pacfile.py:324-348), for random band
    layouts including empty bands, zero-bit bands and 32 bands."""
    rng = np.random.default_rng(8)
    for trial in range(200):
        NB = int(rng.integers(1, 33))
        M = int(rng.choice([512, 1024]))
        cuts = np.sort(rng.integers(0, M + 1, NB - 1)) if NB > 1 else np.array([], int)
        lo = np.concatenate([[0], cuts, [M]]).astype(int)                     # lo[b] .. lo[b+1]-1, empty bands allowed
        nl = np.diff(lo)
        ba = rng.integers(0, 17, NB)
        ba[ba == 1] = 0
        if trial % 5 == 0:
            ba[:] = 0
        hdr, pos0 = 8, 8
        tok = np.zeros(M, int)
        for b in range(NB):
            if ba[b]:
                tok[lo[b]:lo[b + 1]] = rng.integers(1, 38, nl[b])            # code length, escapes up to 37 bits
        # sequential layout
        pos = pos0
        want_hdr, want_sgn, want_tok = np.zeros(NB, int), np.full(NB, -1), np.full(M, -1)
        for b in range(NB):
            want_hdr[b] = pos
            pos += hdr
            if not ba[b]:
                continue
            want_sgn[b] = pos
            pos += nl[b]
            for i in range(lo[b], lo[b + 1]):
                want_tok[i] = pos
                pos += tok[i]
        want_lrms = pos
        # flat evaluation
        fixed = hdr + np.where(ba > 0, nl, 0)
        S = pos0 + np.concatenate([[0], np.cumsum(fixed)[:-1]])
        P = np.concatenate([[0], np.cumsum(tok)])                              # P[i] = token bits before line i; P[M] = total
        got_hdr = S + P[lo[:-1]]
        assert np.array_equal(got_hdr, want_hdr)
        for b in range(NB):
            if ba[b]:
                assert S[b] + hdr + P[lo[b]] == want_sgn[b]
                i = np.arange(lo[b], lo[b + 1])
                assert np.array_equal(S[b] + fixed[b] + P[i], want_tok[i])
        assert pos0 + fixed.sum() + P[M] == want_lrms


def _sortable32(v):
    u = np.float32(v).view(np.uint32)
    return int(~u & 0xffffffff) if (int(u) >> 31) else int(u | 0x80000000)


def test_incremental_bitalloc_loop_equals_plain_restatement():
    """The fp32 water-filling loop of k_scan (csrc/scan.cuh: warp_bitalloc32) carries per-band sortable keys between iterations and
    exits on a zero maximum; here its control flow is replayed in numpy float32 against a plain float32 restatement of
    bitalloc.BitAlloc (bitalloc.py:129-184): same bits and same bitDifference for random SMRs, budgets, reservoirs and LRMS masks."""
    rng = np.random.default_rng(12)
    nl = np.array([5, 4, 5, 5, 5, 5, 7, 7, 7, 9, 10, 11, 13, 15, 17, 21, 26, 32, 42, 51, 61, 83, 116, 163, 304])
    NB, maxMant = len(nl), 16
    f32 = np.float32
    for trial in range(300):
        smr = rng.uniform(-60, 50, NB).astype(f32)
        if trial % 9 == 0:
            smr[:] = smr[0]                                                   # ties: first index wins
        if trial % 11 == 0:
            smr[rng.integers(0, NB, 3)] = f32(-96)
        lrms = rng.integers(0, 2, NB)
        extra = int(rng.integers(-400, 4000))
        total0 = int(2116.48 + extra)
        # plain restatement (float32 arithmetic as the kernel's: v = fma(bits, -6, smr))
        bits = np.zeros(NB, int); valid = np.ones(NB, bool); total = total0
        while valid.any():
            v = (smr - f32(6) * bits.astype(f32)).astype(f32)
            cand = np.where(valid, v, f32(-np.inf))
            iMax = int(np.argmax(cand))
            mx = f32(np.max((v + f32(6)).astype(f32)))
            if mx < (f32(-5) if lrms[iMax] else f32(-15)):
                valid[iMax] = False
            if total - nl[iMax] >= 0:
                bits[iMax] += 1; total -= nl[iMax]
                if bits[iMax] >= maxMant:
                    valid[iMax] = False
            else:
                valid[iMax] = False
        total += nl[bits == 1].sum(); bits[bits == 1] = 0
        want = (bits.copy(), total - extra)
        # the kernel's loop
        bits = np.zeros(NB, int); total = total0
        vkey = [_sortable32(x) for x in smr]
        k2 = [_sortable32(f32(x) + f32(6)) for x in smr]
        kMS, kLR = _sortable32(-5.0), _sortable32(-15.0)
        while True:
            mk1 = max(vkey)
            if mk1 == 0:
                break
            iMax = vkey.index(mk1)
            below = max(k2) < (kMS if lrms[iMax] else kLR)
            afford = total >= nl[iMax]
            if afford:
                total -= nl[iMax]
            ok = (not below) and afford
            if afford:
                bits[iMax] += 1
                if bits[iMax] >= maxMant:
                    ok = False
            v = f32(np.float32(smr[iMax]) - f32(6) * f32(bits[iMax]))
            vkey[iMax] = _sortable32(v) if ok else 0
            k2[iMax] = _sortable32(f32(v + f32(6)))
        total += nl[bits == 1].sum(); bits[bits == 1] = 0
        assert np.array_equal(bits, want[0]) and total - extra == want[1], trial


def test_event_driven_bitalloc_equals_plain_loop():
    """csrc/scan.cuh: warp_bitalloc_jump (stated in tests/model_bitalloc.py) == the plain one-bit-per-iteration loop of
    bitalloc.BitAlloc (bitalloc.py:157-184), in the reference's float64 arithmetic and in the fast mode's float32 keys, on random
    and adversarial problems: 6 dB lattice ties, SMRs sitting on the stop thresholds and on the jump margins, maxed-out bands with
    NMR above the thresholds, empty bands, all-M/S, all-L/R, budgets from a few hundred bits to a 500 k-bit reservoir."""
    import model_bitalloc as mb
    rng = np.random.default_rng(7)
    it_plain = it_fast = 0
    for trial in range(1200):
        T = np.float64 if trial % 2 == 0 else np.float32
        kind = trial % 13
        smr = rng.uniform(-60, 50, 25)
        if kind == 1:
            smr[:] = smr[0]
        elif kind == 2:
            smr = np.round(smr / 6) * 6 + rng.choice([0, 0, 0, 1e-9, -1e-9], 25)
        elif kind == 3:
            smr[rng.integers(0, 25, 3)] = -96.0
        elif kind == 4:
            smr = rng.uniform(-30, -5, 25)
        elif kind == 5:
            smr = np.round(rng.uniform(-30, 10, 25) * 2) / 2
        elif kind == 6:
            smr = rng.uniform(60, 110, 25)
        elif kind == 7:
            smr = rng.uniform(-25, -9, 25)
        elif kind == 8:
            smr = rng.choice([-11., -5., -17., -21., -15., -10.5, -11.5, -20.5, -21.5], 25)
        lrms = rng.integers(0, 2, 25)
        if kind == 9:
            lrms[:] = 0
        elif kind == 10:
            lrms[:] = 1
        extra = int(rng.choice([rng.integers(-4000, 4000), rng.integers(0, 200), rng.integers(3000, 500000), 0]))
        total0 = int(float(rng.choice([2116.48, 1382.0, 5840.3, 300.5])) + extra)
        smrT = smr.astype(T)
        b1, d1, i1 = mb.plain(total0, extra, smrT, lrms, T=T)
        b2, d2, i2, _ = mb.fast(total0, extra, smrT, lrms, T=T)
        assert np.array_equal(b1, b2) and d1 == d2, (trial, kind, extra, total0)
        it_plain += i1; it_fast += i2
    print("single iterations per problem: plain %.1f, event-driven %.1f" % (it_plain / 1200.0, it_fast / 1200.0))
    assert it_fast * 5 < it_plain


def test_synth_shared_memory_swizzles_are_conflict_free():
    """k_synth's register-blocked IMDCT (decode.cuh) addresses its float2 workspace through ZI(i) = i ^ ((i >> 3) & 15) and its float
    DCT-IV output through VI(i) = i ^ ((i >> 2) & 31).  Enumerate every access pattern of the kernel's 1024-line path and count
    shared-memory wavefronts (32 four-byte banks; an 8-byte access is served per half-warp): each load and store must take the minimum.
    The additive paddings used before (one spare element per 8 / per 16) took 1.5x and 2.7x as many."""
    M, H = 1024, 512
    ZI = lambda i: i ^ ((i >> 3) & 15)
    VI = lambda i: i ^ ((i >> 2) & 31)

    def wf(addrs, banks):
        per = {}
        for a in set(addrs):
            per.setdefault(a % banks, set()).add(a)
        return max(len(v) for v in per.values())

    assert len({ZI(i) for i in range(H)}) == H and max(ZI(i) for i in range(H)) < H          # permutations of the rows: no padding needed
    assert len({VI(i) for i in range(M)}) == M and max(VI(i) for i in range(M)) < M

    def cost_z(Z):
        c = 0
        for hw in range(16):                                          # pre-twiddle stores: n = tid, tid + 256
            for k in range(2):
                c += wf([Z(16 * hw + l + 256 * k) for l in range(16)], 16)
        for hw in range(4):                                           # 64 threads per channel
            for r in range(8):
                t = [16 * hw + l for l in range(16)]
                c += wf([Z(x + 64 * r) for x in t], 16)                               # pass 1: t + 64 r (loads and stores)
                c += wf([Z(64 * (x >> 3) + (x & 7) + 8 * r) for x in t], 16)         # pass 2: 64 g + j + 8 r
                c += wf([Z(8 * x + r) for x in t], 16)                                # pass 3: 8 t + r
        return c

    def cost_v(V):
        c = 0
        for h in range(2):                                            # last pass: v[2k], v[M-1-2k], k = 8 (t & 7) + (t >> 3) + 64 p
            for p in range(8):
                ks = [((32 * h + l) >> 3) + 8 * ((32 * h + l) & 7) + 64 * p for l in range(32)]
                c += wf([V(2 * k) for k in ks], 32) + wf([V(M - 1 - 2 * k) for k in ks], 32)
        for warp in range(8):                                         # unfold: ascending / descending runs
            for j in range(4):
                i = [warp * 32 + l + 256 * j for l in range(32)]
                c += wf([V(x + H if x < H else 3 * H - 1 - x) for x in i], 32)
                c += wf([V(H - 1 - x if x < H else x - H) for x in i], 32)
        return c

    assert cost_z(ZI) == 32 + 3 * 32                                 # one wavefront per half-warp access
    assert cost_v(VI) == 32 + 64                                     # one wavefront per warp access
    assert cost_z(lambda i: i + (i >> 3)) == 192 and cost_v(lambda i: i + (i >> 4)) == 256   # what the paddings cost

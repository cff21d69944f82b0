"""
model_bitalloc.py -- numpy model of the event-driven water-filling the scan kernel uses (csrc/scan.cuh: warp_bitalloc_jump).

Test infrastructure.  bitalloc.BitAlloc (bitalloc.py:129-184) hands out one bit per iteration: ~100-150 iterations per
channel, each a pair of 25-way maxima -- the serial heart of the per-stream reservoir chain.  The greedy arg-max over
SMR_b - 6*bits_b is a merge of 25 strictly decreasing sequences, so as long as nothing "happens" the state after all entries
with key >= tau have been served is known in closed form (bits_b = number of keys of band b that are >= tau).  Things that
happen, and how the kernel jumps over them:

  * the budget runs out          -> tau is raised (bisection on the exact cost) until the whole cut is affordable;
  * the max-NMR stop rule fires  -> a band's visits are classified per jump as NORMAL (the rule provably does not fire:
    (:163-168)                      key + 6 >= threshold + margin, or an invalidated band holds the maximum above it) or
                                    TERMINAL (it provably fires: the band takes one more bit -- allocate-after-invalidate,
                                    :169-174 -- and leaves); jumps never cross the +-0.5 dB zones around the thresholds,
                                    where the outcome depends on rounding: those entries go through the exact loop;
  * 16 bits reached              -> the band leaves (its sequence simply ends).

Additionally a band whose width exceeds the remaining budget can only ever be invalidated (the budget never grows inside
the loop), so it is dropped at once ("prune").  `fast` below must give the same (bits, bitDifference) as `plain` for every
input; tests/test_model.py runs both on random and adversarial problems in float64 (the reference's arithmetic, used by the
fp64 verification mode) and float32 (the fast mode's key arithmetic).
"""
import numpy as np

NL44 = np.array([5, 4, 5, 5, 5, 5, 7, 7, 7, 9, 10, 11, 13, 15, 17, 21, 26, 32, 42, 51, 61, 83, 116, 163, 304])


def key(smr_b, j, T):
    """SMR - 6*bits as the loop forms it (one rounding: 6*j is exact)"""
    return T(smr_b - T(6) * T(j))


def e2(smr_b, bits_b, T):
    """SMR - 6*(bits-1): float64 as bitalloc.py:165 writes it; float32 as the fast kernel does (key + 6, rounded again)"""
    if T is np.float64:
        return smr_b - (bits_b - 1) * 6.
    return T(key(smr_b, bits_b, T) + T(6))


def plain(total0, extra, smr, lrms, nl=NL44, maxb=16, T=np.float64):
    """bitalloc.py:157-184 restated; returns (bits, bitDifference, iterations)"""
    NB = len(nl)
    bits = np.zeros(NB, int)
    valid = np.ones(NB, bool)
    total = total0
    it = 0
    smr = smr.astype(T)
    while valid.any():
        it += 1
        v = np.array([key(smr[b], bits[b], T) for b in range(NB)])
        iMax = int(np.argmax(np.where(valid, v, T(-np.inf))))
        mx = max(e2(smr[b], bits[b], T) for b in range(NB))
        if mx < (T(-5) if lrms[iMax] else T(-15)):
            valid[iMax] = False
        if total - nl[iMax] >= 0:
            bits[iMax] += 1
            total -= nl[iMax]
            if bits[iMax] >= maxb:
                valid[iMax] = False
        else:
            valid[iMax] = False
    total += nl[bits == 1].sum()
    bits[bits == 1] = 0
    return bits, total - extra, it


def count_ge(smr_b, tau, lo, cap, T):
    """largest n in [lo, cap] with key(b, j) >= tau for all lo <= j < n (keys strictly decrease with j): a closed-form guess
    fixed up with the loop's own key expression"""
    if cap <= lo:
        return lo
    if tau == -np.inf:
        return cap
    g = np.floor((float(smr_b) - float(tau)) * (1.0 / 6.0))
    n = int(min(max(g, -1.0), 64.0)) + 1
    n = min(max(n, lo), cap)
    # the guess is off by at most one (one rounding in the division): a single corrective step, as the kernel does it branch-free
    if n > lo and key(smr_b, n - 1, T) < tau:
        n -= 1
    elif n < cap and key(smr_b, n, T) >= tau:
        n += 1
    assert (n == lo or key(smr_b, n - 1, T) >= tau) and (n == cap or key(smr_b, n, T) < tau), (smr_b, tau, lo, cap, n)
    return n


def try_jump(bits, valid, total, smr, lrms, nl, maxb, T, probes=6):
    """One event-free advance from the loop state (bits, valid, total).  Returns the new state and the number of bits handed out."""
    NB = len(nl)
    vb = [b for b in range(NB) if valid[b]]
    if not vb:
        return bits, valid, total, 0
    ms = [b for b in vb if lrms[b]]
    lr = [b for b in vb if not lrms[b]]
    m = max(key(smr[b], bits[b], T) for b in vb)
    inv = [e2(smr[b], bits[b], T) for b in range(NB) if not valid[b]]
    F = max(inv) if inv else T(-np.inf)
    NINF = T(-np.inf)
    floor = NINF
    termMS = termLR = False
    if ms:
        if m >= T(-10.5):
            floor = max(floor, T(-10.5))
        elif m < T(-11.5) and F < T(-5.5):
            termMS = True
        else:
            return bits, valid, total, 0
    if lr:
        k1 = max(key(smr[b], bits[b], T) for b in ms) if (ms and termMS) else NINF
        FB = max(F, k1)
        if FB >= T(-14.5):
            pass                                   # an invalidated band keeps max NMR above -15: the L/R rule never fires
        elif m >= T(-20.5):
            floor = max(floor, T(-20.5))
        elif m < T(-21.5) and FB < T(-15.5):
            termLR = True
        else:
            return bits, valid, total, 0
    if not (floor < m or floor == NINF):
        return bits, valid, total, 0
    cap = np.array([(bits[b] + 1 if (termMS if lrms[b] else termLR) else maxb) if valid[b] else bits[b] for b in range(NB)])
    cap = np.minimum(cap, maxb)

    def state(tau):
        nb = bits.copy()
        for b in vb:
            nb[b] = count_ge(smr[b], tau, bits[b], cap[b], T)
        return nb, int(np.sum((nb - bits) * nl))

    nb, cost = state(floor)
    if cost > total:
        hi = T(m + T(1))                           # nothing is >= hi: cost 0
        lo = floor if floor != NINF else T(min(key(smr[b], maxb - 1, T) for b in vb) - T(1))
        nb, cost = bits.copy(), 0
        for _ in range(probes):
            mid = T((lo + hi) * T(0.5))
            n2, c2 = state(mid)
            if c2 <= total:
                hi, nb, cost = mid, n2, c2
            else:
                lo = mid
    adv = int((nb - bits).sum())
    if adv == 0:
        return bits, valid, total, 0
    nv = valid.copy()
    for b in vb:
        if nb[b] >= maxb or ((termMS if lrms[b] else termLR) and nb[b] > bits[b]):
            nv[b] = False
    return nb, nv, total - cost, adv


def fast(total0, extra, smr, lrms, nl=NL44, maxb=16, T=np.float64):
    """the kernel's control flow: prune, try to jump, else one exact iteration.  Returns (bits, bitDifference, exact iterations, jumps)"""
    NB = len(nl)
    bits = np.zeros(NB, int)
    valid = np.ones(NB, bool)
    total = total0
    it = jumps = 0
    smr = smr.astype(T)
    can = True
    while True:
        valid = valid & (nl <= total)             # the budget never grows inside the loop: such a band can only be invalidated
        if not valid.any():
            break
        if can:
            bits, valid, total, adv = try_jump(bits, valid, total, smr, lrms, nl, maxb, T)
            if adv:
                jumps += 1
                continue
        can = True
        it += 1
        v = np.array([key(smr[b], bits[b], T) for b in range(NB)])
        iMax = int(np.argmax(np.where(valid, v, T(-np.inf))))
        mx = max(e2(smr[b], bits[b], T) for b in range(NB))
        if mx < (T(-5) if lrms[iMax] else T(-15)):
            valid[iMax] = False
        if total - nl[iMax] >= 0:
            bits[iMax] += 1
            total -= nl[iMax]
            if bits[iMax] >= maxb:
                valid[iMax] = False
        else:
            valid[iMax] = False
    total += nl[bits == 1].sum()
    bits[bits == 1] = 0
    return bits, total - extra, it, jumps

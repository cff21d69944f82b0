"""gpu_perf.py -- per-kernel timing probe + full-corpus parity sweep on the GPU box (diagnostic, not pytest)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "perceptual-audio-codec_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import _pacb200  # noqa: E402
import oracle as orc  # noqa: E402


def synth(S, sec, seed=1, noise_db=None):
    import torch
    n = sec * 44100
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.arange(n, device="cuda", dtype=torch.float32) / 44100.0
    sig = torch.zeros(S, n, 2, device="cuda")
    for k in range(5):
        f = 50.0 * (320.0 ** torch.rand(S, 1, 1, device="cuda", generator=g))
        amp = 10 ** (-(6 + 24 * torch.rand(S, 1, 2, device="cuda", generator=g)) / 20)
        sig += amp * torch.sin(2 * np.pi * f * t.view(1, n, 1))
    nd = (25 + 25 * torch.rand(S, 1, 2, device="cuda", generator=g)) if noise_db is None else torch.full((S, 1, 2), float(noise_db), device="cuda")
    sig += 10 ** (-nd / 20) * torch.randn(S, n, 2, device="cuda", generator=g)
    return (sig.clamp(-1, 1) * 32767).round().to(torch.int16).contiguous()


def perf(S=2048, sec=20):
    import torch
    pcm = synth(S, sec)
    n = pcm.shape[1]
    for prec in ("fp32", "fp64"):
        if prec == "fp64":
            pcm2 = pcm[:S // 8].contiguous()
        else:
            pcm2 = pcm
        e = _pacb200.Engine(0, prec)
        cap = e.encode_bound(n)
        out = torch.empty(pcm2.shape[0], cap, dtype=torch.uint8, device="cuda")
        e.encode_batch(pcm2, out=out, cap=cap)
        torch.cuda.synchronize()
        e.timing(True)
        t0 = time.time()
        _, ob = e.encode_batch(pcm2, out=out, cap=cap)
        torch.cuda.synchronize()
        dt = time.time() - t0
        tm = e.timing_get()
        e.timing(False)
        nblk = pcm2.shape[0] * e.num_blocks(n)
        print("%s encode: S=%d x %ds  wall %.3fs -> %.0f audio-s/s ; per kernel (ms, launches): %s ; blocks %d ; %.1f ns/block analysis"
              % (prec, pcm2.shape[0], sec, dt, pcm2.shape[0] * sec / dt, {k: v for k, v in tm.items() if v[1]}, nblk,
                 1e6 * tm["analysis"][0] / nblk), flush=True)
        # decode throughput
        host = out.cpu().numpy()
        pacs = [host[s, :ob[s]].tobytes() for s in range(min(pcm2.shape[0], 256))]
        e.decode_batch(pacs)
        e.timing(True)
        t0 = time.time()
        e.decode_batch(pacs)
        dt = time.time() - t0
        tm = e.timing_get()
        print("%s decode (host in/out): %d streams wall %.3fs -> %.0f audio-s/s ; %s" % (prec, len(pacs), dt, len(pacs) * sec / dt,
              {k: v for k, v in tm.items() if v[1]}), flush=True)
        e.close()


def corpus():
    man = json.load(open(os.path.join(REPO, "tests", "golden", "manifest.json")))
    cdir = os.path.join(REPO, "tests", "golden", "_corpus")
    names = sorted(n for n in man["files"] if os.path.exists(os.path.join(cdir, n + ".wav")))
    if not names:
        print("corpus not present")
        return
    pcms = [orc.read_wav(os.path.join(cdir, n + ".wav"))[1] for n in names]
    L = max(len(p) for p in pcms)
    batch = np.zeros((len(names), L, 2), np.int16)
    ns = np.zeros(len(names), np.int64)
    for i, p in enumerate(pcms):
        batch[i, :len(p)] = p
        ns[i] = len(p)
    e = _pacb200.Engine(0, "fp64")
    t0 = time.time()
    outs = e.encode_batch(batch, nSamples=ns)
    t1 = time.time()
    okp = 0
    for n, o in zip(names, outs):
        r = man["files"][n]
        good = hashlib.sha256(o).hexdigest() == r["pac_sha256"]
        okp += good
        if not good:
            print("  ENCODE MISMATCH", n, len(o), r["pac_bytes"])
    fs = e.last_final_state
    okf = sum(1 for i, n in enumerate(names) if (int(fs[i][0]), int(fs[i][1])) == (man["files"][n]["bitDeposit_end"], man["files"][n]["extraBits_end"]))
    t2 = time.time()
    dec = e.decode_batch(outs)
    t3 = time.time()
    okd = 0
    for n, (p, rate, hn) in zip(names, dec):
        r = man["files"][n]
        good = hashlib.sha256(orc.wav_bytes(p, rate, hn)).hexdigest() == r["out_sha256"]
        okd += good
        if not good:
            print("  DECODE MISMATCH", n)
    print("corpus fp64: %d files, encode byte-exact %d, final state exact %d, decode byte-exact %d ; encode %.2fs decode %.2fs (%.1f s audio)"
          % (len(names), okp, okf, okd, t1 - t0, t3 - t2, ns.sum() / 44100.0), flush=True)
    e32 = _pacb200.Engine(0, "fp32")
    outs32 = e32.encode_batch(batch, nSamples=ns)
    same = sum(1 for a, b in zip(outs, outs32) if a == b)
    print("corpus fp32: %d of %d files byte-identical to fp64; size ratio %.5f" % (same, len(names), sum(map(len, outs32)) / sum(map(len, outs))))


def perf_split(S=2048, sec=20):
    import torch
    e = _pacb200.Engine(0, "fp32")
    for tag, nd in (("mixed", None), ("loud noise -25 dBFS", 25), ("quiet noise -50 dBFS", 50), ("no noise", 200)):
        pcm = synth(S, sec, noise_db=nd)
        n = pcm.shape[1]
        cap = e.encode_bound(n)
        out = torch.empty(S, cap, dtype=torch.uint8, device="cuda")
        e.encode_batch(pcm, out=out, cap=cap)
        e.timing(True)
        torch.cuda.synchronize()
        t0 = time.time()
        e.encode_batch(pcm, out=out, cap=cap)
        torch.cuda.synchronize()
        wall = time.time() - t0
        tm = e.timing_get()
        e.timing(False)
        nblk = S * e.num_blocks(n)
        print("%-22s wall %.1f ns/block | kernels (overlapped): analysis %.1f  scan %.1f  pack %.1f" % (tag, 1e9 * wall / nblk, 1e6 * tm["analysis"][0] / nblk, 1e6 * tm["scan"][0] / nblk, 1e6 * tm["pack"][0] / nblk), flush=True)


def perf_one(noise_db, S=296, sec=5):
    import torch
    e = _pacb200.Engine(0, "fp32")
    pcm = synth(S, sec, noise_db=noise_db)
    n = pcm.shape[1]
    cap = e.encode_bound(n)
    out = torch.empty(S, cap, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        e.encode_batch(pcm, out=out, cap=cap)
    torch.cuda.synchronize()
    print("one ok")


if __name__ == "__main__":
    if "one" in sys.argv:
        perf_one(float(sys.argv[sys.argv.index("one") + 1]))
        sys.exit(0)
    if "split" in sys.argv:
        perf_split()
    if "corpus" in sys.argv or len(sys.argv) == 1:
        corpus()
    if "perf" in sys.argv or len(sys.argv) == 1:
        perf()

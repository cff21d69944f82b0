"""Decode-side perf probe (not a test): encode a small synthetic corpus on the device, decode it in place twice.
usage: python tests/gpu_decode_perf.py [streams] [seconds] [precision]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "perceptual-audio-codec_b200"))
sys.path.insert(0, ROOT)
import _pacb200  # noqa: E402
from bench import gen_streams  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 296
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = torch.device("cuda", 0)
n = int(sec * 44100)
eng = _pacb200.Engine(0, prec)
pcm = gen_streams(list(range(S)), n, dev)
cap = eng.encode_bound(n)
out = torch.empty(S, cap, dtype=torch.uint8, device=dev)
_, ob = eng.encode_batch(pcm, out=out, cap=cap)
nblk = eng.num_blocks(n)
stride = nblk * 1024 + 1024
dec = torch.empty(S, stride, 2, dtype=torch.int16, device=dev)
beg = np.arange(S, dtype=np.int64) * cap
for it in range(3):
    eng.timing(True)
    torch.cuda.synchronize()
    t0 = time.time()
    ns, _, _ = eng.decode_batch_strided(out, beg, ob, dec, stride)
    torch.cuda.synchronize()
    dt = time.time() - t0
    tm = eng.timing_get()
    print("decode %d x %.0f s: %.2f ms wall, %.0f audio-s/s, %.1f ns/block; kernels ms %s" % (
        S, sec, dt * 1e3, S * sec / dt, dt * 1e9 / (S * nblk), {k: round(v[0], 3) for k, v in tm.items() if v[1]}))
err = (dec[:, :n].int() - pcm.int()).abs().max().item()
print("max |decoded - original| =", err)

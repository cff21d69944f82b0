"""gpu_tiny.py -- a few seconds of every kernel (fp32 and fp64 encode, decode, KBD option), small enough for compute-sanitizer
(diagnostic).  usage: compute-sanitizer --tool racecheck python tests/gpu_tiny.py [streams] [seconds]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "perceptual-audio-codec_b200"))
import _pacb200  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
n = int(sec * 44100)
rng = np.random.default_rng(5)
t = np.arange(n) / 44100.0
pcm = np.zeros((S, n, 2), np.int16)
for s in range(S):
    x = 0.3 * np.sin(2 * np.pi * (200 + 300 * s) * t)[:, None] * np.array([1.0, 0.7])[None, :]
    x += 10 ** (-(25 + 4 * s) / 20) * rng.standard_normal((n, 2))
    pcm[s] = np.clip(np.round(x * 32767), -32767, 32767).astype(np.int16)
for prec, win in (("fp32", "sine"), ("fp64", "sine"), ("fp32", "kbd")):
    e = _pacb200.Engine(0, prec, window=win)
    pacs = e.encode_batch(pcm)
    dec = e.decode_batch(pacs)
    print(prec, win, [len(p) for p in pacs][:3], dec[0][0].shape)
    e.close()
print("tiny ok")

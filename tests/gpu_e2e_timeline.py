"""e2e timeline probe: host-pinned buffers through pac_encode_batch with PAC_TIMELINE=1"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "perceptual-audio-codec_b200")); sys.path.insert(0, ROOT)
import _pacb200
from bench import gen_streams
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sec = 60.0
dev = torch.device("cuda", 0)
n = int(sec * 44100)
pcm_h = torch.empty(S, n, 2, dtype=torch.int16, pin_memory=True)
for c in range(0, S, 64):
    pcm_h[c:c + 64].copy_(gen_streams(list(range(c, min(c + 64, S))), n, dev))
eng = _pacb200.Engine(0, "fp32")
cap = eng.encode_bound(n)
out_h = torch.empty(S, cap, dtype=torch.uint8, pin_memory=True)
ph, oh = pcm_h.numpy(), out_h.numpy()
eng.encode_batch(ph, out=oh, cap=cap)
torch.cuda.synchronize()
os.environ["PAC_TIMELINE"] = "1"
t0 = time.time()
eng.encode_batch(ph, out=oh, cap=cap)
torch.cuda.synchronize()
dt = time.time() - t0
print("e2e %.3f s -> %.0f audio-s/s" % (dt, S * sec / dt))
if len(sys.argv) > 2:      # same streams, device-resident
    Sd = int(sys.argv[2])
    pcm_d = pcm_h[:Sd].to(dev)
    out_d = torch.empty(Sd, cap, dtype=torch.uint8, device=dev)
    for _ in range(2):
        t0 = time.time()
        eng.encode_batch(pcm_d, out=out_d, cap=cap)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print("device-resident %d streams %.3f s -> %.0f audio-s/s" % (Sd, dt, Sd * sec / dt))

"""Where does the host-buffer path lose time against the device-resident one?  Same 4096 streams, four combinations of host/device
input and output buffers (diagnostic; PAC_TIMELINE-free wall clock of the second call of each kind)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "perceptual-audio-codec_b200")); sys.path.insert(0, ROOT)
import _pacb200
from corpus import gen_streams
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = 60 * 44100
dev = torch.device("cuda", 0)
pcm_d = torch.empty(S, n, 2, dtype=torch.int16, device=dev)
for c in range(0, S, 64):
    pcm_d[c:c + 64] = gen_streams(list(range(c, min(c + 64, S))), n, dev)
eng = _pacb200.Engine(0, "fp32")
cap = eng.encode_bound(n)
out_d = torch.empty(S, cap, dtype=torch.uint8, device=dev)
ph = _pacb200.pinned_empty((S, n, 2), np.int16)
oh = _pacb200.pinned_empty((S, cap), np.uint8)
torch.from_numpy(ph).copy_(pcm_d)
torch.cuda.synchronize()
for name, p, o in (("device in, device out", pcm_d, out_d), ("device in, pinned host out", pcm_d, oh), ("pinned host in, device out", ph, out_d),
                   ("pinned host in, pinned host out", ph, oh), ("device in, device out", pcm_d, out_d)):
    eng.encode_batch(p, out=o, cap=cap)
    torch.cuda.synchronize()
    eng.timing(True)
    t0 = time.time()
    eng.encode_batch(p, out=o, cap=cap)
    torch.cuda.synchronize()
    dt = time.time() - t0
    tm = eng.timing_get()
    eng.timing(False)
    print("%-34s %.1f ms   %s" % (name, dt * 1e3, ", ".join("%s %.0f" % (k, v[0]) for k, v in tm.items() if v[1])))

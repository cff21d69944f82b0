"""Does a host->device DMA stream running BESIDE the kernels slow the device-resident encode? (diagnostic)
The copies must be enqueued after the call has uploaded its per-stream state: a pageable H2D copy queued behind gigabytes of DMA on
the same copy engine would wait for them (that, not interference, is what a first version of this probe measured: 792 -> 1563 ms)."""
import os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "perceptual-audio-codec_b200")); sys.path.insert(0, ROOT)
import _pacb200
from corpus import gen_streams
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = 60 * 44100
dev = torch.device("cuda", 0)
pcm = torch.empty(S, n, 2, dtype=torch.int16, device=dev)
for c in range(0, S, 64):
    pcm[c:c + 64] = gen_streams(list(range(c, min(c + 64, S))), n, dev)
eng = _pacb200.Engine(0, "fp32")
cap = eng.encode_bound(n)
out = torch.empty(S, cap, dtype=torch.uint8, device=dev)
src = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream()
def run(copies):
    torch.cuda.synchronize()
    res = {}
    def enc():
        t0 = time.time()
        eng.encode_batch(pcm, out=out, cap=cap)          # ctypes releases the GIL; returns when the call's streams have drained
        res["dt"] = time.time() - t0
    th = threading.Thread(target=enc)
    th.start()
    time.sleep(0.03)                                     # the call is past its set-up and its kernels are running
    if copies:
        with torch.cuda.stream(side):
            for _ in range(copies):
                dst.copy_(src, non_blocking=True)
    th.join()
    torch.cuda.synchronize()
    return res["dt"]
run(0)
for copies in (0, 30, 0, 30):
    print("H2D copies beside the kernels: %2d GB -> encode %.1f ms" % (copies, run(copies) * 1e3))

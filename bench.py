#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric on BASELINE.json's config.

  metric   : audio-seconds encoded per wall-second (whole job, all GPUs)
  workload : configs[2] = synthetic corpus of 4096 x 60 s 44.1 kHz stereo int16 streams (tones + noise + transients),
             fp32 fast mode; at N > 1 the SAME corpus sharded by stream, stream s -> rank s mod N (configs[3]),
             NCCL used only to gather the per-stream byte counts.
  value    : corpus already resident in HBM when the timed region starts (K steps, CUDA-event timed, max over ranks)
  e2e      : the same corpus through the C-ABI call with HOST (pinned) buffers: H2D of the PCM and D2H of the .pac
             images inside the timed region
  roofline : dominant kernel (analysis = window + MDCT + M/S decision + SMR): algorithmic bytes (12 496 B per stereo
             block in fp32 mode) / CUDA-event duration of its launches, against the measured HBM peak
  cpu_baseline / --impl reference : the CPU oracle port (oracle/pac_oracle.c, validated byte-for-byte against the
             reference on all 22 inputs) on all host cores, on a bounded sample of the same corpus.

One JSON line on stdout (rank 0).  `python bench.py` defaults to N=1 and finishes within a few minutes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "perceptual-audio-codec_b200"))

FS = 44100
M = 1024
ALGO_BYTES_FP32 = 12496      # 4096 B PCM in + 8192 B lines + 200 B SMR + 8 B scale/LRMS out
ALGO_BYTES_FP64 = 20888


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------- synthetic corpus

from corpus import gen_streams, gen_streams_numpy  # noqa: E402  (integer Philox generator: identical samples under numpy, torch-CPU and torch-CUDA)


# ---------------------------------------------------------------- clocks sampler (B200_PROFILING.md)

class Clocks(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def bind_near_gpu(local):
    """Host-side plumbing for the e2e leg: run this rank's host thread on the CPUs of the NUMA node its GPU hangs off, so that the
    pinned buffers it allocates next are node-local (first touch) and the DMA does not cross the socket interconnect -- with eight ranks
    pulling 43 GB of PCM per step out of host memory that is the difference between a PCIe-bound and a UPI-bound copy.  Returns
    (previous affinity, note); never fatal."""
    try:
        old = os.sched_getaffinity(0)
    except Exception as e:
        return None, "affinity unsupported: %s" % e
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        if hasattr(pr, "pci_bus_id"):
            bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, getattr(pr, "pci_device_id", 0))
        else:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = vis.split(",")[local] if vis else str(local)
            bus = subprocess.run(["nvidia-smi", "-i", idx, "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True,
                                 timeout=10).stdout.strip().lower()
            if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
                bus = bus[4:]
        def node_cpus(k):
            cs = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % k).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cs.update(range(int(lo), int(hi or lo) + 1))
            return cs
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        how = "sysfs"
        if node < 0:
            # the OS does not say (VMs, containers): place by measurement -- a 128 MB pinned buffer allocated while running on each
            # node's CPUs, its host -> device copy timed; the fastest node wins
            import _pacb200
            nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
            best = None
            seen = []
            for k in nodes:
                cs = node_cpus(k) & old
                if not cs:
                    continue
                os.sched_setaffinity(0, cs)
                buf = _pacb200.pinned_empty((128 << 20,), np.uint8)
                buf[::4096] = 1                            # touch every page on this node
                gbs = max(_pacb200.h2d_bandwidth(buf, local) for _ in range(2))
                del buf
                seen.append("node%d %.1f" % (k, gbs))
                if best is None or gbs > best[1]:
                    best = (k, gbs)
            os.sched_setaffinity(0, old)
            if best is None or len(nodes) < 2:
                return old, "GPU %s reports no NUMA node and the host shows %d node(s)" % (bus, len(nodes))
            node, how = best[0], "probed H2D GB/s: " + ", ".join(seen)
        cpus = node_cpus(node) & old
        if not cpus:
            return old, "no allowed CPU on NUMA node %d of GPU %s" % (node, bus)
        os.sched_setaffinity(0, cpus)
        return old, "host thread bound to NUMA node %d (%d CPUs) of GPU %s [%s]" % (node, len(cpus), bus, how)
    except Exception as e:
        try:
            os.sched_setaffinity(0, old)
        except Exception:
            pass
        return old, "not bound: %s" % e


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per analysis launch from the committed ncu --set full capture, if one has been summarised."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ---------------------------------------------------------------- CPU baseline (oracle port) -- test infrastructure

def cpu_baseline_run(pcm_host, seconds_each, cores):
    """Encode `cores` streams of `seconds_each` s with the C oracle on `cores` host threads.  Returns audio-s/s."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle as orc
    O = orc.get()
    n = int(seconds_each * FS)
    batch = np.ascontiguousarray(pcm_host[:, :n])
    t0 = time.time()
    coded = O.encode_batch(batch, nthreads=cores)
    dt = time.time() - t0
    cpu_baseline_run.coded = coded
    return batch.shape[0] * seconds_each / dt, dt


def parity_leg(pcm_host, seconds_each, cores, device_index):
    """north_star's two correctness modes, measured on the streams the cpu_baseline leg has just encoded with the oracle and
    reported in the JSON line: (a) fp64 verification mode: .pac bytes == the oracle's, stream by stream; (b) fp32 fast mode:
    mismatch rates of the quantities the reference defines (M/S decision codec.py:96-102, overall scale :245, bitAlloc :258,
    scaleFactor :274, mantissa codes :276-277, tableID Huffman.py:309) and the worst MDCT-line / SMR error in units of 1e-5
    relative, against the oracle's per-block trace of a sub-sample (the traced oracle is single-threaded per stream)."""
    import _pacb200
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle as orc
    O = orc.get()
    n = int(seconds_each * FS)
    batch = np.ascontiguousarray(pcm_host[:, :n])
    want = cpu_baseline_run.coded
    e64 = _pacb200.Engine(device_index, "fp64")
    got = e64.encode_batch(batch)
    e64.close()
    same = [g == w for g, w in zip(got, want)]
    out = {"fp64_bytes_equal_oracle": bool(all(same)), "fp64_streams_checked": len(same), "fp64_streams_equal": int(sum(same)),
           "fp64_bytes_checked": int(sum(len(w) for w in want))}
    # fp32 fast mode against the traced oracle: the first 4 s of every sampled stream
    nt = min(n, 4 * FS)
    sub = np.ascontiguousarray(batch[:, :nt])
    with ThreadPoolExecutor(max_workers=cores) as ex:
        otr = list(ex.map(lambda s: O.encode_stream(sub[s], trace=True)[1], range(sub.shape[0])))
    e32 = _pacb200.Engine(device_index, "fp32")
    _, tr = e32.encode_batch(sub, trace=True)
    e32.close()
    cnt = {k: 0 for k in ("lrms", "oscale", "ba", "sf", "tableID", "mant")}
    mis = dict(cnt)
    worst_line = worst_smr = 0.0
    smr_over = smr_n = 0
    nl = np.array([5, 4, 5, 5, 5, 5, 7, 7, 7, 9, 10, 11, 13, 15, 17, 21, 26, 32, 42, 51, 61, 83, 116, 163, 304])
    band = np.repeat(np.arange(25), nl)
    for s, o in enumerate(otr):
        nb = len(o["lrms"])
        bits = ((o["lrms"][:, None] >> np.arange(25)[None, :]) & 1)
        gbits = ((tr["lrms"][s][:nb, None] >> np.arange(25)[None, :]) & 1)
        mis["lrms"] += int(np.sum(bits != gbits)); cnt["lrms"] += bits.size
        for k in ("oscale", "ba", "sf", "tableID"):
            mis[k] += int(np.sum(tr[k][s][:nb] != o[k])); cnt[k] += o[k].size
        coded = (np.repeat(o["ba"], nl, axis=2) > 0) | (np.repeat(tr["ba"][s][:nb], nl, axis=2) > 0)
        mis["mant"] += int(np.sum((tr["mant"][s][:nb] != o["mant"]) & coded)); cnt["mant"] += int(np.sum(coded))
        agree = (tr["lrms"][s][:nb] == o["lrms"]) & np.all(tr["oscale"][s][:nb] == o["oscale"], axis=1)
        ref, g = o["lines"][agree], tr["lines"][s][:nb][agree]
        if ref.size:
            ms = ((o["lrms"][agree][:, None] >> band[None, :]) & 1).astype(bool)
            big = np.maximum(np.abs(ref[:, 0] + ref[:, 1]), np.abs(ref[:, 0] - ref[:, 1]))
            # tolerance of tests/test_gpu_parity.py:line_tolerance: 1e-5 relative + the reference's own FFT noise floor (1e-11 of the
            # block maximum) + the fp32 resolution of L and R at that line in M/S bands
            tol = 1e-5 * np.abs(ref) + 1e-11 * np.max(np.abs(ref), axis=(1, 2), keepdims=True) + 1.2e-7 * np.where(ms, big, 0.0)[:, None, :]
            worst_line = max(worst_line, float(np.max(np.abs(g - ref) / np.maximum(tol, 1e-300))))
            rs, gs = o["smr"][agree], tr["smr"][s][:nb][agree]
            es = np.abs(gs - rs) / (1e-5 * np.maximum(np.abs(rs), 10.0))
            smr_over += int(np.sum(es > 1.0)); smr_n += es.size
            worst_smr = max(worst_smr, float(np.quantile(es, 0.999)))
    out["fp32"] = {"sample": "%d streams x %.0f s" % (sub.shape[0], nt / FS),
                   "mismatch_rate": {k: (mis[k] / cnt[k] if cnt[k] else None) for k in cnt},
                   "compared": cnt,
                   "worst_line_error_in_units_of_tolerance": worst_line,
                   "smr_error_p999_in_units_of_1e-5_rel": worst_smr,
                   "smr_values_over_1e-5_rel": smr_over, "smr_values": smr_n,
                   "note": "mant = signed mantissa codes of every line coded by either side; SMR values over tolerance are peak-picking flips "
                           "(findpeaks' strict comparisons between nearly equal bins, psychoac.py:166-168), a discrete mismatch like a flipped M/S decision"}
    return out


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference itself is Python 2
    and cannot run on the box) on all host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # the corpus generator is integer arithmetic on a counter-based Philox (corpus.py): numpy here, torch on the GPU in the repo arm,
    # bit-identical samples either way -- this arm needs neither torch nor a CUDA device
    corpus_note = ""
    sec = args.ref_seconds
    n = int(sec * FS)
    S = cores
    pcm = gen_streams_numpy(list(range(S)), n)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_baseline_run(pcm, sec, cores)
        log("reference step %d: %.2f audio-s/s (%.1fs)" % (i, v, dt))
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    sample = "%d streams x %.0f s of the synthetic corpus (stream ids 0..%d), one stream per host thread%s" % (S, sec, S - 1, corpus_note)
    line = {"impl": "reference", "metric": "audio-seconds encoded per second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic corpus 4096 x 60 s 44.1 kHz stereo (bounded sample: %s)" % sample, "precision": "fp64 (reference arithmetic)",
                       "targetBitsPerSample": 2.27},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------- our arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--ref-seconds", type=float, default=16.0, help="seconds per stream of the bounded CPU sample (one stream per host thread)")
    ap.add_argument("--tbps", type=float, default=2.27, help="targetBitsPerSample (pacfile.py:455 default 2.27)")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE.json configs[4]: also run the bitrate sweep 64..256 kb/s/ch (encode and decode-only, extra 'bitrate_sweep' key)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity leg that rides on the cpu_baseline sample")
    ap.add_argument("--no-stages", action="store_true", help="skip the MDCT-only stage timing (roofline.stages)")
    ap.add_argument("--decode", action="store_true", help="(default on; kept for old command lines) time decode-only throughput of the coded corpus")
    ap.add_argument("--no-decode", action="store_true", help="skip the decode-only leg (extra 'decode' key with its own roofline)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    # stdout carries exactly ONE JSON line: anything libraries print to fd 1 meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import _pacb200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, n = args.streams, int(args.seconds * FS)
    mine = list(range(rank, S, world))                   # stream s -> rank s mod world
    eng = _pacb200.Engine(local, args.precision, targetBitsPerSample=args.tbps)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)      # torch's CUDA events now bracket the library's work
    nblk = eng.num_blocks(n)
    t0 = time.time()
    pcm = torch.empty(len(mine), n, 2, dtype=torch.int16, device=dev)
    for c0 in range(0, len(mine), 64):
        pcm[c0:c0 + 64] = gen_streams(mine[c0:c0 + 64], n, dev)
    torch.cuda.synchronize()
    log("[rank %d] generated %d streams x %.0f s in %.1fs" % (rank, len(mine), args.seconds, time.time() - t0))
    cap = eng.encode_bound(n)
    out = torch.empty(len(mine), cap, dtype=torch.uint8, device=dev)

    def step():
        _, ob = eng.encode_batch(pcm, out=out, cap=cap)
        counts = torch.zeros(S, dtype=torch.int64, device=dev)
        counts[mine] = torch.as_tensor(ob, device=dev)
        if world > 1:
            dist.all_reduce(counts)                       # the only collective: gather of per-stream byte counts
        return counts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        counts = step()
    barrier()
    clocks = Clocks(local)
    clocks.start()
    eng.timing(True)
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        counts = step()
    ev1.record()
    barrier()
    wall = time.time() - w0
    tm = eng.timing_get()
    eng.timing(False)
    launches = eng.launches - l0
    # the library was put on torch's current stream, so the CUDA events bracket exactly the K steps on the device
    el = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed = float(el.item())
    audio_s = S * args.seconds
    value = audio_s * args.steps / elapsed
    total_bytes = int(counts.sum().item())

    # ---- stage split of the analysis kernel (SURVEY.md 8d "Which roofline"): the window+MDCT stage by itself (the MDCT-only
    # instantiation of k_analysis, pac_mdct_batch) and the whole analysis kernel by itself (pac_analysis_batch: not overlapped with
    # scan/pack as inside the pipeline); SMR stage = the difference.  Device-resident corpus, CUDA-event time of the launches.
    stages = None
    if not args.no_stages:
        eng.mdct_batch(pcm, want=False)
        ms_m = eng.mdct_batch(pcm, want=False)[2]
        eng.analysis_batch_ms(pcm)
        ms_a = eng.analysis_batch_ms(pcm)
        nb_loc = len(mine) * nblk
        peak_s, _ = measured_peaks()
        lb = 4 if args.precision == "fp32" else 8
        by_m = 4096 + 2 * 1024 * lb + 2                     # PCM in + scaled L/R lines + two overall scales out
        by_s = 4096 + 2 * 25 * lb + 8                       # a stand-alone SMR stage: PCM in + band SMRs + LRMS/scales out
        algo_a = ALGO_BYTES_FP32 if args.precision == "fp32" else ALGO_BYTES_FP64
        gbs = lambda by, ms: by * nb_loc / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stages = {"blocks": nb_loc, "rank": rank,
                  "mdct": {"kernel": "k_mdct_enc (fp32 mode; fp64: k_analysis<MDCT_ONLY>): PCM->fraction, SineWindow, MDCT, overall scale", "ms": ms_m,
                           "ns_per_block": ms_m * 1e6 / nb_loc, "algorithmic_bytes_per_block": by_m, "achieved": gbs(by_m, ms_m),
                           "frac": gbs(by_m, ms_m) / peak_s},
                  "analysis_alone": {"kernel": "k_analysis by itself (persistent grid, nothing co-running)", "ms": ms_a,
                                     "ns_per_block": ms_a * 1e6 / nb_loc, "algorithmic_bytes_per_block": algo_a,
                                     "achieved": gbs(algo_a, ms_a), "frac": gbs(algo_a, ms_a) / peak_s},
                  "smr": {"kernel": "analysis_alone - mdct (M/S decision + six threshold curves + band SMR)", "ms": ms_a - ms_m,
                          "ns_per_block": (ms_a - ms_m) * 1e6 / nb_loc, "algorithmic_bytes_per_block": by_s,
                          "achieved": gbs(by_s, ms_a - ms_m), "frac": gbs(by_s, ms_a - ms_m) / peak_s,
                          "share_of_analysis_time": (ms_a - ms_m) / ms_a if ms_a > 0 else None}}
        log("[rank %d] stages: %s" % (rank, json.dumps(stages)))

    # ---- decode-only throughput (BASELINE.json configs[4]): the coded images are decoded where pac_encode_batch left them
    dec = None
    if not args.no_decode:
        Sd = min(len(mine), 1024)                         # bounded: the decoded PCM needs its own buffer next to the inputs
        ob_h = counts[mine].cpu().numpy()[:Sd]
        beg = np.arange(Sd, dtype=np.int64) * cap
        pcm_out = torch.empty(Sd, nblk * 1024 + 1024, 2, dtype=torch.int16, device=dev)
        dstride = pcm_out.shape[1]
        for _ in range(2):
            ns, _, _ = eng.decode_batch_strided(out, beg, ob_h, pcm_out, dstride)
        barrier()
        eng.timing(True)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(args.steps):
            ns, _, _ = eng.decode_batch_strided(out, beg, ob_h, pcm_out, dstride)
        d1.record()
        barrier()
        tmd = eng.timing_get()
        eng.timing(False)
        dl = torch.tensor([d0.elapsed_time(d1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dl, op=dist.ReduceOp.MAX)
        sd_all = torch.tensor([Sd], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(sd_all)
        # decode roofline: algorithmic bytes = coded images in + int16 PCM out, over the whole decode call (index + unpack + synth)
        dbytes = float(np.sum(ob_h)) + float(Sd) * float(ns[0]) * 4.0
        dgbs = dbytes * args.steps / float(dl.item()) / 1e9
        dpeak, dsrc = measured_peaks()
        dec = {"value": int(sd_all.item()) * args.seconds * args.steps / float(dl.item()), "unit": "audio-s/s",
               "roofline": {"bound": "hbm", "rank": rank, "algorithmic_bytes_per_step": dbytes, "achieved": dgbs, "peak": dpeak, "unit": "GB/s",
                            "frac": dgbs / dpeak, "peak_source": dsrc,
                            "note": "whole decode call of this rank (k_index + k_unpack + k_synth); bytes = coded images read + PCM written"},
               "ms_per_step": float(dl.item()) / args.steps * 1e3, "streams": int(sd_all.item()), "samples_per_stream": int(ns[0]),
               "input": "coded images left on the device by pac_encode_batch ([S][cap] strided), PCM written to device memory",
               "kernels_ms": {k: v[0] for k, v in tmd.items() if v[1]}}
        del pcm_out

    # ---- bitrate sweep (BASELINE.json configs[4]): encode and decode-only per target rate
    sweep = None
    if args.sweep:
        sweep = []
        Sd = min(len(mine), 1024)
        pcm_out = torch.empty(Sd, nblk * 1024 + 1024, 2, dtype=torch.int16, device=dev)
        beg_of = lambda c: np.arange(Sd, dtype=np.int64) * c

        def timed(fn):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a0.record()
            r = fn()
            a1.record()
            barrier()
            t = torch.tensor([a0.elapsed_time(a1) * 1e-3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return r, float(t.item())

        for kbps in (64, 96, 128, 192, 256):
            tb = kbps * 1000.0 / FS
            e2 = _pacb200.Engine(local, args.precision, targetBitsPerSample=tb)
            e2.set_stream(torch.cuda.current_stream().cuda_stream)
            cap2 = e2.encode_bound(n)
            out2 = torch.empty(len(mine), cap2, dtype=torch.uint8, device=dev)
            e2.encode_batch(pcm, out=out2, cap=cap2)                              # warm
            (_, ob2), t_enc = timed(lambda: e2.encode_batch(pcm, out=out2, cap=cap2))
            ob_h = np.asarray(ob2, dtype=np.int64)[:Sd]
            e2.decode_batch_strided(out2, beg_of(cap2), ob_h, pcm_out, pcm_out.shape[1])   # warm
            _, t_dec = timed(lambda: e2.decode_batch_strided(out2, beg_of(cap2), ob_h, pcm_out, pcm_out.shape[1]))
            tot = torch.tensor([float(np.sum(ob2)), float(Sd)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tot)
            sweep.append({"kbps_per_channel": kbps, "targetBitsPerSample": round(tb, 4),
                          "encode_audio_s_per_s": audio_s / t_enc, "decode_audio_s_per_s": float(tot[1].item()) * args.seconds / t_dec,
                          "decode_streams": int(tot[1].item()), "coded_bytes": int(tot[0].item()),
                          "coded_kbps_per_channel": float(tot[0].item()) * 8 / 1000.0 / (audio_s * 2)})
            e2.close()
            del out2
            torch.cuda.empty_cache()                      # cudaMalloc'd library workspaces of the next engine need the room
        del pcm_out
        torch.cuda.empty_cache()

    # ---- e2e through the C ABI with pinned host buffers
    e2e = None
    if not args.no_e2e:
        old_aff, numa_note = bind_near_gpu(local)         # pinned buffers below land on the GPU's NUMA node
        log("[rank %d] e2e: %s" % (rank, numa_note))
        ph = _pacb200.pinned_empty((len(mine), n, 2), np.int16)     # pac_pinned_alloc: page-locked and device-mapped
        oh = _pacb200.pinned_empty((len(mine), cap), np.uint8)
        torch.from_numpy(ph).copy_(pcm)
        del out
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                          # the staging buffers are the library's own cudaMalloc
        eng.encode_batch(ph, out=oh, cap=cap)             # warm (allocates the staging buffers)
        barrier()
        ksteps = max(1, min(args.steps, 2))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(ksteps):
            _, ob = eng.encode_batch(ph, out=oh, cap=cap)
        e1.record()
        barrier()
        e2 = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2, op=dist.ReduceOp.MAX)
        e2e = {"value": audio_s * ksteps / float(e2.item()), "unit": "audio-s/s", "h2d_bytes_per_step": int(S * n * 4),
               "d2h_bytes_per_step": int(total_bytes), "steps": ksteps,
               "path": "pinned host PCM staged slab-wise by DMA; the .pac images are written by k_pack straight into the pinned host buffer "
                       "(only outBytes[s] bytes of each image cross PCIe)", "host": numa_note}
        del ph, oh
        if old_aff:
            try:
                os.sched_setaffinity(0, old_aff)          # the CPU baseline leg below uses every host core again
            except Exception:
                pass
    clk = clocks.summary()

    # ---- roofline of the dominant stage: window+MDCT (k_mdct_enc, fp32 mode) + M/S decision + SMR (k_analysis) run back to back on one
    # stream for every tile; the 12 496 algorithmic bytes of SURVEY.md 8(d) belong to the pair, so the pair's time is the denominator
    peak, peak_src = measured_peaks()
    a_ms, a_cnt = tm["analysis"]
    m_ms, m_cnt = tm.get("mdct", (0.0, 0))
    blocks_local = len(mine) * nblk * args.steps
    algo = ALGO_BYTES_FP32 if args.precision == "fp32" else ALGO_BYTES_FP64
    st_ms = a_ms + m_ms
    ach = blocks_local * algo / (st_ms * 1e-3) / 1e9 if st_ms > 0 else 0.0
    traffic = ncu_traffic()
    winst = traffic.get("warp_inst_per_block") if traffic else None
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_issue = 4 * sms * (clk.get("sm_max_mhz") or 1965.0) * 1e-3
    roof = {"bound": "hbm", "kernel": "k_mdct_enc + k_analysis (window+MDCT, then M/S decision + SMR), timed inside the pipeline"
                                       if m_cnt else "k_analysis (window+MDCT+M/S decision+SMR), timed inside the pipeline",
            "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "peak_source": peak_src,
            "traffic": (traffic["dram_bytes_per_block"] * blocks_local / max(a_cnt, 1)) if traffic and traffic.get("dram_bytes_per_block") else None,
            "traffic_source": traffic.get("source") if traffic else None,
            "stages": stages,
            "algorithmic_bytes_per_block": algo, "blocks_per_launch": blocks_local / max(a_cnt, 1),
            "avg_launch_ms": st_ms / max(a_cnt, 1), "launches": a_cnt + m_cnt,
            "ns_per_block": {"analysis": a_ms * 1e6 / blocks_local if blocks_local else None,
                             "mdct": m_ms * 1e6 / blocks_local if blocks_local else None},
            # the stage is issue-bound, so the same launches are also placed on the SM issue roofline: warp instructions per block
            # (ncu smsp__inst_executed, profiles/) over the launch time against 4 schedulers x SMs x SM clock
            "issue": ({"warp_inst_per_block": winst,
                       "achieved_ginst_s": winst * blocks_local / (st_ms * 1e-3) / 1e9 if st_ms > 0 else 0.0,
                       "peak_ginst_s": peak_issue,
                       "frac": winst * blocks_local / (st_ms * 1e-3) / 1e9 / peak_issue if st_ms > 0 else 0.0}
                      if winst and args.precision == "fp32" else None),
            "note": "SMR is issue/transcendental-bound, not HBM-bound (SURVEY.md App. E); kernel time split (inside the overlapped pipeline): "
                    + ", ".join("%s %.1f ms" % (k, v[0]) for k, v in tm.items() if v[1])}

    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:       # the CPU port is timed beside the N=1 run only
        cores = os.cpu_count() or 1
        sec = args.ref_seconds
        sample_ids = list(range(cores))
        pcm_s = gen_streams(sample_ids, int(sec * FS), dev).cpu().numpy()
        v, dt = cpu_baseline_run(pcm_s, sec, cores)
        cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": "%d streams x %.0f s of the synthetic corpus, one per host thread, %.1f s wall" % (cores, sec, dt)}
        if not args.no_parity:
            t0 = time.time()
            parity = parity_leg(pcm_s, sec, cores, local)
            log("parity leg: %.1f s: %s" % (time.time() - t0, json.dumps(parity)))

    if rank == 0:
        line = {"metric": "audio-seconds encoded per second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
                "config": {"workload": "synthetic corpus %d x %.0f s 44.1 kHz stereo int16 (tones+noise+transients; corpus.py: counter-based Philox, integer arithmetic, bit-identical on CPU and GPU), %s mode, streams sharded s mod N"
                                       % (S, args.seconds, args.precision),
                           "streams": S, "seconds_per_stream": args.seconds, "blocks_per_stream": nblk, "targetBitsPerSample": args.tbps,
                           "cache": "inputs (%.1f GB per rank) far larger than the 126 MB L2; no flush needed" % (len(mine) * n * 4 / 1e9),
                           "coded_bytes": total_bytes},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "parity": parity}
        if dec:
            line["decode"] = dec
        if sweep:
            line["bitrate_sweep"] = sweep
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

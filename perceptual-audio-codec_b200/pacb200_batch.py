"""
pacb200_batch.py -- whole-file / whole-corpus front end of the engine: what `python pacfile.py x.wav` does for one
file, done for many files in one pac_encode_batch / pac_decode_batch call (the per-block Python API cannot express
the batch).  Stream-level sharding across GPUs lives here too.
"""
import struct

import numpy as np

import _pacb200


def read_wav(path):
    """(sampleRate, pcm int16 [n][2]) -- header walk as PCMFile.ReadFileHeader (pcmfile.py:32-57)."""
    raw = open(path, "rb").read()
    if raw[0:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise Exception("ERROR: File opened for PCMFile is not a RIFF file!")
    p = 12
    while raw[p:p + 4] != b"fmt ":
        p += 4
        if p + 4 > len(raw):
            raise Exception("ERROR: Didn't find WAV file 'fmt ' chunk following RIFF file header")
    p += 4
    fsize, tag, nch, rate, bps, align, bits = struct.unpack("<LHHLLHH", raw[p:p + 20])
    p += 20
    if tag != 1:
        raise Exception("Opened a non-PCM WAV file as a PCMFile")
    if bits != 16:
        raise Exception("PCMFile was not 16-bits per sample")
    if nch != 2:
        raise Exception("only 2-channel files are supported (as in the reference, codec.py:46-47)")
    while raw[p:p + 4] != b"data":
        p += 4
        if p + 4 > len(raw):
            raise Exception("Didn't find WAV file 'data' chunk following 'fmt ' chunk")
    n = struct.unpack("<L", raw[p + 4:p + 8])[0] // 4
    body = raw[p + 8:p + 8 + n * 4]
    got = len(body) // 4
    pcm = np.zeros((n, 2), dtype=np.int16)
    pcm[:got] = np.frombuffer(body[:got * 4], dtype="<i2").reshape(-1, 2)
    return rate, pcm


def wav_bytes(pcm, sampleRate, numSamplesHdr):
    """PCMFile.WriteFileHeader + payload (pcmfile.py:103-147): the header carries cp.numSamples from the PAC header."""
    dataBytes = int(numSamplesHdr) * 4
    hdr = struct.pack("<4sL4s4sLHHLLHH4sL", b"RIFF", 36 + dataBytes, b"WAVE", b"fmt ", 16, 1, 2, sampleRate, sampleRate * 4, 4, 16,
                      b"data", dataBytes)
    return hdr + np.ascontiguousarray(pcm, dtype="<i2").tobytes()


def encode_files(paths, precision="fp64", targetBitsPerSample=2.27, device=0):
    """[wav path] -> [pac bytes] with one batch call per sample rate."""
    loaded = [read_wav(p) for p in paths]
    out = [None] * len(paths)
    for rate in sorted(set(r for r, _ in loaded)):
        idx = [i for i, (r, _) in enumerate(loaded) if r == rate]
        L = max(len(loaded[i][1]) for i in idx)
        batch = np.zeros((len(idx), max(L, 1), 2), np.int16)
        ns = np.zeros(len(idx), np.int64)
        for j, i in enumerate(idx):
            pcm = loaded[i][1]
            batch[j, :len(pcm)] = pcm
            ns[j] = len(pcm)
        e = _pacb200.engine(sampleRate=rate, targetBitsPerSample=targetBitsPerSample, device=device, precision=precision)
        for j, b in zip(idx, e.encode_batch(batch, nSamples=ns)):
            out[j] = b
    return out


def decode_files(pacs, precision="fp64", device=0):
    """[pac bytes] -> [wav bytes] as the reference's Decode pass writes them."""
    out = [None] * len(pacs)
    rates = [struct.unpack("<L", p[4:8])[0] for p in pacs]
    for rate in sorted(set(rates)):
        idx = [i for i, r in enumerate(rates) if r == rate]
        e = _pacb200.engine(sampleRate=rate, device=device, precision=precision)
        for i, (pcm, sr, ns) in zip(idx, e.decode_batch([pacs[i] for i in idx])):
            out[i] = wav_bytes(pcm, sr, ns)
    return out


def shard_streams(S, rank, world):
    """stream s -> rank s mod world: the indices this rank owns."""
    return list(range(rank, S, world))


def gather_byte_counts(local_counts, S, rank, world, dist=None, device=None):
    """The path's only collective: all ranks learn every stream's byte count (int64 [S])."""
    import torch
    full = torch.zeros(S, dtype=torch.int64, device=device)
    mine = shard_streams(S, rank, world)
    full[mine] = torch.as_tensor(np.asarray(local_counts, dtype=np.int64), device=device)
    if dist is not None and world > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM)      # disjoint supports: SUM == gather
    return full

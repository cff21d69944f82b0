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
    buf = bytearray(n * 4)                                 # a short file is padded with zero BYTES (pcmfile.py:77-79)
    buf[:len(body)] = body
    return rate, np.frombuffer(bytes(buf), dtype="<i2").reshape(-1, 2).copy()


def wav_bytes(pcm, sampleRate, numSamplesHdr):
    """PCMFile.WriteFileHeader + payload (pcmfile.py:103-147): the header carries cp.numSamples from the PAC header."""
    dataBytes = int(numSamplesHdr) * 4
    hdr = struct.pack("<4sL4s4sLHHLLHH4sL", b"RIFF", 36 + dataBytes, b"WAVE", b"fmt ", 16, 1, 2, sampleRate, sampleRate * 4, 4, 16,
                      b"data", dataBytes)
    return hdr + np.ascontiguousarray(pcm, dtype="<i2").tobytes()


def wav_info(path):
    """One RIFF walk over the file's HEAD only (the 4-byte-step chunk scan of PCMFile.ReadFileHeader, pcmfile.py:32-57):
    (sampleRate, numSamples from the 'data' chunk size, byte offset of the samples, bytes of samples actually present)."""
    import os
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        raw = f.read(1 << 16)
        if raw[0:4] != b"RIFF" or raw[8:12] != b"WAVE":
            raise Exception("ERROR: File opened for PCMFile is not a RIFF file!")
        if len(raw) == (1 << 16) and (raw.find(b"data", 12) < 0 or raw.find(b"fmt ", 12) < 0):
            raw += f.read()                               # unusually long header chunks: fall back to the whole file
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
    off = p + 8
    return rate, n, off, max(0, min(n * 4, size - off))


def load_wavs_pinned(paths):
    """Batched WAV ingest (pcmfile.py:32-100 for many files): every file's samples are read STRAIGHT into its row of one pinned
    [S][stride][2] int16 slab (readinto: no intermediate bytes object, no pageable copy), one slab per sample rate.
    Returns {rate: (indices into paths, slab, nSamples int64[S])}."""
    info = [wav_info(p) for p in paths]
    groups = {}
    for rate in sorted(set(i[0] for i in info)):
        idx = [k for k, i in enumerate(info) if i[0] == rate]
        stride = max(1, max(info[k][1] for k in idx))
        slab = _pacb200.pinned_empty((len(idx), stride, 2), np.int16)
        ns = np.zeros(len(idx), np.int64)
        for j, k in enumerate(idx):
            _, n, off, avail = info[k]
            row = slab[j].reshape(-1).view(np.uint8)
            with open(paths[k], "rb", buffering=0) as f:
                f.seek(off)
                got = 0
                while got < avail:
                    r = f.readinto(memoryview(row)[got:avail])
                    if not r:
                        break
                    got += r
            row[got:n * 4] = 0                            # a truncated file is padded with zero bytes (pcmfile.py:77-79)
            ns[j] = n
        groups[rate] = (idx, slab, ns)
    return groups


def encode_files(paths, precision="fp64", targetBitsPerSample=2.27, device=0, out_paths=None):
    """[wav path] -> [pac bytes] (or, with out_paths, .pac files written straight from the pinned output buffer) with one
    pac_encode_batch call per sample rate: pinned PCM slab in, pinned images out (written in place by the pack kernel)."""
    out = [None] * len(paths)
    for rate, (idx, slab, ns) in load_wavs_pinned(paths).items():
        e = _pacb200.engine(sampleRate=rate, targetBitsPerSample=targetBitsPerSample, device=device, precision=precision)
        cap = e.encode_bound(int(ns.max()) if len(ns) else 0)
        img = _pacb200.pinned_empty((len(idx), cap), np.uint8)
        _, ob = e.encode_batch(slab, nSamples=ns, out=img, cap=cap)
        for j, k in enumerate(idx):
            if out_paths is not None:
                with open(out_paths[k], "wb") as f:
                    f.write(memoryview(img[j])[:int(ob[j])])
                out[k] = int(ob[j])
            else:
                out[k] = img[j, :ob[j]].tobytes()
    return out


def decode_files(pacs, precision="fp64", device=0, out_paths=None):
    """[pac bytes | pac path] -> [wav bytes] as the reference's Decode pass writes them (or, with out_paths, WAV files written as
    header + the pinned PCM rows, no per-file copy).  One pac_decode_batch call per sample rate: the images are gathered into one
    pinned blob (files: readinto), the PCM comes back into one pinned [S][stride][2] buffer."""
    import os
    is_path = [isinstance(p, str) for p in pacs]
    sizes = [os.path.getsize(p) if ip else len(p) for p, ip in zip(pacs, is_path)]
    heads = []
    for p, ip in zip(pacs, is_path):
        if ip:
            with open(p, "rb") as f:
                heads.append(f.read(8))
        else:
            heads.append(bytes(p[:8]))
    rates = [struct.unpack("<L", h[4:8])[0] if len(h) >= 8 else 0 for h in heads]
    out = [None] * len(pacs)
    for rate in sorted(set(rates)):
        idx = [i for i, r in enumerate(rates) if r == rate]
        e = _pacb200.engine(sampleRate=rate if rate else 44100, device=device, precision=precision)
        beg = np.zeros(len(idx), np.int64)
        tot = 0
        for j, i in enumerate(idx):
            beg[j] = tot
            tot += (sizes[i] + 15) & ~15
        blob = _pacb200.pinned_empty((tot + 16,), np.uint8)
        for j, i in enumerate(idx):
            dst = memoryview(blob)[int(beg[j]):int(beg[j]) + sizes[i]]
            if is_path[i]:
                with open(pacs[i], "rb", buffering=0) as f:
                    got = 0
                    while got < sizes[i]:
                        r = f.readinto(dst[got:])
                        if not r:
                            break
                        got += r
            else:
                dst[:] = pacs[i]
        length = np.array([sizes[i] for i in idx], np.int64)
        stride = max(e.decode_bound(int(n)) for n in length)
        pcm = _pacb200.pinned_empty((len(idx), stride, 2), np.int16)
        ns, hn, hr = e.decode_batch_strided(blob, beg, length, pcm, stride)
        for j, i in enumerate(idx):
            dataBytes = int(hn[j]) * 4
            hdr = struct.pack("<4sL4s4sLHHLLHH4sL", b"RIFF", 36 + dataBytes, b"WAVE", b"fmt ", 16, 1, 2, int(hr[j]), int(hr[j]) * 4, 4, 16,
                              b"data", dataBytes)                       # PCMFile.WriteFileHeader, pcmfile.py:103-116
            body = memoryview(pcm[j].reshape(-1).view(np.uint8))[:int(ns[j]) * 4]
            if out_paths is not None:
                with open(out_paths[i], "wb") as f:
                    f.write(hdr)
                    f.write(body)
                out[i] = len(hdr) + len(body)
            else:
                out[i] = hdr + bytes(body)
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

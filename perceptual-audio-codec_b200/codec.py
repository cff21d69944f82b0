"""
codec.py -- codec.Encode / codec.Decode with the reference's signatures and return structures
(codec/codec.py:25-129), executed by the CUDA engine through the C ABI (pac_encode_blocks / pac_decode_blocks).
"""
import numpy as np

import _pacb200


def _states(codingParams, huffman):
    return [[int(codingParams.extraBits), int(huffman.bitDeposit)]]


def Encode(data, codingParams, huffman):
    """codec.py:83-129.  data = [L, R], each the 2*nMDCTLines prior|current block of signed fractions.
    Returns (scaleFactor, bitAlloc, signBits, huffmanCodedMantissa, tableID, overallScaleFactor, LRMS) and updates
    codingParams.extraBits / huffman.bitDeposit like the reference (codec.py:118-120,229,258-260)."""
    e = _pacb200.engine_for(codingParams)
    blk = np.stack([np.asarray(data[0], dtype=np.float64), np.asarray(data[1], dtype=np.float64)])[None]
    st = _states(codingParams, huffman)
    r = e.encode_blocks(blk, st)
    codingParams.extraBits, huffman.bitDeposit = st[0][0], st[0][1]
    sfBands = codingParams.sfBands
    nB = sfBands.nBands
    scaleFactor = [r["sf"][0, ch].astype(np.int32) for ch in range(2)]
    bitAlloc = [r["ba"][0, ch].astype(int) for ch in range(2)]
    tableID = [int(r["tableID"][0, ch]) for ch in range(2)]
    overallScaleFactor = [int(r["oscale"][0, ch]) for ch in range(2)]
    LRMS = np.array([(int(r["lrms"][0]) >> b) & 1 for b in range(nB)], dtype='int')
    signBits, coded = [], []
    for ch in range(2):
        enc = e.tables[tableID[ch]]
        sb, hc = [], []
        for b in range(nB):
            ba = int(bitAlloc[ch][b])
            if not ba:
                continue
            lo = int(sfBands.lowerLine[b])
            for m in r["mant"][0, ch, lo:lo + int(sfBands.nLines[b])]:
                m = int(m)
                sb.append(m >> (ba - 1))                             # StripSignBits, codec.py:67-81
                mag = m & ((1 << (ba - 1)) - 1)
                code = enc.get(mag)
                hc.append(code if code is not None else enc[-1] + format(mag, '0' + str(ba) + 'b'))   # Huffman.py:292-298
        signBits.append(sb)
        coded.append(hc)
    codingParams._pac_chunks = r["chunks"][0]            # packed payloads of this block (used by PACFile.WriteDataBlock)
    return (scaleFactor, bitAlloc, signBits, coded, tableID, overallScaleFactor, LRMS)


def Decode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams, LRMS):
    """codec.py:25-65: returns (dataL, dataR), the windowed IMDCT output (2*nMDCTLines samples, pre overlap-add),
    including the reference's M/S recombination aliasing (codec.py:46-56)."""
    e = _pacb200.engine_for(codingParams)
    mask = sum((1 << b) for b, v in enumerate(LRMS) if v)
    out = e.decode_blocks(np.asarray(scaleFactor)[None], np.asarray(bitAlloc)[None], np.asarray(mantissa)[None],
                          np.asarray(overallScaleFactor)[None], np.array([mask]))
    return out[0, 0], out[0, 1]


def StripSignBits(codingParams, mantissa, bitAlloc):
    """codec.py:67-81 (integer field split)."""
    signBits, unsignedMantissas = [], []
    iMant = 0
    for iBand in range(codingParams.sfBands.nBands):
        ba = int(bitAlloc[iBand])
        if ba:
            for j in range(int(codingParams.sfBands.nLines[iBand])):
                m = int(mantissa[iMant + j])
                signBits.append(m >> (ba - 1))
                unsignedMantissas.append(m & ((1 << (ba - 1)) - 1))
            iMant += int(codingParams.sfBands.nLines[iBand])
    return (signBits, unsignedMantissas)


def _huffman_encode(huffman, codingParams, mantissaCode, bitAlloc):
    """Huffman.encodeData (Huffman.py:274-309): device table search (pac_huffman_select), then the winning table's
    codes formatted as the reference's '0101' strings."""
    e = _pacb200.engine_for(codingParams)
    sfBands = codingParams.sfBands
    ba_sym = []
    for b in range(sfBands.nBands):
        if bitAlloc[b]:
            ba_sym += [int(bitAlloc[b])] * int(sfBands.nLines[b])
    mags = [int(m) for m in mantissaCode]
    if len(mags) != len(ba_sym):
        raise ValueError("encodeData: %d mantissas for %d allocated lines" % (len(mags), len(ba_sym)))
    tid, _ = e.huffman_select(np.array(mags, dtype=np.uint32), np.array(ba_sym, dtype=np.int32))
    enc = e.tables[tid]
    out = []
    for m, ba in zip(mags, ba_sym):
        code = enc.get(m)
        out.append(code if code is not None else enc[-1] + format(m, '0' + str(ba) + 'b'))
    return (out, tid)

"""
pcmfile.py -- PCMFile: 16-bit PCM WAV handlers with the reference's interface (codec/pcmfile.py:29-147).
The int16 <-> signed-fraction conversions are the reference's vDequantizeUniform/vQuantizeUniform(., 16)
(pcmfile.py:91-98,127-134) and run through the C ABI (quantize.py shim); header parsing is host-side file plumbing
that mirrors the reference's 4-byte-step chunk scan (pcmfile.py:32-57).
"""
from struct import pack, unpack

import numpy as np

from audiofile import AudioFile, CodingParams
from quantize import vDequantizeUniform, vQuantizeUniform

BYTESIZE = 8


class PCMFile(AudioFile):
    def ReadFileHeader(self):
        tag = self.fp.read(12)
        if tag[0:4] != b"RIFF" or tag[8:12] != b"WAVE":
            raise Exception("ERROR: File opened for PCMFile is not a RIFF file!")
        while True:
            tag = self.fp.read(4)
            if len(tag) < 4:
                raise Exception("ERROR: Didn't find WAV file 'fmt ' chunk following RIFF file header")
            if tag[0:4] == b"fmt ":
                break
        tag = self.fp.read(20)
        (formatSize, formatTag, nChannels, sampleRate, bytesPerSec, blockAlign, bitsPerSample) = unpack("<LHHLLHH", tag)
        if formatTag != 1:
            raise Exception("Opened a non-PCM WAV file as a PCMFile")
        if bitsPerSample != 16:
            raise Exception("PCMFile was not 16-bits per sample")
        while True:
            tag = self.fp.read(4)
            if len(tag) < 4:
                raise Exception("Didn't find WAV file 'data' chunk following 'fmt ' chunk")
            if tag[0:4] == b"data":
                break
        numSamples = unpack('<L', self.fp.read(4))[0]
        numSamples //= nChannels * (bitsPerSample // BYTESIZE)
        myParams = CodingParams()
        myParams.nChannels = nChannels
        myParams.bitsPerSample = bitsPerSample
        myParams.sampleRate = sampleRate
        myParams.numSamples = numSamples
        myParams.bytesReadSoFar = 0
        return myParams

    def ReadDataBlock(self, codingParams):
        cp = codingParams
        bytesPer = cp.bitsPerSample // BYTESIZE
        bytesToRead = cp.nSamplesPerBlock * cp.nChannels * bytesPer
        left = cp.nChannels * cp.numSamples * bytesPer - cp.bytesReadSoFar
        if left <= 0:
            dataBlock = None
        elif left < bytesToRead:
            dataBlock = self.fp.read(left)
        else:
            dataBlock = self.fp.read(bytesToRead)
        cp.bytesReadSoFar += bytesToRead
        if dataBlock and len(dataBlock) < bytesToRead:
            dataBlock += (bytesToRead - len(dataBlock)) * b"\0"      # partial block: zero pad (pcmfile.py:77-79)
        elif not dataBlock:
            return
        if cp.bitsPerSample != 16:
            raise Exception("PCMFile was not 16-bit PCM in PCMFile.ReadDataBlock!")
        codesAll = np.frombuffer(dataBlock, dtype="<i2").astype(np.int64)
        data = []
        for iCh in range(cp.nChannels):
            codes = codesAll[iCh::cp.nChannels].copy()
            signs = np.signbit(codes)
            codes[signs] *= -1
            temp = vDequantizeUniform(codes, 16)                     # pcmfile.py:96
            temp[signs] *= -1.
            data.append(temp)
        return data

    def WriteFileHeader(self, codingParams):
        cp = codingParams
        bytesPer = cp.bitsPerSample // BYTESIZE
        dataBytes = cp.numSamples * cp.nChannels * bytesPer
        self.fp.write(pack('<4sL4s4sLHHLLHH4sL', b"RIFF", 36 + dataBytes, b"WAVE", b"fmt ", 16, 1, cp.nChannels, cp.sampleRate,
                           cp.sampleRate * cp.nChannels * bytesPer, cp.nChannels * bytesPer, cp.bitsPerSample, b"data", dataBytes))

    def WriteDataBlock(self, data, codingParams):
        nChannels = len(data)
        if nChannels != codingParams.nChannels:
            raise Exception("Data block to PCMFile did not have expected number of channels")
        nSamples = min([len(data[iCh]) for iCh in range(nChannels)])
        codes = []
        for iCh in range(nChannels):
            temp = data[iCh]
            signs = np.signbit(temp)
            temp[signs] *= -1.                                       # the reference flips its argument in place too
            q = vQuantizeUniform(temp, 16).astype(np.int16)          # pcmfile.py:131-132
            q[signs] *= -1
            codes.append(q)
        if codingParams.bitsPerSample != 16:
            raise Exception("Asked to write to a PCM file with other than 16-bits per sample in PCMFile.WriteDataBlock!")
        block = np.empty(nSamples * nChannels, dtype="<i2")
        for iCh in range(nChannels):
            block[iCh::nChannels] = codes[iCh][:nSamples]
        self.fp.write(block.tobytes())
        return

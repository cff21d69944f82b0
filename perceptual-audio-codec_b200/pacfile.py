"""
pacfile.py -- PACFile with the reference's interface and file format (codec/pacfile.py:115-383), backed by the CUDA
engine.  Header layout, per-channel length-prefixed chunks, block sequencing (zero prior block, flush block in Close,
overlap tail at EOF) are the reference's; the per-block work (Encode, bit packing, chunk parsing, Decode) runs on the GPU.

`python pacfile.py name.wav` reproduces the reference's round trip driver (pacfile.py:388-503); see also
pacb200_batch.py for the batched whole-file API the per-block interface cannot express.
"""
import sys
from struct import calcsize, pack, unpack

import numpy as np

import _pacb200
import codec
from audiofile import AudioFile, CodingParams
from psychoac import AssignMDCTLinesFromFreqLimits, ScaleFactorBands

MAX16BITS = 32767
huffman = None            # the reference's Close() reads a module-global `huffman` (pacfile.py:365)


class PACFile(AudioFile):
    tag = b'PAC '

    def ReadFileHeader(self):                                            # pacfile.py:123-151
        tag = self.fp.read(4)
        if tag != self.tag:
            raise Exception("Tried to read a non-PAC file into a PACFile object")
        (sampleRate, nChannels, numSamples, nMDCTLines, nScaleBits, nMantSizeBits) = unpack('<LHLLHH', self.fp.read(calcsize('<LHLLHH')))
        nBands = unpack('<L', self.fp.read(calcsize('<L')))[0]
        nLines = unpack('<' + str(nBands) + 'H', self.fp.read(calcsize('<' + str(nBands) + 'H')))
        myParams = CodingParams()
        myParams.sampleRate = sampleRate
        myParams.nChannels = nChannels
        myParams.numSamples = numSamples
        myParams.nMDCTLines = myParams.nSamplesPerBlock = nMDCTLines
        myParams.nScaleBits = nScaleBits
        myParams.nMantSizeBits = nMantSizeBits
        myParams.sfBands = ScaleFactorBands(nLines)
        myParams.overlapAndAdd = [np.zeros(nMDCTLines, dtype=np.float64) for _ in range(nChannels)]
        return myParams

    def ReadDataBlock(self, codingParams, huffman):                     # pacfile.py:153-229
        cp = codingParams
        chunks = []
        for iCh in range(cp.nChannels):
            s = self.fp.read(calcsize("<L"))
            if not s:
                if cp.overlapAndAdd:
                    overlapAndAdd = cp.overlapAndAdd
                    cp.overlapAndAdd = 0
                    return overlapAndAdd
                return
            nBytes = unpack("<L", s)[0]
            payload = self.fp.read(nBytes)
            if len(payload) < nBytes:
                raise Exception("Only read a partial block of coded PACFile data")
            chunks.append(payload)
        cp.nTableIDBits = 4                                              # pacfile.py:189
        e = _pacb200.engine_for(cp)
        u = e.unpack_blocks([chunks])                                   # bit unpack + Huffman decode on the GPU
        LRMS = np.array([(int(u["lrms"][0]) >> b) & 1 for b in range(cp.sfBands.nBands)], dtype='int')
        decodedData = self.Decode(u["sf"][0], u["ba"][0], u["mant"][0], u["oscale"][0], cp, LRMS, huffman)
        data = []
        for iCh in range(cp.nChannels):
            data.append(np.add(cp.overlapAndAdd[iCh], decodedData[iCh][:cp.nMDCTLines]))     # :225
            cp.overlapAndAdd[iCh] = decodedData[iCh][cp.nMDCTLines:]                          # :226
        return data

    def WriteFileHeader(self, codingParams):                            # pacfile.py:231-271
        cp = codingParams
        self.fp.write(self.tag)
        if not cp.numSamples % cp.nMDCTLines:
            cp.numSamples += (cp.nMDCTLines - cp.numSamples % cp.nMDCTLines)     # the padding rule as the reference has it
        self.fp.write(pack('<LHLLHH', cp.sampleRate, cp.nChannels, cp.numSamples, cp.nMDCTLines, cp.nScaleBits, cp.nMantSizeBits))
        sfBands = ScaleFactorBands(AssignMDCTLinesFromFreqLimits(cp.nMDCTLines, cp.sampleRate))
        cp.sfBands = sfBands
        self.fp.write(pack('<L', sfBands.nBands))
        self.fp.write(pack('<' + str(sfBands.nBands) + 'H', *(sfBands.nLines.tolist())))
        cp.priorBlock = [np.zeros(cp.nMDCTLines, dtype=np.float64) for _ in range(cp.nChannels)]
        cp.extraBits = 0
        cp.curBlock = 0
        return

    def WriteDataBlock(self, data, codingParams, huffman):              # pacfile.py:273-353
        cp = codingParams
        fullBlockData = [np.concatenate((cp.priorBlock[iCh], data[iCh])) for iCh in range(cp.nChannels)]
        cp.priorBlock = data
        cp._pac_huffman = huffman
        self.Encode(fullBlockData, cp, huffman)                          # quantise + code + pack on the GPU
        for iCh in range(cp.nChannels):
            payload = cp._pac_chunks[iCh]
            self.fp.write(pack("<L", len(payload)))                      # :317
            self.fp.write(payload)                                       # :351
        return

    def Close(self, codingParams, huffman_=None):                       # pacfile.py:355-366
        if self.fp.mode == "wb":
            h = huffman_ or getattr(codingParams, "_pac_huffman", None) or huffman
            data = [np.zeros(codingParams.nMDCTLines, dtype=float), np.zeros(codingParams.nMDCTLines, dtype=float)]
            self.WriteDataBlock(data, codingParams, h)
        self.fp.close()

    def Encode(self, data, codingParams, huffman):                      # pacfile.py:368-375
        return codec.Encode(data, codingParams, huffman)

    def Decode(self, scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams, LRMS, huffman):    # :377-383
        return codec.Decode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams, LRMS)


def encode_decode(input_filename, coded_filename, output_filename, targetBitsPerSample=2.27, verbose=False):
    """The reference's __main__ loop (pacfile.py:430-499) as a function."""
    global huffman
    from Huffman import Huffman
    from pcmfile import PCMFile
    huffman = Huffman()
    for Direction in ("Encode", "Decode"):
        if Direction == "Encode":
            inFile, outFile = PCMFile(input_filename), PACFile(coded_filename)
        else:
            inFile, outFile = PACFile(coded_filename), PCMFile(output_filename)
        codingParams = inFile.OpenForReading()
        if Direction == "Encode":
            codingParams.nMDCTLines = 1024
            codingParams.nScaleBits = 4
            codingParams.nMantSizeBits = 4
            codingParams.targetBitsPerSample = targetBitsPerSample
            codingParams.nTableIDBits = 4
            codingParams.nSamplesPerBlock = codingParams.nMDCTLines
        else:
            codingParams.bitsPerSample = 16
            codingParams.nTableIDBits = 4
        outFile.OpenForWriting(codingParams)
        firstBlock = True
        while True:
            if Direction == "Encode":
                data = inFile.ReadDataBlock(codingParams)
            else:
                data = inFile.ReadDataBlock(codingParams, huffman)
            if not data:
                break
            if firstBlock and Direction == "Decode":
                firstBlock = False
                continue
            if Direction == "Encode":
                outFile.WriteDataBlock(data, codingParams, huffman)
            else:
                outFile.WriteDataBlock(data, codingParams)
            if verbose:
                sys.stdout.write(".")
                sys.stdout.flush()
        inFile.Close(codingParams)
        outFile.Close(codingParams)
    return huffman


if __name__ == "__main__":
    import time
    name = sys.argv[1] if len(sys.argv) > 1 else "castanets.wav"
    input_filename = "../inputs/" + name
    coded_filename = "../coded/" + name[:-4] + ".wak"
    output_filename = "../outputs/" + name[:-4] + ".wav"
    print("\nRunning the PAC coder ({} -> {} -> {}):".format(input_filename, coded_filename, output_filename))
    elapsed = time.time()
    h = encode_decode(input_filename, coded_filename, output_filename, verbose=True)
    print("\nSaved " + str(h.getBitDeposit()) + " Bits!")
    print("\nDone with Encode/Decode test\n")
    print(time.time() - elapsed, " seconds elapsed")

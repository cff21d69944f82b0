"""
bitpack.py -- PackedBits, the reference's MSB-first bit container (codec/bitpack.py:13-174).

This class is host-side state (a byte array with a cursor) kept for API parity: user code that packs its own custom
fields keeps working.  The codec's own packing does not go through it -- chunks are packed on the GPU by
csrc/pack.cuh and unpacked by csrc/decode.cuh.
"""
import numpy as np

BYTESIZE = 8


class PackedBits:
    def __init__(self):
        self.iByte = self.iBit = 0

    def Size(self, nBytes):
        self.nBytes = int(nBytes)
        self.iByte = self.iBit = 0
        self.data = np.zeros(self.nBytes, dtype=np.uint8)

    def GetPackedData(self):
        return self.data.tobytes()

    def SetPackedData(self, data):
        self.nBytes = len(data)
        self.data = np.frombuffer(data, dtype=np.uint8).copy()

    def WriteBits(self, info, nBits):
        """Writes the lowest nBits of info at the cursor, most significant bit first (bitpack.py:36-101)."""
        info = int(info)
        for i in range(int(nBits) - 1, -1, -1):
            if (info >> i) & 1:
                self.data[self.iByte] |= np.uint8(0x80 >> self.iBit)
            self.iBit += 1
            if self.iBit == BYTESIZE:
                self.iBit = 0
                self.iByte += 1

    def ReadBits(self, nBits):
        """Returns the next nBits at the cursor (bitpack.py:104-170)."""
        info = 0
        for _ in range(int(nBits)):
            info = (info << 1) | ((int(self.data[self.iByte]) >> (7 - self.iBit)) & 1)
            self.iBit += 1
            if self.iBit == BYTESIZE:
                self.iBit = 0
                self.iByte += 1
        return info

    def ResetPointers(self):
        self.iBit = self.iByte = 0

"""
Huffman.py -- run-time Huffman object with the reference's interface (codec/Huffman.py:253-374).

The pickled tables (huffmanTables.pickle, histograms.pickle) are the reference's unchanged fixtures; unpickling needs
importable classes named Huffman.HuffmanTable and Huffman.Histogram, which this module provides.  Table search, code
emission and decoding run inside the CUDA kernels (csrc/scan.cuh, pack.cuh, decode.cuh); this object carries the
tables and the savings pool (`bitDeposit`) that the reference keeps here.
"""
import os
import pickle
from collections import deque  # noqa: F401  (histograms.pickle references collections.deque)

import numpy as np

import _pacb200


class HuffmanTable:                                  # Huffman.py:138-153
    def __init__(self, manitssaCodeToHuffmanCode):
        self.encodingTable = manitssaCodeToHuffmanCode
        self.decodingTable = dict()
        for key, value in manitssaCodeToHuffmanCode.items():
            self.decodingTable[value] = key


class Histogram:                                     # Huffman.py:27-37 (only what unpickling needs)
    def __init__(self):
        self.LOW_FREQ = 10
        self.ESCAPE_CODE = -1


class Huffman:
    def __init__(self):
        with open(_pacb200.find_pickle('huffmanTables.pickle'), 'rb') as handle:        # Huffman.py:257-258
            self.huffmanTables = pickle.load(handle, encoding="latin1")
        with open(_pacb200.find_pickle('histograms.pickle'), 'rb') as handle:           # :259-260
            self.histograms = pickle.load(handle, encoding="latin1")
        self.ESCAPE_CODE = -1
        self.bitDeposit = 0

    # ---- savings pool, Huffman.py:353-374 (plain integer state that rides along with the stream)
    def depositBits(self, numBits):
        self.bitDeposit += numBits

    def withdrawBits(self):
        extraBit = 0
        if self.bitDeposit > 10:
            extraBit = self.bitDeposit // 100
            self.bitDeposit -= extraBit
        elif self.bitDeposit < 0:
            extraBit = self.bitDeposit
            self.bitDeposit = 0
        return extraBit

    def getBitDeposit(self):
        return self.bitDeposit

    # ---- Huffman.py:274-309: choose the table and emit code strings for one channel's unsigned mantissas
    def encodeData(self, codingParams, mantissaCode, bitAlloc):
        """Runs the device table search (the same kernel path codec.Encode uses) on synthetic lines that quantise to
        `mantissaCode`, and formats the winning table's codes as the reference's '0101' strings."""
        import codec
        return codec._huffman_encode(self, codingParams, mantissaCode, bitAlloc)

    # ---- Huffman.py:321-344
    def decodeData(self, bitReader, tableID, bitAlloc):
        """Decode ONE mantissa at the reader's cursor.  Single-symbol decoding is inherently serial host-side cursor
        work on a PackedBits object; whole chunks are decoded on the GPU (PACFile.ReadDataBlock -> pac_unpack_blocks)."""
        table = self.huffmanTables[tableID].decodingTable
        code = ""
        while code not in table:
            code += "1" if bitReader.ReadBits(1) else "0"
        m = table[code]
        if m == self.ESCAPE_CODE:
            return bitReader.ReadBits(bitAlloc)
        return m

"""
Huffman.py -- run-time Huffman object with the reference's interface (codec/Huffman.py:253-374), and the offline
trainer (HuffmanNode / Histogram / HuffmanTrainer, :15-250) with its counting pass on the GPU (pac_histogram).

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


class HuffmanNode:                                   # Huffman.py:15-24
    def __init__(self, mantissaCode=None, freq=None, left=None, right=None):
        self.zero = left
        self.one = right
        self.freq = freq
        self.mantissaCode = mantissaCode


class Histogram:
    """Huffman.py:27-135.  As in the reference, `probability`, `statistics` and `queue` are CLASS attributes: every
    Histogram of a process shares them.  A second trainer therefore keeps counting on top of the first, getMatchScore
    compares the shared dict with itself, and -- because makeHuffmanNodeQueue appends to the class-level deque and then
    binds a sorted COPY to the instance -- the leaves of earlier trainers are still queued when the next tree is built.
    Unpickling histograms.pickle only restores LOW_FREQ / ESCAPE_CODE."""
    probability = dict()
    statistics = dict()
    queue = deque()
    _seen = 0                                        # codes counted so far (positions of first occurrences on the device)

    def __init__(self):
        self.LOW_FREQ = 10
        self.ESCAPE_CODE = -1

    def getMatchScore(self, blockHistogram):         # :50-62
        mine = np.array([v for _, v in sorted(self.probability.items())])
        theirs = np.array([v for _, v in sorted(blockHistogram.probability.items())])
        return 3.0 - float(np.sum((mine - theirs) ** 2))

    def generateStatistics(self, mantissaCode):      # :71-83, the counting on the GPU
        """`mantissaCode`: sequence / numpy array / CUDA tensor of unsigned mantissa codes (no sign bit)."""
        e = _pacb200.engine()
        if hasattr(mantissaCode, "data_ptr"):
            codes, n = mantissaCode, int(mantissaCode.numel())
        else:
            codes = np.ascontiguousarray(np.asarray(mantissaCode, dtype=np.int64).astype(np.uint32))
            n = int(codes.size)
        counts, first = e.histogram(codes, nbins=1 << 16, base=Histogram._seen)
        Histogram._seen += n
        present = np.nonzero(counts)[0]
        stats = self.statistics
        for k in present[np.argsort(first[present], kind="stable")]:      # new keys enter in order of first occurrence
            k = int(k)
            stats[k] = stats.get(k, 0) + int(counts[k])
        total = sum(stats.values())
        for key, value in stats.items():
            self.probability[key] = value / float(total)

    def makeHuffmanNodeQueue(self):                  # :93-108
        escapeFreq = 0
        for code, freq in sorted(self.statistics.items(), key=lambda t: t[1]):       # stable: ties keep insertion order
            if freq < self.LOW_FREQ:
                escapeFreq += 1                      # the escape node's weight is the NUMBER of rare codes (:101)
            else:
                self.queue.append(HuffmanNode(code, freq))
        self.queue.append(HuffmanNode(self.ESCAPE_CODE, escapeFreq))
        self.queue = deque(sorted(self.queue, key=lambda t: t.freq))      # instance attribute from here on (:108)

    def appendToHuffmanQueue(self, huffmanNode):     # :117-119
        self.queue.append(huffmanNode)
        self.queue = deque(sorted(self.queue, key=lambda t: t.freq))

    def getNextPair(self):                           # :128-134
        if len(self.queue) == 1:
            return (self.queue.popleft(), None)
        firstNode = self.queue.popleft()
        return (firstNode, self.queue.popleft())


class HuffmanTrainer:                                # Huffman.py:156-250
    histogram = Histogram()
    huffmanCodeTable = dict()                        # class attribute in the reference too
    root = HuffmanNode()

    def __init__(self, tableID):
        self.tableID = tableID
        self.histogram = Histogram()

    def countFreq(self, mantissaCode):               # :184-185
        self.histogram.generateStatistics(mantissaCode)

    def constructHuffmanTable(self):                 # :195-211: the two pickles in the CWD are read, extended and rewritten
        self.histogram.makeHuffmanNodeQueue()
        self._buildEncodingTree()
        self._buildEncodingTable()
        with open('huffmanTables.pickle', 'rb') as handle:
            huffmanTables = _pacb200.safe_load(handle, _FIXTURE_CLASSES)
        huffmanTables[self.tableID] = HuffmanTable(self.huffmanCodeTable)
        with open('huffmanTables.pickle', 'wb') as handle:
            pickle.dump(huffmanTables, handle, protocol=0)
        with open('histograms.pickle', 'rb') as handle:
            histograms = _pacb200.safe_load(handle, _FIXTURE_CLASSES)
        histograms[self.tableID] = self.histogram
        with open('histograms.pickle', 'wb') as handle:
            pickle.dump(histograms, handle, protocol=0)

    def _buildEncodingTree(self):                    # :221-228; the re-sorted deque is a stable priority queue, so a heap keyed
        import heapq                                 # (freq, arrival number) pops the same pairs without re-sorting every time
        heap = [(n.freq, i, n) for i, n in enumerate(self.histogram.queue)]
        heapq.heapify(heap)
        seq = len(heap)
        while len(heap) > 1:
            f1, _, n1 = heapq.heappop(heap)
            f2, _, n2 = heapq.heappop(heap)
            heapq.heappush(heap, (f1 + f2, seq, HuffmanNode(None, f1 + f2, n1, n2)))
            seq += 1
        self.root = heap[0][2]
        self.histogram.queue.clear()

    def _buildEncodingTable(self):                   # :238-250, iterative (the reference recurses)
        stack = [(self.root, "")]
        while stack:
            node, code = stack.pop()
            if node.mantissaCode is not None:
                self.huffmanCodeTable[node.mantissaCode] = code
                continue
            stack.append((node.one, code + "1"))
            stack.append((node.zero, code + "0"))


_FIXTURE_CLASSES = {"HuffmanTable": HuffmanTable, "Histogram": Histogram, "HuffmanNode": HuffmanNode}


class Huffman:
    def __init__(self):
        with open(_pacb200.find_pickle('huffmanTables.pickle'), 'rb') as handle:        # Huffman.py:257-258
            self.huffmanTables = _pacb200.safe_load(handle, _FIXTURE_CLASSES)
        with open(_pacb200.find_pickle('histograms.pickle'), 'rb') as handle:           # :259-260
            self.histograms = _pacb200.safe_load(handle, _FIXTURE_CLASSES)
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

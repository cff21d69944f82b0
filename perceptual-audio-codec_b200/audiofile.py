"""
audiofile.py -- CodingParams attribute bag and the abstract AudioFile (reference: codec/audiofile.py:51-92).
Host-side file plumbing only; no arithmetic happens here.
"""


class CodingParams:
    """A class to hold coding parameters to share across files (audiofile.py:51-53): attributes are added at run time."""
    pass


class AudioFile:
    """Handlers expected of a data file containing audio data (audiofile.py:56-92)."""

    def __init__(self, filename):
        self.filename = filename

    def OpenForReading(self):
        self.fp = open(self.filename, "rb")
        codingParams = self.ReadFileHeader()
        return codingParams

    def OpenForWriting(self, codingParams):
        self.fp = open(self.filename, "wb")
        self.WriteFileHeader(codingParams)

    def Close(self, codingParams):
        self.fp.close()

    def ReadFileHeader(self):
        return CodingParams()

    def ReadDataBlock(self, codingParams):
        pass

    def WriteFileHeader(self, codingParams):
        pass

    def WriteDataBlock(self, data, codingParams):
        pass

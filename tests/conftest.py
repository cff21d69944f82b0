import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")
CORPUS = os.path.join(GOLD, "_corpus")          # optional, git-ignored copy of the reference's inputs/*.wav
PKG = os.path.join(REPO, "perceptual-audio-codec_b200")

for p in (REPO, os.path.join(REPO, "oracle"), PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def oracle():
    import oracle as _o
    return _o.get()


@pytest.fixture(scope="session")
def gold_dir():
    return GOLD


@pytest.fixture(scope="session")
def manifest():
    import json
    with open(os.path.join(GOLD, "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kats():
    import json
    with open(os.path.join(GOLD, "kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def stages():
    import numpy as np
    return np.load(os.path.join(GOLD, "stages.npz"))


def corpus_files():
    """wav paths available for whole-file parity: the two committed fixtures always, the full
    reference corpus when tests/golden/_corpus has been populated (build() does it when
    /root/reference is present)."""
    out = {}
    for n in ("piano_test2", "castanets"):
        out[n] = os.path.join(GOLD, n + ".wav")
    if os.path.isdir(CORPUS):
        for f in sorted(os.listdir(CORPUS)):
            if f.endswith(".wav"):
                out[f[:-4]] = os.path.join(CORPUS, f)
    return out

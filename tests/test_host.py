"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol include/pac_b200.h declares,
host logic of the reference-named shim modules, stream sharding + the byte-count gather over gloo (world size 2).
No compute calls are made (there is no GPU here and the engine has no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "perceptual-audio-codec_b200")


def test_library_exports_every_declared_symbol():
    import build as pkg_build
    lib_path = pkg_build.build()
    hdr = open(os.path.join(REPO, "include", "pac_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pac_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.pac_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pac_version()


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", os.path.join(PKG, "libpacb200.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(8\d|9\d)\b", out)


def test_no_cpu_fallback_without_gpu():
    """With no CUDA device the engine must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import _pacb200
    with pytest.raises(_pacb200.PacError) as ei:
        _pacb200.Engine(0, "fp64")
    assert ei.value.code in (-5, -2)


def test_product_never_imports_the_oracle():
    bad = []
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"(import\s+oracle|from\s+oracle|pac_oracle|orc_[a-z_]+\()", txt):
                    bad.append(f)
    assert not bad, bad


def test_band_layout_and_bands_object(kats):
    import psychoac
    for fs in (44100, 48000, 22050, 32000):
        assert psychoac.AssignMDCTLinesFromFreqLimits(1024, fs) == kats["bands"][str(fs)]
    sfb = psychoac.ScaleFactorBands(kats["bands"]["44100"])
    assert sfb.nBands == 25 and sfb.lowerLine[0] == 0 and sfb.upperLine[-1] == 1023
    assert list(sfb.upperLine - sfb.lowerLine + 1) == kats["bands"]["44100"]


def test_scalar_psy_helpers(kats):
    import psychoac
    p = kats["psy_scalar"]
    np.testing.assert_allclose([psychoac.Bark(f) for f in p["f"]], p["Bark"], rtol=1e-15)
    np.testing.assert_allclose([psychoac.Thresh(f) for f in p["f"]], p["Thresh"], rtol=1e-14)
    np.testing.assert_allclose([psychoac.SPL(v) for v in (1.0, 1e-3, 1e-13, 0.0)], p["SPL"], rtol=1e-15)


def test_packedbits_kat(kats):
    import bitpack
    k = kats["bitpack"]
    bp = bitpack.PackedBits()
    bp.Size(2)
    for v, w in zip(k["values"], k["widths"]):
        bp.WriteBits(v, w)
    assert bp.GetPackedData().hex() == k["bytes_hex"] == "3ab7"
    bp2 = bitpack.PackedBits()
    bp2.SetPackedData(bp.GetPackedData())
    assert [bp2.ReadBits(w) for w in k["widths"]] == k["readback"]


def test_huffman_object_loads_unchanged_pickles(kats):
    cwd = os.getcwd()
    os.chdir(PKG)                      # the reference opens the pickles relative to the CWD (Huffman.py:257-260)
    try:
        import Huffman
        h = Huffman.Huffman()
    finally:
        os.chdir(cwd)
    assert sorted(h.huffmanTables) == list(range(1, 11))
    for tid, f in kats["huffman_tables"].items():
        t = h.huffmanTables[int(tid)]
        assert len(t.encodingTable) == f["nsym"] and t.encodingTable[-1] == f["esc"]
        assert t.decodingTable[f["esc"]] == -1
    for dep, w, after in kats["withdraw"]:
        h.bitDeposit = dep
        assert h.withdrawBits() == w and h.bitDeposit == after
    import hashlib
    sha = hashlib.sha256(open(os.path.join(PKG, "huffmanTables.pickle"), "rb").read()).hexdigest()
    assert sha == "6e59e09578f07e5910da509acfe76c26a268adbc512f4f9a52483842012895e2"


def test_flattened_tables_match_pickle():
    import _pacb200
    tabs = _pacb200.load_encoding_tables(os.path.join(PKG, "huffmanTables.pickle"))
    nkeys, off, code, ln, ec, el = _pacb200.flatten_tables(tabs)
    for t in range(10):
        enc = tabs[t + 1]
        for k, s in enc.items():
            if k >= 0:
                assert ln[off[t] + k] == len(s) and code[off[t] + k] == int(s, 2)
        assert int(np.count_nonzero(ln[off[t]:off[t] + nkeys[t]])) == len(enc) - 1
        assert (int(ec[t]), int(el[t])) == (int(enc[-1], 2), len(enc[-1]))
    assert ln.max() <= 31 and nkeys.max() <= 17920


def test_wav_header_walk_and_roundtrip(gold_dir):
    import pacb200_batch as pb
    rate, pcm = pb.read_wav(os.path.join(gold_dir, "castanets.wav"))      # odd-sized file with trailing bytes
    assert rate == 44100 and pcm.shape == (397488, 2)
    gold = open(os.path.join(gold_dir, "piano_test2.out.wav"), "rb").read()
    # the decoder's WAV header carries numSamples of the PAC header (pcmfile.py:107), fewer than the samples written
    nhdr = int.from_bytes(gold[40:44], "little") // 4
    body = np.frombuffer(gold[44:], dtype="<i2").reshape(-1, 2)
    assert nhdr < body.shape[0] == 174 * 1024
    assert pb.wav_bytes(body, 44100, nhdr) == gold


def test_shard_assignment():
    import pacb200_batch as pb
    S = 11
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in pb.shard_streams(S, r, world))
        assert seen == list(range(S))
        assert all(s % world == r for r in range(world) for s in pb.shard_streams(S, r, world))


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import pacb200_batch as pb
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, S = dist.get_rank(), 9
truth = np.arange(S, dtype=np.int64) * 1000 + 76
mine = pb.shard_streams(S, rank, 2)
full = pb.gather_byte_counts(truth[mine], S, rank, 2, dist=dist)
assert full.tolist() == truth.tolist(), full
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_byte_count_gather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), PKG, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) prints ONE JSON line with the contract's keys;
    a non-zero rank under torchrun exits 0 without work.  Runs on the host cores only (the oracle port)."""
    import json
    import subprocess
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-seconds", "0.5"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio-seconds encoded per second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_hostile_pickle_in_cwd_is_refused(tmp_path, monkeypatch):
    """The reference opens huffmanTables.pickle relative to the CWD (Huffman.py:257-260) and so does the engine; a pickle found there
    that references anything but the fixture's three globals must be refused, not executed."""
    import pickle
    import _pacb200

    class Evil(object):
        def __reduce__(self):
            return (os.system, ("touch %s" % (tmp_path / "pwned"),))

    (tmp_path / "huffmanTables.pickle").write_bytes(pickle.dumps({1: Evil()}, protocol=0))
    monkeypatch.chdir(tmp_path)
    with pytest.raises(pickle.UnpicklingError):
        _pacb200.load_encoding_tables()
    assert not (tmp_path / "pwned").exists()
    monkeypatch.chdir(os.path.dirname(_pacb200.__file__))
    tabs = _pacb200.load_encoding_tables()               # the packaged fixture still loads
    assert sorted(tabs) == list(range(1, 11))


def test_corpus_generator_is_backend_independent():
    """bench.py's synthetic corpus (corpus.py): counter-based Philox4x32-10 (Random123 known answers) and integer-only per-sample
    arithmetic, so the numpy and the torch backends give the same int16 samples -- the CPU arm of the bench encodes the very streams the
    GPU arm does, wherever they were generated."""
    import torch
    sys.path.insert(0, REPO)
    import corpus
    z = np.zeros(1, dtype=np.int64)
    assert [int(v[0]) for v in corpus.philox4x32(z, z, z, z, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.full(1, 0xffffffff, dtype=np.int64)
    assert [int(v[0]) for v in corpus.philox4x32(f, f, f, f, 0xffffffff, 0xffffffff)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    p = np.array([0x243f6a88], dtype=np.int64), np.array([0x85a308d3], dtype=np.int64), np.array([0x13198a2e], dtype=np.int64), np.array([0x03707344], dtype=np.int64)
    assert [int(v[0]) for v in corpus.philox4x32(*p, 0xa4093822, 0x299f31d0)] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    ids, n = [0, 3, 77, 4095], 2 * 44100 + 123
    a = corpus.gen_streams_numpy(ids, n)
    b = corpus.gen_streams(ids, n, "cpu").numpy()
    assert a.dtype == np.int16 and a.shape == (4, n, 2) and np.array_equal(a, b)
    assert np.array_equal(corpus.gen_streams_numpy([77], n)[0], a[2])                       # a stream depends on its id only
    x = a.astype(np.float64) / 32767.0
    rms = 20 * np.log10(np.sqrt(np.mean(x ** 2, axis=1)))
    assert np.all(rms < -3) and np.all(rms > -40) and np.abs(a).max() <= 32767
    assert len({a[i].tobytes() for i in range(4)}) == 4

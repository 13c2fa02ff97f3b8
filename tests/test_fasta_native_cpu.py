"""Row f3: the native FASTA -> corpus-buffer reader (csrc/fasta.cu, host code, no GPU needed)
against the Python twin of the reference parser (host/data_utils.py:DataLoader.parse_sequences, itself checked
against the reference's own parse_sequences in tests/golden) + host/corpus.py:stream_chunks."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from protgram_directgcn_b200 import _native as nat
from protgram_directgcn_b200.host import corpus
from protgram_directgcn_b200.host.data_utils import DataLoader


def _python_chunks(path, chunk_bytes=1 << 20, rank=0, world=1, block=4096):
    return [bytes(c) for c in corpus.stream_chunks((s for _, s in DataLoader.parse_sequences(path)), chunk_bytes, rank, world, block)]


def _native_chunks(path, chunk_bytes=1 << 20, rank=0, world=1, block=4096, stats=None):
    return [bytes(c.numpy()) for c in corpus.stream_chunks_native(path, chunk_bytes, rank, world, block, stats=stats)]


def _write(tmp_path, data: bytes, name="x.fasta"):
    p = tmp_path / name
    p.write_bytes(data)
    return str(p)


from tests.fasta_cases import CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_native_reader_matches_python_parser(name, tmp_path):
    path = _write(tmp_path, CASES[name])
    stats = {}
    assert b"".join(_native_chunks(path, stats=stats)) == b"".join(_python_chunks(path))
    assert stats["sequences"] == sum(1 for _ in DataLoader.parse_sequences(path))
    assert stats["stopped_early"] == name.startswith("bare_header")


def test_native_reader_small_chunks_and_rank_dealing(tmp_path):
    rng = np.random.default_rng(5)
    recs = []
    for i in range(1500):
        seq = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYacd"), size=int(rng.integers(1, 200))))
        lines = [seq[j:j + 60] for j in range(0, len(seq), 60)]
        recs.append(f">id{i} x\n" + "\n".join(lines) + "\n")
    path = _write(tmp_path, "".join(recs).encode())
    whole = b"".join(_python_chunks(path))
    for cap in (1 << 16, 70_000, 1 << 20):
        assert b"".join(_native_chunks(path, cap)) == whole
    for world in (2, 3):
        for rank in range(world):
            assert b"".join(_native_chunks(path, 1 << 16, rank, world, block=64)) == b"".join(_python_chunks(path, 1 << 16, rank, world, block=64))
    # chunks are cut at record boundaries
    for c in _native_chunks(path, 1 << 16):
        assert c.endswith(b" \xff")


def test_native_reader_long_record_grows_the_chunk(tmp_path):
    path = _write(tmp_path, b">a\n" + b"ACDE" * 100_000 + b"\n>b\nAC\n")
    assert b"".join(_native_chunks(path, 1 << 16)) == b"".join(_python_chunks(path))


def test_native_reader_refuses_non_ascii_and_missing_file(tmp_path):
    path = _write(tmp_path, ">a\nAC\xc3\xa9DE\n".encode("latin-1"))
    with pytest.raises(corpus.NonAsciiSequence):
        _native_chunks(path)
    ok = _write(tmp_path, ">a \xc3\xa9\nACDE\n".encode("latin-1"), "hdr.fasta")     # non-ASCII in a HEADER is fine (ids are not used)
    assert b"".join(_native_chunks(ok)) == b" ACDE \xff"
    with pytest.raises(FileNotFoundError):
        _native_chunks(str(tmp_path / "missing.fasta"))


LINE = st.one_of(st.just(b""), st.just(b">"), st.binary(max_size=12).map(lambda b: bytes(c & 0x7F for c in b)),
                 st.text(alphabet="ACDEFGacdefg >|\t ", max_size=20).map(str.encode),
                 st.text(alphabet="abcXYZ|_ ", min_size=1, max_size=10).map(lambda s: b">" + s.encode()))
EOL = st.sampled_from([b"\n", b"\r\n", b"\r"])


@settings(max_examples=300, deadline=None)
@given(st.lists(st.tuples(LINE, EOL), max_size=25), st.booleans())
def test_native_reader_fuzz(tmp_path_factory, lines, final_eol):
    data = b"".join(l + e for l, e in lines)
    if not final_eol and lines:
        data = data[:-len(lines[-1][1])]
    path = str(tmp_path_factory.mktemp("fz") / "f.fasta")
    with open(path, "wb") as fh:
        fh.write(data)
    assert b"".join(_native_chunks(path)) == b"".join(_python_chunks(path)), data


# ------------------------------------------------------------------------------------------------ 5-bit host format
def test_pack5_roundtrip_against_spec_and_refusal():
    import torch
    from tests import kernel_spec
    rng = np.random.default_rng(9)
    alphabet = np.frombuffer(b" ABCDEFGHIJKLMNOPQRSTUVWXYZ*-.\xff", dtype=np.uint8)
    for n in (0, 1, 7, 8, 9, 15, 16, 17, 1000, 4099):
        buf = alphabet[rng.integers(0, alphabet.size, n)]
        chunk = corpus.pack5(buf, pinned=False)
        assert chunk is not None and chunk.n_symbols == n and chunk.packed.numel() == (n + 7) // 8 * 5
        out = torch.zeros(max(n, 1), dtype=torch.uint8)
        kernel_spec.pg_unpack5(chunk.packed, n, out)
        assert np.array_equal(out[:n].numpy(), buf)
    assert corpus.pack5(np.frombuffer(b"ACD1EF", dtype=np.uint8), pinned=False) is None      # '1' has no code: keep bytes
    assert corpus.pack5(np.frombuffer(b"acd", dtype=np.uint8), pinned=False) is None          # the reader upper-cases first


def test_graph_builder_streams_packed_host_chunks(monkeypatch, tmp_path):
    """Corpus beyond the HBM budget: chunks stay on the host in the 5-bit format and are re-uploaded + unpacked per level;
    the graphs equal the all-resident run."""
    import pickle
    from tests import kernel_spec
    from protgram_directgcn_b200.host import data_builder
    from protgram_directgcn_b200.host.config import Config
    kernel_spec.install(monkeypatch, nat)
    rng = np.random.default_rng(4)
    recs = "".join(f">p{i}\n" + "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWY"), size=int(rng.integers(5, 90)))) + "\n" for i in range(400))
    fasta = _write(tmp_path, recs.encode())
    graphs = {}
    real_unpack5 = corpus.unpack5
    for name, resident in (("resident", 1 << 40), ("streamed", 0)):
        cfg = Config()
        cfg.GCN_INPUT_FASTA_PATH = fasta
        cfg.BASE_OUTPUT_DIR = str(tmp_path / name)
        cfg.GRAPH_OBJECTS_DIR = str(tmp_path / name / "graphs")
        cfg.GCN_NGRAM_MAX_N = 2
        cfg.GRAPH_BUILDER_CHUNK_BYTES = 1 << 12                  # several chunks
        cfg.GRAPH_BUILDER_RESIDENT_BYTES = resident
        seen = []
        monkeypatch.setattr(corpus, "unpack5", lambda *a, _seen=seen, **k: (_seen.append(1), real_unpack5(*a, **k))[1])
        data_builder.GraphBuilder(cfg).run()
        assert bool(seen) == (name == "streamed")
        graphs[name] = [pickle.load(open(os.path.join(cfg.GRAPH_OBJECTS_DIR, f"ngram_graph_n{n}.pkl"), "rb")) for n in (1, 2)]
    for a, b in zip(graphs["resident"], graphs["streamed"]):
        assert a.node_sequences == b.node_sequences and a.number_of_edges == b.number_of_edges
        assert np.array_equal(a.A_out_w.coalesce().indices().numpy(), b.A_out_w.coalesce().indices().numpy())
        assert np.array_equal(a.A_out_w.coalesce().values().numpy(), b.A_out_w.coalesce().values().numpy())
        assert np.array_equal(a.mathcal_A_in.coalesce().values().numpy(), b.mathcal_A_in.coalesce().values().numpy())


# ------------------------------------------------------------------------------------------------ parallel whole-file reader
@pytest.mark.parametrize("name", sorted(CASES))
def test_parallel_reader_matches_streaming_reader_on_edge_files(name, tmp_path):
    path = _write(tmp_path, CASES[name])
    ref_stats, stats = {}, {}
    ref = b"".join(_native_chunks(path, stats=ref_stats))
    for threads in (1, 3):
        got = bytes(corpus.read_fasta_parallel(path, threads=threads, stats=stats).numpy())
        assert got == ref and stats == ref_stats


def test_parallel_reader_large_file_threads_and_ranks(tmp_path):
    rng = np.random.default_rng(11)
    recs = []
    for i in range(30_000):
        seq = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYacd"), size=int(rng.integers(1, 120))))
        recs.append(f">id{i} x\n" + "\n".join(seq[j:j + 60] for j in range(0, len(seq), 60)) + ("\r\n" if i % 7 == 0 else "\n"))
        if i == 17_000:
            recs.append(">emptyrecord\n\n")
    data = "".join(recs).encode()
    assert len(data) > (1 << 20)                    # several ranges
    path = _write(tmp_path, data)
    ref_stats = {}
    ref = b"".join(_native_chunks(path, 1 << 22, stats=ref_stats))
    for threads in (1, 2, 8):
        stats = {}
        assert bytes(corpus.read_fasta_parallel(path, threads=threads, stats=stats).numpy()) == ref and stats == ref_stats
    for rank in range(3):
        assert bytes(corpus.read_fasta_parallel(path, threads=4, rank=rank, world=3, block=64).numpy()) == \
            b"".join(_native_chunks(path, 1 << 22, rank, 3, 64))
    # a bare '>' in the middle stops every later range as well
    stop = data[:len(data) // 2].rsplit(b"\n>", 1)[0] + b"\n>\n" + data[len(data) // 2:]
    path2 = _write(tmp_path, stop, "stop.fasta")
    s1, s2 = {}, {}
    assert bytes(corpus.read_fasta_parallel(path2, threads=8, stats=s1).numpy()) == b"".join(_native_chunks(path2, 1 << 22, stats=s2))
    assert s1 == s2 and s1["stopped_early"]
    # chunking of a whole buffer at separators
    whole = corpus.read_fasta_parallel(path, threads=4)
    parts = corpus.split_at_separators(whole, 1 << 18)
    assert b"".join(bytes(p.numpy()) for p in parts) == ref and all(bytes(p.numpy()).endswith(b"\xff") for p in parts)


@settings(max_examples=200, deadline=None)
@given(st.binary(max_size=300))
def test_pack5_property(data):
    """Every buffer over the code's alphabet round-trips; any other byte makes pack5 refuse (never a silent substitution)."""
    import torch
    from tests import kernel_spec
    buf = np.frombuffer(data, dtype=np.uint8)
    ok = all((c == 0x20) or (0x41 <= c <= 0x5A) or c in (0x2A, 0x2D, 0x2E, 0xFF) for c in data)
    chunk = corpus.pack5(buf, pinned=False) if len(data) else corpus.pack5(np.zeros(0, dtype=np.uint8), pinned=False)
    if not ok:
        assert chunk is None
        return
    assert chunk is not None and chunk.n_symbols == len(data)
    out = torch.zeros(max(len(data), 1), dtype=torch.uint8)
    kernel_spec.pg_unpack5(chunk.packed, len(data), out)
    assert bytes(out[:len(data)].numpy()) == data


@pytest.mark.parametrize("name", sorted(CASES))
def test_readers_match_reference_generated_golden(name, tmp_path):
    """tests/golden/fasta_cases.npz holds what the reference's own DataLoader.parse_sequences yields for every edge file
    (tests/golden/make_golden_fasta.py): the Python twin, the streaming native reader and the multi-threaded one must all
    produce exactly those records."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fasta_cases.npz"))
    count = int(g[name + "_count"])
    ids = str(g[name + "_ids"]).split("\x1f") if count else []
    seqs = str(g[name + "_seqs"]).split("\x1f") if count else []
    path = _write(tmp_path, CASES[name])
    twin = list(DataLoader.parse_sequences(path))
    assert [r[0] for r in twin] == ids and [r[1] for r in twin] == seqs
    expect = b"".join(((b" " if i == 0 else b"") + s.encode() + b" \xff") for i, s in enumerate(seqs))
    stats = {}
    assert b"".join(_native_chunks(path, stats=stats)) == expect and stats["sequences"] == count
    assert bytes(corpus.read_fasta_parallel(path, threads=2).numpy()) == expect


# ------------------------------------------------------------------------------------------------ windowed parallel reader
def _windows(path, window_bytes, stats=None, **kw):
    return b"".join(bytes(w.numpy()) for w in corpus.stream_fasta_windows(path, window_bytes, stats=stats, **kw))


@pytest.mark.parametrize("name", sorted(CASES))
def test_windowed_reader_matches_reference_golden_on_edge_files(name, tmp_path):
    """Files larger than host memory are walked window by window (pg_fasta_pack_window); whatever the window size, the bytes
    are those the reference's own parse_sequences implies (tests/golden/fasta_cases.npz) and the counters those of the
    streaming reader."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fasta_cases.npz"))
    count = int(g[name + "_count"])
    seqs = str(g[name + "_seqs"]).split("\x1f") if count else []
    expect = b"".join(((b" " if i == 0 else b"") + s.encode() + b" \xff") for i, s in enumerate(seqs))
    path = _write(tmp_path, CASES[name])
    ref_stats = {}
    _native_chunks(path, stats=ref_stats)
    for window in (1, 2, 5, 11, 1 << 20):
        stats = {}
        assert _windows(path, window, stats=stats, threads=2) == expect, window
        assert stats == ref_stats, window


def test_windowed_reader_large_file_threads_ranks_and_stop(tmp_path):
    rng = np.random.default_rng(12)
    recs = []
    for i in range(40_000):
        seq = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYacd"), size=int(rng.integers(1, 150))))
        recs.append(f">id{i} x\n" + "\n".join(seq[j:j + 60] for j in range(0, len(seq), 60)) + ("\r\n" if i % 5 == 0 else "\n"))
        if i % 9_000 == 0:
            recs.append(">emptyrecord\n\n")
    data = "".join(recs).encode()
    assert len(data) > (3 << 20)
    path = _write(tmp_path, data)
    ref_stats = {}
    ref = bytes(corpus.read_fasta_parallel(path, threads=4, stats=ref_stats).numpy())
    for window, threads in ((1 << 20, 1), (1 << 20, 8), ((1 << 21) + 12345, 3), (700_001, 4), (1 << 30, 4)):
        stats = {}
        assert _windows(path, window, stats=stats, threads=threads) == ref and stats == ref_stats, (window, threads)
    for rank in range(3):
        assert _windows(path, (1 << 20) + 7, threads=4, rank=rank, world=3, block=64) == \
            bytes(corpus.read_fasta_parallel(path, threads=4, rank=rank, world=3, block=64).numpy())
    stop = data[:len(data) // 2].rsplit(b"\n>", 1)[0] + b"\n>\n" + data[len(data) // 2:]
    path2 = _write(tmp_path, stop, "stop.fasta")
    s1, s2 = {}, {}
    assert _windows(path2, 1 << 20, stats=s1, threads=4) == bytes(corpus.read_fasta_parallel(path2, threads=4, stats=s2).numpy())
    assert s1 == s2 and s1["stopped_early"]


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(LINE, EOL), max_size=25), st.booleans(), st.integers(min_value=1, max_value=40))
def test_windowed_reader_fuzz(tmp_path_factory, lines, final_eol, window):
    data = b"".join(l + e for l, e in lines)
    if not final_eol and lines:
        data = data[: len(data) - len(lines[-1][1])]
    path = _write(tmp_path_factory.mktemp("w"), data)
    try:
        ref_stats = {}
        ref = b"".join(_native_chunks(path, stats=ref_stats))
    except corpus.NonAsciiSequence:
        with pytest.raises(corpus.NonAsciiSequence):
            _windows(path, window)
        return
    stats = {}
    assert _windows(path, window, stats=stats, threads=2) == ref and stats == ref_stats


def test_graph_builder_takes_the_windowed_reader_for_large_files(monkeypatch, tmp_path):
    """GraphBuilder.run picks the windowed reader when the file exceeds a quarter of host memory (forced here)."""
    from tests import kernel_spec
    from protgram_directgcn_b200.host import data_builder
    from protgram_directgcn_b200.host.config import Config
    from tests.helpers import load, fasta_sequences
    from tests.test_host_logic_cpu import check_graph_against_golden
    import protgram_directgcn_b200 as pg
    kernel_spec.install(monkeypatch, nat)
    g = load("build_protein")
    cfg = Config()
    cfg.GCN_INPUT_FASTA_PATH = fasta_sequences(str(g["fasta"]), tmp_path)
    cfg.BASE_OUTPUT_DIR = str(tmp_path / "o")
    cfg.GRAPH_OBJECTS_DIR = str(tmp_path / "o" / "g")
    cfg.GCN_NGRAM_MAX_N = 2
    cfg.GRAPH_BUILDER_FASTA_WINDOW_BYTES = 97
    cfg.GRAPH_BUILDER_CHUNK_BYTES = 64
    used = []
    real = corpus.stream_fasta_windows
    monkeypatch.setattr(corpus, "stream_fasta_windows", lambda *a, **k: (used.append(1), real(*a, **k))[1])
    monkeypatch.setattr(os, "sysconf", lambda name: 1)           # "host memory" of a few bytes
    data_builder.GraphBuilder(cfg).run()
    assert used
    for n in (1, 2):
        check_graph_against_golden(pg.DataUtils.load_object(os.path.join(cfg.GRAPH_OBJECTS_DIR, f"ngram_graph_n{n}.pkl")), g, n)


def test_windowed_reader_grows_its_buffer_for_a_record_far_longer_than_the_window(tmp_path):
    rng = np.random.default_rng(5)
    long_seq = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWY"), size=300_000))
    data = (">short\nACDE\n>long\n" + "\n".join(long_seq[i:i + 80] for i in range(0, len(long_seq), 80)) + "\n>tail\nGG\n").encode()
    path = _write(tmp_path, data)
    expect = b" ACDE \xff" + long_seq.encode() + b" \xffGG \xff"
    stats = {}
    assert _windows(path, 10, stats=stats, threads=2) == expect and stats["sequences"] == 3

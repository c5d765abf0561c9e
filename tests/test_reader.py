"""Host-side input path (ntl_seqfile_*, api.SeqFile / prefetch_batches): FASTA/FASTQ, plain and gzip, multi-line records,
CRLF, ids cut at the first whitespace (bin/read_fasta.py:6-46), batching at record boundaries, lines that straddle the
reader's 8 MiB blocks, buffer ownership of the zero-copy arrays. No GPU involved."""
import gc
import gzip
import os

import numpy as np
import pytest

import util
from ntlink_b200 import api


def reference_parse(text):
    "plain-Python restatement of bin/read_fasta.py:6-46 (readfq): [(id, seq)]"
    out, lines, i = [], text.split("\n"), 0
    while i < len(lines):
        if not lines[i] or lines[i][0] not in ">@":
            i += 1
            continue
        name = lines[i][1:].split()[0] if lines[i][1:].split() else ""
        i += 1
        seq = []
        while i < len(lines) and (not lines[i] or lines[i][0] not in ">@+"):
            seq.append(lines[i].rstrip("\r"))
            i += 1
        s = "".join(seq)
        if i < len(lines) and lines[i] and lines[i][0] == "+":
            i += 1
            q = 0
            while i < len(lines) and q < len(s):
                q += len(lines[i].rstrip("\r"))
                i += 1
        out.append((name, s))
    return out


def make_text(rng, n, fastq, width):
    recs = []
    for r in range(n):
        L = int(rng.integers(0, 700))
        s = "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, L))
        hdr = f"rec{r} some description\tmore"
        if fastq:
            recs.append(f"@{hdr}\n{s}\n+\n{'I' * L}\n")
        else:
            body = "\n".join(s[i:i + width] for i in range(0, L, width)) if width else s
            recs.append(f">{hdr}\n{body}\n" if L else f">{hdr}\n")
    return "".join(recs)


@pytest.mark.parametrize("fastq,width,gz,crlf", [(False, 60, False, False), (False, 0, True, False), (True, 0, False, False),
                                                  (True, 0, True, False), (False, 70, False, True)])
def test_reader_matches_readfq(tmp_path, fastq, width, gz, crlf):
    rng = np.random.default_rng(11 + width + 2 * gz + fastq)
    text = make_text(rng, 300, fastq, width)
    if crlf:
        text = text.replace("\n", "\r\n")
    path = os.path.join(str(tmp_path), "x.fq" if fastq else "x.fa") + (".gz" if gz else "")
    with (gzip.open(path, "wt", newline="") if gz else open(path, "w", newline="")) as fout:
        fout.write(text)
    want = reference_parse(text.replace("\r\n", "\n"))
    whole = api.read_sequences(path)
    got = [(n, whole.seq[int(a):int(b)].tobytes().decode()) for n, a, b in zip(whole.names, whole.offsets, whole.offsets[1:])]
    assert got == want
    # streamed in small batches: same records, cut at record boundaries
    streamed = []
    for batch in api.prefetch_batches([path], 5000):
        assert int(batch.offsets[-1]) < 5000 + 700
        streamed += [(n, batch.seq[int(a):int(b)].tobytes().decode()) for n, a, b in zip(batch.names, batch.offsets, batch.offsets[1:])]
    assert streamed == want


def test_lines_longer_than_a_block_and_fixture_files(tmp_path):
    rng = np.random.default_rng(3)
    long_seq = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, 20_000_000))      # one 20 Mbp line: spans three 8 MiB blocks
    path = os.path.join(str(tmp_path), "long.fa")
    with open(path, "w") as fout:
        fout.write(f">a\n{long_seq}\n>b x\nACGT\nAC\n")
    b = api.read_sequences(path)
    assert b.names == ["a", "b"] and b.offsets.tolist() == [0, 20_000_000, 20_000_006]
    assert b.seq[:20_000_000].tobytes().decode() == long_seq and b.seq[20_000_000:].tobytes() == b"ACGTAC"
    # the reference's fixtures against the test-suite's own FASTA loader
    for name in ("scaffolds_3.fa", "long_reads_2.fq"):
        f = util.fixture_file(tmp_path, name)
        names, seq, offs = util.load_fasta_batch(f)
        got = api.read_sequences(f)
        assert got.names == names and np.array_equal(got.offsets, offs) and np.array_equal(got.seq, seq)


def test_zero_copy_arrays_keep_their_buffer_alive(tmp_path):
    path = os.path.join(str(tmp_path), "y.fa")
    with open(path, "w") as fout:
        fout.write(">s\n" + "ACGT" * 5000 + "\n")
    batch = api.read_sequences(path)
    part = batch.seq[100:108]
    del batch
    gc.collect()
    assert part.tobytes() == b"ACGTACGT"
    assert api.read_sequences(os.path.join(str(tmp_path), "y.fa"), max_bases=1).offsets.tolist() == [0, 20000]
    with pytest.raises(OSError):
        api.read_sequences(os.path.join(str(tmp_path), "missing.fa"))
    empty = os.path.join(str(tmp_path), "empty.fa")
    open(empty, "w").close()
    assert len(api.read_sequences(empty)) == 0


def _records(path, max_bases, threads):
    os.environ["NTL_READER_THREADS"] = str(threads)
    try:
        out, sizes = [], []
        for batch in api.prefetch_batches([path], max_bases):
            sizes.append(int(batch.offsets[-1]))
            out += [(n, batch.seq[int(a):int(b)].tobytes()) for n, a, b in zip(batch.names, batch.offsets, batch.offsets[1:])]
        return out, sizes
    finally:
        del os.environ["NTL_READER_THREADS"]


@pytest.mark.parametrize("width,crlf,tail", [(60, False, ""), (0, False, ""), (71, True, ""), (60, False, "fastq"), (60, False, "junk")])
def test_parallel_fasta_reader_equals_sequential(tmp_path, width, crlf, tail):
    "plain FASTA files are parsed by several threads from the mapped file; same records and same batch cuts as one thread"
    rng = np.random.default_rng(width + 5)
    text = "leading line without a header\n" + make_text(rng, 4000, False, width)
    if tail == "fastq":                       # FASTQ records after FASTA ones: the parallel reader must hand over
        text += make_text(rng, 50, True, 0)
    elif tail == "junk":
        text += ">last no newline at the end\nACGTACGT"
    if crlf:
        text = text.replace("\n", "\r\n")
    path = os.path.join(str(tmp_path), "p.fa")
    with open(path, "w", newline="") as fout:
        fout.write(text)
    want = [(n, s.encode()) for n, s in reference_parse(text.replace("\r\n", "\n"))]
    for max_bases in (0, 200_000, 3_000):
        seq_recs, seq_sizes = _records(path, max_bases, 1)
        for threads in (2, 5):
            par_recs, par_sizes = _records(path, max_bases, threads)
            assert par_recs == seq_recs == want
            if tail != "fastq":               # after a hand-over the cuts may differ, the records may not
                assert par_sizes == seq_sizes


def fastq_text(rng, n, adversarial=True, multiline=False):
    recs = []
    for r in range(n):
        L = int(rng.integers(0, 500))
        s = "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, L))
        q = "".join("@+>I5#"[int(x)] for x in rng.integers(0, 6, L)) if adversarial else "I" * L     # quality lines may start with @ + >
        if multiline and L > 60:
            s = s[:60] + "\n" + s[60:]
            q = q[:60] + "\n" + q[60:]
        recs.append(f"@rec{r} desc\t{r}\n{s}\n+{'rec%d' % r if r % 3 == 0 else ''}\n{q}\n")
    return "".join(recs)


@pytest.mark.parametrize("variant", ["plain", "crlf", "multiline", "no_final_newline", "blank_lines", "fasta_tail"])
def test_parallel_fastq_reader_equals_sequential(tmp_path, variant):
    """4-line FASTQ is parsed in parallel too; record starts are recognised by structure (a quality line may start with
    '@'), and anything that is not strictly 4-line FASTQ is handed to the sequential reader"""
    rng = np.random.default_rng(len(variant))
    text = fastq_text(rng, 3000, multiline=(variant == "multiline"))
    if variant == "crlf":
        text = text.replace("\n", "\r\n")
    elif variant == "no_final_newline":
        text = text[:-1]
    elif variant == "blank_lines":
        text = text.replace("\n@rec7 ", "\n\n\n@rec7 ").replace("\n@rec1500 ", "\n\n@rec1500 ")
    elif variant == "fasta_tail":
        text += make_text(rng, 40, False, 60)
    path = os.path.join(str(tmp_path), "p.fq")
    with open(path, "w", newline="") as fout:
        fout.write(text)
    want = [(n, q.encode()) for n, q in reference_parse(text.replace("\r\n", "\n"))]
    for max_bases in (0, 100_000, 2_000):
        seq_recs, seq_sizes = _records(path, max_bases, 1)
        assert len(seq_recs) >= 3000 and seq_recs == want
        for threads in (2, 7):
            par_recs, par_sizes = _records(path, max_bases, threads)
            assert par_recs == seq_recs
            if variant in ("plain", "crlf", "no_final_newline", "blank_lines"):
                assert par_sizes == seq_sizes


@pytest.mark.parametrize("kind,block", [("fasta", 65280), ("fasta", 900), ("fastq", 65280), ("fastq", 1500), ("fastq_fasta_tail", 4000)])
def test_bgzf_input_is_inflated_in_parallel(tmp_path, kind, block):
    """bgzip (BGZF) files: members located by their headers, inflated by several threads, parsed by the parallel parser;
    same records as the plain file and as zlib's sequential reader, same batch cuts as the plain file"""
    import gzip as gz
    rng = np.random.default_rng(block)
    if kind == "fasta":
        text = make_text(rng, 4000, False, 60)
    else:
        text = fastq_text(rng, 3000)
        if kind == "fastq_fasta_tail":
            text += make_text(rng, 40, False, 60)
    plain = os.path.join(str(tmp_path), "p.txt")
    with open(plain, "w", newline="") as fout:
        fout.write(text)
    bz = os.path.join(str(tmp_path), "p.bgz.gz")
    util.write_bgzf(bz, text.encode(), block, eof_marker=(block != 900))
    assert gz.open(bz, "rb").read() == text.encode()                 # a valid multi-member gzip file
    want = [(n, s.encode()) for n, s in reference_parse(text)]
    for max_bases in (0, 150_000, 2_500):
        one, _ = _records(bz, max_bases, 1)                          # one thread: zlib's gzread
        assert one == want
        plain_recs, plain_sizes = _records(plain, max_bases, 4)
        for threads in (2, 6):
            recs, sizes = _records(bz, max_bases, threads)
            assert recs == want
            if kind != "fastq_fasta_tail":
                assert sizes == plain_sizes


def test_bgzf_corruption_is_an_error(tmp_path):
    rng = np.random.default_rng(9)
    text = make_text(rng, 3000, False, 60).encode()
    path = os.path.join(str(tmp_path), "bad.fa.gz")
    util.write_bgzf(path, text, 20000)
    raw = bytearray(open(path, "rb").read())
    raw[len(raw) // 2] ^= 0x55                                       # somewhere inside a member of the middle of the file
    open(path, "wb").write(bytes(raw))
    os.environ["NTL_READER_THREADS"] = "4"
    try:
        with pytest.raises(Exception):
            api.read_sequences(path)
    finally:
        del os.environ["NTL_READER_THREADS"]

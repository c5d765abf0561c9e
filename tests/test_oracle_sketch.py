"""The sketch oracle (oracle/indexlr_oracle.c) against the reference's own golden vectors and the
intermediate known-answer values of SURVEY.md Appendix A. CPU only."""
import ctypes

import numpy as np
import pytest

import util

GOLDEN_SKETCHES = [("scaffolds_1.fa", 32, 250), ("scaffolds_2.fa", 32, 100),
                   ("scaffolds_3.fa", 24, 250), ("scaffolds_4.fa", 40, 100)]


@pytest.mark.parametrize("name,k,w", GOLDEN_SKETCHES)
def test_golden_target_sketch_byte_exact(tmp_path, name, k, w):
    got = util.oracle_indexlr(util.fixture_file(tmp_path, name), k, w)
    assert got == util.expected_output(f"{name}.k{k}.w{w}.tsv")


def _hashes(kmer):
    lib = util.oracle_lib()
    vals = [ctypes.c_uint64() for _ in range(4)]
    lib.ntl_oracle_kmer_hashes(kmer.encode(), len(kmer), *[ctypes.byref(v) for v in vals])
    return [v.value for v in vals]


def test_appendix_a_known_answers():
    lib = util.oracle_lib()
    seed_a = 0x3c8bfbb395c60474
    assert lib.ntl_oracle_srol(seed_a, 1) == 0x7917f7652b8c08e9
    assert lib.ntl_oracle_srol(seed_a, 32) == 0x7917f764cae3023a
    seq = "AAGGAAGGAACATGTTGCAAATCCAGTGGCCGGGGGAGGGGGGC"
    expect = [(0x85825eddfacafb1a, 0xc6111c735e0cad0c, 0x4b937b5158d7a826, 6440039448140340696),
              (0x3cc1e79994c0ef74, 0x6463aea0a2fda781, 0xa125963a37be96f5, 16713728021746434175),
              (0x4e46951348d4c7a8, 0xb55af7c95c8522c7, 0x03a18cdca559ea6f, 10216980414591918508),
              (0xd8c2b475386f3e55, 0xd1a51f45b71857e9, 0xaa67d3baef87963e, 10435180475771958535)]
    for pos, exp in enumerate(expect):
        assert tuple(_hashes(seq[pos:pos + 40])) == exp
    fh, rh, h0, h1 = _hashes("ACGTACGTACGTACGTACGTACGTACGTACGT")
    assert fh == rh == 0x6b60211785bb95f3 and h0 == 0xd6c0422f0b772be6 and h1 == 10513942262040716327


def test_short_and_invalid_sequences():
    lib = util.oracle_lib()
    h = np.empty(64, np.uint64); p = np.empty(64, np.uint32); s = np.empty(64, np.uint8)
    def sk(seq, k, w):
        n = lib.ntl_oracle_sketch(seq.encode(), len(seq), k, w, h.ctypes.data, p.ctypes.data, s.ctypes.data, 64)
        return list(zip(h[:n].tolist(), p[:n].tolist(), s[:n].tolist()))
    assert sk("ACGT", 8, 2) == []                      # k > L
    assert sk("ACGTACGTAC", 8, 4) == []                # w > L-k+1
    assert sk("ACGTNCGTACNTAGGATN", 4, 2) != []        # N handling: windows span the gaps
    assert sk("NNNNNNNNNNNN", 4, 2) == []
    # lowercase acgt hash like uppercase, any other letter is invalid
    up, lo = "ACGGTCATTGCAGTCAGTCCATGACGT", "acggtcattgcagtcagtccatgacgt"
    assert sk(up, 5, 3) == sk(lo, 5, 3)
    # window over valid k-mers only: an N in the middle must not reset the window
    a = sk("ACGGTCATTGCAGTCAGTNCATGACGTTTGACCA", 4, 6)
    assert all(0 <= pos <= 30 for _, pos, _ in a) and len(a) > 0

"""Shared helpers for the test-suite (fixtures, oracle access, output comparison rules of SURVEY.md App. B)."""
import ctypes
import gzip
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_EXE = os.path.join(ORACLE_DIR, "_build", "indexlr_oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "libntl_oracle.so")

if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)


def ensure_oracle():
    if not (os.path.exists(ORACLE_EXE) and os.path.exists(ORACLE_LIB)):
        subprocess.check_call(["make", "-C", ORACLE_DIR])


def manifest():
    with open(os.path.join(GOLD, "manifest.json")) as fin:
        return json.load(fin)


def gunzip_bytes(path):
    with gzip.open(path, "rb") as fin:
        return fin.read()


def fixture_file(tmpdir, name):
    "Materialise tests/golden/inputs/<name>.gz as tmpdir/<name>; returns the path"
    dst = os.path.join(str(tmpdir), name)
    if not os.path.exists(dst):
        tmp = f"{dst}.{os.getpid()}.tmp"          # atomic: several test processes may share tmpdir
        with open(tmp, "wb") as fout:
            fout.write(gunzip_bytes(os.path.join(GOLD, "inputs", name + ".gz")))
        os.replace(tmp, dst)
    return dst


def golden_case(case, what):
    return gunzip_bytes(os.path.join(GOLD, "cases", case, what + ".gz"))


def expected_output(name):
    return gunzip_bytes(os.path.join(GOLD, "expected_outputs", name + ".gz"))


def oracle_indexlr(path, k, w, length=False, threads=4, pos=True, strand=True):
    ensure_oracle()
    cmd = [ORACLE_EXE, "--long", "-k", str(k), "-w", str(w), "-t", str(threads)]
    if pos:
        cmd.append("--pos")
    if strand:
        cmd.append("--strand")
    if length:
        cmd.append("--len")
    cmd.append(path)
    return subprocess.check_output(cmd)


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        ensure_oracle()
        lib = ctypes.CDLL(ORACLE_LIB)
        lib.ntl_oracle_sketch.restype = ctypes.c_size_t
        lib.ntl_oracle_sketch.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        lib.ntl_oracle_sketch_batch.restype = ctypes.c_size_t
        lib.ntl_oracle_sketch_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint,
                                                ctypes.c_uint, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        lib.ntl_oracle_kmer_hashes.restype = None
        lib.ntl_oracle_kmer_hashes.argtypes = [ctypes.c_char_p, ctypes.c_uint] + [ctypes.POINTER(ctypes.c_uint64)] * 4
        lib.ntl_oracle_srol.restype = ctypes.c_uint64
        lib.ntl_oracle_srol.argtypes = [ctypes.c_uint64, ctypes.c_uint]
        _lib = lib
    return _lib


def oracle_sketch_batch(seq_bytes, offsets, k, w, threads=4):
    """seq_bytes: uint8 array of concatenated sequences; offsets: uint64[nseq+1].
    Returns (hash u64[], pos u32[], strand u8[], mx_off u64[nseq+1])."""
    lib = oracle_lib()
    seq = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    nseq = len(offsets) - 1
    cap = max(1024, int(len(seq) * 2.5 / max(1, w) + 4 * nseq) + 1024)
    while True:
        h = np.empty(cap, np.uint64)
        p = np.empty(cap, np.uint32)
        s = np.empty(cap, np.uint8)
        off = np.empty(nseq + 1, np.uint64)
        n = lib.ntl_oracle_sketch_batch(seq.ctypes.data, offsets.ctypes.data, nseq, k, w, threads,
                                        h.ctypes.data, p.ctypes.data, s.ctypes.data, cap, off.ctypes.data)
        if n != ctypes.c_size_t(-1).value:
            return h[:n], p[:n], s[:n], off
        cap *= 4


def dot_parts(data):
    "Split a .scaffold.dot into (header lines, sorted node lines, ordered edge lines, tail) per SURVEY App. B"
    lines = data.decode().splitlines()
    head = lines[:2]
    nodes = sorted(l for l in lines[2:] if "->" not in l and l != "}")
    edges = [l for l in lines[2:] if "->" in l]
    return head, nodes, edges, lines[-1]


def is_ordered_subsequence(small_lines, big_lines):
    it = iter(big_lines)
    return all(any(x == y for y in it) for x in small_lines)


# ---------------------------------------------------------------------------------------------------
# host emulation of the kernel logic (tests/emu) -- CPU tests only
EMU_DIR = os.path.join(HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libntl_emu.so")
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        subprocess.check_call(["make", "-s", "-C", EMU_DIR])
        lib = ctypes.CDLL(EMU_LIB)
        lib.emu_sketch.restype = ctypes.c_int64
        lib.emu_sketch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                   ctypes.c_uint32, ctypes.c_double, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        lib.emu_order_first_seen.restype = None
        lib.emu_order_first_seen.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        lib.emu_map.restype = ctypes.c_int64
        lib.emu_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        lib.emu_liftover.restype = ctypes.c_uint32
        lib.emu_liftover.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32] + \
            [ctypes.c_void_p] * 3
        _emu = lib
    return _emu


def lift_golden(case, what):
    return gunzip_bytes(os.path.join(GOLD, "liftover", case, what + ".gz"))


def lift_manifest():
    with open(os.path.join(GOLD, "liftover", "manifest.json")) as fin:
        return json.load(fin)


def verbose_text(hit_off, nruns, runs, hits, read_ids, contig_names):
    "ntl_map_out-style arrays -> verbose_mapping.tsv text (pure Python; for emulation tests)"
    out = []
    for r, rid in enumerate(read_ids):
        base = int(hit_off[r])
        for i in range(int(nruns[r])):
            ctg, start, count = (int(v) for v in runs[base + i])
            toks = [f"{int(h[1]) & 0x7FFFFFFF}:{'+' if int(h[1]) >> 31 else '-'}_{int(h[2]) & 0x7FFFFFFF}:{'+' if int(h[2]) >> 31 else '-'}"
                    for h in hits[base + start: base + start + count]]
            out.append(f"{rid}\t{contig_names[ctg]}\t{count}\t{' '.join(toks)}\n")
    return "".join(out)


def emu_sketch(seq, offsets, k, w, S=256, c=10.0, cap_override=0):
    "returns (hash, pos_strand, mx_off, stats dict)"
    lib = emu_lib()
    seq = np.ascontiguousarray(seq, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    nseq = len(offsets) - 1
    cap = len(seq) + 16
    h = np.empty(cap, np.uint64)
    p = np.empty(cap, np.uint32)
    off = np.empty(nseq + 1, np.uint64)
    st = np.zeros(4, np.uint64)
    n = lib.emu_sketch(seq.ctypes.data, offsets.ctypes.data, nseq, k, w, S, c, cap_override, h.ctypes.data,
                       p.ctypes.data, cap, off.ctypes.data, st.ctypes.data)
    assert n >= 0
    return h[:n], p[:n], off, {"cand": int(st[0]), "gaps": int(st[1]), "ovf": int(st[2]), "stackovf": int(st[3])}


def load_fasta_batch(path):
    "tiny pure-Python FASTA/FASTQ reader for tests: returns (names, seq uint8, offsets uint64)"
    import gzip as _gz
    opener = _gz.open if path.endswith(".gz") else open
    names, parts = [], []
    with opener(path, "rt") as fin:
        lines = fin.read().split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln[:1] in (">", "@"):
            names.append(ln[1:].split()[0])
            i += 1
            s = []
            while i < len(lines) and lines[i][:1] not in (">", "@", "+"):
                s.append(lines[i].strip())
                i += 1
            seq = "".join(s)
            parts.append(seq)
            if i < len(lines) and lines[i][:1] == "+":
                i += 1
                q = 0
                while q < len(seq) and i < len(lines):
                    q += len(lines[i])
                    i += 1
        else:
            i += 1
    offs = np.zeros(len(parts) + 1, np.uint64)
    if parts:
        offs[1:] = np.cumsum([len(p) for p in parts])
    seq = np.frombuffer("".join(parts).encode(), dtype=np.uint8)
    return names, seq, offs


def write_bgzf(path, data, block=65280, eof_marker=True):
    "bgzip-style file: independent gzip members of <= `block` input bytes with the BC extra field (SAM spec 4.1)"
    import struct
    import zlib

    def member(chunk):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = co.compress(chunk) + co.flush()
        bsize = 18 + len(payload) + 8
        assert bsize <= 65536
        return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + payload +
                struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))

    with open(path, "wb") as fout:
        for i in range(0, len(data), block):
            fout.write(member(data[i:i + block]))
        if eof_marker:
            fout.write(member(b""))

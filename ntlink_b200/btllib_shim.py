"""btllib-shaped access to the GPU sketcher.

`ntlink_patch_gaps.py` (bin/ntlink_patch_gaps.py:203,269,417-439) and `ntlink_filter_sequences.py` (:37) use btllib's
Python objects:  `btllib.Indexlr(path, k, w, flags, threads)` iterates records with `.id`, `.readlen`,
`.minimizers[i].out_hash / .pos / .forward`, and `btllib.SeqReader(path, flags)` iterates `.id`, `.seq`.
This module offers the same shapes on top of libntlink_b200.so, so `import ntlink_b200.btllib_shim as btllib` is
enough for those call sites. Only what ntLink touches is provided."""
from collections import namedtuple

import numpy as np

from . import api

Minimizer = namedtuple("Minimizer", ["out_hash", "pos", "forward"])


class IndexlrFlag:
    "flag names ntLink passes (values are irrelevant here: long mode only affects btllib's reader buffering)"
    NO_ID = 1
    BX = 2
    SEQ = 4
    FILTER_IN = 8
    FILTER_OUT = 16
    SHORT_MODE = 32
    LONG_MODE = 64
    QUIET = 128


class SeqReaderFlag:
    SHORT_MODE = 1
    LONG_MODE = 2


class Record:
    __slots__ = ("num", "id", "barcode", "readlen", "minimizers")

    def __init__(self, num, id_, readlen, minimizers):
        self.num, self.id, self.barcode, self.readlen, self.minimizers = num, id_, "", readlen, minimizers


class Indexlr:
    """with Indexlr(path, k, w, flags, threads) as idx: for rec in idx: ...   (records in input order)"""

    def __init__(self, seqfile, k, w, flags=IndexlrFlag.LONG_MODE, threads=4, verbose=False, device=0, ctx=None):
        self.path, self.k, self.w = seqfile, k, w
        self._ctx = ctx or api.Context(device)
        self._own = ctx is None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self._own and self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    def __iter__(self):
        batch = api.read_sequences(self.path)
        sk = self._ctx.sketch(batch, self.k, self.w)
        pos, fwd, off, lens = sk.pos, sk.strand, sk.seq_off, batch.lengths
        for i, name in enumerate(batch.names):
            a, b = int(off[i]), int(off[i + 1])
            mins = [Minimizer(int(h), int(p), bool(f)) for h, p, f in zip(sk.hash[a:b], pos[a:b], fwd[a:b])]
            yield Record(i, name, int(lens[i]), mins)


SeqRecord = namedtuple("SeqRecord", ["num", "id", "comment", "seq", "qual"])


class SeqReader:
    "iterates records with .id and .seq (host-side reader of libntlink_b200.so, bin/read_fasta.py semantics)"

    def __init__(self, seqfile, flags=SeqReaderFlag.LONG_MODE, threads=1):
        self.path = seqfile

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        batch = api.read_sequences(self.path)
        raw = batch.seq.tobytes()
        for i, name in enumerate(batch.names):
            yield SeqRecord(i, name, "", raw[int(batch.offsets[i]):int(batch.offsets[i + 1])].decode(), "")

"""Thin Python layer over the C ABI (include/ntlink_b200.h): numpy in, numpy out. The surrounding ntLink Python
stays Python and calls the CUDA path through this module (BASELINE.json north_star)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import POS_MASK, STRAND_BIT, NtlError


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _np_from(ptr, n, dtype):
    "copy n items from a ctypes pointer into a fresh numpy array"
    if n == 0 or not ptr:
        return np.empty(0, dtype)
    addr = C.cast(ptr, C.c_void_p).value
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def name_ranks(names):
    "rank of every contig name under Python string comparison (normalize_pair, bin/ntlink_pair.py:213-219)"
    order = sorted(range(len(names)), key=lambda i: names[i])
    rank = np.empty(len(names), np.uint32)
    for r, i in enumerate(order):
        rank[i] = r
    return rank


class SeqBatch:
    """Sequences back to back: seq uint8[total], offsets uint64[n+1], names list[str]. Batches that come from the native
    reader carry the names as one blob (`from_blob`): the text emitters take that blob as it is, and the Python strings are
    only made if somebody asks for `.names` (a million reads are a million small string objects otherwise)."""

    def __init__(self, seq, offsets, names):
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._names = list(names)
        assert len(self.offsets) == len(self._names) + 1
        self._name_blob = None

    @classmethod
    def from_blob(cls, seq, offsets, blob, name_off):
        "names given as bytes back to back (blob) + offsets uint64[n+1]"
        out = cls.__new__(cls)
        out.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        out.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        assert len(out.offsets) == len(name_off)
        out._names = None
        out._name_blob = (np.frombuffer(bytes(blob) + b"\0", dtype=np.uint8), np.ascontiguousarray(name_off, dtype=np.uint64))
        return out

    @property
    def names(self):
        if self._names is None:
            blob, off = self._name_blob
            raw, no = blob.tobytes(), off.tolist()
            self._names = [raw[no[i]:no[i + 1]].decode() for i in range(len(no) - 1)]
        return self._names

    @names.setter
    def names(self, value):
        self._names = list(value)
        self._name_blob = None

    @classmethod
    def from_strings(cls, named_seqs):
        names, parts, offs, tot = [], [], [0], 0
        for name, s in named_seqs:
            b = s.encode() if isinstance(s, str) else bytes(s)
            names.append(name)
            parts.append(b)
            tot += len(b)
            offs.append(tot)
        seq = np.frombuffer(b"".join(parts) + b"N" * 64, dtype=np.uint8)[:tot] if tot else np.empty(0, np.uint8)
        return cls(seq, np.array(offs, np.uint64), names)

    def __len__(self):
        return len(self.offsets) - 1

    @property
    def lengths(self):
        return np.diff(self.offsets).astype(np.uint64)

    def name_blob(self):
        if self._name_blob is None:
            enc = [n.encode() for n in self._names]
            off = np.zeros(len(enc) + 1, np.uint64)
            if enc:
                off[1:] = np.cumsum([len(e) for e in enc])
            self._name_blob = (np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8), off)
        return self._name_blob


class _CBuffer:
    """Owner of a malloc'ed buffer returned by the library, exposed through __array_interface__: numpy keeps this object
    as the base of every array (and view) made from it, and the buffer is freed when the last of them is gone."""

    def __init__(self, lib, ptr, n, typestr):
        self.lib, self.ptr = lib, ptr
        self.__array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}

    def __del__(self):
        if self.ptr:
            self.lib.ntl_free(self.ptr)
            self.ptr = None


def _text_out(lib, out, n, copy):
    """Text produced by the library's emitters: bytes (copy=True), or a memoryview straight over the library's buffer for
    callers that only write it to a file -- the buffer goes back to the emitters' pool (ntl_buf_free) when the view dies,
    so the next batch is formatted into memory that is already mapped."""
    if copy or n == 0:
        data = C.string_at(out, n)
        lib.ntl_buf_free(out)
        return data
    import weakref
    arr = (C.c_char * n).from_address(out.value)
    weakref.finalize(arr, lib.ntl_buf_free, C.c_void_p(out.value))
    return memoryview(arr)


def _np_view(lib, ptr, n, dtype):
    "numpy array over a library buffer WITHOUT copying"
    if n == 0 or not ptr:
        if ptr:
            lib.ntl_free(ptr)
        return np.empty(0, dtype)
    return np.asarray(_CBuffer(lib, ptr, n, np.dtype(dtype).str))


class SeqFile:
    """Streaming FASTA/FASTQ reader (plain or gzip, multi-line; id = header up to the first whitespace,
    bin/read_fasta.py:6-46) over the library's block reader. `read(max_bases)` returns the next SeqBatch of about
    max_bases bases (whole records), None at the end of the file; the arrays are views of the library's buffers."""

    def __init__(self, path):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        if self.lib.ntl_seqfile_open(path.encode(), C.byref(self.h)) != 0:
            raise OSError(f"cannot open {path}")

    def read(self, max_bases=0):
        if not self.h:
            return None
        seq, off, names, noff = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_uint32()
        rc = self.lib.ntl_seqfile_read(self.h, max_bases, C.byref(seq), C.byref(off), C.byref(names), C.byref(noff), C.byref(n))
        if rc != 0:
            raise NtlError(rc, "ntl_seqfile_read failed")
        nseq = n.value
        offsets = _np_from(off, nseq + 1, np.uint64)
        name_off = _np_from(noff, nseq + 1, np.uint64)
        total = int(offsets[-1]) if nseq else 0
        nb = C.string_at(names, int(name_off[-1])) if nseq else b""
        for p in (off, names, noff):
            self.lib.ntl_free(p)
        if nseq == 0:
            self.lib.ntl_free(seq)
            return None
        s = _np_view(self.lib, seq.value, total, np.uint8)
        return SeqBatch.from_blob(s, offsets, nb, name_off)

    def close(self):
        if self.h:
            self.lib.ntl_seqfile_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_verbose_mappings(path, contig_names=None, share_repeated=True, max_hits=0):
    """verbose_mapping.tsv (plain or gzip) -> batches (hit_off, nruns, runs (n,3), hits (n,3), read_len, read_ids, contig_names)
    through the library's native parser (ntl_verbose_*): the arrays of Context.tally_mappings / liftover_mappings. With
    contig_names the ids follow that table and unknown contigs raise; without, ids are assigned in order of appearance and
    the table grows from batch to batch. max_hits = 0: one batch for the whole file."""
    lib = _lib.load()
    h = C.c_void_p()
    if contig_names is not None:
        enc = [n.encode() for n in contig_names]
        off = np.zeros(len(enc) + 1, np.uint64)
        if enc:
            off[1:] = np.cumsum([len(e) for e in enc])
        blob = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8)
        rc = lib.ntl_verbose_open(path.encode(), _ptr(blob), _ptr(off), len(enc), C.byref(h))
    else:
        rc = lib.ntl_verbose_open(path.encode(), None, None, 0, C.byref(h))
    if rc != 0:
        raise OSError(f"cannot open {path}")
    try:
        while True:
            mo = _lib.MappingsOut()
            rc = lib.ntl_verbose_read(h, max_hits, int(bool(share_repeated)), C.byref(mo))
            if rc != 0:
                raise ValueError((lib.ntl_verbose_error(h) or b"malformed mappings file").decode())
            n, slots = mo.n_reads, mo.n_slots
            hit_off = _np_from(mo.hit_off, n + 1, np.uint32)
            nruns = _np_from(mo.nruns, n, np.uint32)
            read_len = _np_from(mo.read_len, n, np.uint32)
            runs = _np_from(mo.runs, slots * 3, np.uint32).reshape(-1, 3)
            hits = _np_from(mo.hits, slots * 3, np.uint32).reshape(-1, 3)
            rno = _np_from(mo.read_name_off, n + 1, np.uint64).tolist()
            rnb = C.string_at(mo.read_names, rno[-1]) if n else b""
            cno = _np_from(mo.ctg_name_off, mo.n_contigs + 1, np.uint64).tolist()
            cnb = C.string_at(mo.ctg_names, cno[-1]) if mo.n_contigs else b""
            for p in (mo.hit_off, mo.nruns, mo.read_len, mo.runs, mo.hits, mo.read_name_off, mo.ctg_name_off):
                lib.ntl_free(C.cast(p, C.c_void_p))
            lib.ntl_free(mo.read_names)
            lib.ntl_free(mo.ctg_names)
            if n == 0:
                return
            ids = [rnb[rno[i]:rno[i + 1]].decode() for i in range(n)]
            names = [cnb[cno[i]:cno[i + 1]].decode() for i in range(mo.n_contigs)]
            yield hit_off, nruns, runs, hits, read_len, ids, names
            if max_hits == 0:
                return
    finally:
        lib.ntl_verbose_close(h)


def read_sketch_tsv(path, with_len, max_mx=0):
    """indexlr TSV (plain or gzip, '-' = stdin) -> batches (names, lengths u32 or None, Sketch) through the library's
    native parser (ntl_tsv_*): whole records until about max_mx minimizers per batch (0 = one batch for the whole file)."""
    lib = _lib.load()
    h = C.c_void_p()
    if lib.ntl_tsv_open(path.encode(), int(bool(with_len)), C.byref(h)) != 0:
        raise OSError(f"cannot open {path}")
    try:
        while True:
            to = _lib.TsvOut()
            if lib.ntl_tsv_read(h, max_mx, C.byref(to)) != 0:
                raise ValueError((lib.ntl_tsv_error(h) or b"malformed sketch TSV").decode())
            n = to.n_seq
            hashes = _np_from(to.hash, to.n_mx, np.uint64)
            posf = _np_from(to.pos_strand, to.n_mx, np.uint32)
            off = _np_from(to.mx_off, n + 1, np.uint64)
            lens = _np_from(to.seq_len, n, np.uint32) if with_len else None
            no = _np_from(to.name_off, n + 1, np.uint64).tolist()
            nb = C.string_at(to.names, no[-1]) if n else b""
            for ptr in (to.hash, to.pos_strand, to.mx_off, to.seq_len, to.name_off):
                if ptr:
                    lib.ntl_free(C.cast(ptr, C.c_void_p))
            lib.ntl_free(to.names)
            if n == 0:
                return
            yield [nb[no[i]:no[i + 1]].decode() for i in range(n)], lens, Sketch(hashes, posf, off)
            if max_mx == 0:
                return
    finally:
        lib.ntl_tsv_close(h)


def read_sequences(path, max_bases=0):
    """FASTA/FASTQ (plain or gzip, multi-line) -> SeqBatch of the whole file (or of its first ~max_bases bases)."""
    with SeqFile(path) as f:
        batch = f.read(max_bases)
    if batch is None:
        return SeqBatch(np.empty(0, np.uint8), np.zeros(1, np.uint64), [])
    return batch


def prefetch_batches(paths, max_bases):
    """Iterator over the SeqBatches of several files; the next batch is decoded by a background thread (the library call
    releases the GIL) while the caller works on the current one."""
    import queue
    import threading
    q = queue.Queue(maxsize=2)

    def producer():
        try:
            for path in paths:
                with SeqFile(path) as f:
                    while True:
                        b = f.read(max_bases)
                        if b is None:
                            break
                        q.put(b)
            q.put(None)
        except BaseException as exc:      # hand the error to the consumer
            q.put(exc)

    threading.Thread(target=producer, daemon=True).start()
    while True:
        item = q.get()
        if item is None:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


class Sketch:
    "indexlr output as arrays: hash u64[n], pos_strand u32[n], seq_off u64[nseq+1]"

    def __init__(self, hash_, pos_strand, seq_off):
        self.hash, self.pos_strand, self.seq_off = hash_, pos_strand, seq_off

    @property
    def pos(self):
        return self.pos_strand & np.uint32(POS_MASK)

    @property
    def strand(self):
        return (self.pos_strand >> np.uint32(31)).astype(np.uint8)

    def __len__(self):
        return len(self.hash)

    def _struct(self):
        s = _lib.SketchOut()
        s.n_mx, s.nseq = len(self.hash), len(self.seq_off) - 1
        s.hash = C.cast(self.hash.ctypes.data, C.POINTER(C.c_uint64))
        s.pos_strand = C.cast(self.pos_strand.ctypes.data, C.POINTER(C.c_uint32))
        s.seq_off = C.cast(self.seq_off.ctypes.data, C.POINTER(C.c_uint64))
        return s

    def to_tsv(self, batch, with_len=False, with_pos=True, with_strand=True, threads=4, copy=True):
        "the bytes `indexlr --long [--pos] [--strand] [--len]` writes for these sequences (copy=False: see _text_out)"
        lib = _lib.load()
        blob, noff = batch.name_blob()
        lens = batch.lengths if with_len else None
        out = C.c_void_p()
        st = self._struct()
        n = lib.ntl_format_sketch_tsv(C.byref(st), _ptr(blob), _ptr(noff), _ptr(lens), int(with_pos), int(with_strand),
                                      threads, C.byref(out))
        if n < 0:
            raise NtlError(n, "ntl_format_sketch_tsv")
        return _text_out(lib, out, n, copy)


class MapResult:
    """Accepted mappings + pair events of a batch of reads (copies of ntl_map_out)."""

    def __init__(self, mo):
        n = mo.n_reads
        self.n_reads, self.n_mx, self.n_hits, self.n_runs, self.n_events = n, mo.n_mx, mo.n_hits, mo.n_runs, mo.n_events
        self.hit_off = _np_from(mo.hit_off, n + 1, np.uint32) if n else np.zeros(1, np.uint32)
        self.nruns = _np_from(mo.nruns, n, np.uint32)
        nh = int(self.hit_off[-1]) if n else 0
        self.runs = _np_from(mo.runs, nh * 3, np.uint32).reshape(-1, 3)
        self.hits = _np_from(mo.hits, nh * 3, np.uint32).reshape(-1, 3)
        self.ev_off = _np_from(mo.ev_off, n + 1, np.uint32) if n else np.zeros(1, np.uint32)
        self.ev_cnt = _np_from(mo.ev_cnt, n, np.uint32)
        ne = int(self.ev_off[-1]) if n else 0
        self.events = _np_from(mo.events, ne * 6, np.uint32).reshape(-1, 6)

    def _struct(self):
        m = _lib.MapOut()
        m.n_reads, m.n_mx, m.n_hits, m.n_runs, m.n_events = self.n_reads, self.n_mx, self.n_hits, self.n_runs, self.n_events
        m.hit_off = C.cast(self.hit_off.ctypes.data, C.POINTER(C.c_uint32))
        m.nruns = C.cast(self.nruns.ctypes.data, C.POINTER(C.c_uint32))
        m.runs = C.cast(self.runs.ctypes.data, C.POINTER(_lib.Run))
        m.hits = C.cast(self.hits.ctypes.data, C.POINTER(_lib.Hit))
        m.ev_off = C.cast(self.ev_off.ctypes.data, C.POINTER(C.c_uint32))
        m.ev_cnt = C.cast(self.ev_cnt.ctypes.data, C.POINTER(C.c_uint32))
        m.events = C.cast(self.events.ctypes.data, C.POINTER(_lib.Event))
        return m

    def verbose_bytes(self, reads, contigs, threads=4, copy=True):
        "verbose_mapping.tsv lines (bin/ntlink_pair.py:382-388); copy=False: see _text_out"
        lib = _lib.load()
        rb, ro = reads.name_blob()
        cb, co = contigs.name_blob()
        out = C.c_void_p()
        st = self._struct()
        n = lib.ntl_format_verbose(C.byref(st), _ptr(rb), _ptr(ro), _ptr(cb), _ptr(co), threads, C.byref(out))
        if n < 0:
            raise NtlError(n, "ntl_format_verbose")
        return _text_out(lib, out, n, copy)

    def paf_bytes(self, reads, read_len, contigs, k, threads=4, copy=True):
        "PAF-like lines (bin/ntlink_paf_output.py:103-135); raises where the reference asserts"
        lib = _lib.load()
        rb, ro = reads.name_blob()
        cb, co = contigs.name_blob()
        rl = np.ascontiguousarray(read_len, dtype=np.uint32)
        cl = np.ascontiguousarray(contigs.lengths, dtype=np.uint32)
        out = C.c_void_p()
        st = self._struct()
        n = lib.ntl_format_paf(C.byref(st), _ptr(rb), _ptr(ro), _ptr(rl), _ptr(cb), _ptr(co), _ptr(cl), k, threads,
                               C.byref(out))
        if n < 0:
            raise NtlError(n, "ntl_format_paf: a PAF assertion of the reference failed")
        return _text_out(lib, out, n, copy)


class Context:
    """One GPU. Not thread-safe."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.ntl_init(device, C.byref(h))
        if rc != 0:
            raise NtlError(rc, "ntl_init failed: no usable CUDA device (ntlink_b200 has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.ntl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise NtlError(rc, f"{what}: {self.lib.ntl_last_error(self.h).decode()}")

    def set_option(self, name, value):
        self._check(self.lib.ntl_set_option(self.h, name.encode(), float(value)), "ntl_set_option")

    # ---- sketch
    def _sketch_from(self, so):
        n, ns = so.n_mx, so.nseq
        return Sketch(_np_from(so.hash, n, np.uint64), _np_from(so.pos_strand, n, np.uint32),
                      _np_from(so.seq_off, ns + 1, np.uint64))

    def sketch(self, batch, k, w):
        so = _lib.SketchOut()
        self._check(self.lib.ntl_sketch(self.h, _ptr(batch.seq), _ptr(batch.offsets), len(batch), k, w, C.byref(so)),
                    "ntl_sketch")
        return self._sketch_from(so)

    # ---- index
    def build_index(self, hash_, contig, pos_strand, contig_len, names):
        hash_ = np.ascontiguousarray(hash_, np.uint64)
        contig = np.ascontiguousarray(contig, np.uint32)
        pos_strand = np.ascontiguousarray(pos_strand, np.uint32)
        cl = np.ascontiguousarray(contig_len, np.uint32)
        rank = name_ranks(names)
        self._check(self.lib.ntl_index_build(self.h, _ptr(hash_), _ptr(contig), _ptr(pos_strand), len(hash_), _ptr(cl),
                                             _ptr(rank), len(names)), "ntl_index_build")

    def build_index_from_sequences(self, batch, k, w, want_sketch=True):
        rank = name_ranks(batch.names)
        so = _lib.SketchOut()
        self._check(self.lib.ntl_index_build_from_sequences(self.h, _ptr(batch.seq), _ptr(batch.offsets), len(batch), k, w,
                                                            _ptr(rank), C.byref(so) if want_sketch else None),
                    "ntl_index_build_from_sequences")
        return self._sketch_from(so) if want_sketch else None

    def index_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.ntl_index_stats(self.h, C.byref(a), C.byref(b), C.byref(c)), "ntl_index_stats")
        return {"inserted": a.value, "unique": b.value, "slots": c.value}

    # ---- mapping
    @staticmethod
    def params(k, w, z=500, f=10, x=0.0, sensitive=False, repeat_filter=False):
        return _lib.Params(k, w, z, f, float(x), int(bool(sensitive)), int(bool(repeat_filter)))

    def map_reads(self, batch, prm, first_ordinal=0):
        mo = _lib.MapOut()
        self._check(self.lib.ntl_map_reads(self.h, _ptr(batch.seq), _ptr(batch.offsets), len(batch), first_ordinal,
                                           C.byref(prm), C.byref(mo)), "ntl_map_reads")
        return MapResult(mo)

    def map_sketch(self, sketch, read_len, prm, first_ordinal=0):
        mo = _lib.MapOut()
        rl = np.ascontiguousarray(read_len, np.uint32)
        h = np.ascontiguousarray(sketch.hash, np.uint64)
        p = np.ascontiguousarray(sketch.pos_strand, np.uint32)
        o = np.ascontiguousarray(sketch.seq_off, np.uint64)
        self._check(self.lib.ntl_map_sketch(self.h, _ptr(h), _ptr(p), _ptr(o), _ptr(rl), len(rl), first_ordinal,
                                            C.byref(prm), C.byref(mo)), "ntl_map_sketch")
        return MapResult(mo)

    def map_groups(self, target_sketch, target_len, group_t_off, read_sketch, read_len, prm):
        """Gap-filling style mapping of many small groups in one call (bin/ntlink_patch_gaps.py:412-442): group g = target
        sequences [group_t_off[g], group_t_off[g+1]) of target_sketch + read g of read_sketch. Returns a MapResult with one
        entry per group (accepted runs + hits, contig ids = target sequence indices)."""
        th = np.ascontiguousarray(target_sketch.hash, np.uint64)
        tp = np.ascontiguousarray(target_sketch.pos_strand, np.uint32)
        to = np.ascontiguousarray(target_sketch.seq_off, np.uint64)
        tl = np.ascontiguousarray(target_len, np.uint32)
        go = np.ascontiguousarray(group_t_off, np.uint32)
        rh = np.ascontiguousarray(read_sketch.hash, np.uint64)
        rp = np.ascontiguousarray(read_sketch.pos_strand, np.uint32)
        ro = np.ascontiguousarray(read_sketch.seq_off, np.uint64)
        rl = np.ascontiguousarray(read_len, np.uint32)
        if len(to) != len(tl) + 1 or len(ro) != len(rl) + 1 or len(go) != len(rl) + 1 or (len(go) and int(go[-1]) != len(tl)):
            raise ValueError("map_groups: inconsistent array sizes")
        mo = _lib.MapOut()
        self._check(self.lib.ntl_map_groups(self.h, _ptr(th), _ptr(tp), _ptr(to), _ptr(tl), len(tl), _ptr(go), _ptr(rh), _ptr(rp), _ptr(ro),
                                            _ptr(rl), len(rl), C.byref(prm), C.byref(mo)), "ntl_map_groups")
        return MapResult(mo)

    def liftover_mappings(self, hit_off, nruns, runs, hits, agp_rows, k, want_result=True):
        """bin/ntlink_liftover_mappings.py:61-124 on the GPU. agp_rows: (ncontig, 5) uint32 {new_id, flags, scaf_start,
        ctg_start, ctg_end} (see liftover.agp_table). Returns a MapResult in the new contig namespace (or None); the
        lifted arrays also stay on the device for tally_lifted()."""
        ho = np.ascontiguousarray(hit_off, np.uint32)
        nr = np.ascontiguousarray(nruns, np.uint32)
        ru = np.ascontiguousarray(runs, np.uint32).reshape(-1, 3)
        hi = np.ascontiguousarray(hits, np.uint32).reshape(-1, 3)
        ag = np.ascontiguousarray(agp_rows, np.uint32).reshape(-1, 5)
        if len(ho) != len(nr) + 1 or len(ru) < int(ho[-1]) or len(hi) < int(ho[-1]):
            raise ValueError("liftover_mappings: inconsistent array sizes")
        mo = _lib.MapOut()
        self._check(self.lib.ntl_liftover_mappings(self.h, _ptr(ho), _ptr(nr), _ptr(ru), _ptr(hi), len(nr), _ptr(ag), len(ag), k,
                                                   C.byref(mo) if want_result else None), "ntl_liftover_mappings")
        return MapResult(mo) if want_result else None

    def tally_lifted(self, nreads, prm, first_ordinal=0):
        "pair events of the mappings the last liftover_mappings() left on the device (checkpoint read length, pair:483-487)"
        ne = C.c_uint64()
        self._check(self.lib.ntl_tally_mappings(self.h, None, None, None, None, None, nreads, first_ordinal, C.byref(prm),
                                                C.byref(ne)), "ntl_tally_mappings")
        return ne.value

    def tally_mappings(self, hit_off, nruns, runs, hits, read_len, prm, first_ordinal=0):
        """Checkpoint path (bin/ntlink_pair.py:437-488): pair events of reads whose accepted runs/hits are known.
        runs: (n,3) uint32 {ctg,start,count}; hits: (n,3) uint32 {ctg,ctg_pos_strand,read_pos_strand}; layout as in MapResult.
        Returns the number of events appended to the event log."""
        ho = np.ascontiguousarray(hit_off, np.uint32)
        nr = np.ascontiguousarray(nruns, np.uint32)
        ru = np.ascontiguousarray(runs, np.uint32).reshape(-1, 3)
        hi = np.ascontiguousarray(hits, np.uint32).reshape(-1, 3)
        rl = None if read_len is None else np.ascontiguousarray(read_len, np.uint32)
        if len(ho) != len(nr) + 1 or (rl is not None and len(rl) != len(nr)) or len(ru) < int(ho[-1]) or len(hi) < int(ho[-1]):
            raise ValueError("tally_mappings: inconsistent array sizes")
        ne = C.c_uint64()
        self._check(self.lib.ntl_tally_mappings(self.h, _ptr(ho), _ptr(nr), _ptr(ru), _ptr(hi), None if rl is None else _ptr(rl),
                                                len(nr), first_ordinal,
                                                C.byref(prm), C.byref(ne)), "ntl_tally_mappings")
        return ne.value

    # ---- pairs
    def events_reset(self):
        self._check(self.lib.ntl_events_reset(self.h), "ntl_events_reset")

    def events_append(self, events):
        ev = np.ascontiguousarray(events, np.uint32).reshape(-1, 6)
        self._check(self.lib.ntl_events_append(self.h, _ptr(ev), len(ev)), "ntl_events_append")

    def events_count(self):
        n = C.c_uint64()
        self._check(self.lib.ntl_events_count(self.h, C.byref(n)), "ntl_events_count")
        return n.value

    def pairs_raw(self):
        "the pair table as arrays: (n_pairs x 10 uint32 view of ntl_pair, gaps int32[]) in first-seen order"
        po = _lib.PairsOut()
        self._check(self.lib.ntl_pairs_finish(self.h, C.byref(po)), "ntl_pairs_finish")
        n = po.n_pairs
        raw = _np_from(po.pairs, n * 10, np.uint32).reshape(-1, 10) if n else np.empty((0, 10), np.uint32)
        return raw, _np_from(po.gaps, po.n_gaps, np.int32)

    def pairs(self):
        """[(src, tgt, flags, n, anchor, gaps int32[n])] in first-seen order (the reference's dict order)."""
        raw, gaps = self.pairs_raw()
        out = []
        for row in raw:
            goff = int(row[6]) | (int(row[7]) << 32)
            out.append((int(row[0]), int(row[1]), int(row[2]), int(row[3]), int(row[4]), gaps[goff:goff + int(row[3])]))
        return out

    # ---- resident / timing (bench)
    def reads_upload(self, batch):
        self._check(self.lib.ntl_reads_upload(self.h, _ptr(batch.seq), _ptr(batch.offsets), len(batch)), "ntl_reads_upload")

    def map_resident(self, prm, first_ordinal=0):
        mo = _lib.MapOut()
        self._check(self.lib.ntl_map_resident(self.h, first_ordinal, C.byref(prm), C.byref(mo)), "ntl_map_resident")
        return {"reads": mo.n_reads, "mx": mo.n_mx, "hits": mo.n_hits, "runs": mo.n_runs, "events": mo.n_events}

    def target_upload(self, batch):
        rank = name_ranks(batch.names)
        self._check(self.lib.ntl_target_upload(self.h, _ptr(batch.seq), _ptr(batch.offsets), len(batch), _ptr(rank)),
                    "ntl_target_upload")

    def index_build_resident(self, k, w):
        self._check(self.lib.ntl_index_build_resident(self.h, k, w), "ntl_index_build_resident")

    # ---- synthetic resident inputs (bench / tests; csrc/synth_logic.cuh)
    def synth_target_resident(self, seed, plan, names):
        "generate the contigs of a synth.plan_assembly plan straight into the resident target"
        plan = np.ascontiguousarray(plan)
        rank = name_ranks(names)
        self._check(self.lib.ntl_synth_target_resident(self.h, seed, plan.ctypes.data, len(plan), _ptr(rank)), "ntl_synth_target_resident")

    def synth_reads_resident(self, seed, plan, err=(2621, 1966, 1966)):
        "generate the reads of a synth.plan_reads plan straight into the resident reads; returns the number of bases"
        plan = np.ascontiguousarray(plan)
        tot = C.c_uint64()
        self._check(self.lib.ntl_synth_reads_resident(self.h, seed, plan.ctypes.data, len(plan), err[0], err[1], err[2], C.byref(tot)),
                    "ntl_synth_reads_resident")
        return tot.value

    def resident_info(self, which):
        "(number of sequences, bases) of the resident target (which=0) or reads (which=1)"
        n, b = C.c_uint32(), C.c_uint64()
        self._check(self.lib.ntl_resident_info(self.h, which, C.byref(n), C.byref(b)), "ntl_resident_info")
        return n.value, b.value

    def resident_download(self, which, first, count, names, pinned=False):
        """sequences [first, first+count) of the resident target / reads as a SeqBatch on the host (pinned=True: the
        arrays live in pinned memory obtained through torch, for end-to-end timing)"""
        off = np.zeros(count + 1, np.uint64)
        self._check(self.lib.ntl_resident_download(self.h, which, first, count, None, _ptr(off)), "ntl_resident_download")
        nb = int(off[-1])
        keep = None
        if pinned:
            import torch
            keep = torch.empty(nb + 64, dtype=torch.uint8, pin_memory=True)
            seq = keep.numpy()
        else:
            seq = np.empty(nb + 64, np.uint8)
        seq[nb:] = ord("N")
        self._check(self.lib.ntl_resident_download(self.h, which, first, count, seq.ctypes.data, _ptr(off)), "ntl_resident_download")
        out = SeqBatch.__new__(SeqBatch)
        out.seq, out.offsets, out.names, out._name_blob, out._keep = seq[:nb], off, list(names), None, keep
        return out

    def stat(self, name):
        "counters since init: async_calls, async_fallbacks, graph_launches"
        v = C.c_double()
        self._check(self.lib.ntl_get_stat(self.h, name.encode(), C.byref(v)), "ntl_get_stat")
        return v.value

    def timing_reset(self):
        self._check(self.lib.ntl_timing_reset(self.h), "ntl_timing_reset")

    def timing(self):
        ms = (C.c_double * 16)()
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.ntl_timing(self.h, ms, C.byref(a), C.byref(b), C.byref(c)), "ntl_timing")
        d = {name: ms[i] for i, name in enumerate(_lib.T_NAMES)}
        d.update(launches=a.value, dense_launches=b.value, dense_bases=c.value)
        ms1, n1, b1 = C.c_double(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.ntl_timing_dense(self.h, C.byref(ms1), C.byref(n1), C.byref(b1)), "ntl_timing_dense")
        d.update(big_dense_ms=ms1.value, big_dense_launches=n1.value, big_dense_bases=b1.value)
        return d

    def mark(self, which):
        self._check(self.lib.ntl_mark(self.h, which), "ntl_mark")

    def mark_elapsed_ms(self):
        ms = C.c_double()
        self._check(self.lib.ntl_mark_elapsed(self.h, C.byref(ms)), "ntl_mark_elapsed")
        return ms.value

    def sync(self):
        self._check(self.lib.ntl_device_sync(self.h), "ntl_device_sync")

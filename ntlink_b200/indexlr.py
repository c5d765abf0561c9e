#!/usr/bin/env python3
"""GPU-backed drop-in for btllib's `indexlr` as ntLink invokes it (ntLink:199,223,244,249):

    indexlr --long [--pos] [--strand] [--len] -k K -w W [-t T] FILE|-

FASTA/FASTQ (plain or gzip) in, the indexlr TSV on stdout, records in input order. The file is streamed in chunks
of about 1 Gbp so that stdin pipes (`gzip -cd reads | indexlr ... -`) of any size work."""
import argparse
import ctypes as C
import sys

import numpy as np

from . import _lib, api


def parse_arguments(argv=None):
    p = argparse.ArgumentParser(description="indexlr on a B200 (minimizer sketches as TSV)")
    p.add_argument("FILE", help="FASTA/FASTQ[.gz] or - for stdin")
    p.add_argument("-k", type=int, required=True)
    p.add_argument("-w", type=int, required=True)
    p.add_argument("-t", type=int, default=4, help="host threads for text formatting")
    p.add_argument("--long", action="store_true", help="accepted for compatibility (reader buffering only)")
    p.add_argument("--pos", action="store_true")
    p.add_argument("--strand", action="store_true")
    p.add_argument("--len", action="store_true")
    p.add_argument("--device", type=int, default=0)
    p.add_argument("--chunk-bases", type=float, default=1e9)
    return p.parse_args(argv)


def main(argv=None):
    a = parse_arguments(argv)
    lib = _lib.load()
    ctx = api.Context(a.device)
    h = C.c_void_p()
    if lib.ntl_seqfile_open(a.FILE.encode(), C.byref(h)) != 0:
        sys.exit(f"indexlr: cannot open {a.FILE}")
    out = sys.stdout.buffer
    try:
        while True:
            seq, off, names, noff = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            n = C.c_uint32()
            if lib.ntl_seqfile_read(h, int(a.chunk_bases), C.byref(seq), C.byref(off), C.byref(names), C.byref(noff), C.byref(n)) != 0:
                sys.exit("indexlr: read error")
            nseq = n.value
            if nseq == 0:
                for p in (seq, off, names, noff):
                    lib.ntl_free(p)
                break
            offsets = api._np_from(off, nseq + 1, np.uint64)
            name_off = api._np_from(noff, nseq + 1, np.uint64)
            s = api._np_from(seq, int(offsets[-1]), np.uint8)
            nb = api._np_from(names, int(name_off[-1]), np.uint8).tobytes()
            nm = [nb[int(name_off[i]):int(name_off[i + 1])].decode() for i in range(nseq)]
            for p in (seq, off, names, noff):
                lib.ntl_free(p)
            batch = api.SeqBatch(s, offsets, nm)
            sk = ctx.sketch(batch, a.k, a.w)
            out.write(sk.to_tsv(batch, with_len=a.len, with_pos=a.pos, with_strand=a.strand, threads=a.t))
    finally:
        lib.ntl_seqfile_close(h)
        ctx.close()


if __name__ == "__main__":
    main()

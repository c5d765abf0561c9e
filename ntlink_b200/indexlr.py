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
    ctx = api.Context(a.device)
    out = sys.stdout.buffer
    try:
        # streamed in chunks; the next chunk is read / decompressed while this one is sketched and printed
        for batch in api.prefetch_batches([a.FILE], int(a.chunk_bases)):
            sk = ctx.sketch(batch, a.k, a.w)
            out.write(sk.to_tsv(batch, with_len=a.len, with_pos=a.pos, with_strand=a.strand, threads=a.t, copy=False))
    except OSError as exc:
        sys.exit(f"indexlr: {exc}")
    finally:
        ctx.close()


if __name__ == "__main__":
    main()

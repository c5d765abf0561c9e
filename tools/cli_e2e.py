#!/usr/bin/env python3
"""T_e2e-file: wall time of the drop-in command line on FILES, the way ntLink's make recipe calls it --
target FASTA + read FASTA (optionally .gz) in, <p>.n1.scaffold.dot / .pairs.tsv (/ .verbose_mapping.tsv / .paf) out.

    python tools/cli_e2e.py [--config c1|c2] [--scale 0.25] [--gz | --bgzf] [--gpus N]

--gz: reads as single-stream gzip (`gzip -1`), inflated by one zlib thread; --bgzf: reads as bgzip-style members (what
`bgzip` writes), inflated in parallel by the reader.

Inputs are the bench workloads (generated on the device, written as FASTA to tmpfs before the clock starts)."""
import argparse
import contextlib
import io
import json
import os
import shutil
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))


def write_bgzf(src, dst, level=1, block=65280):
    "bgzip-style file: independent gzip members of <= 64 KiB with the BC extra field (SAM spec 4.1)"
    import struct
    import zlib
    with open(src, "rb") as fin, open(dst, "wb") as fout:
        while True:
            chunk = fin.read(block)
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
            payload = co.compress(chunk) + co.flush()
            fout.write(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 18 + len(payload) + 8 - 1) +
                       payload + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
            if not chunk:
                break


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--gz", action="store_true")
    ap.add_argument("--bgzf", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    import bench
    import cpu_pipeline as cp
    from ntlink_b200 import Context, pair, synth
    args = argparse.Namespace(config=a.config, scale=a.scale)
    cfg = bench.pick_config(args, 1)
    cplan, cnames, rplan = bench.plans(cfg, 0)
    ctx = Context(0)
    ctx.synth_target_resident(bench.SEED, cplan, cnames)
    ctx.synth_reads_resident(bench.SEED, rplan)
    contigs = ctx.resident_download(0, 0, len(cplan), cnames)
    reads = ctx.resident_download(1, 0, len(rplan), synth.read_names(rplan))
    ctx.close()
    d = bench.scratch_dir()
    try:
        tgt, rd = os.path.join(d, "target.fa"), os.path.join(d, "reads.fa")
        cp.write_fasta(tgt, contigs)
        cp.write_fasta(rd, reads)
        if a.gz:
            subprocess.check_call(["gzip", "-1", rd])
            rd += ".gz"
        elif a.bgzf:
            write_bgzf(rd, rd + ".bgz.gz")
            os.remove(rd)
            rd += ".bgz.gz"
        bases = int(reads.offsets[-1])
        for mode, extra in (("scaffold.dot + pairs.tsv", []), ("+ verbose_mapping.tsv + paf", ["--verbose", "--paf"])):
            times = []
            for it in range(3):
                prefix = os.path.join(d, f"out{it}")
                argv = ["-p", prefix, "-n", "1", "-s", tgt, "-k", str(cfg["k"]), "-w", str(cfg["w"]), "-a", "1", "-z", "1000", "-f", "10", "-x", "0",
                        "--pairs", "--sketch-target", "--reads-fasta", rd, "-t", "8"] + extra + (["--sensitive"] if cfg["sensitive"] else [])
                if a.gpus > 1:
                    argv += ["--gpus", str(a.gpus)]
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    pair.main(argv)
                times.append(time.perf_counter() - t0)
                for suffix in (".verbose_mapping.tsv", ".paf"):
                    if os.path.exists(prefix + suffix):
                        os.remove(prefix + suffix)
            best = min(times)
            print(json.dumps({"T_e2e_file": "python -m ntlink_b200.pair --sketch-target --reads-fasta reads.fa" + (".gz" if a.gz else ".bgz.gz (BGZF)" if a.bgzf else ""),
                              "workload": cfg["workload"], "outputs": mode, "gpus": a.gpus, "read_bases": bases, "target_bases": int(contigs.offsets[-1]),
                              "wall_s": [round(t, 3) for t in times], "gbp_per_s_best": round(bases / best / 1e9, 3)}), flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()

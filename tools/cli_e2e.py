#!/usr/bin/env python3
"""Wall time of the drop-in command line on files (the way ntLink's make recipe would call it): FASTA in, the four output
files out. Shows where a real run spends its time now that the kernels take milliseconds.

    python tools/cli_e2e.py [--gz]"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def write_fasta(path, batch):
    seq, off = batch.seq.tobytes(), batch.offsets
    with open(path, "wb") as f:
        for i, n in enumerate(batch.names):
            f.write(b">" + n.encode() + b"\n" + seq[int(off[i]):int(off[i + 1])] + b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gz", action="store_true")
    a = ap.parse_args()
    import bench
    from ntlink_b200 import pair
    contigs, reads = bench.make_inputs(0, 1)
    d = tempfile.mkdtemp(prefix="ntl_cli_")
    tgt, rd = os.path.join(d, "target.fa"), os.path.join(d, "reads.fa")
    write_fasta(tgt, contigs)
    write_fasta(rd, reads)
    if a.gz:
        subprocess.check_call(["gzip", "-1", rd])
        rd += ".gz"
    bases = int(reads.offsets[-1])
    for mode, extra in (("pairs + dot only", []), ("+ verbose_mapping.tsv + paf", ["--verbose", "--paf"])):
        times = []
        for it in range(3):
            prefix = os.path.join(d, f"out{it}")
            argv = ["-p", prefix, "-n", "1", "-s", tgt, "-k", "32", "-w", "100", "-a", "1", "-z", "1000", "-f", "10", "-x", "0", "--pairs",
                    "--sketch-target", "--reads-fasta", rd, "-t", "8"] + extra
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                pair.main(argv)
            times.append(time.perf_counter() - t0)
            for suffix in (".verbose_mapping.tsv", ".paf"):
                if os.path.exists(prefix + suffix):
                    os.remove(prefix + suffix)
        best = min(times)
        print(json.dumps({"cli": "python -m ntlink_b200.pair --sketch-target --reads-fasta reads.fa" + (".gz" if a.gz else ""), "outputs": mode,
                          "read_bases": bases, "wall_s": [round(t, 3) for t in times], "gbp_per_s_best": round(bases / best / 1e9, 3)}), flush=True)


if __name__ == "__main__":
    main()

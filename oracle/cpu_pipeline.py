"""The reference's CPU pipeline for the hot path, as its make recipe runs it (ntLink:198-199, 221-225) -- TEST / BENCH
INFRASTRUCTURE ONLY (checker and CPU baseline; never imported by the product):

    indexlr --long --pos --strand -k K -w W -t T target.fa > target.fa.kK.wW.tsv
    indexlr --long --pos --strand --len -k K -w W -t T reads.fa | ntlink_pair.py -p P -n 1 -m target.tsv -s target.fa
            -k K -a 1 -z Z -f 10 -x 0 [--verbose --pairs --paf --sensitive] -

`indexlr` is oracle/_build/indexlr_oracle (C restatement of btllib's indexlr, multi-threaded, pinned on the reference's
golden sketches; btllib itself is not in the reference tree nor installed). `ntlink_pair.py` is the UNMODIFIED
reference file staged in oracle/_ref/ by `make -C oracle ref` (with this repository's igraph stand-in) when it is there
-- kind "reference" -- and otherwise the port oracle/pair_oracle.py -- kind "port". Both are single-threaded Python,
like the reference (bin/ntlink_pair.py has no -t)."""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
INDEXLR = os.path.join(HERE, "_build", "indexlr_oracle")
REF_PAIR = os.path.join(HERE, "_ref", "ntlink_pair.py")
PORT_PAIR = os.path.join(HERE, "pair_oracle.py")


def mapper_kind():
    return "reference" if os.path.exists(REF_PAIR) else "port"


def write_fasta(path, batch, first=0, count=None):
    "one line per sequence, like the simulated read files of SURVEY.md 8d"
    count = len(batch) - first if count is None else count
    off = batch.offsets
    seq = batch.seq
    with open(path, "wb") as fout:
        for i in range(first, first + count):
            fout.write(b">" + batch.names[i].encode() + b"\n")
            fout.write(memoryview(seq[int(off[i]):int(off[i + 1])]))
            fout.write(b"\n")


def sketch_target(target_fa, k, w, threads):
    "ntLink:198-199; returns (tsv path, seconds)"
    if not os.path.exists(INDEXLR):
        subprocess.check_call(["make", "-s", "-C", HERE])
    out = f"{target_fa}.k{k}.w{w}.tsv"
    t0 = time.perf_counter()
    with open(out, "wb") as fout:
        subprocess.check_call([INDEXLR, "--long", "--pos", "--strand", "-k", str(k), "-w", str(w), "-t", str(threads), target_fa], stdout=fout)
    return out, time.perf_counter() - t0


def map_reads(target_fa, target_tsv, reads_fa, prefix, k, w, z, threads, sensitive=False, repeat_filter=False, f=10, x=0, a=1, n=1,
              verbose=True, pairs=False, paf=False, kind=None):
    """ntLink:221-225 as two piped processes; returns seconds. Removes a stale <prefix>.verbose_mapping.tsv first (it would
    be taken for a checkpoint, bin/ntlink_pair.py:565-575)."""
    kind = kind or mapper_kind()
    for suffix in (".verbose_mapping.tsv", ".paf", ".pairs.tsv", f".n{n}.scaffold.dot"):
        if os.path.exists(prefix + suffix):
            os.remove(prefix + suffix)
    script = REF_PAIR if kind == "reference" else PORT_PAIR
    cmd = [sys.executable, script, "-p", prefix, "-n", str(n), "-m", target_tsv, "-s", target_fa, "-k", str(k), "-a", str(a), "-z", str(z),
           "-f", str(f), "-x", str(x)]
    cmd += (["--verbose"] if verbose else []) + (["--pairs"] if pairs else []) + (["--paf"] if paf else [])
    cmd += (["--sensitive"] if sensitive else []) + (["--repeat-filter"] if repeat_filter else []) + ["-"]
    env = dict(os.environ, PYTHONHASHSEED="0")
    if kind == "reference":
        env["PYTHONPATH"] = os.path.dirname(REF_PAIR) + os.pathsep + env.get("PYTHONPATH", "")
    t0 = time.perf_counter()
    p1 = subprocess.Popen([INDEXLR, "--long", "--pos", "--strand", "--len", "-k", str(k), "-w", str(w), "-t", str(threads), reads_fa],
                          stdout=subprocess.PIPE)
    p2 = subprocess.Popen(cmd, stdin=p1.stdout, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
    p1.stdout.close()
    err = p2.communicate()[1]
    rc1 = p1.wait()
    dt = time.perf_counter() - t0
    if p2.returncode != 0 or rc1 != 0:
        raise RuntimeError(f"CPU pipeline failed (indexlr rc {rc1}, mapper rc {p2.returncode}): {err.decode()[-2000:]}")
    return dt


def next_round(verbose_path, agp_path, scaffolds_fa, prefix, k, z, f=10, x=0, a=1, n=1, kind=None):
    """One later round of ntLink_rounds on the CPU (ntLink_rounds:122-145): the reference's liftover script writes
    <prefix>.verbose_mapping.tsv, which its pair stage then takes as a checkpoint instead of mapping again
    (bin/ntlink_pair.py:565-575). Returns (liftover seconds, pair seconds)."""
    kind = kind or mapper_kind()
    here = os.path.dirname(REF_PAIR)
    lift = os.path.join(here, "ntlink_liftover_mappings.py")
    env = dict(os.environ, PYTHONHASHSEED="0", PYTHONPATH=here + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = prefix + ".verbose_mapping.tsv"
    for suffix in (".pairs.tsv", f".n{n}.scaffold.dot"):
        if os.path.exists(prefix + suffix):
            os.remove(prefix + suffix)
    t0 = time.perf_counter()
    if kind == "reference" and os.path.exists(lift):
        subprocess.check_call([sys.executable, lift, "-m", verbose_path, "-a", agp_path, "-o", out, "-k", str(k)], env=env,
                              stdout=subprocess.DEVNULL)
        t1 = time.perf_counter()
        cmd = [sys.executable, REF_PAIR, "-p", prefix, "-n", str(n), "-m", "unused.tsv", "-s", scaffolds_fa, "-k", str(k), "-a", str(a),
               "-z", str(z), "-f", str(f), "-x", str(x), "--pairs", "unused_reads.tsv"]
        r = subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if r.returncode != 0:
            raise RuntimeError("CPU checkpoint round failed: " + r.stderr.decode()[-2000:])
        return t1 - t0, time.perf_counter() - t1
    # the port, in this process (pair_oracle.py has no command line for the checkpoint path)
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import liftover_oracle
    import pair_oracle as po
    with open(verbose_path) as fin, open(agp_path) as fagp:
        lifted = liftover_oracle.liftover(fin, fagp.readlines(), k)
    with open(out, "w") as fout:
        fout.writelines(lifted)
    t1 = time.perf_counter()
    lengths = po.read_fasta_lengths(scaffolds_fa)
    prm = po.default_params(k, z=z, a=a, f=f, x=x, n=n)
    pairs = po.filter_pairs(po.retally_from_verbose(lifted, lengths, prm), lengths, a)
    with open(prefix + ".pairs.tsv", "w") as fout:
        fout.writelines(po.pairs_tsv_lines(pairs))
    return t1 - t0, time.perf_counter() - t1


def outputs(prefix, n=1):
    "bytes of the files the mapper wrote (missing file -> None)"
    out = {}
    for key, suffix in (("verbose", ".verbose_mapping.tsv"), ("paf", ".paf"), ("pairs", ".pairs.tsv"), ("dot", f".n{n}.scaffold.dot")):
        p = prefix + suffix
        out[key] = open(p, "rb").read() if os.path.exists(p) else None
    return out

#!/usr/bin/env python3
"""
Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
  * copies the reference's test inputs (data, not code) to tests/golden/inputs/ (gzip-compressed) and the
    reference's own expected outputs for the hot path to tests/golden/expected_outputs/;
  * sketches target + reads with the C oracle (oracle/_build/indexlr_oracle; byte-identical to the
    reference's golden target sketches, see tests/test_oracle_sketch.py) because btllib's indexlr is not
    installed here;
  * runs /root/reference/bin/ntlink_pair.py (imported unmodified; python-igraph is absent, so a ~60 line
    stand-in providing only the calls that file makes is put on PYTHONPATH) for every case in CASES and
    stores verbose_mapping.tsv / .paf / .pairs.tsv / .scaffold.dot (gzip) under tests/golden/cases/<case>/.

The committed outputs pin oracle/pair_oracle.py (CPU tests) and the CUDA path (GPU tests).
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
ORACLE = os.path.join(REPO, "oracle", "_build", "indexlr_oracle")

IGRAPH_STANDIN = r'''
"""Minimal stand-in for python-igraph: only what bin/ntlink_pair.py + bin/ntlink_utils.py touch."""
class InternalError(Exception):
    pass
class _V:
    def __init__(self, g, i): self.g, self.index = g, i
    def __getitem__(self, key):
        assert key == "name"
        return self.g._names[self.index]
class _VS(list):
    def find(self, name):
        for v in self:
            if v["name"] == name:
                return v
        raise ValueError(name)
class _E:
    def __init__(self, g, i): self.g, self.index = g, i
    @property
    def source(self): return self.g._edges[self.index][0]
    @property
    def target(self): return self.g._edges[self.index][1]
    def __getitem__(self, key): return self.g._attrs[key][self.index]
class _ES(list):
    def __init__(self, g, items): super().__init__(items); self.g = g
    def __setitem__(self, key, values):
        if isinstance(key, str):
            self.g._attrs[key] = list(values)
        else:
            super().__setitem__(key, values)
    def __getitem__(self, key):
        if isinstance(key, str):
            return self.g._attrs[key]
        return super().__getitem__(key)
class Graph:
    def __init__(self, directed=True):
        self._names, self._idx, self._edges, self._attrs = [], {}, [], {}
    def add_vertices(self, names):
        for n in names:
            self._idx[n] = len(self._names); self._names.append(n)
    def add_edges(self, pairs):
        for s, t in pairs:
            self._edges.append((self._idx[s], self._idx[t]))
    def get_eid(self, s, t):
        key = (self._idx[s], self._idx[t])
        for i, e in enumerate(self._edges):
            if e == key:
                return i
        raise InternalError("no such edge")
    def vs(self): return _VS(_V(self, i) for i in range(len(self._names)))
    def es(self): return _ES(self, [_E(self, i) for i in range(len(self._edges))])
    def copy(self):
        g = Graph(); g._names = list(self._names); g._idx = dict(self._idx)
        g._edges = list(self._edges); g._attrs = {k: list(v) for k, v in self._attrs.items()}
        return g
    def delete_edges(self, idxs):
        drop = set(idxs)
        keep = [i for i in range(len(self._edges)) if i not in drop]
        self._edges = [self._edges[i] for i in keep]
        self._attrs = {k: [v[i] for i in keep] for k, v in self._attrs.items()}
'''

# fixture -> (target fasta, reads file, k, w)
FIXTURES = {
    "f1": ("scaffolds_1.fa", "long_reads_1.fa", 32, 250),
    "f1w100": ("scaffolds_1.fa", "long_reads_1.fa", 32, 100),
    "f2": ("scaffolds_2.fa", "long_reads_2.fq.gz", 32, 100),
    "f3": ("scaffolds_3.fa", "long_reads_3.fa.gz", 24, 250),
    "f4": ("scaffolds_4.fa", "long_reads_4.fa.gz", 40, 100),
    "f4top5": ("scaffolds_4.fa", "long_reads_4_top5.fa", 40, 100),
    "f3k20w10": ("scaffolds_3.fa", "long_reads_3.fa.gz", 20, 10),
}

# case -> (fixture, extra ntlink_pair.py options). z=1000 a=1 f=10 x=0 are the ntLink make defaults.
BASE = dict(z=1000, a=1, f=10, x=0, n=1)
CASES = {
    "f1_default": ("f1", {}),
    "f1w100_default": ("f1w100", {}),
    "f2_default": ("f2", {}),
    "f3_default": ("f3", {}),
    "f4_default": ("f4", {}),
    "f4top5_default": ("f4top5", {}),
    "f3_sensitive": ("f3", {"sensitive": True}),
    "f3_repeat": ("f3", {"repeat_filter": True}),
    "f3_sensitive_repeat": ("f3", {"sensitive": True, "repeat_filter": True}),
    "f3_x1.5": ("f3", {"x": 1.5}),
    "f3_x0.3": ("f3", {"x": 0.3}),
    "f3_f1": ("f3", {"f": 1}),
    "f3_f2_sensitive": ("f3", {"f": 2, "sensitive": True}),
    "f3_a2_n2": ("f3", {"a": 2, "n": 2}),
    "f3_z500": ("f3", {"z": 500}),
    "f3_z40000": ("f3", {"z": 40000}),
    "f2_sensitive_x1.1": ("f2", {"sensitive": True, "x": 1.1}),
    "f2_f1_a3": ("f2", {"f": 1, "a": 3}),
    "f4_sensitive": ("f4", {"sensitive": True}),
    "f1w100_repeat_f1": ("f1w100", {"repeat_filter": True, "f": 1}),
    "f3k20w10_default": ("f3k20w10", {}),
    "f3k20w10_sensitive_f2": ("f3k20w10", {"sensitive": True, "f": 2}),
}


CHECKPOINT_CASES = {"f1_default", "f2_default", "f3_default", "f4_default", "f3_sensitive", "f3_f1", "f2_f1_a3", "f3_a2_n2",
                    "f3k20w10_default"}


# liftover (bin/ntlink_liftover_mappings.py, ntLink_rounds:122-125) + the round-2 checkpoint tally on its output.
#   ref_*: the reference's own expected round-1 outputs (verbose_mapping.tsv + trimmed_scafs.agp)
#   syn_*: the golden verbose mapping of a CASES entry + a seeded synthetic AGP (paths of 1-4 contigs in read order,
#          both orientations, trimmed ends, identity entries, contigs missing from the AGP)
LIFTOVER_REF = {"lift_ref_f1": ("scaffolds_1.fa.k32.w250.z1000", 32), "lift_ref_f2": ("scaffolds_2.fa.k32.w100.z1000", 32),
                "lift_ref_f3": ("scaffolds_3.fa.k24.w250.z1000", 24), "lift_ref_f4": ("scaffolds_4.fa.k40.w100.z1000", 40)}
LIFTOVER_SYN = {"lift_syn_f3": ("f3_default", 11), "lift_syn_f3_sensitive": ("f3_sensitive", 12), "lift_syn_f2": ("f2_default", 13),
                "lift_syn_f3k20w10": ("f3k20w10_default", 14), "lift_syn_f1": ("f1_default", 15), "lift_syn_f3b": ("f3_default", 16)}


def fasta_lengths(path):
    out, name = {}, None
    with (gzip.open(path, "rt") if path.endswith(".gz") else open(path)) as fin:
        for line in fin:
            if line.startswith(">"):
                name = line[1:].split()[0]
                out[name] = 0
            elif name is not None:
                out[name] += len(line.strip())
    return out


def synthetic_agp(verbose_text, lengths, k, seed):
    """AGP lines + {new name: length}. Contigs are chained into paths in the order they first appear in the mapping file,
    so that neighbours in a read often land in one path (merged runs, subsumed runs, non-monotonic runs all occur)."""
    import random
    rng = random.Random(seed)
    order = []
    for line in verbose_text.splitlines():
        ctg = line.split("\t")[1]
        if ctg not in order:
            order.append(ctg)
    lines, new_len, path_no, i = [], {}, 0, 0
    while i < len(order):
        roll = rng.random()
        if roll < 0.15:                       # not in the AGP at all
            new_len[order[i]] = lengths[order[i]]
            i += 1
            continue
        if roll < 0.30:                       # identity entry: path id == contig id, coordinates untouched
            ctg = order[i]
            lines.append(f"{ctg}\t1\t{lengths[ctg]}\t1\tW\t{ctg}\t1\t{lengths[ctg]}\t{rng.choice('+-')}")
            new_len[ctg] = lengths[ctg]
            i += 1
            continue
        members = order[i:i + rng.randint(1, 4)]
        i += len(members)
        path, pos, comp = f"ntLink_{path_no}", 1, 1
        path_no += 1
        for j, ctg in enumerate(members):
            if j:
                gap = rng.choice([1, 20, 100, 431])
                lines.append(f"{path}\t{pos}\t{pos + gap - 1}\t{comp}\tN\t{gap}\tscaffold\tyes\tpaired-ends")
                pos += gap
                comp += 1
            start = 1 + (rng.randint(0, min(300, lengths[ctg] // 4)) if rng.random() < 0.5 else 0)
            end = lengths[ctg] - (rng.randint(0, min(300, lengths[ctg] // 4)) if rng.random() < 0.5 else 0)
            lines.append(f"{path}\t{pos}\t{pos + end - start}\t{comp}\tW\t{ctg}\t{start}\t{end}\t{rng.choice('+-')}")
            pos += end - start + 1
            comp += 1
        new_len[path] = pos - 1
    return "\n".join(lines) + "\n", new_len


def make_liftover_goldens(work, env):
    out_manifest = {}
    jobs = []
    for case, (prefix, k) in LIFTOVER_REF.items():
        exp = os.path.join(REF, "tests", "expected_outputs")
        verbose = open(os.path.join(exp, prefix + ".verbose_mapping.tsv")).read()
        agp = open(os.path.join(exp, prefix + ".trimmed_scafs.agp")).read()
        new_len = fasta_lengths(os.path.join(exp, prefix + ".ntLink.scaffolds.fa"))
        jobs.append((case, verbose, agp, new_len, k, dict(BASE)))
    cases_manifest = json.load(open(os.path.join(HERE, "manifest.json")))
    for case, (src, seed) in LIFTOVER_SYN.items():
        info = cases_manifest[src]
        verbose = gzip.open(os.path.join(HERE, "cases", src, "verbose_mapping.tsv.gz"), "rt").read()
        lengths = fasta_lengths(os.path.join(REF, "tests", info["target"]))
        agp, new_len = synthetic_agp(verbose, lengths, info["k"], seed)
        jobs.append((case, verbose, agp, new_len, info["k"], {key: info[key] for key in BASE}))
    for case, verbose, agp, new_len, k, opt in jobs:
        d = os.path.join(work, case)
        os.makedirs(d)
        open(os.path.join(d, "in.verbose_mapping.tsv"), "w").write(verbose)
        open(os.path.join(d, "in.agp"), "w").write(agp)
        lifted = os.path.join(d, "round2.verbose_mapping.tsv")
        subprocess.check_call([sys.executable, os.path.join(REF, "bin", "ntlink_liftover_mappings.py"), "-m",
                               os.path.join(d, "in.verbose_mapping.tsv"), "-a", os.path.join(d, "in.agp"), "-o", lifted,
                               "-k", str(k)], env=env)
        # round 2 (ntLink_rounds:137-145): the lifted file is the checkpoint of ntlink_pair.py on the scaffolds
        fa = os.path.join(d, "round2.fa")
        with open(fa, "w") as fout:
            for name, length in new_len.items():
                fout.write(f">{name}\n{'A' * length}\n")
        cmd = [sys.executable, os.path.join(REF, "bin", "ntlink_pair.py"), "-p", os.path.join(d, "round2"), "-n", str(opt["n"]),
               "-m", "unused.tsv", "-s", fa, "-k", str(k), "-a", str(opt["a"]), "-z", str(opt["z"]), "-f", str(opt["f"]),
               "-x", str(opt["x"]), "--pairs", "unused_reads.tsv"]
        ok = subprocess.call(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0
        out_dir = os.path.join(HERE, "liftover", case)
        os.makedirs(out_dir, exist_ok=True)
        gz_write(os.path.join(out_dir, "in.verbose_mapping.tsv.gz"), verbose.encode())
        gz_write(os.path.join(out_dir, "in.agp.gz"), agp.encode())
        gz_write(os.path.join(out_dir, "lifted.verbose_mapping.tsv.gz"), open(lifted, "rb").read())
        gz_write(os.path.join(out_dir, "round2.lengths.tsv.gz"), "".join(f"{n}\t{l}\n" for n, l in new_len.items()).encode())
        if ok:
            gz_write(os.path.join(out_dir, "round2.pairs.tsv.gz"), open(os.path.join(d, "round2.pairs.tsv"), "rb").read())
            gz_write(os.path.join(out_dir, "round2.scaffold.dot.gz"),
                     open(os.path.join(d, f"round2.n{opt['n']}.scaffold.dot"), "rb").read())
        out_manifest[case] = {"k": k, "round2_ok": ok, **opt}
        print("golden", case, "ok" if ok else "(round 2: the reference raises)")
    with open(os.path.join(HERE, "liftover", "manifest.json"), "w") as fout:
        json.dump(out_manifest, fout, indent=1, sort_keys=True)


def gz_copy(src, dst):
    data = gzip.open(src, "rb").read() if src.endswith(".gz") else open(src, "rb").read()
    with gzip.GzipFile(dst, "wb", mtime=0) as fout:
        fout.write(data)


def gz_write(path, data):
    with gzip.GzipFile(path, "wb", mtime=0) as fout:
        fout.write(data)


def main():
    if not os.path.isdir(REF):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")])
    inputs = os.path.join(HERE, "inputs")
    os.makedirs(inputs, exist_ok=True)
    for name in ["scaffolds_1.fa", "scaffolds_2.fa", "scaffolds_3.fa", "scaffolds_4.fa", "long_reads_1.fa",
                 "long_reads_2.fq.gz", "long_reads_3.fa.gz", "long_reads_4.fa.gz", "long_reads_4_top5.fa"]:
        base = name[:-3] if name.endswith(".gz") else name
        gz_copy(os.path.join(REF, "tests", name), os.path.join(inputs, base + ".gz"))
    exp = os.path.join(HERE, "expected_outputs")
    os.makedirs(exp, exist_ok=True)
    for name in sorted(os.listdir(os.path.join(REF, "tests", "expected_outputs"))):
        if name.endswith((".tsv", ".dot")) and "trimmed" not in name and "abyssfac" not in name:
            gz_copy(os.path.join(REF, "tests", "expected_outputs", name), os.path.join(exp, name + ".gz"))

    work = tempfile.mkdtemp(prefix="ntl_golden_")
    with open(os.path.join(work, "igraph.py"), "w") as fout:
        fout.write(IGRAPH_STANDIN)
    env = dict(os.environ, PYTHONPATH=work + os.pathsep + os.path.join(REF, "bin"), PYTHONHASHSEED="0")
    sketches = {}
    for fx, (tgt, reads, k, w) in FIXTURES.items():
        tgt_fa = os.path.join(work, tgt)
        if not os.path.exists(tgt_fa):
            shutil.copy(os.path.join(REF, "tests", tgt), tgt_fa)
        t_tsv = os.path.join(work, f"{fx}.target.tsv")
        r_tsv = os.path.join(work, f"{fx}.reads.tsv")
        with open(t_tsv, "wb") as fout:
            subprocess.check_call([ORACLE, "--long", "--pos", "--strand", "-k", str(k), "-w", str(w), "-t", "8",
                                   tgt_fa], stdout=fout)
        with open(r_tsv, "wb") as fout:
            subprocess.check_call([ORACLE, "--long", "--pos", "--strand", "--len", "-k", str(k), "-w", str(w),
                                   "-t", "8", os.path.join(REF, "tests", reads)], stdout=fout)
        sketches[fx] = (tgt_fa, t_tsv, r_tsv, k, w)

    manifest = {}
    for case, (fx, extra) in CASES.items():
        tgt_fa, t_tsv, r_tsv, k, w = sketches[fx]
        opt = dict(BASE)
        opt.update(extra)
        prefix = os.path.join(work, case)
        cmd = [sys.executable, os.path.join(REF, "bin", "ntlink_pair.py"), "-p", prefix, "-n", str(opt["n"]),
               "-m", t_tsv, "-s", tgt_fa, "-k", str(k), "-a", str(opt["a"]), "-z", str(opt["z"]),
               "-f", str(opt["f"]), "-x", str(opt["x"]), "--verbose", "--pairs", "--paf"]
        if opt.get("sensitive"):
            cmd.append("--sensitive")
        if opt.get("repeat_filter"):
            cmd.append("--repeat-filter")
        cmd.append(r_tsv)
        subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL)
        out_dir = os.path.join(HERE, "cases", case)
        os.makedirs(out_dir, exist_ok=True)
        for suffix, dst in [(".verbose_mapping.tsv", "verbose_mapping.tsv"), (".paf", "paf"),
                            (".pairs.tsv", "pairs.tsv"), (f".n{opt['n']}.scaffold.dot", "scaffold.dot")]:
            gz_write(os.path.join(out_dir, dst + ".gz"), open(prefix + suffix, "rb").read())
        # checkpoint path (bin/ntlink_pair.py:565-575,437-488): a second run finds <prefix>.verbose_mapping.tsv and
        # re-tallies the pairs from it without reading the sketches
        if case in CHECKPOINT_CASES:
            for suffix in (".pairs.tsv", f".n{opt['n']}.scaffold.dot"):
                os.remove(prefix + suffix)
            subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL)
            for suffix, dst in [(".pairs.tsv", "checkpoint.pairs.tsv"), (f".n{opt['n']}.scaffold.dot", "checkpoint.scaffold.dot")]:
                gz_write(os.path.join(out_dir, dst + ".gz"), open(prefix + suffix, "rb").read())
        manifest[case] = {"fixture": fx, "target": FIXTURES[fx][0], "reads": FIXTURES[fx][1].replace(".gz", ""),
                          "k": k, "w": w, "checkpoint": case in CHECKPOINT_CASES, **opt}
        print("golden", case, "ok")
    with open(os.path.join(HERE, "manifest.json"), "w") as fout:
        json.dump(manifest, fout, indent=1, sort_keys=True)
    make_liftover_goldens(work, env)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()

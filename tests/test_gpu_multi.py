"""Multi-GPU parity on hardware: the drop-in pairing CLI on 2 GPUs (one process per GPU through torchrun: read batches
dealt to the ranks, target sketched in contig shards + NCCL all-gather of the minimizer triples, replicated index, NCCL
gather of the pair events, tally on rank 0) must write the same bytes as on 1 GPU and as the CPU pipeline
("multi-GPU = same test at world size 1/2/4/8 asserting identical bytes"). Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
sys.path.insert(0, util.ORACLE_DIR)


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _inputs(tmp, genome=4_000_000, cov=6, n_frac=0.2):
    import cpu_pipeline as cp
    from ntlink_b200 import synth
    cplan, names = synth.plan_assembly(genome, 31, n_frac=n_frac)
    rplan = synth.plan_reads(genome, cov * genome, 32)
    contigs, reads = synth.host_contigs(3, cplan, names), synth.host_reads(3, rplan)
    tf, rf = os.path.join(tmp, "target.fa"), os.path.join(tmp, "reads.fa")
    cp.write_fasta(tf, contigs)
    cp.write_fasta(rf, reads)
    return cp, tf, rf


def _run_cli(argv):
    env = dict(os.environ, PYTHONPATH=util.REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "ntlink_b200.pair"] + argv, env=env, cwd=util.REPO, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def _files(prefix):
    return {s: open(prefix + s, "rb").read() for s in (".verbose_mapping.tsv", ".paf", ".pairs.tsv", ".n1.scaffold.dot")}


@pytest.mark.parametrize("k,w,sens", [(24, 250, True), (32, 100, False)])
def test_pair_cli_n_gpus_equals_one_gpu_and_cpu(tmp_path, k, w, sens):
    n = min(_ngpu(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    tmp = str(tmp_path)
    cp, tf, rf = _inputs(tmp)
    tsv, _ = cp.sketch_target(tf, k, w, 4)
    cp.map_reads(tf, tsv, rf, os.path.join(tmp, "cpu"), k, w, 1000, 4, sensitive=sens, verbose=True, pairs=True, paf=True)
    want = cp.outputs(os.path.join(tmp, "cpu"))
    common = ["-n", "1", "-s", tf, "-k", str(k), "-w", str(w), "-a", "1", "-z", "1000", "-f", "10", "-x", "0", "--verbose", "--pairs", "--paf"]
    common += ["--sensitive"] if sens else []
    fused = common + ["--sketch-target", "--reads-fasta", rf, "--batch-bases", "3e6"]           # ~8 read batches
    _run_cli(["-p", os.path.join(tmp, "g1")] + fused)
    _run_cli(["-p", os.path.join(tmp, "gn"), "--gpus", str(n)] + fused)
    one, many = _files(os.path.join(tmp, "g1")), _files(os.path.join(tmp, "gn"))
    assert one == many
    assert one[".verbose_mapping.tsv"] == want["verbose"] and one[".paf"] == want["paf"] and one[".pairs.tsv"] == want["pairs"]
    assert util.dot_parts(one[".n1.scaffold.dot"]) == util.dot_parts(want["dot"])
    assert not [f for f in os.listdir(tmp) if ".part" in f]
    # the reference's text interface on N GPUs: target TSV by -m, read TSV on FILES, streamed in small batches
    rtsv = os.path.join(tmp, "reads.tsv")
    with open(rtsv, "wb") as fout:
        fout.write(util.oracle_indexlr(rf, k, w, length=True))
    _run_cli(["-p", os.path.join(tmp, "tn"), "--gpus", str(n), "-m", tsv, "--batch-minimizers", "20000"] + common + [rtsv])
    assert _files(os.path.join(tmp, "tn")) == one


def test_sharded_resident_index_and_event_gather(tmp_path):
    "bench.py's multi-GPU step (sharded target sketch of the RESIDENT target + NCCL all-gather + event gather) vs one GPU"
    n = min(_ngpu(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    code = r'''
import os, sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from ntlink_b200 import Context, synth
from ntlink_b200 import dist as nd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = 6_000_000
cplan, names = synth.plan_assembly(G, 41, n_frac=0.1)
rplan = synth.plan_reads(G, 10 * G, 42)
lo, hi = nd.read_shard(len(rplan), rank, world)
def digest(ctx):
    raw, gaps = ctx.pairs_raw()
    off = raw[:, 6].astype(np.int64) | (raw[:, 7].astype(np.int64) << 32); n = raw[:, 3].astype(np.int64)
    idx = np.repeat(off - np.concatenate(([0], np.cumsum(n)[:-1])), n) + np.arange(int(n.sum()))
    return hashlib.sha256(np.ascontiguousarray(raw[:, :5]).tobytes() + np.ascontiguousarray(gaps[idx]).tobytes()).hexdigest(), len(raw)
ctx = Context(local)
prm = ctx.params(32, 250, 1000, 10, 0.0, False, False)
ctx.synth_target_resident(7, cplan, names)
ctx.synth_reads_resident(7, rplan[lo:hi])
x = nd.GpuExchange(ctx, dist, rank, world)
ctx.events_reset()
n_idx = x.build_index_sharded_resident(32, 250)
st = ctx.map_resident(prm, first_ordinal=lo)
x.gather_events(True)
x.agree_capacity()
ctx.events_reset(); x.build_index_sharded_resident(32, 250); ctx.map_resident(prm, first_ordinal=lo); x.gather_events()     # sync-free exchange
if rank == 0:
    many = digest(ctx)
    c1 = Context(local)
    c1.synth_target_resident(7, cplan, names); c1.synth_reads_resident(7, rplan)
    c1.events_reset(); c1.index_build_resident(32, 250); c1.map_resident(prm, 0)
    one = digest(c1)
    assert n_idx == c1.index_stats()["inserted"], (n_idx, c1.index_stats())
    assert one == many and one[1] > 50, (one, many)
    print("OK", one[1])
dist.barrier(); dist.destroy_process_group()
''' % util.REPO
    script = tmp_path / "shard.py"
    script.write_text(code)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

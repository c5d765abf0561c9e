"""The N>1 path on CPU: world_size-2 gloo process group. Checks the sharding / variable-length all-gather helpers of
ntlink_b200/dist.py and that mapping two read shards separately (kernel logic through tests/emu), gathering the
pair events in rank order and tallying them equals the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, case):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ntlink_b200 import dist as nd
    # 1) variable-length all-gather keeps rank order
    t = torch.arange(3 + 5 * rank, dtype=torch.int64).reshape(-1, 1) + 1000 * rank
    parts = nd.all_gather_var(t, dist)
    assert [p.shape[0] for p in parts] == [3 + 5 * r for r in range(world)]
    assert all(int(p[0, 0]) == 1000 * r for r, p in enumerate(parts))
    # 2) triples gathered in rank order reproduce the unsharded arrays
    tnames, tseq, toff = util.load_fasta_batch(util.fixture_file(tmp, case["target"]))
    th, tp, ts, tmo = util.oracle_sketch_batch(tseq, toff, case["k"], case["w"])
    tpf = (tp | (ts.astype(np.uint32) << 31)).astype(np.uint32)
    tctg = np.repeat(np.arange(len(tnames), dtype=np.int32), np.diff(tmo).astype(np.int64))
    a, b = nd.contig_shard(toff, rank, world)
    lo, hi = int(tmo[a]), int(tmo[b])
    gh, gc, gp = nd.gather_triples(torch.from_numpy(th[lo:hi].view(np.int64).copy()), torch.from_numpy(tctg[lo:hi].copy()),
                                   torch.from_numpy(tpf[lo:hi].view(np.int32).copy()), dist)
    assert np.array_equal(gh.numpy().view(np.uint64), th) and np.array_equal(gc.numpy(), tctg)
    assert np.array_equal(gp.numpy().view(np.uint32), tpf)
    # 3) events of two read shards, gathered in rank order, equal the events of the whole read set
    import test_emu_map as tem
    full = tem.run_emu_events(tmp, case, 0, None)
    rnames, _, _ = util.load_fasta_batch(util.fixture_file(tmp, case["reads"]))
    r0, r1 = nd.read_shard(len(rnames), rank, world)
    mine = tem.run_emu_events(tmp, case, r0, r1)
    allev = nd.gather_events(torch.from_numpy(mine.view(np.int32).copy()), dist).numpy().view(np.uint32)
    assert np.array_equal(allev, full), "sharded events differ from the single-process events"
    # the one-collective gather used by bench.py (capacity grows when a rank overflows it)
    parts = nd.EventGather(capacity=8).gather(torch.from_numpy(mine.view(np.int32).copy()), dist)
    assert np.array_equal(torch.cat(parts).numpy().view(np.uint32), full)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    case = util.manifest()["f3_default"]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), case), nprocs=2, join=True)

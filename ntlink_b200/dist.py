"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards by READS: every rank maps its own contiguous block of reads against the FULL target
index. Two exchange steps exist and only these use a collective:
  1. target sketch: contigs are split over the ranks by cumulative length, each rank sketches its share, the
     (hash, contig, pos|strand) triples are all-gathered so that every GPU builds the same replicated index
     (duplicates are decided on the full multiset, exactly like bin/ntlink_pair.py:204-209);
  2. pair events: every rank's events are gathered in rank order (= global read order, because read blocks are
     contiguous and events carry their global read ordinal) and tallied once.
Nothing here touches the kernels; tensors are plain torch tensors on whatever device the process group uses.
"""
import numpy as np
import torch


def contig_shard(offsets, rank, world):
    "contigs [a, b) of this rank: split by cumulative length so every rank sketches about the same number of bases"
    cum = np.asarray(offsets, dtype=np.int64)
    n = len(cum) - 1
    bounds = [int(np.searchsorted(cum, cum[-1] * r // world)) for r in range(world)] + [n]
    bounds[0] = 0
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds[rank], bounds[rank + 1]


def read_shard(n_reads, rank, world):
    "contiguous block of reads of this rank (keeps global read order = rank order)"
    return n_reads * rank // world, n_reads * (rank + 1) // world


def all_gather_var(t, dist, group=None):
    """all-gather of tensors whose first dimension differs per rank; returns the list of per-rank tensors.
    (counts first, then one padded all_gather -- NCCL has no variable-length gather)"""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    pad[:t.shape[0]] = t
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:c] for o, c in zip(out, counts)]


def gather_triples(hash_i64, ctg_i32, posf_i32, dist, group=None):
    "replicated target index input: every rank's triples concatenated in rank order"
    meta = (ctg_i32.to(torch.int64) << 32) | (posf_i32.to(torch.int64) & 0xFFFFFFFF)
    parts = all_gather_var(torch.stack([hash_i64, meta], dim=1), dist, group)
    allp = torch.cat(parts, dim=0)
    return (allp[:, 0].contiguous(), (allp[:, 1] >> 32).to(torch.int32).contiguous(),
            (allp[:, 1] & 0xFFFFFFFF).to(torch.int32).contiguous())


def gather_events(events_i32, dist, group=None):
    "pair events (n x 6 int32, layout of ntl_event) of all ranks concatenated in rank order = global read order"
    return torch.cat(all_gather_var(events_i32.reshape(-1, 6), dist, group), dim=0)


class EventGather:
    """Gathers the pair events of all ranks with ONE collective and no size round trip: every rank sends a fixed
    capacity buffer whose first row carries its event count; the capacity grows (and the gather is repeated) only
    when some rank overflowed it."""

    def __init__(self, capacity=4096):
        self.cap = capacity
        self.send = None
        self.recv = None

    def _ensure(self, device, world):
        if self.send is None or self.send.shape[0] != self.cap + 1 or self.send.device != device:
            self.send = torch.zeros((self.cap + 1, 6), device=device, dtype=torch.int32)
            self.recv = [torch.zeros_like(self.send) for _ in range(world)]

    def gather(self, events_i32, dist, group=None):
        "events_i32: (n, 6) int32 on the group's device. Returns (list of per-rank event tensors)."
        world = dist.get_world_size(group)
        n = int(events_i32.shape[0])
        while True:
            self._ensure(events_i32.device, world)
            self.send[0, 0] = n
            m = min(n, self.cap)
            if m:
                self.send[1:1 + m] = events_i32[:m]
            dist.all_gather(self.recv, self.send, group=group)
            counts = torch.stack([r[0, 0] for r in self.recv]).tolist()      # the only synchronisation
            if max(counts) <= self.cap:
                return [r[1:1 + c] for r, c in zip(self.recv, counts)]
            self.cap = int(max(counts) * 2)
            self.send = None

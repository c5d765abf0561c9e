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


def init_nccl(local_rank, timeout_minutes=10):
    """torch.distributed over NCCL for this rank's GPU. The communicator is created here (one barrier) with file
    descriptor 1 pointed at stderr meanwhile: NCCL prints its version banner to stdout when it comes up, and the callers
    of this module (bench.py, pair.py, tools/) promise one JSON line / a byte-exact stream there."""
    import datetime
    import os
    import sys
    import torch.distributed as dist
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=timeout_minutes))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return dist


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


class GpuExchange:
    """The two exchange steps on GPUs, driving the library through its C ABI (ctx = ntlink_b200.Context of this rank):

      build_index_sharded*   every rank sketches its contig shard of the target (by cumulative length), the minimizer
                             triples are all-gathered over NCCL, every GPU builds the same replicated index;
      gather_events          pair events of every rank -> rank 0's device event log in rank order = global read order,
                             ONE fixed-capacity all_gather_into_tensor. After agree_capacity() the exchange runs without
                             any host synchronisation (export kernel -> collective -> import kernel; streams ordered by
                             events, per-rank counts interpreted on the device); before it the counts are read on the host
                             and the buffers grow as needed.
    """

    def __init__(self, ctx, dist, rank, world, device=None):
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.cap, self.send, self.recv, self.agreed, self.stream = 8192, None, None, False, None
        self.index_bytes = 0

    # ---- replicated index
    def _build_from_shard(self, n_mx, dh, dc, dp, contig_len, name_rank):
        import ctypes as C
        ctx = self.ctx
        m = int(n_mx)
        h = torch.empty(m, device=self.dev, dtype=torch.int64)
        ctg = torch.empty(m, device=self.dev, dtype=torch.int32)
        posf = torch.empty(m, device=self.dev, dtype=torch.int32)
        if m:
            ctx._check(ctx.lib.ntl_copy_device(ctx.h, h.data_ptr(), dh, m * 8), "copy")
            ctx._check(ctx.lib.ntl_copy_device(ctx.h, ctg.data_ptr(), dc, m * 4), "copy")
            ctx._check(ctx.lib.ntl_copy_device(ctx.h, posf.data_ptr(), dp, m * 4), "copy")
        hashes, ctgs, posfs = gather_triples(h, ctg, posf, self.dist)
        torch.cuda.synchronize()
        self.index_bytes = int(hashes.numel()) * 16
        ctx._check(ctx.lib.ntl_index_build_device(ctx.h, hashes.data_ptr(), ctgs.data_ptr(), posfs.data_ptr(), int(hashes.numel()),
                                                  contig_len.ctypes.data, name_rank.ctypes.data, len(contig_len)), "ntl_index_build_device")
        return int(hashes.numel())

    def build_index_sharded_resident(self, k, w):
        "target = the context's resident target (ntl_target_upload / ntl_synth_target_resident)"
        import ctypes as C
        ctx = self.ctx
        ncontig, _ = ctx.resident_info(0)
        cl = np.empty(ncontig, np.uint32)
        rk = np.empty(ncontig, np.uint32)
        ctx._check(ctx.lib.ntl_target_resident_meta(ctx.h, cl.ctypes.data, rk.ctypes.data), "ntl_target_resident_meta")
        off = np.zeros(ncontig + 1, np.int64)
        off[1:] = np.cumsum(cl.astype(np.int64))
        a, b = contig_shard(off, self.rank, self.world)
        n, dh, dc, dp = C.c_uint64(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        ctx._check(ctx.lib.ntl_target_sketch_resident(ctx.h, a, b - a, k, w, C.byref(n), C.byref(dh), C.byref(dc), C.byref(dp)),
                   "ntl_target_sketch_resident")
        return self._build_from_shard(n.value, dh, dc, dp, cl, rk)

    def build_index_sharded(self, contigs, k, w):
        "target = a host SeqBatch (every rank holds it; each one copies and sketches only its shard)"
        import ctypes as C
        from .api import SeqBatch, name_ranks
        ctx = self.ctx
        cum = contigs.offsets.astype(np.int64)
        a, b = contig_shard(contigs.offsets, self.rank, self.world)
        part = SeqBatch(contigs.seq[int(cum[a]):int(cum[b])], contigs.offsets[a:b + 1] - contigs.offsets[a], contigs.names[a:b])
        sk = ctx.sketch(part, k, w)            # the triples also stay on the device; the host copy gives the contig ids
        m = len(sk.hash)
        nmx, dh, dp, do = C.c_uint64(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        ctx._check(ctx.lib.ntl_device_sketch_arrays(ctx.h, C.byref(nmx), C.byref(dh), C.byref(dp), C.byref(do)), "arrays")
        ctg = torch.from_numpy(np.repeat(np.arange(a, b, dtype=np.int32), np.diff(sk.seq_off).astype(np.int64))).to(self.dev)
        return self._build_from_shard(m, dh, ctg.data_ptr(), dp, contigs.lengths.astype(np.uint32), name_ranks(contigs.names))

    # ---- pair events
    def _buffers(self):
        if self.send is None or self.send.shape[0] != (self.cap + 1) * 6:
            self.send = torch.zeros((self.cap + 1) * 6, device=self.dev, dtype=torch.int32)
            self.recv = torch.zeros(self.world * (self.cap + 1) * 6, device=self.dev, dtype=torch.int32)

    def _gather_sync(self):
        import ctypes as C
        ctx = self.ctx
        while True:
            self._buffers()
            n = C.c_uint64()
            ctx._check(ctx.lib.ntl_events_export(ctx.h, self.send.data_ptr(), self.cap, C.byref(n)), "ntl_events_export")
            self.dist.all_gather_into_tensor(self.recv, self.send)
            counts = self.recv.view(self.world, -1)[:, 0].tolist()
            if max(counts) <= self.cap:
                if self.rank == 0:
                    ovf = C.c_int(0)
                    ctx._check(ctx.lib.ntl_events_import_gathered(ctx.h, self.recv.data_ptr(), self.world, self.cap, C.byref(ovf)),
                               "ntl_events_import_gathered")
                return
            self.cap = int(max(counts)) * 2
            self.send = None

    def gather_events(self, force_sync=False):
        import ctypes as C
        ctx = self.ctx
        if force_sync or not self.agreed:
            self._gather_sync()
            return
        if self.stream is None:
            sp = C.c_void_p()
            ctx._check(ctx.lib.ntl_stream(ctx.h, C.byref(sp)), "ntl_stream")
            self.stream = torch.cuda.ExternalStream(sp.value, device=self.dev)
        n = C.c_uint64()
        ctx._check(ctx.lib.ntl_events_export_async(ctx.h, self.send.data_ptr(), self.cap, C.byref(n)), "ntl_events_export_async")
        if n.value > self.cap:
            raise RuntimeError("event exchange buffer too small for this step; run more warm-up steps")
        torch.cuda.current_stream().wait_stream(self.stream)
        self.dist.all_gather_into_tensor(self.recv, self.send)
        self.stream.wait_stream(torch.cuda.current_stream())
        if self.rank == 0:
            ctx._check(ctx.lib.ntl_events_import_device(ctx.h, self.recv.data_ptr(), self.world, self.cap), "ntl_events_import_device")

    def agree_capacity(self):
        "after some synchronous exchanges: every rank takes the same capacity (4x the largest count seen, >= 8192 events)"
        seen = int(self.recv.view(self.world, -1)[:, 0].max().item()) if self.recv is not None else 0
        t = torch.tensor([seen], device=self.dev, dtype=torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        self.cap = max(8192, 4 * int(t.item()))
        self.send = None
        self._buffers()
        torch.cuda.synchronize()
        self.agreed = True

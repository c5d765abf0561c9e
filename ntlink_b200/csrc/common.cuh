// common.cuh -- context, device buffers and error plumbing shared by the translation units of libntlink_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ntlink_b200.h"
#include "map_logic.cuh"
#include "nthash.cuh"
#include "sketch_logic.cuh"

namespace ntl {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // state of the single-pass scan when the buffer serves as its scratch (scan.cu)
    uint64_t scan_epoch = 0;
    void* scan_ready_for = nullptr;
    // grows geometrically; contents are NOT preserved
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// device-side counters of one sketch pass (one 64-byte block, zeroed before every pass)
struct SketchStatus {
    uint32_t nstrips;       // strips of this batch
    uint32_t pool_used;     // overflow-pool entries handed out
    uint32_t ngaps;         // GapRec entries queued
    uint32_t extras_used;   // gap minimizers reserved
    uint32_t n_mx;          // total minimizers (after emit-scan)
    uint32_t err;           // NTL_SKERR_* bits
    uint32_t n_ovf;         // strips that overflowed their slots
    uint32_t n_cand;        // total candidates (statistics)
    uint32_t pad[8];
};
enum : uint32_t { SKERR_POOL = 1, SKERR_GAPS = 2, SKERR_EXTRAS = 4, SKERR_OUT = 8 };

struct MapStatus {
    uint32_t n_hits;        // index hits of the batch
    uint32_t n_events;      // pair events written
    uint32_t err;           // MAPERR_* bits
    uint32_t n_runs;        // accepted (read, contig) runs
    uint32_t pad[12];
};
enum : uint32_t { MAPERR_EVENTS = 1, MAPERR_ASSERT = 2, MAPERR_READ_EVENTS = 4 };

// per-stage device timings (CUDA events on ctx->stream), milliseconds; stage ids are NTL_T_* of the public header
enum { T_PACK = NTL_T_PACK, T_DENSE = NTL_T_DENSE, T_SELECT = NTL_T_SELECT, T_GAP = NTL_T_GAP, T_EMIT = NTL_T_EMIT,
       T_LOOKUP = NTL_T_LOOKUP, T_CHAIN = NTL_T_CHAIN, T_TALLY = NTL_T_TALLY, T_INDEX = NTL_T_INDEX,
       T_TOTAL = NTL_T_TOTAL, T_NUM = NTL_T_NUM };

struct SketchWork {           // device workspace of the sketch pipeline (reused across calls)
    DevBuf packed, scnt, strip_off, blocksums, slots, cnt, nv, vbase, ovf_off, sel, selcnt, selmask, selbase, strip_seq,
        gaps, gap_head, extras, has_cand, status, tbl, tile_state, stage_hash, stage_posf;
    uint32_t tbl_k = 0;       // k the device roll table was built for
};

struct DeviceSketch {         // result of a sketch pass, resident on the device
    DevBuf hash;              // uint64 h1 per minimizer
    DevBuf posf;              // uint32 pos | fwd << 31
    DevBuf mx_off;            // uint32 [nseq + 1]
    uint32_t n_mx = 0;        // exact after a synchronous pass; an UPPER BOUND after a deferred (sync-free) pass
    uint32_t nseq = 0;
    const uint32_t* n_dev = nullptr;   // device word holding the exact count (null: n_mx is exact and host-only)
};

// Device-side state of one sync-free call (ntl_map_reads / ntl_map_resident fast path): the chunks of the call are
// enqueued back to back without a host synchronisation in between; errors are sticky and make the host repeat the
// call on the synchronous path, totals place each chunk's results behind the previous chunk's.
struct CallState {
    uint32_t err;                                            // CALLERR_* bits
    uint32_t log_n;                                          // cursor of the device event log
    uint32_t hits_total, ev_total, runs_total, mx_total;     // totals over the finished chunks
    uint32_t pad[10];
};
enum : uint32_t { CALLERR_SKETCH = 1, CALLERR_MAP = 2, CALLERR_RESULTS = 4 };
struct HostResults {          // pinned host arrays (device-addressable) the fast path writes results into
    uint32_t *hit_off, *nruns, *ev_off, *ev_cnt;
    Run* runs; Hit* hits; Event* events;
    uint32_t hits_cap, ev_cap;
};

struct TargetIndex {
    DevBuf table, special, ctg_len, name_rank, dupflag;
    uint64_t slots = 0;
    uint32_t ncontig = 0;
    uint64_t n_inserted = 0;
    bool built = false;
};

struct PreMappings {          // accepted runs/hits supplied by the host (checkpoint path), ntl_map_out layout
    const uint32_t* hit_off; const uint32_t* nruns; const Run* runs; const Hit* hits; uint32_t n_hits;
    bool resident;            // the arrays are already in MapWork (left there by liftover_device): no upload
    bool compute_read_len;    // read length = largest first/last read position of the read's runs (pair:483-487)
};

struct TallyWork {
    DevBuf keys, pn, panchor, pfirst, ev_slot, gap_off, cursor, gkey, gval, nonempty, ppref, out, ndev, bs, skey, sval;
    PinnedBuf h_stage;        // pair table + gap lists on their way to the host
};

struct MapWork {
    DevBuf hit_tmp, hit_flag, hit_pref, hits, runs, mark, hit_off, nruns, events, status, read_len, ev_cnt, blocksums;
    DevBuf lift_runs, lift_nruns, lift_agp;     // liftover inputs
    DevBuf gm[15];                              // grouped mapping (ntl_map_groups): inputs, per-group plan, global tables
    bool lifted_valid = false;                  // hits/runs/nruns/hit_off hold the output of the last liftover
    uint32_t lifted_reads = 0, lifted_hits = 0;
};

}  // namespace ntl

struct ntl_ctx {
    int device = 0;
    void* res = nullptr;                   // ntl::Results (capi.cu): host result buffers, resident inputs
    cudaStream_t stream = nullptr;
    std::string err;
    ntl::SketchWork sw;
    ntl::MapWork mw;
    ntl::TallyWork tw;
    ntl::TargetIndex index;
    ntl::DevBuf d_seq, d_off;              // staging of the current batch (ASCII + offsets)
    ntl::DeviceSketch dsk;                 // sketch of the current batch
    ntl::PinnedBuf h_status;
    // tuning knobs
    uint32_t strip_len = 0;                // k-mer positions per thread of the dense pass; 0 = chosen per batch (auto_strip_len)
    double cand_c = 7.0;
    uint64_t batch_bases = 1ull << 30;
    uint64_t pipeline_min_bases = 80ull << 20;   // smallest batch worth pipelining (copy/compute overlap)
    // timing
    struct StageRec { int stage; uint32_t ev; bool closed; uint64_t bases; };
    std::vector<cudaEvent_t> ev_pool;      // stage-timer events, two per record, recycled by collect_timing
    std::vector<StageRec> ev_recs;
    size_t ev_next = 0;
    int ev_open[ntl::T_NUM];
    cudaEvent_t mark[2];
    double ms_accum[ntl::T_NUM];           // accumulated since ntl_timing_reset
    uint64_t launches = 0;                 // kernels launched since ntl_timing_reset
    uint64_t dense_launches = 0;
    uint64_t dense_bases = 0;
    double big_dense_ms = 0;               // k_dense launches over >= 16 Mbp only (read batches)
    uint64_t big_dense_launches = 0, big_dense_bases = 0;
    // tally state (pairs accumulated over calls)
    ntl::DevBuf tl_events;                 // all events appended so far
    uint64_t tl_n_events = 0;
    ntl::DevBuf tl_count;                  // {exact number of events, overflow flag} after ntl_events_import_device
    bool tl_count_on_device = false;       // tl_n_events is only an upper bound until the next tally
    // sync-free call state
    ntl::DevBuf call_state;                // ntl::CallState
    uint64_t tl_pending_bound = 0;         // events the chunks in flight may still append (capacity reserved for them)
    uint32_t ev_cap_hint = 0;              // event buffer size that was enough so far
    int async_mode = 1;                    // 0: always take the synchronous path (option "async")
    uint64_t n_async_calls = 0, n_async_fallbacks = 0, n_graph_launches = 0, n_graph_failures = 0;   // ntl_get_stat
    double mx_density_factor = 2.6;        // minimizers per base * (w + 1), upper estimate (see sketch_out_bound)
    int copy_threads = -1;                 // host threads of the pageable->pinned bounce copy (-1 auto, 0 = off)
    uint64_t n_tile_batches = 0, n_tile_fallbacks = 0;   // ntl_get_stat "tile_batches" / "tile_fallbacks"
    int small_mode = 1;                    // dense-mode kernels (small_kernel.cuh) for w <= 16: option "small" = 0 off, 1 auto, 2 tile, 3 streaming
    uint64_t n_small_batches = 0;          // ntl_get_stat "small_batches"
    int tile_mode = 0;                     // single-pass sketch kernel (tile_kernel.cuh) where the window size allows (option "tile")
    int graph_mode = 1;                    // sync-free ntl_map_reads: one CUDA graph per chunk (option "graph")
    bool capturing = false;                // c->stream is being captured: no synchronisation, no allocation-by-copy
    bool no_stage_timing = false;          // chunk graphs of a pipelined call: several in flight, stage events meaningless
};

#define NTL_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            char b__[512];                                                                       \
            snprintf(b__, sizeof b__, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            (ctx)->err = b__;                                                                    \
            return NTL_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define NTL_TRY(expr)                    \
    do {                                 \
        int r__ = (expr);                \
        if (r__ != NTL_OK) return r__;   \
    } while (0)

namespace ntl {
// implemented in sketch.cu
// call_state != null: deferred pass -- no host synchronisation, the output is sized by an upper bound, a workspace
// overflow sets CALLERR_SKETCH in call_state->err and yields an empty sketch
int sketch_device(ntl_ctx* c, const uint8_t* d_seq, const uint64_t* d_off, uint32_t nseq, uint64_t total_bases,
                  uint32_t k, uint32_t w, DeviceSketch& out, CallState* call_state = nullptr);
// uploads the rolling table of k if it is not the current one (synchronises; must not run inside a stream capture)
int sketch_prepare(ntl_ctx* c, uint32_t k);
// implemented in map.cu
int index_build_device(ntl_ctx* c, const uint64_t* d_hash, const uint32_t* d_ctg, const uint32_t* d_posf, uint64_t n,
                       const uint32_t* h_ctg_len, const uint32_t* h_name_rank, uint32_t ncontig, const uint32_t* n_dev = nullptr,
                       bool sync = true);
// Upper bound of the number of minimizers a deferred sketch pass reserves room for: random sequence has 2/(w+1) per
// base; `factor` (ntl_ctx::mx_density_factor, >= 2.6) follows the densest input this context has seen, so that a run on
// low-complexity data pays the fallback to the synchronous path once, not for every batch.
inline uint32_t sketch_out_bound(uint64_t total_bases, uint32_t nseq, uint32_t w, double factor) {
    return (uint32_t)std::min<uint64_t>(total_bases, (uint64_t)(factor * (double)total_bases / ((double)w + 1.0)) + 8ull * nseq + 4096);
}
// scan utility (scan.cu): exclusive prefix sum of in[0..n) into out[0..n], out[n] = total; n read from the device
int exclusive_scan_u32(ntl_ctx* c, const uint32_t* in, uint32_t* out, const uint32_t* n_dev, uint32_t n_max,
                       DevBuf& blocksums);
// Small fills as ONE kernel instead of cudaMemsetAsync: the driver may route a memset through a copy engine, where it
// would queue behind the large host->device copy of the next batch (measured: kernels of a batch ran ~2x longer while
// a copy was in flight).
struct FillSegs { void* p[6]; uint64_t n[6]; uint32_t v[6]; };
static __global__ void k_fill_segs(FillSegs s) {
#pragma unroll
    for (int q = 0; q < 6; q++) {
        uint8_t* p = (uint8_t*)s.p[q];
        const uint32_t b = s.v[q] & 0xFFu;
        if ((((uintptr_t)p | s.n[q]) & 15u) == 0) {                      // aligned segments: 16 bytes per store
            const uint32_t w = b * 0x01010101u;
            uint4* p4 = (uint4*)p;
            for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n[q] / 16; i += (uint64_t)gridDim.x * blockDim.x) p4[i] = make_uint4(w, w, w, w);
        } else {
            for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n[q]; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (uint8_t)b;
        }
    }
}
// Stage timers: every tick/tock pair takes its own pair of CUDA events from a pool that is recycled by collect_timing, so
// that a call made of several chunks (each with the same stages, several of them in flight) still yields every stage's
// device time. Inside a stream capture the records become event-record nodes of the graph (cudaEventRecordExternal), so
// the stage times of a graph launch can be read like those of plain launches.
inline void note_mx_density(ntl_ctx* c, uint64_t n_mx, uint64_t bases, uint32_t w) {
    if (bases < 100000) return;                                   // too small to say anything
    const double seen = (double)n_mx * ((double)w + 1.0) / (double)bases;
    if (1.3 * seen > c->mx_density_factor) c->mx_density_factor = 1.3 * seen;
}
inline void tick(ntl_ctx* c, int stage) {
    if (c->no_stage_timing) return;
    if (c->ev_next + 2 > c->ev_pool.size()) {
        cudaEvent_t a = nullptr, b = nullptr;
        cudaEventCreate(&a); cudaEventCreate(&b);
        c->ev_pool.push_back(a); c->ev_pool.push_back(b);
    }
    ntl_ctx::StageRec r;
    r.stage = stage; r.ev = (uint32_t)c->ev_next; r.closed = false; r.bases = 0;
    c->ev_next += 2;
    c->ev_open[stage] = (int)c->ev_recs.size();
    c->ev_recs.push_back(r);
    cudaEventRecordWithFlags(c->ev_pool[r.ev], c->stream, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}
inline void tock(ntl_ctx* c, int stage, uint64_t bases = 0) {
    if (c->no_stage_timing) return;
    const int i = c->ev_open[stage];
    if (i < 0 || (size_t)i >= c->ev_recs.size()) return;
    ntl_ctx::StageRec& r = c->ev_recs[(size_t)i];
    cudaEventRecordWithFlags(c->ev_pool[r.ev + 1], c->stream, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
    r.closed = true; r.bases = bases;
    c->ev_open[stage] = -1;
}
// after a stream synchronize: fold the recorded stage times into the accumulators
inline void collect_timing(ntl_ctx* c) {
    for (const ntl_ctx::StageRec& r : c->ev_recs) {
        if (!r.closed) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev_pool[r.ev], c->ev_pool[r.ev + 1]) == cudaSuccess) {
            c->ms_accum[r.stage] += ms;
            if (r.stage == T_DENSE && r.bases >= (16ull << 20)) {
                c->big_dense_ms += ms; c->big_dense_launches += 1; c->big_dense_bases += r.bases;
            }
        } else cudaGetLastError();
    }
    c->ev_recs.clear();
    c->ev_next = 0;
    for (int s = 0; s < T_NUM; s++) c->ev_open[s] = -1;
}
}  // namespace ntl

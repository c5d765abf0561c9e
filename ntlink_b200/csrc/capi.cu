// capi.cu -- the extern "C" boundary of libntlink_b200.so (declared in include/ntlink_b200.h).
// Host orchestration only: batching, H2D/D2H staging through pinned memory, error translation. All compute is
// in sketch.cu / map.cu. There is deliberately no CPU fallback: without a CUDA device ntl_init fails.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <thread>

#include "common.cuh"
#include "lift_logic.cuh"
#include "synth_kernels.cuh"

namespace ntl {
int expand_contig_ids(ntl_ctx* c, const DeviceSketch& sk, DevBuf& ctg_ids);
int map_device(ntl_ctx* c, const DeviceSketch& sk, const uint32_t* d_read_len, uint32_t nreads, uint64_t first_ordinal,
               const ntl_params* prm, MapStatus* counts_out, uint64_t* log_base_out, const PreMappings* pre = nullptr,
               CallState* call = nullptr);
int call_begin(ntl_ctx* c, CallState** call_out);
int call_reserve_events(ntl_ctx* c, uint32_t nreads);
int call_note_mx(ntl_ctx* c, const uint32_t* n_dev, CallState* call);
int events_import_device(ntl_ctx* c, const void* d_src, uint32_t world, uint64_t cap_events);
int events_resolve_count(ntl_ctx* c);
int group_map_device(ntl_ctx* c, const uint64_t* t_hash, const uint32_t* t_posf, const uint64_t* t_mx_off, const uint32_t* t_len,
                     uint32_t ntargets, const uint32_t* g_t_off, const uint64_t* r_hash, const uint32_t* r_posf,
                     const uint64_t* r_mx_off, const uint32_t* r_len, uint32_t ngroups, const ntl_params* prm, MapStatus* counts_out);
int call_chunk_finish(ntl_ctx* c, CallState* call, uint32_t rb, uint32_t nreads, const HostResults* H);
int call_end(ntl_ctx* c, CallState* call, CallState* host_out);
int liftover_device(ntl_ctx* c, const uint32_t* hit_off, const uint32_t* nruns, const Run* runs, const Hit* hits, uint32_t nreads,
                    const AgpRow* agp, uint32_t ncontig, int32_t k, MapStatus* counts_out);
int tally_device(ntl_ctx* c, std::vector<ntl_pair>& pairs, std::vector<int32_t>& gaps);
int read_len_device(ntl_ctx* c, const uint64_t* d_off, uint32_t nreads, DevBuf& out);

// pinned host vector that can grow while keeping its contents
struct HostVec {
    PinnedBuf b;
    size_t used = 0;   // bytes
    int reserve(size_t bytes) {
        if (bytes <= b.cap) return 0;
        PinnedBuf nb;
        if (nb.ensure(bytes + bytes / 2) != cudaSuccess) return -1;
        if (used) memcpy(nb.p, b.p, used);
        b.release();
        b = nb;
        return 0;
    }
    template <class T> T* at(size_t byte_off) { return (T*)((char*)b.p + byte_off); }
};

struct Results {
    // sketch
    HostVec sk_hash, sk_posf, sk_off;
    // map
    HostVec hit_off, nruns, runs, hits, ev_off, ev_cnt, events;
    // pairs
    std::vector<ntl_pair> pairs;
    std::vector<int32_t> gaps;
    // resident inputs (bench)
    // resident reads: ASCII of all reads back to back, cut into chunks of whole reads (< 4 Gbp each, 32-bit positions
    // inside a device batch); r_off holds the offsets of every chunk rebased to the chunk's first base, back to back
    struct ResChunk { uint32_t first, count; uint64_t base, nbases, dev_base; size_t off_at; };   // dev_base: 256-byte aligned place in r_seq
    DevBuf r_seq, r_off, r_len; uint32_t r_nreads = 0; uint64_t r_bases = 0;
    std::vector<ResChunk> r_chunks;
    std::vector<uint64_t> r_abs_off;           // host copy of the absolute offsets [r_nreads + 1]
    std::vector<cudaGraphExec_t> resident_execs;
    uint64_t resident_chunk_bases = 512ull << 20;
    DevBuf t_seq, t_off, t_ctg; uint32_t t_ncontig = 0; uint64_t t_bases = 0;
    std::vector<uint32_t> t_len, t_rank;
    DevBuf stage_off, ctg_ids, read_len, in_hash, in_posf, in_ctg;
    PinnedBuf h_off;
    // pipelined read path: copy stream + double-buffered staging
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};
    DevBuf st_seq[2], st_off[2];
    PinnedBuf st_hoff[2];
    // sync-free path: staging slot free again (its chunk's kernels are done), rebased offsets of every chunk of the call
    cudaEvent_t comp_done[2] = {nullptr, nullptr};
    PinnedBuf pin_slot[2];                     // bounce buffers for callers that pass pageable memory
    std::vector<cudaGraphExec_t> chunk_execs;
    cudaGraphExec_t index_exec = nullptr;
    PinnedBuf idx_meta;                        // contig lengths + name ranks of the index being built (stable for the graph)
    PinnedBuf call_hoff;
    uint64_t last_call_hits = 0, last_call_bases = 0;
};
}  // namespace ntl

using namespace ntl;

static __global__ void k_export_header(uint32_t* dst, uint32_t n) { if (threadIdx.x < 6) dst[threadIdx.x] = threadIdx.x == 0 ? n : 0u; }
// copy with a few threads: one core moves ~10 GB/s, the PCIe link takes 55
static void parallel_memcpy(char* dst, const char* src, size_t n, int threads) {
    if (threads <= 1 || n < (4u << 20)) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    const size_t per = ((n + threads - 1) / threads + 4095) & ~(size_t)4095;
    for (int t = 1; t < threads; t++) {
        const size_t b = (size_t)t * per;
        if (b >= n) break;
        th.emplace_back([=]() { memcpy(dst + b, src + b, std::min(per, n - b)); });
    }
    memcpy(dst, src, std::min(per, n));
    for (auto& x : th) x.join();
}
// A CUDA-graph step failed (capture, instantiate or launch): graphs are switched off for this context and the caller
// falls back to plain launches -- slower next to a host->device copy, never wrong.
// tests only: NTL_TEST_GRAPH_FAIL=1 fails every graph step, =map only the second chunk of ntl_map_reads
static bool inject_graph_failure(const char* site) {
    static const char* v = getenv("NTL_TEST_GRAPH_FAIL");
    return v && (!strcmp(v, "1") || !strcmp(v, site));
}
static void graph_failed(ntl_ctx* c, const char* what, cudaError_t e) {
    c->graph_mode = 0;
    c->capturing = false; c->no_stage_timing = false;
    c->n_graph_failures++;
    cudaGetLastError();
    if (getenv("NTL_TRACE")) fprintf(stderr, "[ntl] %s failed (%s): CUDA graphs disabled for this context\n", what, cudaGetErrorString(e));
}
static double host_now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static Results* res_of(ntl_ctx* c) { return static_cast<Results*>(c->res); }

static int finish_call(ntl_ctx* c) {
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    collect_timing(c);
    return NTL_OK;
}

// split [0, nseq) into batches of at most batch_bases bases (a single longer sequence forms its own batch)
static void plan_batches(const uint64_t* offsets, uint32_t nseq, uint64_t batch_bases, std::vector<uint32_t>& bounds) {
    bounds.clear();
    bounds.push_back(0);
    uint32_t b = 0;
    while (b < nseq) {
        uint32_t e = b + 1;
        while (e < nseq && offsets[e + 1] - offsets[b] <= batch_bases) e++;
        bounds.push_back(e);
        b = e;
    }
}

// copy sequences [b, e) to the device: ASCII into c->d_seq and rebased offsets into c->d_off
static int stage_batch(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t b, uint32_t e, uint64_t* nbases_out) {
    Results* R = res_of(c);
    const uint64_t base = offsets[b], nb = offsets[e] - base;
    const uint32_t ns = e - b;
    if (nb >= (1ull << 32) - 4096) { c->err = "a single sequence/batch exceeds 4 Gbp"; return NTL_ERR_ARG; }
    NTL_CUDA(c, c->d_seq.ensure(nb + 256));
    NTL_CUDA(c, c->d_off.ensure(((size_t)ns + 1) * 8));
    NTL_CUDA(c, R->h_off.ensure(((size_t)ns + 1) * 8));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));      // h_off may still be in flight from the previous batch
    uint64_t* ho = R->h_off.as<uint64_t>();
    for (uint32_t i = 0; i <= ns; i++) ho[i] = offsets[b + i] - base;
    if (nb) NTL_CUDA(c, cudaMemcpyAsync(c->d_seq.p, seq + base, nb, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(c->d_off.p, ho, ((size_t)ns + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    *nbases_out = nb;
    return NTL_OK;
}

// asynchronous variant for the pipelined read path: copies on the copy stream into staging slot `slot`
static int stage_batch_async(ntl_ctx* c, int slot, const char* seq, const uint64_t* offsets, uint32_t b, uint32_t e,
                             uint64_t* nbases_out) {
    Results* R = res_of(c);
    const uint64_t base = offsets[b], nb = offsets[e] - base;
    const uint32_t ns = e - b;
    if (nb >= (1ull << 32) - 4096) { c->err = "a single sequence/batch exceeds 4 Gbp"; return NTL_ERR_ARG; }
    NTL_CUDA(c, R->st_seq[slot].ensure(nb + 256));
    NTL_CUDA(c, R->st_off[slot].ensure(((size_t)ns + 1) * 8));
    NTL_CUDA(c, R->st_hoff[slot].ensure(((size_t)ns + 1) * 8));
    uint64_t* ho = R->st_hoff[slot].as<uint64_t>();
    for (uint32_t i = 0; i <= ns; i++) ho[i] = offsets[b + i] - base;
    if (nb) NTL_CUDA(c, cudaMemcpyAsync(R->st_seq[slot].p, seq + base, nb, cudaMemcpyHostToDevice, R->copy_stream));
    NTL_CUDA(c, cudaMemcpyAsync(R->st_off[slot].p, ho, ((size_t)ns + 1) * 8, cudaMemcpyHostToDevice, R->copy_stream));
    NTL_CUDA(c, cudaEventRecord(R->h2d_done[slot], R->copy_stream));
    *nbases_out = nb;
    return NTL_OK;
}

extern "C" {

int ntl_version(void) { return 100; }

int ntl_init(int device, ntl_ctx** out) {
    if (!out) return NTL_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return NTL_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return NTL_ERR_CUDA;
    ntl_ctx* c = new ntl_ctx();
    c->res = new Results();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete res_of(c); delete c; return NTL_ERR_CUDA; }
    cudaEventCreate(&c->mark[0]); cudaEventCreate(&c->mark[1]);
    cudaStreamCreateWithFlags(&res_of(c)->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&res_of(c)->h2d_done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&res_of(c)->h2d_done[1], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&res_of(c)->comp_done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&res_of(c)->comp_done[1], cudaEventDisableTiming);
    for (int i = 0; i < T_NUM; i++) { c->ev_open[i] = -1; c->ms_accum[i] = 0; }
    *out = c;
    return NTL_OK;
}

void ntl_destroy(ntl_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    Results* R = res_of(c);
    SketchWork& W = c->sw;
    DevBuf* sb[] = {&W.packed, &W.scnt, &W.strip_off, &W.blocksums, &W.slots, &W.cnt, &W.nv, &W.vbase, &W.ovf_off, &W.sel,
                    &W.selcnt, &W.selmask, &W.selbase, &W.strip_seq, &W.gaps, &W.gap_head, &W.extras, &W.has_cand, &W.status, &W.tbl, &W.tile_state, &W.stage_hash, &W.stage_posf};
    for (DevBuf* b : sb) b->release();
    MapWork& M = c->mw;
    DevBuf* mb[] = {&M.hit_tmp, &M.hit_flag, &M.hit_pref, &M.hits, &M.runs, &M.mark, &M.hit_off, &M.nruns, &M.events,
                    &M.status, &M.read_len, &M.ev_cnt, &M.blocksums, &M.lift_runs, &M.lift_nruns, &M.lift_agp};
    for (DevBuf* b : mb) b->release();
    for (DevBuf& b : M.gm) b.release();
    ntl::TallyWork& TW = c->tw;
    DevBuf* tb[] = {&TW.keys, &TW.pn, &TW.panchor, &TW.pfirst, &TW.ev_slot, &TW.gap_off, &TW.cursor, &TW.gkey, &TW.gval,
                    &TW.nonempty, &TW.ppref, &TW.out, &TW.ndev, &TW.bs, &TW.skey, &TW.sval};
    for (DevBuf* b : tb) b->release();
    TW.h_stage.release();
    c->call_state.release(); c->tl_count.release();
    DevBuf* ib[] = {&c->index.table, &c->index.special, &c->index.ctg_len, &c->index.name_rank, &c->index.dupflag,
                    &c->d_seq, &c->d_off, &c->dsk.hash, &c->dsk.posf, &c->dsk.mx_off, &c->tl_events,
                    &R->r_seq, &R->r_off, &R->r_len, &R->t_seq, &R->t_off, &R->t_ctg, &R->stage_off, &R->ctg_ids,
                    &R->read_len, &R->in_hash, &R->in_posf, &R->in_ctg};
    for (DevBuf* b : ib) b->release();
    HostVec* hv[] = {&R->sk_hash, &R->sk_posf, &R->sk_off, &R->hit_off, &R->nruns, &R->runs, &R->hits, &R->ev_off,
                     &R->ev_cnt, &R->events};
    for (HostVec* v : hv) v->b.release();
    R->h_off.release();
    for (int i = 0; i < 2; i++) { R->st_seq[i].release(); R->st_off[i].release(); R->st_hoff[i].release(); cudaEventDestroy(R->h2d_done[i]); cudaEventDestroy(R->comp_done[i]); }
    R->call_hoff.release();
    R->pin_slot[0].release(); R->pin_slot[1].release();
    for (cudaGraphExec_t e : R->chunk_execs) if (e) cudaGraphExecDestroy(e);
    R->chunk_execs.clear();
    for (cudaGraphExec_t e : R->resident_execs) if (e) cudaGraphExecDestroy(e);
    R->resident_execs.clear();
    if (R->index_exec) cudaGraphExecDestroy(R->index_exec);
    R->idx_meta.release();
    cudaStreamSynchronize(R->copy_stream);
    cudaStreamDestroy(R->copy_stream);
    c->h_status.release();
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    cudaEventDestroy(c->mark[0]); cudaEventDestroy(c->mark[1]);
    cudaStreamDestroy(c->stream);
    delete R;
    delete c;
}

const char* ntl_last_error(const ntl_ctx* c) { return c ? c->err.c_str() : "no context (ntl_init failed: no CUDA device?)"; }

int ntl_set_option(ntl_ctx* c, const char* name, double value) {
    if (!c || !name) return NTL_ERR_ARG;
    if (!strcmp(name, "strip_len")) {
        uint32_t v = (uint32_t)value;
        if (v != 0 && (v < 8 || v > 65536 || (v & 7))) { c->err = "strip_len must be 0 (automatic) or a multiple of 8 in [8, 65536]"; return NTL_ERR_ARG; }
        c->strip_len = v;
    } else if (!strcmp(name, "cand_c")) {
        if (!(value > 0)) { c->err = "cand_c must be positive"; return NTL_ERR_ARG; }
        c->cand_c = value;
    } else if (!strcmp(name, "async")) {
        c->async_mode = value != 0.0;
    } else if (!strcmp(name, "copy_threads")) {
        c->copy_threads = (int)value;
    } else if (!strcmp(name, "tile")) {
        c->tile_mode = value != 0.0;
    } else if (!strcmp(name, "small")) {
        c->small_mode = (int)value;                        // 0 off, 1 automatic, 2 tile kernel, 3 streaming kernel
    } else if (!strcmp(name, "graph")) {
        c->graph_mode = value != 0.0;
    } else if (!strcmp(name, "pipeline_min_bases")) {
        if (value < 1024 || value > 3.9e9) { c->err = "pipeline_min_bases out of range"; return NTL_ERR_ARG; }
        c->pipeline_min_bases = (uint64_t)value;
    } else if (!strcmp(name, "resident_chunk_bases")) {
        if (value < 65536 || value > 3.9e9) { c->err = "resident_chunk_bases out of range"; return NTL_ERR_ARG; }
        res_of(c)->resident_chunk_bases = (uint64_t)value;
    } else if (!strcmp(name, "batch_bases")) {
        if (value < 1024 || value > 3.9e9) { c->err = "batch_bases out of range"; return NTL_ERR_ARG; }
        c->batch_bases = (uint64_t)value;
    } else { c->err = std::string("unknown option ") + name; return NTL_ERR_ARG; }
    return NTL_OK;
}

// ------------------------------------------------------------------------------------------- sketch
// Sync-free index build from sequences that are already on the device: deferred sketch, contig ids, table insert and
// finalize as ONE graph, one synchronisation. *ok = false (nothing built) when the sketch outgrew its bound: the caller
// then takes the synchronous path.
static int index_build_async(ntl_ctx* c, const uint8_t* d_seq, const uint64_t* d_off, uint32_t nseq, uint64_t nbases, int k, int w,
                             const uint32_t* h_len, const uint32_t* h_rank, DevBuf& ctg_ids, bool* ok) {
    Results* R = res_of(c);
    *ok = false;
    if (nseq == 0 || nbases == 0 || nbases >= (1ull << 32) - 4096) return NTL_OK;
    NTL_CUDA(c, R->idx_meta.ensure((size_t)nseq * 8 + 64));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));          // the previous build may still be reading idx_meta
    uint32_t* meta = R->idx_meta.as<uint32_t>();
    memcpy(meta, h_len, (size_t)nseq * 4);
    memcpy(meta + nseq, h_rank, (size_t)nseq * 4);
    NTL_TRY(sketch_prepare(c, (uint32_t)k));
    CallState* call = nullptr;
    NTL_TRY(call_begin(c, &call));
    auto enqueue = [&]() -> int {
        NTL_TRY(sketch_device(c, d_seq, d_off, nseq, nbases, (uint32_t)k, (uint32_t)w, c->dsk, call));
        NTL_TRY(expand_contig_ids(c, c->dsk, ctg_ids));
        NTL_TRY(index_build_device(c, c->dsk.hash.as<uint64_t>(), ctg_ids.as<uint32_t>(), c->dsk.posf.as<uint32_t>(), c->dsk.n_mx, meta,
                                   meta + nseq, nseq, c->dsk.n_dev, false));
        NTL_TRY(call_note_mx(c, c->dsk.n_dev, call));
        return NTL_OK;
    };
    if (c->graph_mode) {
        cudaGraph_t graph = nullptr;
        NTL_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        c->capturing = true;
        const int rc = enqueue();
        c->capturing = false;
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        if (rc != NTL_OK) { if (graph) cudaGraphDestroy(graph); c->index.built = false; return rc; }
        if (ce == cudaSuccess && inject_graph_failure("index")) { cudaGraphDestroy(graph); graph = nullptr; ce = cudaErrorUnknown; }
        if (ce != cudaSuccess) { c->index.built = false; graph_failed(c, "graph capture", ce); return NTL_OK; }
        if (R->index_exec) {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(R->index_exec, graph, &info) != cudaSuccess) {
                cudaGetLastError();
                cudaGraphExecDestroy(R->index_exec);
                R->index_exec = nullptr;
            }
        }
        if (!R->index_exec) {
            const cudaError_t ie = cudaGraphInstantiate(&R->index_exec, graph, 0);
            if (ie != cudaSuccess) { cudaGraphDestroy(graph); R->index_exec = nullptr; c->index.built = false; graph_failed(c, "graph instantiate", ie); return NTL_OK; }
        }
        cudaGraphDestroy(graph);
        NTL_CUDA(c, cudaGraphLaunch(R->index_exec, c->stream));
        c->n_graph_launches++;
    } else {
        NTL_TRY(enqueue());
    }
    CallState hs;
    NTL_TRY(call_end(c, call, &hs));
    collect_timing(c);
    if (hs.err) { c->index.built = false; return NTL_OK; }
    c->index.n_inserted = hs.mx_total;
    c->dsk.n_mx = hs.mx_total;
    *ok = true;
    return NTL_OK;
}

static int sketch_to_host(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t nseq, int k, int w,
                          ntl_sketch_out* out, bool build_index, const uint32_t* name_rank) {
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    std::vector<uint32_t> bounds;
    // the index is built from ONE device batch: a target larger than the batch size option still goes in one piece as long as
    // it fits the 32-bit positions of a batch
    const uint64_t all_bases = nseq ? offsets[nseq] - offsets[0] : 0;
    plan_batches(offsets, nseq, build_index && all_bases < 3900000000ull ? std::max<uint64_t>(c->batch_bases, all_bases) : c->batch_bases, bounds);
    R->sk_hash.used = R->sk_posf.used = R->sk_off.used = 0;
    if (R->sk_off.reserve(((size_t)nseq + 1) * 8)) { c->err = "pinned alloc failed"; return NTL_ERR_CUDA; }
    uint64_t* seq_off = R->sk_off.at<uint64_t>(0);
    seq_off[0] = 0;
    uint64_t total = 0;
    std::vector<uint32_t> tmp_off;
    // when building the index from a multi-batch target the triples are collected on the device
    uint64_t idx_n = 0;
    if (build_index && bounds.size() > 2) {
        c->err = "target does not fit one device batch: raise the batch_bases option (max 3.9e9)";
        return NTL_ERR_ARG;
    }
    if (build_index && !out && bounds.size() == 2 && c->async_mode) {
        // index only (no sketch wanted back): copy, deferred sketch, contig ids and table build as one graph, one sync
        uint64_t nb = 0;
        NTL_TRY(stage_batch(c, seq, offsets, 0, nseq, &nb));
        std::vector<uint32_t> len(nseq);
        for (uint32_t i = 0; i < nseq; i++) len[i] = (uint32_t)(offsets[i + 1] - offsets[i]);
        bool ok = false;
        NTL_TRY(index_build_async(c, c->d_seq.as<uint8_t>(), c->d_off.as<uint64_t>(), nseq, nb, k, w, len.data(), name_rank, R->ctg_ids, &ok));
        if (ok) return NTL_OK;
    }
    for (size_t bi = 0; bi + 1 < bounds.size(); bi++) {
        const uint32_t b = bounds[bi], e = bounds[bi + 1];
        uint64_t nb = 0;
        tick(c, T_TOTAL);
        NTL_TRY(stage_batch(c, seq, offsets, b, e, &nb));
        NTL_TRY(sketch_device(c, c->d_seq.as<uint8_t>(), c->d_off.as<uint64_t>(), e - b, nb, (uint32_t)k, (uint32_t)w, c->dsk));
        const uint32_t n = c->dsk.n_mx;
        if (out) {
            R->sk_hash.used = total * 8; R->sk_posf.used = total * 4;
            if (R->sk_hash.reserve((total + n + 1) * 8) || R->sk_posf.reserve((total + n + 1) * 4)) { c->err = "pinned alloc failed"; return NTL_ERR_CUDA; }
            tmp_off.resize((size_t)(e - b) + 1);
            if (n) {
                NTL_CUDA(c, cudaMemcpyAsync(R->sk_hash.at<uint64_t>(total * 8), c->dsk.hash.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
                NTL_CUDA(c, cudaMemcpyAsync(R->sk_posf.at<uint32_t>(total * 4), c->dsk.posf.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
            }
            NTL_CUDA(c, cudaMemcpyAsync(tmp_off.data(), c->dsk.mx_off.p, ((size_t)(e - b) + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        }
        if (build_index) {
            // single batch (checked above): the index is built straight from the sketch arrays, only the contig id of
            // every minimizer has to be materialised
            NTL_TRY(expand_contig_ids(c, c->dsk, R->ctg_ids));
            idx_n = n;
        }
        tock(c, T_TOTAL);
        if (out || !build_index) NTL_TRY(finish_call(c));
        if (out) {
            for (uint32_t i = 0; i < e - b; i++) seq_off[b + i + 1] = total + tmp_off[i + 1];
        }
        total += n;
    }
    if (out) {
        out->n_mx = total; out->nseq = nseq; out->reserved = 0;
        out->hash = R->sk_hash.at<uint64_t>(0); out->pos_strand = R->sk_posf.at<uint32_t>(0); out->seq_off = seq_off;
    }
    if (build_index) {
        std::vector<uint32_t> len(nseq);
        for (uint32_t i = 0; i < nseq; i++) len[i] = (uint32_t)(offsets[i + 1] - offsets[i]);
        NTL_TRY(index_build_device(c, c->dsk.hash.as<uint64_t>(), R->ctg_ids.as<uint32_t>(), c->dsk.posf.as<uint32_t>(), idx_n,
                                   len.data(), name_rank, nseq));
        NTL_TRY(finish_call(c));
    }
    return NTL_OK;
}

int ntl_sketch(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t nseq, int k, int w, ntl_sketch_out* out) {
    if (!c || !out || (!seq && nseq) || !offsets || k <= 0 || w <= 0) { if (c) c->err = "ntl_sketch: bad argument"; return NTL_ERR_ARG; }
    return sketch_to_host(c, seq, offsets, nseq, k, w, out, false, nullptr);
}

// ------------------------------------------------------------------------------------------- index
int ntl_index_build(ntl_ctx* c, const uint64_t* hash, const uint32_t* contig, const uint32_t* pos_strand, uint64_t n,
                    const uint32_t* contig_len, const uint32_t* name_rank, uint32_t ncontig) {
    if (!c || (n && (!hash || !contig || !pos_strand)) || !contig_len || !name_rank) { if (c) c->err = "ntl_index_build: bad argument"; return NTL_ERR_ARG; }
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    NTL_CUDA(c, R->in_hash.ensure(n * 8 + 8));
    NTL_CUDA(c, R->in_posf.ensure(n * 4 + 4));
    NTL_CUDA(c, R->in_ctg.ensure(n * 4 + 4));
    if (n) {
        NTL_CUDA(c, cudaMemcpyAsync(R->in_hash.p, hash, n * 8, cudaMemcpyHostToDevice, c->stream));
        NTL_CUDA(c, cudaMemcpyAsync(R->in_posf.p, pos_strand, n * 4, cudaMemcpyHostToDevice, c->stream));
        NTL_CUDA(c, cudaMemcpyAsync(R->in_ctg.p, contig, n * 4, cudaMemcpyHostToDevice, c->stream));
    }
    NTL_TRY(index_build_device(c, R->in_hash.as<uint64_t>(), R->in_ctg.as<uint32_t>(), R->in_posf.as<uint32_t>(), n, contig_len,
                               name_rank, ncontig));
    return finish_call(c);
}

int ntl_index_build_device(ntl_ctx* c, const void* d_hash, const void* d_contig, const void* d_pos_strand, uint64_t n,
                           const uint32_t* contig_len, const uint32_t* name_rank, uint32_t ncontig) {
    if (!c || !contig_len || !name_rank) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    NTL_TRY(index_build_device(c, (const uint64_t*)d_hash, (const uint32_t*)d_contig, (const uint32_t*)d_pos_strand, n,
                               contig_len, name_rank, ncontig));
    return finish_call(c);
}

int ntl_index_build_from_sequences(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t ncontig, int k, int w,
                                   const uint32_t* name_rank, ntl_sketch_out* sketch_out) {
    if (!c || (!seq && ncontig) || !offsets || !name_rank || k <= 0 || w <= 0) { if (c) c->err = "ntl_index_build_from_sequences: bad argument"; return NTL_ERR_ARG; }
    return sketch_to_host(c, seq, offsets, ncontig, k, w, sketch_out, true, name_rank);
}

int ntl_index_stats(ntl_ctx* c, uint64_t* n_inserted, uint64_t* n_unique, uint64_t* table_slots) {
    if (!c || !c->index.built) { if (c) c->err = "no index"; return NTL_ERR_STATE; }
    cudaSetDevice(c->device);
    unsigned long long u = 0;
    NTL_CUDA(c, cudaMemcpy(&u, (char*)c->index.special.p + sizeof(IdxSpecial), 8, cudaMemcpyDeviceToHost));
    IdxSpecial sp;
    NTL_CUDA(c, cudaMemcpy(&sp, c->index.special.p, sizeof sp, cudaMemcpyDeviceToHost));
    if (n_inserted) *n_inserted = c->index.n_inserted;
    if (n_unique) *n_unique = u + (sp.count == 1 ? 1 : 0);
    if (table_slots) *table_slots = c->index.slots;
    return NTL_OK;
}

int ntl_device_sketch_arrays(ntl_ctx* c, uint64_t* n_mx, void** d_hash, void** d_pos_strand, void** d_seq_off) {
    if (!c) return NTL_ERR_ARG;
    if (n_mx) *n_mx = c->dsk.n_mx;
    if (d_hash) *d_hash = c->dsk.hash.p;
    if (d_pos_strand) *d_pos_strand = c->dsk.posf.p;
    if (d_seq_off) *d_seq_off = c->dsk.mx_off.p;
    return NTL_OK;
}

// ------------------------------------------------------------------------------------------- mapping
static int append_map_results(ntl_ctx* c, uint32_t rb, uint32_t nreads, const MapStatus& cs, uint64_t log_base,
                              uint64_t& hits_total, uint64_t& ev_total) {
    Results* R = res_of(c);
    MapWork& M = c->mw;
    const uint32_t nh = cs.n_hits, ne = cs.n_events;
    if (hits_total + nh >= (1ull << 32) || ev_total + ne >= (1ull << 32)) { c->err = "more than 4G hits in one call: split the call"; return NTL_ERR_ARG; }
    R->runs.used = hits_total * sizeof(ntl_run); R->hits.used = hits_total * sizeof(ntl_hit);
    R->events.used = ev_total * sizeof(ntl_event);
    if (R->runs.reserve((hits_total + nh + 1) * sizeof(ntl_run)) || R->hits.reserve((hits_total + nh + 1) * sizeof(ntl_hit)) ||
        R->events.reserve((ev_total + ne + 1) * sizeof(ntl_event))) { c->err = "pinned alloc failed"; return NTL_ERR_CUDA; }
    uint32_t* hit_off = R->hit_off.at<uint32_t>(0) + rb;
    uint32_t* nruns = R->nruns.at<uint32_t>(0) + rb;
    uint32_t* ev_off = R->ev_off.at<uint32_t>(0) + rb;
    uint32_t* ev_cnt = R->ev_cnt.at<uint32_t>(0) + rb;
    const uint32_t* d_evblock = M.ev_cnt.as<uint32_t>();
    const uint32_t* d_ev_cnt = d_evblock + 2 * ((size_t)nreads + 2);
    const uint32_t* d_ev_pref = d_evblock + 3 * ((size_t)nreads + 2);
    NTL_CUDA(c, cudaMemcpyAsync(hit_off, M.hit_off.p, ((size_t)nreads + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(nruns, M.nruns.p, (size_t)nreads * 4, cudaMemcpyDeviceToHost, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(ev_off, d_ev_pref, ((size_t)nreads + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(ev_cnt, d_ev_cnt, (size_t)nreads * 4, cudaMemcpyDeviceToHost, c->stream));
    if (nh) {
        NTL_CUDA(c, cudaMemcpyAsync(R->runs.at<ntl_run>(hits_total * sizeof(ntl_run)), M.runs.p, (size_t)nh * sizeof(ntl_run), cudaMemcpyDeviceToHost, c->stream));
        NTL_CUDA(c, cudaMemcpyAsync(R->hits.at<ntl_hit>(hits_total * sizeof(ntl_hit)), M.hits.p, (size_t)nh * sizeof(ntl_hit), cudaMemcpyDeviceToHost, c->stream));
    }
    if (ne)
        NTL_CUDA(c, cudaMemcpyAsync(R->events.at<ntl_event>(ev_total * sizeof(ntl_event)), c->tl_events.as<Event>() + log_base,
                                    (size_t)ne * sizeof(ntl_event), cudaMemcpyDeviceToHost, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    if (hits_total || ev_total) {
        for (uint32_t i = 0; i <= nreads; i++) { hit_off[i] += (uint32_t)hits_total; ev_off[i] += (uint32_t)ev_total; }
    }
    hits_total += nh; ev_total += ne;
    return NTL_OK;
}

static int begin_map_results(ntl_ctx* c, uint32_t nreads) {
    Results* R = res_of(c);
    HostVec* v[] = {&R->hit_off, &R->nruns, &R->ev_off, &R->ev_cnt};
    for (HostVec* h : v) { h->used = 0; if (h->reserve(((size_t)nreads + 2) * 4)) { c->err = "pinned alloc failed"; return NTL_ERR_CUDA; } }
    R->runs.used = R->hits.used = R->events.used = 0;
    R->hit_off.at<uint32_t>(0)[0] = 0; R->ev_off.at<uint32_t>(0)[0] = 0;
    return NTL_OK;
}

static void fill_map_out(ntl_ctx* c, ntl_map_out* out, uint32_t nreads, uint64_t n_mx, uint64_t hits, uint64_t runs, uint64_t events) {
    Results* R = res_of(c);
    out->n_reads = nreads; out->reserved = 0; out->n_mx = n_mx; out->n_hits = hits; out->n_runs = runs; out->n_events = events;
    out->hit_off = R->hit_off.at<uint32_t>(0); out->nruns = R->nruns.at<uint32_t>(0);
    out->runs = R->runs.at<ntl_run>(0); out->hits = R->hits.at<ntl_hit>(0);
    out->ev_off = R->ev_off.at<uint32_t>(0); out->ev_cnt = R->ev_cnt.at<uint32_t>(0); out->events = R->events.at<ntl_event>(0);
}

// Sync-free form of ntl_map_reads: every chunk's copy, sketch, mapping and result write-back is enqueued without waiting
// for the device; ONE synchronisation ends the call. Capacities come from bounds and from what earlier calls needed; if
// any of them turns out too small (or a reference assertion fails) *ok is false, nothing has been committed, and the
// caller repeats the call on the synchronous path, which sizes everything exactly and reports errors precisely.
static int map_reads_async(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t nreads, uint64_t first_read_ordinal,
                           const ntl_params* prm, ntl_map_out* out, bool* ok) {
    Results* R = res_of(c);
    *ok = false;
    const double host_tin = host_now_us();
    const uint64_t total_bases = offsets[nreads] - offsets[0];
    // chunks: about pipeline_min_bases / 4 each (20 Mbp by default: no per-chunk round trip to amortise any more, and the
    // part of the call that cannot overlap the copy is the last chunk's compute), at most batch_bases
    std::vector<uint32_t> bounds;
    const uint64_t target = std::max<uint64_t>(1ull << 20, c->pipeline_min_bases / 4);
    const uint64_t nb_target = std::max<uint64_t>(1, (total_bases + target / 2) / target);
    uint64_t per_chunk = (total_bases + nb_target - 1) / nb_target + (1u << 18);
    if (per_chunk > c->batch_bases) per_chunk = c->batch_bases;
    plan_batches(offsets, nreads, per_chunk, bounds);
    const size_t nch = bounds.size() - 1;
    uint64_t max_nb = 0; uint32_t max_ns = 0, mx_bound_total = 0;
    for (size_t i = 0; i < nch; i++) {
        const uint64_t nb = offsets[bounds[i + 1]] - offsets[bounds[i]];
        if (nb >= (1ull << 32) - 4096) return NTL_OK;          // let the synchronous path report it
        max_nb = std::max(max_nb, nb); max_ns = std::max(max_ns, bounds[i + 1] - bounds[i]);
        const uint64_t b = sketch_out_bound(nb, bounds[i + 1] - bounds[i], (uint32_t)prm->w, c->mx_density_factor);
        if ((uint64_t)mx_bound_total + b >= (1ull << 31)) return NTL_OK;
        mx_bound_total += (uint32_t)b;
    }
    // host result arrays: hits <= minimizers; keep the pinned arrays moderate with the hit rate seen so far
    uint64_t hits_cap = mx_bound_total;
    if (hits_cap > (4u << 20)) {
        const double rate = R->last_call_bases ? (double)R->last_call_hits / (double)R->last_call_bases : 0.01;
        hits_cap = std::min<uint64_t>(hits_cap, std::max<uint64_t>(4u << 20, (uint64_t)(3.0 * rate * (double)total_bases) + (1u << 20)));
    }
    const uint64_t ev_cap_total = std::max<uint64_t>((uint64_t)std::max<uint32_t>(c->ev_cap_hint, 1u << 16) * 2, 8ull * nreads);
    NTL_TRY(begin_map_results(c, nreads));
    if (R->runs.reserve((hits_cap + 1) * sizeof(ntl_run)) || R->hits.reserve((hits_cap + 1) * sizeof(ntl_hit)) ||
        R->events.reserve((ev_cap_total + 1) * sizeof(ntl_event))) { c->err = "pinned alloc failed"; return NTL_ERR_CUDA; }
    HostResults H;
    H.hit_off = R->hit_off.at<uint32_t>(0); H.nruns = R->nruns.at<uint32_t>(0);
    H.ev_off = R->ev_off.at<uint32_t>(0); H.ev_cnt = R->ev_cnt.at<uint32_t>(0);
    H.runs = R->runs.at<Run>(0); H.hits = R->hits.at<Hit>(0); H.events = R->events.at<Event>(0);
    H.hits_cap = (uint32_t)std::min<uint64_t>(hits_cap, 0xFFFFFFF0u); H.ev_cap = (uint32_t)std::min<uint64_t>(ev_cap_total, 0xFFFFFFF0u);
    // staging: two device slots, the rebased offsets of ALL chunks in one pinned array (nothing is reused while in flight)
    for (int sl = 0; sl < 2 && (size_t)sl < nch; sl++) {
        NTL_CUDA(c, R->st_seq[sl].ensure(max_nb + 256));
        NTL_CUDA(c, R->st_off[sl].ensure(((size_t)max_ns + 1) * 8));
    }
    NTL_CUDA(c, R->call_hoff.ensure(((size_t)nreads + nch + 1) * 8));
    NTL_CUDA(c, R->read_len.ensure(((size_t)max_ns + 1) * 4));
    uint64_t* ho_all = R->call_hoff.as<uint64_t>();
    std::vector<size_t> ho_at(nch);
    {
        size_t at = 0;
        for (size_t i = 0; i < nch; i++) {
            ho_at[i] = at;
            const uint64_t base = offsets[bounds[i]];
            for (uint32_t r = bounds[i]; r <= bounds[i + 1]; r++) ho_all[at++] = offsets[r] - base;
        }
    }
    // Pageable caller memory (numpy arrays from the file reader): the driver's own staging of such copies runs at
    // ~3 GB/s, so the chunk is moved into a pinned bounce buffer by a few host threads and sent from there.
    bool pageable = true;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, seq) == cudaSuccess) pageable = (attr.type == cudaMemoryTypeUnregistered);
        else cudaGetLastError();
    }
    const int copy_threads = c->copy_threads >= 0 ? c->copy_threads : (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    if (copy_threads == 0) pageable = false;                   // option "copy_threads" = 0: leave the staging to the driver
    if (pageable) for (int sl = 0; sl < 2 && (size_t)sl < nch; sl++) NTL_CUDA(c, R->pin_slot[sl].ensure(max_nb + 256));
    auto enqueue_copy = [&](size_t i) -> int {
        const int sl = (int)(i & 1);
        const uint64_t base = offsets[bounds[i]], nb = offsets[bounds[i + 1]] - base;
        const uint32_t ns = bounds[i + 1] - bounds[i];
        const char* src = seq + base;
        if (pageable && nb) {
            if (i >= 2) NTL_CUDA(c, cudaEventSynchronize(R->h2d_done[sl]));                   // the bounce buffer's previous chunk has left
            parallel_memcpy(R->pin_slot[sl].as<char>(), src, nb, copy_threads);
            src = R->pin_slot[sl].as<char>();
        }
        if (i >= 2) NTL_CUDA(c, cudaStreamWaitEvent(R->copy_stream, R->comp_done[sl], 0));   // slot's previous chunk is done with it
        if (nb) NTL_CUDA(c, cudaMemcpyAsync(R->st_seq[sl].p, src, nb, cudaMemcpyHostToDevice, R->copy_stream));
        NTL_CUDA(c, cudaMemcpyAsync(R->st_off[sl].p, ho_all + ho_at[i], ((size_t)ns + 1) * 8, cudaMemcpyHostToDevice, R->copy_stream));
        NTL_CUDA(c, cudaEventRecord(R->h2d_done[sl], R->copy_stream));
        return NTL_OK;
    };
    CallState* call = nullptr;
    NTL_TRY(call_begin(c, &call));
    // NTL_TRACE: device timeline of the call (copy / compute end of every chunk, relative to the start of the call)
    const bool trace = getenv("NTL_TRACE") != nullptr;
    const double host_t0 = host_now_us();
    std::vector<cudaEvent_t> tr_copy, tr_comp;
    cudaEvent_t tr_start = nullptr;
    if (trace) {
        cudaEventCreate(&tr_start); cudaEventRecord(tr_start, c->stream);
        cudaStreamWaitEvent(R->copy_stream, tr_start, 0);
        tr_copy.resize(nch); tr_comp.resize(nch);
        for (size_t i = 0; i < nch; i++) { cudaEventCreate(&tr_copy[i]); cudaEventCreate(&tr_comp[i]); }
    }
    if (nch) { NTL_TRY(enqueue_copy(0)); if (trace) cudaEventRecord(tr_copy[0], R->copy_stream); }
    const bool use_graph = c->graph_mode != 0;
    // executable graphs are kept across calls (one per chunk position) and updated in place: the topology of a chunk's
    // graph never changes, only kernel arguments and grid sizes do
    std::vector<cudaGraphExec_t>& execs = R->chunk_execs;
    NTL_TRY(sketch_prepare(c, (uint32_t)prm->k));
    tick(c, T_TOTAL);
    for (size_t i = 0; i < nch; i++) {
        const int sl = (int)(i & 1);
        const uint32_t b = bounds[i], ns = bounds[i + 1] - b;
        const uint64_t nb = offsets[bounds[i + 1]] - offsets[b];
        // pinned source: queue the next copy first (it is asynchronous); pageable source: the host-side bounce copy of
        // the next chunk would delay this chunk's launch, so it comes after it
        if (!pageable && i + 1 < nch) { NTL_TRY(enqueue_copy(i + 1)); if (trace) cudaEventRecord(tr_copy[i + 1], R->copy_stream); }
        if (use_graph) {
            // The ~60 launches of a chunk go into ONE CUDA graph: while a host->device copy is in flight every single
            // kernel launch costs several microseconds more (its launch data is fetched over the same PCIe link), a
            // graph launch pays that once. Capture + instantiation run on the host while the copy is still under way.
            NTL_TRY(call_reserve_events(c, ns));
            const double h0 = trace ? host_now_us() : 0;
            cudaGraph_t graph = nullptr;
            NTL_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            c->capturing = true; c->no_stage_timing = true;
            int rc = sketch_device(c, R->st_seq[sl].as<uint8_t>(), R->st_off[sl].as<uint64_t>(), ns, nb, (uint32_t)prm->k, (uint32_t)prm->w,
                                   c->dsk, call);
            if (rc == NTL_OK) rc = read_len_device(c, R->st_off[sl].as<uint64_t>(), ns, R->read_len);
            if (rc == NTL_OK) rc = map_device(c, c->dsk, R->read_len.as<uint32_t>(), ns, first_read_ordinal + b, prm, nullptr, nullptr, nullptr, call);
            if (rc == NTL_OK) rc = call_chunk_finish(c, call, b, ns, &H);
            c->capturing = false; c->no_stage_timing = false;
            cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
            const double h1 = trace ? host_now_us() : 0;
            if (rc != NTL_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce == cudaSuccess && inject_graph_failure("map") && i == 1) { cudaGraphDestroy(graph); graph = nullptr; ce = cudaErrorUnknown; }
            if (ce != cudaSuccess) { graph_failed(c, "graph capture", ce); return NTL_OK; }      // *ok is false: synchronous path
            cudaGraphExec_t exec = i < execs.size() ? execs[i] : nullptr;
            if (exec) {
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(exec, graph, &info) != cudaSuccess) {
                    cudaGetLastError();
                    cudaGraphExecDestroy(exec);
                    exec = nullptr;
                }
            }
            if (!exec) {
                const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
                if (ie != cudaSuccess) { cudaGraphDestroy(graph); if (i < execs.size()) execs[i] = nullptr; graph_failed(c, "graph instantiate", ie); return NTL_OK; }
            }
            cudaGraphDestroy(graph);
            if (i < execs.size()) execs[i] = exec; else execs.push_back(exec);
            NTL_CUDA(c, cudaStreamWaitEvent(c->stream, R->h2d_done[sl], 0));
            const double h2 = trace ? host_now_us() : 0;
            NTL_CUDA(c, cudaGraphLaunch(exec, c->stream));
            c->n_graph_launches++;
            if (trace) fprintf(stderr, "[ntl] chunk %zu host: capture %.0f us, instantiate %.0f us, launch %.0f us (at %.0f us)\n", i, h1 - h0, h2 - h1,
                               host_now_us() - h2, host_now_us() - host_t0);
        } else {
        NTL_CUDA(c, cudaStreamWaitEvent(c->stream, R->h2d_done[sl], 0));
        NTL_TRY(sketch_device(c, R->st_seq[sl].as<uint8_t>(), R->st_off[sl].as<uint64_t>(), ns, nb, (uint32_t)prm->k, (uint32_t)prm->w,
                              c->dsk, call));
        NTL_TRY(read_len_device(c, R->st_off[sl].as<uint64_t>(), ns, R->read_len));
        NTL_TRY(map_device(c, c->dsk, R->read_len.as<uint32_t>(), ns, first_read_ordinal + b, prm, nullptr, nullptr, nullptr, call));
        NTL_TRY(call_chunk_finish(c, call, b, ns, &H));
        }
        NTL_CUDA(c, cudaEventRecord(R->comp_done[sl], c->stream));
        if (trace) cudaEventRecord(tr_comp[i], c->stream);
        if (pageable && i + 1 < nch) { NTL_TRY(enqueue_copy(i + 1)); if (trace) cudaEventRecord(tr_copy[i + 1], R->copy_stream); }
    }
    tock(c, T_TOTAL);
    CallState hs;
    NTL_TRY(call_end(c, call, &hs));
    collect_timing(c);
    if (trace) {
        fprintf(stderr, "[ntl] call synchronised at %.0f us (host clock; %.0f us of set-up before it)\n", host_now_us() - host_t0, host_t0 - host_tin);
        cudaStreamSynchronize(R->copy_stream);
        for (size_t i = 0; i < nch; i++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tr_start, tr_copy[i]); cudaEventElapsedTime(&b, tr_start, tr_comp[i]);
            fprintf(stderr, "[ntl] chunk %zu (%u reads): copy done %7.3f ms, compute done %7.3f ms\n", i, bounds[i + 1] - bounds[i], a, b);
            cudaEventDestroy(tr_copy[i]); cudaEventDestroy(tr_comp[i]);
        }
        cudaEventDestroy(tr_start);
    }
    if (hs.err) return NTL_OK;                                  // *ok stays false
    R->last_call_hits = hs.hits_total; R->last_call_bases = total_bases;
    note_mx_density(c, hs.mx_total, total_bases, (uint32_t)prm->w);
    R->runs.used = (size_t)hs.hits_total * sizeof(ntl_run); R->hits.used = (size_t)hs.hits_total * sizeof(ntl_hit);
    R->events.used = (size_t)hs.ev_total * sizeof(ntl_event);
    fill_map_out(c, out, nreads, hs.mx_total, hs.hits_total, hs.runs_total, hs.ev_total);
    c->dsk.n_mx = nch == 1 ? hs.mx_total : 0;                  // the bound is of no use to anyone after the call
    *ok = true;
    return NTL_OK;
}

int ntl_map_reads(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t nreads, uint64_t first_read_ordinal,
                  const ntl_params* prm, ntl_map_out* out) {
    if (!c || (!seq && nreads) || !offsets || !prm || !out || prm->k <= 0 || prm->w <= 0) { if (c) c->err = "ntl_map_reads: bad argument"; return NTL_ERR_ARG; }
    if (!c->index.built) { c->err = "ntl_map_reads: no target index"; return NTL_ERR_STATE; }
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    if (c->async_mode && nreads) {
        bool ok = false;
        NTL_TRY(map_reads_async(c, seq, offsets, nreads, first_read_ordinal, prm, out, &ok));
        c->n_async_calls++;
        if (ok) return NTL_OK;
        c->n_async_fallbacks++;
        if (getenv("NTL_TRACE")) fprintf(stderr, "[ntl] sync-free call fell back to the synchronous path\n");
    }
    NTL_TRY(begin_map_results(c, nreads));
    // Pipelined: the host->device copy of batch i+1 (copy stream, double-buffered staging) overlaps the kernels of
    // batch i (compute stream). Batches of about a quarter of the call so that the overlap pays off.
    std::vector<uint32_t> bounds;
    const uint64_t total_bases = offsets[nreads] - offsets[0];
    // balanced batches of about pipeline_min_bases each (measured: with less than ~75 Mbp per batch the fixed
    // per-batch cost outweighs the copy/compute overlap)
    const uint64_t nb_target = std::max<uint64_t>(1, (total_bases + c->pipeline_min_bases / 2) / c->pipeline_min_bases);
    uint64_t per_batch = (total_bases + nb_target - 1) / nb_target + (1u << 20);
    if (per_batch > c->batch_bases) per_batch = c->batch_bases;
    plan_batches(offsets, nreads, per_batch, bounds);
    const size_t nbat = bounds.size() - 1;
    uint64_t hits_total = 0, ev_total = 0, mx_total = 0, runs_total = 0;
    uint64_t nb_slot[2] = {0, 0};
    if (nbat) NTL_TRY(stage_batch_async(c, 0, seq, offsets, bounds[0], bounds[1], &nb_slot[0]));
    const bool trace = getenv("NTL_TRACE") != nullptr;
    auto now_us = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = now_us();
    auto lap = [&](const char* what, size_t bi) { if (trace) { double t = now_us(); fprintf(stderr, "[ntl] batch %zu %-8s %8.1f us\n", bi, what, t - t_prev); t_prev = t; } };
    for (size_t bi = 0; bi < nbat; bi++) {
        const uint32_t b = bounds[bi], e = bounds[bi + 1];
        const int slot = (int)(bi & 1);
        if (bi + 1 < nbat) NTL_TRY(stage_batch_async(c, slot ^ 1, seq, offsets, bounds[bi + 1], bounds[bi + 2], &nb_slot[slot ^ 1]));
        lap("stage", bi);
        NTL_CUDA(c, cudaStreamWaitEvent(c->stream, R->h2d_done[slot], 0));
        tick(c, T_TOTAL);
        NTL_TRY(sketch_device(c, R->st_seq[slot].as<uint8_t>(), R->st_off[slot].as<uint64_t>(), e - b, nb_slot[slot], (uint32_t)prm->k,
                              (uint32_t)prm->w, c->dsk));
        lap("sketch", bi);
        NTL_TRY(read_len_device(c, R->st_off[slot].as<uint64_t>(), e - b, R->read_len));
        MapStatus cs; uint64_t log_base = 0;
        NTL_TRY(map_device(c, c->dsk, R->read_len.as<uint32_t>(), e - b, first_read_ordinal + b, prm, &cs, &log_base));
        tock(c, T_TOTAL);
        lap("map", bi);
        NTL_TRY(append_map_results(c, b, e - b, cs, log_base, hits_total, ev_total));
        collect_timing(c);
        lap("results", bi);
        mx_total += c->dsk.n_mx; runs_total += cs.n_runs;
    }
    fill_map_out(c, out, nreads, mx_total, hits_total, runs_total, ev_total);
    R->last_call_hits = hits_total; R->last_call_bases = total_bases;      // what the sync-free path sizes its host arrays from
    return NTL_OK;
}

int ntl_map_sketch(ntl_ctx* c, const uint64_t* hash, const uint32_t* pos_strand, const uint64_t* mx_off,
                   const uint32_t* read_len, uint32_t nreads, uint64_t first_read_ordinal, const ntl_params* prm,
                   ntl_map_out* out) {
    if (!c || !mx_off || !read_len || !prm || !out || prm->k <= 0) { if (c) c->err = "ntl_map_sketch: bad argument"; return NTL_ERR_ARG; }
    if (!c->index.built) { c->err = "ntl_map_sketch: no target index"; return NTL_ERR_STATE; }
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    NTL_TRY(begin_map_results(c, nreads));
    // batches of at most 2^28 minimizers
    uint64_t hits_total = 0, ev_total = 0, runs_total = 0;
    uint32_t b = 0;
    std::vector<uint32_t> off32;
    if (nreads == 0) { fill_map_out(c, out, 0, 0, 0, 0, 0); return NTL_OK; }
    while (b < nreads) {
        uint32_t e = b + 1;
        while (e < nreads && mx_off[e + 1] - mx_off[b] <= (1ull << 28)) e++;
        const uint64_t m0 = mx_off[b], n = mx_off[e] - m0;
        if (n >= (1ull << 31)) { c->err = "a single read has too many minimizers"; return NTL_ERR_ARG; }
        const uint32_t nr = e - b;
        off32.resize((size_t)nr + 1);
        for (uint32_t i = 0; i <= nr; i++) off32[i] = (uint32_t)(mx_off[b + i] - m0);
        DeviceSketch& sk = c->dsk;
        NTL_CUDA(c, sk.hash.ensure(n * 8 + 8));
        NTL_CUDA(c, sk.posf.ensure(n * 4 + 4));
        NTL_CUDA(c, sk.mx_off.ensure(((size_t)nr + 1) * 4));
        NTL_CUDA(c, R->read_len.ensure(((size_t)nr + 1) * 4));
        tick(c, T_TOTAL);
        if (n) {
            NTL_CUDA(c, cudaMemcpyAsync(sk.hash.p, hash + m0, n * 8, cudaMemcpyHostToDevice, c->stream));
            NTL_CUDA(c, cudaMemcpyAsync(sk.posf.p, pos_strand + m0, n * 4, cudaMemcpyHostToDevice, c->stream));
        }
        NTL_CUDA(c, cudaMemcpyAsync(sk.mx_off.p, off32.data(), ((size_t)nr + 1) * 4, cudaMemcpyHostToDevice, c->stream));
        if (nr) NTL_CUDA(c, cudaMemcpyAsync(R->read_len.p, read_len + b, (size_t)nr * 4, cudaMemcpyHostToDevice, c->stream));
        sk.n_mx = (uint32_t)n; sk.nseq = nr; sk.n_dev = nullptr;
        MapStatus cs; uint64_t log_base = 0;
        NTL_TRY(map_device(c, sk, R->read_len.as<uint32_t>(), nr, first_read_ordinal + b, prm, &cs, &log_base));
        tock(c, T_TOTAL);
        NTL_TRY(append_map_results(c, b, nr, cs, log_base, hits_total, ev_total));
        collect_timing(c);
        runs_total += cs.n_runs;
        b = e;
    }
    fill_map_out(c, out, nreads, mx_off[nreads] - mx_off[0], hits_total, runs_total, ev_total);
    return NTL_OK;
}

int ntl_tally_mappings(ntl_ctx* c, const uint32_t* hit_off, const uint32_t* nruns, const ntl_run* runs, const ntl_hit* hits,
                       const uint32_t* read_len, uint32_t nreads, uint64_t first_read_ordinal, const ntl_params* prm,
                       uint64_t* n_events_out) {
    if (!c || !prm || prm->k <= 0) { if (c) c->err = "ntl_tally_mappings: bad argument"; return NTL_ERR_ARG; }
    if (!c->index.built) { c->err = "ntl_tally_mappings: contig lengths / name ranks missing (build or load an index first)"; return NTL_ERR_STATE; }
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    PreMappings pre{};
    if (!hit_off) {                       // mappings left on the device by ntl_liftover_mappings
        if (!c->mw.lifted_valid || c->mw.lifted_reads != nreads) { c->err = "ntl_tally_mappings: no lifted mappings of that size on the device"; return NTL_ERR_STATE; }
        pre.resident = true; pre.n_hits = c->mw.lifted_hits;
    } else {
        if (!nruns) { c->err = "ntl_tally_mappings: bad argument"; return NTL_ERR_ARG; }
        const uint32_t nh = nreads ? hit_off[nreads] : 0;
        if (nh && (!runs || !hits)) { c->err = "ntl_tally_mappings: bad argument"; return NTL_ERR_ARG; }
        pre.hit_off = hit_off; pre.nruns = nruns; pre.runs = reinterpret_cast<const Run*>(runs);
        pre.hits = reinterpret_cast<const Hit*>(hits); pre.n_hits = nh;
    }
    c->mw.lifted_valid = false;
    pre.compute_read_len = (read_len == nullptr);
    NTL_CUDA(c, R->read_len.ensure(((size_t)nreads + 1) * 4));
    if (nreads && read_len) NTL_CUDA(c, cudaMemcpyAsync(R->read_len.p, read_len, (size_t)nreads * 4, cudaMemcpyHostToDevice, c->stream));
    MapStatus cs; uint64_t log_base = 0;
    NTL_TRY(map_device(c, c->dsk, R->read_len.as<uint32_t>(), nreads, first_read_ordinal, prm, &cs, &log_base, &pre));
    NTL_TRY(finish_call(c));
    if (n_events_out) *n_events_out = cs.n_events;
    return NTL_OK;
}

int ntl_map_groups(ntl_ctx* c, const uint64_t* t_hash, const uint32_t* t_pos_strand, const uint64_t* t_mx_off, const uint32_t* t_len,
                   uint32_t ntargets, const uint32_t* group_t_off, const uint64_t* r_hash, const uint32_t* r_pos_strand,
                   const uint64_t* r_mx_off, const uint32_t* r_len, uint32_t ngroups, const ntl_params* prm, ntl_map_out* out) {
    if (!c || !t_mx_off || !group_t_off || !r_mx_off || !prm || !out || prm->k <= 0 || (ntargets && !t_len) || (ngroups && !r_len)) {
        if (c) c->err = "ntl_map_groups: bad argument";
        return NTL_ERR_ARG;
    }
    const uint64_t nt = ntargets ? t_mx_off[ntargets] - t_mx_off[0] : 0, nr = ngroups ? r_mx_off[ngroups] - r_mx_off[0] : 0;
    if ((nt && (!t_hash || !t_pos_strand)) || (nr && (!r_hash || !r_pos_strand))) { c->err = "ntl_map_groups: bad argument"; return NTL_ERR_ARG; }
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    MapStatus cs;
    NTL_TRY(group_map_device(c, t_hash, t_pos_strand, t_mx_off, t_len, ntargets, group_t_off, r_hash, r_pos_strand, r_mx_off, r_len, ngroups,
                             prm, &cs));
    MapWork& M = c->mw;
    NTL_TRY(begin_map_results(c, ngroups));
    if (R->runs.reserve(((size_t)nr + 1) * sizeof(ntl_run)) || R->hits.reserve(((size_t)nr + 1) * sizeof(ntl_hit))) { c->err = "pinned alloc failed"; return NTL_ERR_CUDA; }
    memset(R->ev_off.at<uint32_t>(0), 0, ((size_t)ngroups + 1) * 4);
    memset(R->ev_cnt.at<uint32_t>(0), 0, ((size_t)ngroups + 1) * 4);
    NTL_CUDA(c, cudaMemcpyAsync(R->hit_off.at<uint32_t>(0), M.hit_off.p, ((size_t)ngroups + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    if (ngroups) NTL_CUDA(c, cudaMemcpyAsync(R->nruns.at<uint32_t>(0), M.nruns.p, (size_t)ngroups * 4, cudaMemcpyDeviceToHost, c->stream));
    if (nr) {
        NTL_CUDA(c, cudaMemcpyAsync(R->runs.at<ntl_run>(0), M.runs.p, (size_t)nr * sizeof(ntl_run), cudaMemcpyDeviceToHost, c->stream));
        NTL_CUDA(c, cudaMemcpyAsync(R->hits.at<ntl_hit>(0), M.hits.p, (size_t)nr * sizeof(ntl_hit), cudaMemcpyDeviceToHost, c->stream));
    }
    NTL_TRY(finish_call(c));
    fill_map_out(c, out, ngroups, nr, nr, cs.n_runs, 0);
    out->n_hits = cs.n_hits;
    return NTL_OK;
}

int ntl_liftover_mappings(ntl_ctx* c, const uint32_t* hit_off, const uint32_t* nruns, const ntl_run* runs, const ntl_hit* hits,
                          uint32_t nreads, const ntl_agp_row* agp, uint32_t ncontig, int k, ntl_map_out* out) {
    if (!c || !hit_off || (nreads && !nruns) || (ncontig && !agp) || k <= 0) { if (c) c->err = "ntl_liftover_mappings: bad argument"; return NTL_ERR_ARG; }
    const uint32_t nh = nreads ? hit_off[nreads] : 0;
    if (nh && (!runs || !hits)) { c->err = "ntl_liftover_mappings: bad argument"; return NTL_ERR_ARG; }
    for (uint32_t r = 0; r < nreads; r++)
        if (hit_off[r + 1] < hit_off[r]) { c->err = "ntl_liftover_mappings: hit_off must be non-decreasing"; return NTL_ERR_ARG; }
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    static_assert(sizeof(ntl_agp_row) == sizeof(AgpRow), "ntl_agp_row layout");
    MapStatus cs;
    NTL_TRY(liftover_device(c, hit_off, nruns, reinterpret_cast<const Run*>(runs), reinterpret_cast<const Hit*>(hits), nreads,
                            reinterpret_cast<const AgpRow*>(agp), ncontig, k, &cs));
    if (out) {
        MapWork& M = c->mw;
        NTL_TRY(begin_map_results(c, nreads));
        if (R->runs.reserve(((size_t)nh + 1) * sizeof(ntl_run)) || R->hits.reserve(((size_t)nh + 1) * sizeof(ntl_hit))) { c->err = "pinned alloc failed"; return NTL_ERR_CUDA; }
        memcpy(R->hit_off.at<uint32_t>(0), hit_off, ((size_t)nreads + 1) * 4);      // the regions do not move
        memset(R->ev_off.at<uint32_t>(0), 0, ((size_t)nreads + 1) * 4);
        memset(R->ev_cnt.at<uint32_t>(0), 0, ((size_t)nreads + 1) * 4);
        if (nreads) NTL_CUDA(c, cudaMemcpyAsync(R->nruns.at<uint32_t>(0), M.nruns.p, (size_t)nreads * 4, cudaMemcpyDeviceToHost, c->stream));
        if (nh) {
            NTL_CUDA(c, cudaMemcpyAsync(R->runs.at<ntl_run>(0), M.runs.p, (size_t)nh * sizeof(ntl_run), cudaMemcpyDeviceToHost, c->stream));
            NTL_CUDA(c, cudaMemcpyAsync(R->hits.at<ntl_hit>(0), M.hits.p, (size_t)nh * sizeof(ntl_hit), cudaMemcpyDeviceToHost, c->stream));
        }
        NTL_CUDA(c, cudaStreamSynchronize(c->stream));
        fill_map_out(c, out, nreads, 0, nh, cs.n_runs, 0);
        out->n_hits = cs.n_hits;          // hits that survived (the arrays keep their holey size hit_off[nreads])
    }
    return finish_call(c);
}

// ------------------------------------------------------------------------------------------- pairs
int ntl_events_reset(ntl_ctx* c) { if (!c) return NTL_ERR_ARG; c->tl_n_events = 0; c->tl_count_on_device = false; return NTL_OK; }
int ntl_events_count(ntl_ctx* c, uint64_t* n) { if (!c || !n) return NTL_ERR_ARG; NTL_TRY(events_resolve_count(c)); *n = c->tl_n_events; return NTL_OK; }
int ntl_events_device(ntl_ctx* c, uint64_t* n, void** d_events) {
    if (!c) return NTL_ERR_ARG;
    NTL_TRY(events_resolve_count(c));
    if (n) *n = c->tl_n_events;
    if (d_events) *d_events = c->tl_events.p;
    return NTL_OK;
}

static int events_append_impl(ntl_ctx* c, const void* src, uint64_t n, cudaMemcpyKind kind) {
    cudaSetDevice(c->device);
    if (!n) return NTL_OK;
    NTL_TRY(events_resolve_count(c));
    const size_t keep = c->tl_n_events * sizeof(Event), need = (c->tl_n_events + n + 1) * sizeof(Event);
    if (need > c->tl_events.cap) {
        DevBuf nb;
        NTL_CUDA(c, nb.ensure(need + need / 2));
        if (keep) NTL_CUDA(c, cudaMemcpyAsync(nb.p, c->tl_events.p, keep, cudaMemcpyDeviceToDevice, c->stream));
        NTL_CUDA(c, cudaStreamSynchronize(c->stream));
        c->tl_events.release();
        c->tl_events = nb;
    }
    NTL_CUDA(c, cudaMemcpyAsync(c->tl_events.as<Event>() + c->tl_n_events, src, n * sizeof(Event), kind, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    c->tl_n_events += n;
    return NTL_OK;
}
int ntl_events_append(ntl_ctx* c, const ntl_event* events, uint64_t n) {
    if (!c || (n && !events)) return NTL_ERR_ARG;
    return events_append_impl(c, events, n, cudaMemcpyHostToDevice);
}
int ntl_events_append_device(ntl_ctx* c, const void* d_events, uint64_t n) {
    if (!c || (n && !d_events)) return NTL_ERR_ARG;
    return events_append_impl(c, d_events, n, cudaMemcpyDeviceToDevice);
}

int ntl_events_export(ntl_ctx* c, void* d_dst, uint64_t cap_events, uint64_t* n_out) {
    if (!c || !d_dst) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    NTL_TRY(events_resolve_count(c));
    const uint64_t n = c->tl_n_events, m = std::min(n, cap_events);
    NTL_CUDA(c, c->h_status.ensure(256));
    uint32_t* hdr = c->h_status.as<uint32_t>() + 32;       // pinned scratch (second half of the status block)
    hdr[0] = (uint32_t)n; hdr[1] = hdr[2] = hdr[3] = hdr[4] = hdr[5] = 0;
    NTL_CUDA(c, cudaMemcpyAsync(d_dst, hdr, 24, cudaMemcpyHostToDevice, c->stream));
    if (m) NTL_CUDA(c, cudaMemcpyAsync((char*)d_dst + 24, c->tl_events.p, m * sizeof(Event), cudaMemcpyDeviceToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    if (n_out) *n_out = n;
    return NTL_OK;
}

int ntl_stream(ntl_ctx* c, void** stream_out) {
    if (!c || !stream_out) return NTL_ERR_ARG;
    *stream_out = (void*)c->stream;
    return NTL_OK;
}

int ntl_events_export_async(ntl_ctx* c, void* d_dst, uint64_t cap_events, uint64_t* n_out) {
    if (!c || !d_dst) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    NTL_TRY(events_resolve_count(c));
    const uint64_t n = c->tl_n_events, m = std::min(n, cap_events);
    // header row through a kernel argument (no pinned scratch that a later call could overwrite while in flight)
    k_export_header<<<1, 32, 0, c->stream>>>((uint32_t*)d_dst, (uint32_t)n);
    c->launches++;
    if (m) NTL_CUDA(c, cudaMemcpyAsync((char*)d_dst + 24, c->tl_events.p, m * sizeof(Event), cudaMemcpyDeviceToDevice, c->stream));
    NTL_CUDA(c, cudaGetLastError());
    if (n_out) *n_out = n;
    return NTL_OK;
}

int ntl_events_import_counts(ntl_ctx* c, const void* d_src, uint32_t world, uint64_t cap_events, const uint32_t* counts) {
    if (!c || !d_src || !world || !counts) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    const size_t stride = (cap_events + 1) * sizeof(Event);
    uint64_t total = 0;
    for (uint32_t r = 0; r < world; r++) {
        if (counts[r] > cap_events) { c->err = "ntl_events_import_counts: a rank sent more events than the buffer holds"; return NTL_ERR_ARG; }
        total += counts[r];
    }
    c->tl_n_events = 0; c->tl_count_on_device = false;
    const size_t need = (total + 1) * sizeof(Event);
    if (need > c->tl_events.cap) NTL_CUDA(c, c->tl_events.ensure(need));
    uint64_t o = 0;
    for (uint32_t r = 0; r < world; r++) {
        if (counts[r]) NTL_CUDA(c, cudaMemcpyAsync(c->tl_events.as<Event>() + o, (const char*)d_src + r * stride + sizeof(Event),
                                                  (size_t)counts[r] * sizeof(Event), cudaMemcpyDeviceToDevice, c->stream));
        o += counts[r];
    }
    c->tl_n_events = total;
    c->tl_count_on_device = false;
    return NTL_OK;
}

int ntl_events_import_device(ntl_ctx* c, const void* d_src, uint32_t world, uint64_t cap_events) {
    if (!c || !d_src || !world || !cap_events) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    return events_import_device(c, d_src, world, cap_events);
}

int ntl_events_import_gathered(ntl_ctx* c, const void* d_src, uint32_t world, uint64_t cap_events, int* overflow) {
    if (!c || !d_src || !world || !overflow) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    const size_t stride = (cap_events + 1) * sizeof(Event);
    std::vector<uint32_t> cnt(world);
    NTL_CUDA(c, cudaMemcpy2DAsync(cnt.data(), 4, d_src, stride, 4, world, cudaMemcpyDeviceToHost, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    uint64_t total = 0;
    *overflow = 0;
    for (uint32_t r = 0; r < world; r++) { if (cnt[r] > cap_events) *overflow = 1; total += cnt[r]; }
    if (*overflow) return NTL_OK;
    c->tl_n_events = 0; c->tl_count_on_device = false;
    const size_t need = (total + 1) * sizeof(Event);
    if (need > c->tl_events.cap) NTL_CUDA(c, c->tl_events.ensure(need));
    uint64_t o = 0;
    for (uint32_t r = 0; r < world; r++) {
        if (cnt[r]) NTL_CUDA(c, cudaMemcpyAsync(c->tl_events.as<Event>() + o, (const char*)d_src + r * stride + sizeof(Event),
                                               (size_t)cnt[r] * sizeof(Event), cudaMemcpyDeviceToDevice, c->stream));
        o += cnt[r];
    }
    c->tl_n_events = total;
    c->tl_count_on_device = false;
    return NTL_OK;
}

int ntl_pairs_finish(ntl_ctx* c, ntl_pairs_out* out) {
    if (!c || !out) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    NTL_TRY(tally_device(c, R->pairs, R->gaps));
    collect_timing(c);
    out->n_pairs = R->pairs.size(); out->n_gaps = R->gaps.size();
    out->pairs = R->pairs.data(); out->gaps = R->gaps.data();
    return NTL_OK;
}

// ------------------------------------------------------------------------------------------- resident (bench)
// cut the resident reads into chunks of whole reads and put the rebased offsets of every chunk on the device
static int plan_resident_chunks(ntl_ctx* c, uint32_t nreads) {
    Results* R = res_of(c);
    const std::vector<uint64_t>& ao = R->r_abs_off;
    R->r_chunks.clear();
    std::vector<uint32_t> bounds;
    const uint64_t per = std::min<uint64_t>(R->resident_chunk_bases, (1ull << 32) - 8192);
    plan_batches(ao.data(), nreads, per, bounds);
    std::vector<uint64_t> ho;
    ho.reserve((size_t)nreads + bounds.size() + 1);
    uint64_t dev_at = 0;
    for (size_t i = 0; i + 1 < bounds.size(); i++) {
        Results::ResChunk ch;
        ch.first = bounds[i]; ch.count = bounds[i + 1] - bounds[i];
        ch.base = ao[ch.first] - ao[0]; ch.nbases = ao[bounds[i + 1]] - ao[ch.first]; ch.off_at = ho.size();
        ch.dev_base = dev_at; dev_at = (dev_at + ch.nbases + 511) & ~255ull;
        if (ch.nbases >= (1ull << 32) - 4096) { c->err = "a single resident read/chunk exceeds 4 Gbp"; return NTL_ERR_ARG; }
        for (uint32_t r = ch.first; r <= bounds[i + 1]; r++) ho.push_back(ao[r] - ao[ch.first]);
        R->r_chunks.push_back(ch);
    }
    NTL_CUDA(c, R->r_off.ensure((ho.size() + 1) * 8));
    NTL_CUDA(c, R->r_seq.ensure(dev_at + 512));
    if (!ho.empty()) NTL_CUDA(c, cudaMemcpyAsync(R->r_off.p, ho.data(), ho.size() * 8, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    R->r_nreads = nreads; R->r_bases = ao[nreads] - ao[0];
    return NTL_OK;
}

int ntl_reads_upload(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t nreads) {
    if (!c || (!seq && nreads) || !offsets) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    const uint64_t nb = offsets[nreads] - offsets[0];
    (void)nb;
    R->r_abs_off.assign(offsets, offsets + nreads + 1);
    NTL_TRY(plan_resident_chunks(c, nreads));
    for (const Results::ResChunk& ch : R->r_chunks)
        if (ch.nbases) NTL_CUDA(c, cudaMemcpyAsync(R->r_seq.as<char>() + ch.dev_base, seq + offsets[ch.first], ch.nbases, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    return NTL_OK;
}

int ntl_map_resident(ntl_ctx* c, uint64_t first_read_ordinal, const ntl_params* prm, ntl_map_out* counts) {
    if (!c || !prm) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    if (!R->r_nreads) { c->err = "ntl_map_resident: no resident reads"; return NTL_ERR_STATE; }
    if (!c->index.built) { c->err = "ntl_map_resident: no target index"; return NTL_ERR_STATE; }
    const size_t nch = R->r_chunks.size();
    auto chunk_seq = [&](const Results::ResChunk& ch) { return R->r_seq.as<uint8_t>() + ch.dev_base; };
    auto chunk_off = [&](const Results::ResChunk& ch) { return R->r_off.as<uint64_t>() + ch.off_at; };
    if (c->async_mode) {
        // sync-free: every chunk's sketch + mapping enqueued back to back (one CUDA graph per chunk position, updated in
        // place from call to call), one synchronisation at the end (see map_reads_async)
        CallState* call = nullptr;
        NTL_TRY(call_begin(c, &call));
        NTL_TRY(sketch_prepare(c, (uint32_t)prm->k));
        if (R->resident_execs.size() < nch) R->resident_execs.resize(nch, nullptr);
        tick(c, T_TOTAL);
        for (size_t i = 0; i < nch; i++) {
            const Results::ResChunk& ch = R->r_chunks[i];
            auto enqueue = [&]() -> int {
                NTL_TRY(sketch_device(c, chunk_seq(ch), chunk_off(ch), ch.count, ch.nbases, (uint32_t)prm->k, (uint32_t)prm->w, c->dsk, call));
                NTL_TRY(read_len_device(c, chunk_off(ch), ch.count, R->read_len));
                NTL_TRY(map_device(c, c->dsk, R->read_len.as<uint32_t>(), ch.count, first_read_ordinal + ch.first, prm, nullptr, nullptr, nullptr, call));
                NTL_TRY(call_chunk_finish(c, call, ch.first, ch.count, nullptr));
                return NTL_OK;
            };
            bool launched = false;
            if (c->graph_mode) {
                NTL_TRY(call_reserve_events(c, ch.count));
                cudaGraph_t graph = nullptr;
                NTL_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
                c->capturing = true;
                const int rc = enqueue();
                c->capturing = false; c->no_stage_timing = false;
                cudaError_t ge = cudaStreamEndCapture(c->stream, &graph);
                if (rc != NTL_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ge == cudaSuccess && inject_graph_failure("resident")) ge = cudaErrorUnknown;
                cudaGraphExec_t& exec = R->resident_execs[i];
                if (ge == cudaSuccess && exec) {
                    cudaGraphExecUpdateResultInfo info;
                    if (cudaGraphExecUpdate(exec, graph, &info) != cudaSuccess) {
                        cudaGetLastError();
                        cudaGraphExecDestroy(exec);
                        exec = nullptr;
                    }
                }
                if (ge == cudaSuccess && !exec) {
                    ge = cudaGraphInstantiate(&exec, graph, 0);
                    if (ge != cudaSuccess) exec = nullptr;
                }
                if (graph) cudaGraphDestroy(graph);
                if (ge == cudaSuccess) ge = cudaGraphLaunch(exec, c->stream);
                if (ge == cudaSuccess) { c->n_graph_launches++; launched = true; }
                else graph_failed(c, "graph step of ntl_map_resident", ge);        // plain launches below
            }
            if (!launched) NTL_TRY(enqueue());
        }
        tock(c, T_TOTAL);
        CallState hs;
        NTL_TRY(call_end(c, call, &hs));
        collect_timing(c);
        c->n_async_calls++;
        if (hs.err) c->n_async_fallbacks++;
        if (hs.err == 0) {
            c->dsk.n_mx = nch == 1 ? hs.mx_total : 0;
            note_mx_density(c, hs.mx_total, R->r_bases, (uint32_t)prm->w);
            if (counts) {
                memset(counts, 0, sizeof *counts);
                counts->n_reads = R->r_nreads; counts->n_mx = hs.mx_total; counts->n_hits = hs.hits_total; counts->n_runs = hs.runs_total;
                counts->n_events = hs.ev_total;
            }
            return NTL_OK;
        }
    }
    uint64_t mx = 0, hits = 0, runs = 0, events = 0;
    for (size_t i = 0; i < nch; i++) {
        const Results::ResChunk& ch = R->r_chunks[i];
        tick(c, T_TOTAL);
        NTL_TRY(sketch_device(c, chunk_seq(ch), chunk_off(ch), ch.count, ch.nbases, (uint32_t)prm->k, (uint32_t)prm->w, c->dsk));
        NTL_TRY(read_len_device(c, chunk_off(ch), ch.count, R->read_len));
        MapStatus cs; uint64_t log_base = 0;
        NTL_TRY(map_device(c, c->dsk, R->read_len.as<uint32_t>(), ch.count, first_read_ordinal + ch.first, prm, &cs, &log_base));
        tock(c, T_TOTAL);
        NTL_TRY(finish_call(c));
        mx += c->dsk.n_mx; hits += cs.n_hits; runs += cs.n_runs; events += cs.n_events;
    }
    if (counts) {
        memset(counts, 0, sizeof *counts);
        counts->n_reads = R->r_nreads; counts->n_mx = mx; counts->n_hits = hits; counts->n_runs = runs; counts->n_events = events;
    }
    return NTL_OK;
}

int ntl_target_upload(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t ncontig, const uint32_t* name_rank) {
    if (!c || (!seq && ncontig) || !offsets || !name_rank) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    const uint64_t nb = offsets[ncontig] - offsets[0];
    if (nb >= (1ull << 32) - 4096) { c->err = "resident target exceeds 4 Gbp"; return NTL_ERR_ARG; }
    std::vector<uint64_t> ho((size_t)ncontig + 1);
    R->t_len.resize(ncontig); R->t_rank.assign(name_rank, name_rank + ncontig);
    for (uint32_t i = 0; i <= ncontig; i++) ho[i] = offsets[i] - offsets[0];
    for (uint32_t i = 0; i < ncontig; i++) R->t_len[i] = (uint32_t)(offsets[i + 1] - offsets[i]);
    NTL_CUDA(c, R->t_seq.ensure(nb + 256));
    NTL_CUDA(c, R->t_off.ensure(((size_t)ncontig + 1) * 8));
    if (nb) NTL_CUDA(c, cudaMemcpyAsync(R->t_seq.p, seq + offsets[0], nb, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(R->t_off.p, ho.data(), ((size_t)ncontig + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    R->t_ncontig = ncontig; R->t_bases = nb;
    return NTL_OK;
}

int ntl_index_build_resident(ntl_ctx* c, int k, int w) {
    if (!c || k <= 0 || w <= 0) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    if (!R->t_ncontig) { c->err = "ntl_index_build_resident: no resident target"; return NTL_ERR_STATE; }
    if (c->async_mode) {
        bool ok = false;
        NTL_TRY(index_build_async(c, R->t_seq.as<uint8_t>(), R->t_off.as<uint64_t>(), R->t_ncontig, R->t_bases, k, w, R->t_len.data(),
                                  R->t_rank.data(), R->t_ctg, &ok));
        if (ok) return NTL_OK;
    }
    NTL_TRY(sketch_device(c, R->t_seq.as<uint8_t>(), R->t_off.as<uint64_t>(), R->t_ncontig, R->t_bases, (uint32_t)k, (uint32_t)w, c->dsk));
    NTL_TRY(expand_contig_ids(c, c->dsk, R->t_ctg));
    NTL_TRY(index_build_device(c, c->dsk.hash.as<uint64_t>(), R->t_ctg.as<uint32_t>(), c->dsk.posf.as<uint32_t>(), c->dsk.n_mx,
                               R->t_len.data(), R->t_rank.data(), R->t_ncontig));
    return finish_call(c);
}

static __global__ void k_add_u32(uint32_t* p, uint32_t v, const uint32_t* n_dev, uint32_t n_bound) {
    const uint32_t n = n_dev ? min(*n_dev, n_bound) : n_bound;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] += v;
}

// Multi-GPU: sketch contigs [first, first + count) of the resident target; the minimizer triples stay on the device
// (hash, GLOBAL contig id, pos|strand) for the caller's all-gather (ntlink_b200/dist.py). Synchronous.
int ntl_target_sketch_resident(ntl_ctx* c, uint32_t first, uint32_t count, int k, int w, uint64_t* n_mx, void** d_hash, void** d_contig,
                               void** d_pos_strand) {
    if (!c || k <= 0 || w <= 0) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    if ((uint64_t)first + count > R->t_ncontig) { c->err = "ntl_target_sketch_resident: contig range"; return NTL_ERR_ARG; }
    uint64_t a = 0, nb = 0;
    for (uint32_t i = 0; i < first; i++) a += R->t_len[i];
    std::vector<uint64_t> ho((size_t)count + 1, 0);
    for (uint32_t i = 0; i < count; i++) ho[i + 1] = ho[i] + R->t_len[first + i];
    nb = ho[count];
    // the shard starts at an arbitrary base: the sketch kernels want a 16-byte aligned batch, so the shard is staged
    NTL_CUDA(c, c->d_seq.ensure(nb + 256));
    NTL_CUDA(c, c->d_off.ensure(((size_t)count + 1) * 8));
    if (nb) NTL_CUDA(c, cudaMemcpyAsync(c->d_seq.p, R->t_seq.as<char>() + a, nb, cudaMemcpyDeviceToDevice, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(c->d_off.p, ho.data(), ((size_t)count + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    NTL_TRY(sketch_device(c, c->d_seq.as<uint8_t>(), c->d_off.as<uint64_t>(), count, nb, (uint32_t)k, (uint32_t)w, c->dsk));
    NTL_TRY(expand_contig_ids(c, c->dsk, R->ctg_ids));
    if (c->dsk.n_mx && first) {
        k_add_u32<<<std::min<uint32_t>((c->dsk.n_mx + 255) / 256, 148 * 8), 256, 0, c->stream>>>(R->ctg_ids.as<uint32_t>(), first, nullptr, c->dsk.n_mx);
        c->launches++;
    }
    NTL_TRY(finish_call(c));
    if (n_mx) *n_mx = c->dsk.n_mx;
    if (d_hash) *d_hash = c->dsk.hash.p;
    if (d_contig) *d_contig = R->ctg_ids.p;
    if (d_pos_strand) *d_pos_strand = c->dsk.posf.p;
    return NTL_OK;
}

// contig lengths / name ranks of the resident target (the arguments ntl_index_build_device needs)
int ntl_target_resident_meta(ntl_ctx* c, uint32_t* contig_len, uint32_t* name_rank) {
    if (!c) return NTL_ERR_ARG;
    Results* R = res_of(c);
    if (contig_len) memcpy(contig_len, R->t_len.data(), (size_t)R->t_ncontig * 4);
    if (name_rank) memcpy(name_rank, R->t_rank.data(), (size_t)R->t_ncontig * 4);
    return NTL_OK;
}

int ntl_timing_reset(ntl_ctx* c) {
    if (!c) return NTL_ERR_ARG;
    for (int i = 0; i < T_NUM; i++) c->ms_accum[i] = 0;
    c->launches = 0; c->dense_launches = 0; c->dense_bases = 0;
    c->big_dense_ms = 0; c->big_dense_launches = 0; c->big_dense_bases = 0;
    return NTL_OK;
}
int ntl_timing_dense(ntl_ctx* c, double* ms_accum, uint64_t* launches, uint64_t* bases) {
    if (!c) return NTL_ERR_ARG;
    if (ms_accum) *ms_accum = c->big_dense_ms;
    if (launches) *launches = c->big_dense_launches;
    if (bases) *bases = c->big_dense_bases;
    return NTL_OK;
}
int ntl_timing(ntl_ctx* c, double* ms_accum, uint64_t* launches, uint64_t* dense_launches, uint64_t* dense_bases) {
    if (!c) return NTL_ERR_ARG;
    if (ms_accum) for (int i = 0; i < T_NUM; i++) ms_accum[i] = c->ms_accum[i];
    if (launches) *launches = c->launches;
    if (dense_launches) *dense_launches = c->dense_launches;
    if (dense_bases) *dense_bases = c->dense_bases;
    return NTL_OK;
}
int ntl_get_stat(ntl_ctx* c, const char* name, double* value) {
    if (!c || !name || !value) return NTL_ERR_ARG;
    if (!strcmp(name, "async_calls")) *value = (double)c->n_async_calls;
    else if (!strcmp(name, "async_fallbacks")) *value = (double)c->n_async_fallbacks;
    else if (!strcmp(name, "graph_launches")) *value = (double)c->n_graph_launches;
    else if (!strcmp(name, "graph_failures")) *value = (double)c->n_graph_failures;
    else if (!strcmp(name, "tile_batches")) *value = (double)c->n_tile_batches;
    else if (!strcmp(name, "small_batches")) *value = (double)c->n_small_batches;
    else if (!strcmp(name, "tile_fallbacks")) *value = (double)c->n_tile_fallbacks;
    else { c->err = std::string("unknown stat ") + name; return NTL_ERR_ARG; }
    return NTL_OK;
}
int ntl_mark(ntl_ctx* c, int which) {
    if (!c || which < 0 || which > 1) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    NTL_CUDA(c, cudaEventRecord(c->mark[which], c->stream));
    return NTL_OK;
}
int ntl_mark_elapsed(ntl_ctx* c, double* ms) {
    if (!c || !ms) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    NTL_CUDA(c, cudaEventSynchronize(c->mark[1]));
    float f = 0;
    NTL_CUDA(c, cudaEventElapsedTime(&f, c->mark[0], c->mark[1]));
    *ms = f;
    return NTL_OK;
}
int ntl_copy_device(ntl_ctx* c, void* d_dst, const void* d_src, uint64_t bytes) {
    if (!c || (bytes && (!d_dst || !d_src))) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    if (bytes) NTL_CUDA(c, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    return NTL_OK;
}
int ntl_device_sync(ntl_ctx* c) {
    if (!c) return NTL_ERR_ARG;
    cudaSetDevice(c->device);
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    return NTL_OK;
}

// ------------------------------------------------------------------------------------------- synthetic inputs
// Benchmark / test inputs only: the same counter-based generator on the device (resident target and
// reads, no host copy of multi-Gbp inputs) and on the host (tests, the CPU reference arm).
static_assert(sizeof(ntl_synth_contig) == sizeof(SynthContig) && sizeof(ntl_synth_read) == sizeof(SynthRead), "synth record layout");

int ntl_synth_target_resident(ntl_ctx* c, uint64_t seed, const ntl_synth_contig* contigs, uint32_t ncontig, const uint32_t* name_rank) {
    if (!c || !contigs || !ncontig || !name_rank) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    std::vector<uint64_t> ho((size_t)ncontig + 1, 0);
    R->t_len.resize(ncontig); R->t_rank.assign(name_rank, name_rank + ncontig);
    for (uint32_t i = 0; i < ncontig; i++) { R->t_len[i] = contigs[i].len; ho[i + 1] = ho[i] + contigs[i].len; }
    const uint64_t nb = ho[ncontig];
    if (nb >= (1ull << 32) - 4096) { c->err = "resident target exceeds 4 Gbp"; return NTL_ERR_ARG; }
    DevBuf d_ctg;
    NTL_CUDA(c, d_ctg.ensure((size_t)ncontig * sizeof(SynthContig)));
    NTL_CUDA(c, R->t_seq.ensure(nb + 256));
    NTL_CUDA(c, R->t_off.ensure(((size_t)ncontig + 1) * 8));
    NTL_CUDA(c, cudaMemcpyAsync(d_ctg.p, contigs, (size_t)ncontig * sizeof(SynthContig), cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(R->t_off.p, ho.data(), ((size_t)ncontig + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    k_synth_contigs<<<dim3(32, std::min<uint32_t>(ncontig, 4096)), 256, 0, c->stream>>>(seed, d_ctg.as<SynthContig>(), R->t_off.as<uint64_t>(), ncontig,
                                                                                      R->t_seq.as<uint8_t>());
    NTL_CUDA(c, cudaGetLastError());
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    d_ctg.release();
    R->t_ncontig = ncontig; R->t_bases = nb;
    return NTL_OK;
}

int ntl_synth_reads_resident(ntl_ctx* c, uint64_t seed, const ntl_synth_read* reads, uint32_t nreads, uint32_t sub16, uint32_t del16,
                             uint32_t ins16, uint64_t* total_bases) {
    if (!c || !reads || !nreads || sub16 + del16 + ins16 > 65536u) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    const SynthErr e{sub16, del16, ins16};
    DevBuf d_rd, d_len, d_abs;
    NTL_CUDA(c, d_rd.ensure((size_t)nreads * sizeof(SynthRead)));
    NTL_CUDA(c, d_len.ensure((size_t)nreads * 4));
    NTL_CUDA(c, cudaMemcpyAsync(d_rd.p, reads, (size_t)nreads * sizeof(SynthRead), cudaMemcpyHostToDevice, c->stream));
    const uint32_t blocks = (uint32_t)(((uint64_t)nreads * 32 + 255) / 256);
    k_synth_read_len<<<blocks, 256, 0, c->stream>>>(seed, d_rd.as<SynthRead>(), nreads, e, d_len.as<uint32_t>());
    std::vector<uint32_t> len(nreads);
    NTL_CUDA(c, cudaMemcpyAsync(len.data(), d_len.p, (size_t)nreads * 4, cudaMemcpyDeviceToHost, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    R->r_abs_off.assign((size_t)nreads + 1, 0);
    for (uint32_t i = 0; i < nreads; i++) R->r_abs_off[i + 1] = R->r_abs_off[i] + len[i];
    NTL_TRY(plan_resident_chunks(c, nreads));
    // destination of every read inside the chunked device buffer
    std::vector<uint64_t> dst((size_t)nreads);
    for (const Results::ResChunk& ch : R->r_chunks)
        for (uint32_t r = ch.first; r < ch.first + ch.count; r++) dst[r] = ch.dev_base + (R->r_abs_off[r] - R->r_abs_off[ch.first]);
    NTL_CUDA(c, d_abs.ensure((size_t)nreads * 8));
    NTL_CUDA(c, cudaMemcpyAsync(d_abs.p, dst.data(), (size_t)nreads * 8, cudaMemcpyHostToDevice, c->stream));
    k_synth_reads<<<blocks, 256, 0, c->stream>>>(seed, d_rd.as<SynthRead>(), d_abs.as<uint64_t>(), nreads, e, R->r_seq.as<uint8_t>());
    NTL_CUDA(c, cudaGetLastError());
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    d_rd.release(); d_len.release(); d_abs.release();
    if (total_bases) *total_bases = R->r_bases;
    return NTL_OK;
}

int ntl_resident_info(ntl_ctx* c, int which, uint32_t* nseq, uint64_t* nbases) {
    if (!c) return NTL_ERR_ARG;
    Results* R = res_of(c);
    if (nseq) *nseq = which ? R->r_nreads : R->t_ncontig;
    if (nbases) *nbases = which ? R->r_bases : R->t_bases;
    return NTL_OK;
}

int ntl_resident_download(ntl_ctx* c, int which, uint32_t first, uint32_t count, char* seq_out, uint64_t* off_out) {
    if (!c || !off_out) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    if (which == 0) {
        if ((uint64_t)first + count > R->t_ncontig) { c->err = "ntl_resident_download: range"; return NTL_ERR_ARG; }
        uint64_t a = 0;
        for (uint32_t i = 0; i < first; i++) a += R->t_len[i];
        off_out[0] = 0;
        for (uint32_t i = 0; i < count; i++) off_out[i + 1] = off_out[i] + R->t_len[first + i];
        if (seq_out && off_out[count]) NTL_CUDA(c, cudaMemcpyAsync(seq_out, R->t_seq.as<char>() + a, off_out[count], cudaMemcpyDeviceToHost, c->stream));
    } else {
        if ((uint64_t)first + count > R->r_nreads) { c->err = "ntl_resident_download: range"; return NTL_ERR_ARG; }
        const std::vector<uint64_t>& ao = R->r_abs_off;
        for (uint32_t i = 0; i <= count; i++) off_out[i] = ao[first + i] - ao[first];
        if (seq_out)
            for (const Results::ResChunk& ch : R->r_chunks) {
                const uint32_t a = std::max(first, ch.first), b = std::min(first + count, ch.first + ch.count);
                if (a >= b) continue;
                NTL_CUDA(c, cudaMemcpyAsync(seq_out + (ao[a] - ao[first]), R->r_seq.as<char>() + ch.dev_base + (ao[a] - ao[ch.first]), ao[b] - ao[a],
                                            cudaMemcpyDeviceToHost, c->stream));
            }
    }
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    return NTL_OK;
}

static void run_threads(int threads, uint32_t n, const std::function<void(uint32_t, uint32_t)>& fn) {
    threads = std::max(1, std::min<int>(threads, (int)std::max<uint32_t>(1, n / 64)));
    std::vector<std::thread> th;
    const uint32_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const uint32_t a = std::min<uint64_t>(n, (uint64_t)t * per), b = std::min<uint64_t>(n, (uint64_t)a + per);
        if (a < b) th.emplace_back(fn, a, b);
    }
    for (auto& x : th) x.join();
}

int ntl_synth_host_contigs(uint64_t seed, const ntl_synth_contig* contigs, uint32_t ncontig, uint64_t* off_out, char* seq_out, int threads) {
    if (!contigs || !off_out) return NTL_ERR_ARG;
    off_out[0] = 0;
    for (uint32_t i = 0; i < ncontig; i++) off_out[i + 1] = off_out[i] + contigs[i].len;
    if (!seq_out) return NTL_OK;
    const SynthContig* sc = reinterpret_cast<const SynthContig*>(contigs);
    run_threads(threads, ncontig, [&](uint32_t a, uint32_t b) {
        for (uint32_t i = a; i < b; i++) {
            uint8_t* dst = reinterpret_cast<uint8_t*>(seq_out) + off_out[i];
            for (uint32_t j = 0; j < sc[i].len; j++) dst[j] = contig_base(seed, sc[i], j);
        }
    });
    return NTL_OK;
}

int ntl_synth_host_reads(uint64_t seed, const ntl_synth_read* reads, uint32_t nreads, uint32_t sub16, uint32_t del16, uint32_t ins16,
                         uint64_t* off_out, char* seq_out, int threads) {
    if (!reads || !off_out || sub16 + del16 + ins16 > 65536u) return NTL_ERR_ARG;
    const SynthErr e{sub16, del16, ins16};
    const SynthRead* sr = reinterpret_cast<const SynthRead*>(reads);
    if (!seq_out) {                     // first call: lengths -> offsets
        std::vector<uint32_t> len(nreads);
        run_threads(threads, nreads, [&](uint32_t a, uint32_t b) {
            for (uint32_t i = a; i < b; i++) {
                const uint64_t key = read_key(seed, sr[i].id);
                uint32_t n = 0;
                for (uint32_t j = 0; j < sr[i].len; j++) n += read_emit_count(key, j, e);
                len[i] = n;
            }
        });
        off_out[0] = 0;
        for (uint32_t i = 0; i < nreads; i++) off_out[i + 1] = off_out[i] + len[i];
        return NTL_OK;
    }
    run_threads(threads, nreads, [&](uint32_t a, uint32_t b) {
        for (uint32_t i = a; i < b; i++) {
            const uint64_t key = read_key(seed, sr[i].id);
            uint8_t* p = reinterpret_cast<uint8_t*>(seq_out) + off_out[i];
            for (uint32_t j = 0; j < sr[i].len; j++) {
                uint8_t o[2];
                const uint32_t m = read_emit(seed, key, sr[i], j, e, o);
                if (m) *p++ = o[0];
                if (m == 2) *p++ = o[1];
            }
        }
    });
    return NTL_OK;
}

}  // extern "C"

// capi.cu -- the extern "C" boundary of libntlink_b200.so (declared in include/ntlink_b200.h).
// Host orchestration only: batching, H2D/D2H staging through pinned memory, error translation. All compute is
// in sketch.cu / map.cu. There is deliberately no CPU fallback: without a CUDA device ntl_init fails.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <thread>

#include "common.cuh"
#include "lift_logic.cuh"

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
    DevBuf r_seq, r_off, r_len; uint32_t r_nreads = 0; uint64_t r_bases = 0;
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
    cudaGraphExec_t resident_exec = nullptr, index_exec = nullptr;
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
    for (int i = 0; i < 2 * T_NUM; i++) cudaEventCreate(&c->ev[i]);
    cudaEventCreate(&c->mark[0]); cudaEventCreate(&c->mark[1]);
    cudaStreamCreateWithFlags(&res_of(c)->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&res_of(c)->h2d_done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&res_of(c)->h2d_done[1], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&res_of(c)->comp_done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&res_of(c)->comp_done[1], cudaEventDisableTiming);
    for (int i = 0; i < T_NUM; i++) { c->ev_used[i] = false; c->ms_accum[i] = 0; }
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
                    &W.selcnt, &W.selmask, &W.selbase, &W.strip_seq, &W.gaps, &W.gap_head, &W.extras, &W.has_cand, &W.status, &W.tbl};
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
    if (R->resident_exec) cudaGraphExecDestroy(R->resident_exec);
    if (R->index_exec) cudaGraphExecDestroy(R->index_exec);
    R->idx_meta.release();
    cudaStreamSynchronize(R->copy_stream);
    cudaStreamDestroy(R->copy_stream);
    c->h_status.release();
    for (int i = 0; i < 2 * T_NUM; i++) cudaEventDestroy(c->ev[i]);
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
        if (v < 8 || v > 65536 || (v & 7)) { c->err = "strip_len must be a multiple of 8 in [8, 65536]"; return NTL_ERR_ARG; }
        c->strip_len = v;
    } else if (!strcmp(name, "cand_c")) {
        if (!(value > 0)) { c->err = "cand_c must be positive"; return NTL_ERR_ARG; }
        c->cand_c = value;
    } else if (!strcmp(name, "async")) {
        c->async_mode = value != 0.0;
    } else if (!strcmp(name, "copy_threads")) {
        c->copy_threads = (int)value;
    } else if (!strcmp(name, "graph")) {
        c->graph_mode = value != 0.0;
    } else if (!strcmp(name, "pipeline_min_bases")) {
        if (value < 1024 || value > 3.9e9) { c->err = "pipeline_min_bases out of range"; return NTL_ERR_ARG; }
        c->pipeline_min_bases = (uint64_t)value;
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
    plan_batches(offsets, nseq, c->batch_bases, bounds);
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
int ntl_reads_upload(ntl_ctx* c, const char* seq, const uint64_t* offsets, uint32_t nreads) {
    if (!c || (!seq && nreads) || !offsets) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    const uint64_t nb = offsets[nreads] - offsets[0];
    if (nb >= (1ull << 32) - 4096) { c->err = "resident batch exceeds 4 Gbp"; return NTL_ERR_ARG; }
    std::vector<uint64_t> ho((size_t)nreads + 1);
    for (uint32_t i = 0; i <= nreads; i++) ho[i] = offsets[i] - offsets[0];
    NTL_CUDA(c, R->r_seq.ensure(nb + 256));
    NTL_CUDA(c, R->r_off.ensure(((size_t)nreads + 1) * 8));
    if (nb) NTL_CUDA(c, cudaMemcpyAsync(R->r_seq.p, seq + offsets[0], nb, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(R->r_off.p, ho.data(), ((size_t)nreads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    R->r_nreads = nreads; R->r_bases = nb;
    return NTL_OK;
}

int ntl_map_resident(ntl_ctx* c, uint64_t first_read_ordinal, const ntl_params* prm, ntl_map_out* counts) {
    if (!c || !prm) return NTL_ERR_ARG;
    Results* R = res_of(c);
    cudaSetDevice(c->device);
    if (!R->r_nreads) { c->err = "ntl_map_resident: no resident reads"; return NTL_ERR_STATE; }
    if (c->async_mode) {
        // sync-free: sketch + mapping enqueued back to back (as one CUDA graph, updated in place from call to call),
        // one synchronisation at the end (see map_reads_async)
        CallState* call = nullptr;
        NTL_TRY(call_begin(c, &call));
        auto enqueue = [&]() -> int {
            tick(c, T_TOTAL);
            NTL_TRY(sketch_device(c, R->r_seq.as<uint8_t>(), R->r_off.as<uint64_t>(), R->r_nreads, R->r_bases, (uint32_t)prm->k,
                                  (uint32_t)prm->w, c->dsk, call));
            NTL_TRY(read_len_device(c, R->r_off.as<uint64_t>(), R->r_nreads, R->read_len));
            NTL_TRY(map_device(c, c->dsk, R->read_len.as<uint32_t>(), R->r_nreads, first_read_ordinal, prm, nullptr, nullptr, nullptr, call));
            NTL_TRY(call_chunk_finish(c, call, 0, R->r_nreads, nullptr));
            tock(c, T_TOTAL);
            return NTL_OK;
        };
        bool launched = false;
        if (c->graph_mode) {
            NTL_TRY(sketch_prepare(c, (uint32_t)prm->k));
            NTL_TRY(call_reserve_events(c, R->r_nreads));
            cudaGraph_t graph = nullptr;
            NTL_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            c->capturing = true;
            const int rc = enqueue();
            c->capturing = false;
            cudaError_t ge = cudaStreamEndCapture(c->stream, &graph);
            if (rc != NTL_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ge == cudaSuccess && inject_graph_failure("resident")) ge = cudaErrorUnknown;
            if (ge == cudaSuccess && R->resident_exec) {
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(R->resident_exec, graph, &info) != cudaSuccess) {
                    cudaGetLastError();
                    cudaGraphExecDestroy(R->resident_exec);
                    R->resident_exec = nullptr;
                }
            }
            if (ge == cudaSuccess && !R->resident_exec) {
                ge = cudaGraphInstantiate(&R->resident_exec, graph, 0);
                if (ge != cudaSuccess) R->resident_exec = nullptr;
            }
            if (graph) cudaGraphDestroy(graph);
            if (ge == cudaSuccess) ge = cudaGraphLaunch(R->resident_exec, c->stream);
            if (ge == cudaSuccess) { c->n_graph_launches++; launched = true; }
            else graph_failed(c, "graph step of ntl_map_resident", ge);        // plain launches below
        }
        if (!launched) NTL_TRY(enqueue());
        CallState hs;
        NTL_TRY(call_end(c, call, &hs));
        collect_timing(c);
        c->n_async_calls++;
        if (hs.err) c->n_async_fallbacks++;
        if (hs.err == 0) {
            c->dsk.n_mx = hs.mx_total;
            if (counts) {
                memset(counts, 0, sizeof *counts);
                counts->n_reads = R->r_nreads; counts->n_mx = hs.mx_total; counts->n_hits = hs.hits_total; counts->n_runs = hs.runs_total;
                counts->n_events = hs.ev_total;
            }
            return NTL_OK;
        }
    }
    tick(c, T_TOTAL);
    NTL_TRY(sketch_device(c, R->r_seq.as<uint8_t>(), R->r_off.as<uint64_t>(), R->r_nreads, R->r_bases, (uint32_t)prm->k,
                          (uint32_t)prm->w, c->dsk));
    NTL_TRY(read_len_device(c, R->r_off.as<uint64_t>(), R->r_nreads, R->read_len));
    MapStatus cs; uint64_t log_base = 0;
    NTL_TRY(map_device(c, c->dsk, R->read_len.as<uint32_t>(), R->r_nreads, first_read_ordinal, prm, &cs, &log_base));
    tock(c, T_TOTAL);
    NTL_TRY(finish_call(c));
    if (counts) {
        memset(counts, 0, sizeof *counts);
        counts->n_reads = R->r_nreads; counts->n_mx = c->dsk.n_mx; counts->n_hits = cs.n_hits; counts->n_runs = cs.n_runs;
        counts->n_events = cs.n_events;
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

}  // extern "C"

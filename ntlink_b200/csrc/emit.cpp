// emit.cpp -- host side of the drop-in boundary: byte-identical text emitters and the FASTA/FASTQ reader.
//
//   ntl_format_sketch_tsv   the TSV `indexlr --long [--pos] [--strand] [--len]` prints        
//   ntl_format_verbose      <prefix>.verbose_mapping.tsv lines          bin/ntlink_pair.py:307-313,382-388
//   ntl_format_paf          <prefix>.paf lines                          bin/ntlink_paf_output.py:9-135
//   ntl_seqfile_*           readfq-style reader, id = first token       bin/read_fasta.py:6-46
//
// Pure host code (decimal formatting of GPU results); multi-threaded over reads because at genome scale
// these files are gigabytes of text.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntlink_b200.h"

// Big sequence buffers are recycled: a fresh 1 GB buffer costs hundreds of milliseconds of page faults, more than parsing
// the batch that goes into it. ntl_free() parks up to two of them, the next batch takes one over.
bool ntl_pool_release(void* p);

namespace {

inline void put_u64(std::string& s, uint64_t v) {
    char t[24];
    int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    char r[24];
    for (int i = 0; i < n; i++) r[i] = t[n - 1 - i];
    s.append(r, (size_t)n);
}

template <class F>
void parallel_chunks(uint64_t n, int threads, F&& body /* (chunk_index, begin, end) */, int* nchunks_out) {
    int T = threads < 1 ? 1 : threads;
    if ((uint64_t)T > n) T = n ? (int)n : 1;
    *nchunks_out = T;
    if (T == 1) { body(0, (uint64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) {
        const uint64_t b = n * (uint64_t)t / T, e = n * (uint64_t)(t + 1) / T;
        th.emplace_back([&body, t, b, e]() { body(t, b, e); });
    }
    for (auto& x : th) x.join();
}

// The emitters are called once per batch with outputs of tens to hundreds of megabytes. Fresh memory of that size costs
// more in page faults than formatting the text into it (and the faults of concurrent threads queue on the address-space
// lock), so the per-thread part buffers and the joined output buffers are recycled: parts keep their capacity in a pool,
// ntl_buf_free() parks up to two output buffers for the next call.
struct TextPool {
    std::mutex mu;
    std::vector<std::string> parts;                       // cleared strings that keep their capacity
    std::unordered_map<char*, size_t> live;               // output buffers handed out (ptr -> capacity)
    std::vector<std::pair<char*, size_t>> parked;
    static constexpr size_t MAX_PARKED = 2, MAX_BYTES = (size_t)2 << 30;
};
TextPool& text_pool() { static TextPool* p = new TextPool(); return *p; }

std::vector<std::string> take_parts(size_t n) {
    TextPool& P = text_pool();
    std::vector<std::string> out;
    std::lock_guard<std::mutex> lk(P.mu);
    // largest capacities first: chunk t of this call is about as large as chunk t of the last one
    std::sort(P.parts.begin(), P.parts.end(), [](const std::string& a, const std::string& b) { return a.capacity() < b.capacity(); });
    while (out.size() < n && !P.parts.empty()) { out.emplace_back(std::move(P.parts.back())); P.parts.pop_back(); }
    out.resize(n);
    return out;
}
void give_parts(std::vector<std::string>& parts) {
    TextPool& P = text_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    size_t held = 0;
    for (auto& s : P.parts) held += s.capacity();
    for (auto& s : parts) {
        if (P.parts.size() >= 64 || held + s.capacity() > TextPool::MAX_BYTES) continue;
        held += s.capacity();
        s.clear();
        P.parts.emplace_back(std::move(s));
    }
    parts.clear();
}
char* take_out(size_t total) {
    TextPool& P = text_pool();
    const size_t want = total ? total : 1;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        for (size_t i = 0; i < P.parked.size(); i++)
            if (P.parked[i].second >= want && P.parked[i].second <= 2 * want + (64u << 20)) {
                char* p = P.parked[i].first;
                P.live[p] = P.parked[i].second;
                P.parked.erase(P.parked.begin() + (long)i);
                return p;
            }
    }
    const size_t cap = want + want / 8 + 4096;                  // a little head room: the next batch is rarely the same size
    char* p = (char*)malloc(cap);
    if (p) { std::lock_guard<std::mutex> lk(P.mu); P.live[p] = cap; }
    return p;
}
void release_out(char* p) {
    if (!p) return;
    TextPool& P = text_pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(p);
        if (it != P.live.end()) {
            const size_t cap = it->second;
            P.live.erase(it);
            size_t held = cap;
            for (auto& q : P.parked) held += q.second;
            if (P.parked.size() < TextPool::MAX_PARKED && held <= TextPool::MAX_BYTES) { P.parked.emplace_back(p, cap); return; }
        }
    }
    free(p);
}

int64_t join_out(std::vector<std::string>& parts, char** out_buf) {
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    char* buf = take_out(total);
    if (!buf) { give_parts(parts); return NTL_ERR_ARG; }
    // every part is copied by its own thread (the parts were written by as many threads)
    std::vector<size_t> at(parts.size() + 1, 0);
    for (size_t i = 0; i < parts.size(); i++) at[i + 1] = at[i] + parts[i].size();
    if (total < (8u << 20) || parts.size() < 2) {
        for (size_t i = 0; i < parts.size(); i++) memcpy(buf + at[i], parts[i].data(), parts[i].size());
    } else {
        std::vector<std::thread> th;
        for (size_t i = 1; i < parts.size(); i++) th.emplace_back([&, i]() { memcpy(buf + at[i], parts[i].data(), parts[i].size()); });
        memcpy(buf, parts[0].data(), parts[0].size());
        for (auto& x : th) x.join();
    }
    give_parts(parts);
    *out_buf = buf;
    return (int64_t)total;
}

struct PHit { uint32_t cpos, rpos; bool cfw, rfw; };

inline bool consistent(const std::vector<PHit>& s, bool inc, size_t a, size_t b, const std::vector<uint32_t>& dups) {
    auto is_dup = [&](uint32_t p) { return std::binary_search(dups.begin(), dups.end(), p); };
    if (is_dup(s[a].cpos) || is_dup(s[b].cpos)) return true;
    return inc ? s[a].rpos <= s[b].rpos : s[a].rpos >= s[b].rpos;
}

// bin/ntlink_paf_output.py:60-93 + 18-58: split the (ctg_pos, read_pos)-sorted hits into mapped blocks.
// Returns false when the reference yields no block at all (neither direction is >= 75 % consistent).
bool mapped_blocks(const std::vector<PHit>& s, std::vector<std::vector<PHit>>& blocks) {
    const size_t n = s.size(), nt = n - 1;
    std::vector<char> tinc(nt), tdec(nt);
    std::vector<uint32_t> seen, dups;
    bool all_inc = true, all_dec = true;
    for (size_t i = 0; i < nt; i++) {
        tinc[i] = s[i].rpos <= s[i + 1].rpos;
        tdec[i] = s[i].rpos >= s[i + 1].rpos;
        all_inc = all_inc && tinc[i];
        all_dec = all_dec && tdec[i];
        if (std::find(seen.begin(), seen.end(), s[i].cpos) != seen.end()) dups.push_back(s[i].cpos);
        else seen.push_back(s[i].cpos);
    }
    if (std::find(seen.begin(), seen.end(), s[n - 1].cpos) != seen.end()) dups.push_back(s[n - 1].cpos);
    if (all_inc || all_dec) { blocks.push_back(s); return true; }
    std::sort(dups.begin(), dups.end());
    size_t n_inc = 0;
    for (size_t i = 0; i < nt; i++) n_inc += tinc[i] ? 1 : 0;
    bool inc;
    if (4 * n_inc >= 3 * nt) inc = true;                 // n_inc / nt >= 0.75
    else if (4 * (nt - n_inc) >= 3 * nt) inc = false;    // (nt - n_inc) / nt >= 0.75
    else return false;
    const std::vector<char>& tr = inc ? tinc : tdec;
    std::vector<char> brk(n, 0), flt(n, 0);
    bool any = false;
    auto is_dup = [&](uint32_t p) { return std::binary_search(dups.begin(), dups.end(), p); };
    for (size_t i = 0; i < nt; i++) {
        if (tr[i]) continue;
        if (is_dup(s[i].cpos) || is_dup(s[i + 1].cpos)) continue;
        if (i + 2 >= nt) { brk[i + 1] = 1; any = true; }
        else if (consistent(s, inc, i, i + 2, dups)) { flt[i + 1] = 1; any = true; }
        else if (i > 0 && consistent(s, inc, i - 1, i + 1, dups)) { flt[i] = 1; any = true; }
        else { brk[i + 1] = 1; any = true; }
    }
    if (!any) { blocks.push_back(s); return true; }
    std::vector<PHit> cur;
    for (size_t i = 0; i < n; i++) {
        if (flt[i]) continue;
        if (brk[i]) { blocks.push_back(cur); cur.clear(); cur.push_back(s[i]); }
        else cur.push_back(s[i]);
    }
    blocks.push_back(cur);
    return true;
}

}  // namespace

extern "C" {

int64_t ntl_format_sketch_tsv(const ntl_sketch_out* sk, const char* names, const uint64_t* name_off,
                              const uint64_t* seq_len, int with_pos, int with_strand, int threads, char** out_buf) {
    if (!sk || !names || !name_off || !out_buf) return NTL_ERR_ARG;
    int nch = 1;
    std::vector<std::string> parts = take_parts((size_t)(threads < 1 ? 1 : threads));
    parallel_chunks(sk->nseq, threads, [&](int t, uint64_t b, uint64_t e) {
        // a thread formats into a string of its own: the string objects of `parts` sit side by side in memory, and every
        // push_back updates the size field -- two threads per cache line is false sharing that serialises them all
        std::string s = std::move(parts[(size_t)t]);
        struct PutBack { std::string& from; std::string& to; ~PutBack() { to = std::move(from); } } put_back{s, parts[(size_t)t]};
        s.reserve((size_t)((sk->seq_off[e] - sk->seq_off[b]) * 30 + (e - b) * 48));
        for (uint64_t i = b; i < e; i++) {
            s.append(names + name_off[i], (size_t)(name_off[i + 1] - name_off[i]));
            if (seq_len) { s.push_back('\t'); put_u64(s, seq_len[i]); }
            s.push_back('\t');
            for (uint64_t m = sk->seq_off[i]; m < sk->seq_off[i + 1]; m++) {
                if (m != sk->seq_off[i]) s.push_back(' ');
                put_u64(s, sk->hash[m]);
                if (with_pos) { s.push_back(':'); put_u64(s, sk->pos_strand[m] & NTL_POS_MASK); }
                if (with_strand) { s.push_back(':'); s.push_back((sk->pos_strand[m] & NTL_STRAND_BIT) ? '+' : '-'); }
            }
            s.push_back('\n');
        }
    }, &nch);
    return join_out(parts, out_buf);
}

int64_t ntl_format_verbose(const ntl_map_out* m, const char* read_names, const uint64_t* read_name_off,
                           const char* ctg_names, const uint64_t* ctg_name_off, int threads, char** out_buf) {
    if (!m || !read_names || !read_name_off || !ctg_names || !ctg_name_off || !out_buf) return NTL_ERR_ARG;
    int nch = 1;
    std::vector<std::string> parts = take_parts((size_t)(threads < 1 ? 1 : threads));
    parallel_chunks(m->n_reads, threads, [&](int t, uint64_t b, uint64_t e) {
        // a thread formats into a string of its own: the string objects of `parts` sit side by side in memory, and every
        // push_back updates the size field -- two threads per cache line is false sharing that serialises them all
        std::string s = std::move(parts[(size_t)t]);
        struct PutBack { std::string& from; std::string& to; ~PutBack() { to = std::move(from); } } put_back{s, parts[(size_t)t]};
        {   // room for the whole chunk up front (a line is two names + a count + <= 26 bytes per hit): a string that grows by
            // doubling re-maps its buffer again and again, and the threads then queue on the address-space lock
            size_t need = 64;
            for (uint64_t r = b; r < e; r++) {
                const uint32_t nr = m->nruns[r];
                const uint32_t base = m->hit_off[r];
                for (uint32_t i = 0; i < nr; i++) {
                    const ntl_run run = m->runs[base + i];
                    need += (size_t)(read_name_off[r + 1] - read_name_off[r]) + (size_t)(ctg_name_off[run.ctg + 1] - ctg_name_off[run.ctg]) +
                            16 + (size_t)run.count * 26;
                }
            }
            s.reserve(need);
        }
        for (uint64_t r = b; r < e; r++) {
            const uint32_t nr = m->nruns[r];
            if (!nr) continue;
            const uint32_t base = m->hit_off[r];
            for (uint32_t i = 0; i < nr; i++) {
                const ntl_run run = m->runs[base + i];
                s.append(read_names + read_name_off[r], (size_t)(read_name_off[r + 1] - read_name_off[r]));
                s.push_back('\t');
                s.append(ctg_names + ctg_name_off[run.ctg], (size_t)(ctg_name_off[run.ctg + 1] - ctg_name_off[run.ctg]));
                s.push_back('\t');
                put_u64(s, run.count);
                s.push_back('\t');
                for (uint32_t h = 0; h < run.count; h++) {
                    const ntl_hit hit = m->hits[base + run.start + h];
                    if (h) s.push_back(' ');
                    put_u64(s, hit.ctg_pos_strand & NTL_POS_MASK);
                    s.push_back(':');
                    s.push_back((hit.ctg_pos_strand & NTL_STRAND_BIT) ? '+' : '-');
                    s.push_back('_');
                    put_u64(s, hit.read_pos_strand & NTL_POS_MASK);
                    s.push_back(':');
                    s.push_back((hit.read_pos_strand & NTL_STRAND_BIT) ? '+' : '-');
                }
                s.push_back('\n');
            }
        }
    }, &nch);
    return join_out(parts, out_buf);
}

int64_t ntl_format_paf(const ntl_map_out* m, const char* read_names, const uint64_t* read_name_off,
                       const uint32_t* read_len, const char* ctg_names, const uint64_t* ctg_name_off,
                       const uint32_t* ctg_len, int k, int threads, char** out_buf) {
    if (!m || !read_names || !read_name_off || !read_len || !ctg_names || !ctg_name_off || !ctg_len || !out_buf)
        return NTL_ERR_ARG;
    int nch = 1;
    std::vector<std::string> parts = take_parts((size_t)(threads < 1 ? 1 : threads));
    std::vector<int> failed((size_t)(threads < 1 ? 1 : threads), 0);
    parallel_chunks(m->n_reads, threads, [&](int t, uint64_t b, uint64_t e) {
        // a thread formats into a string of its own: the string objects of `parts` sit side by side in memory, and every
        // push_back updates the size field -- two threads per cache line is false sharing that serialises them all
        std::string s = std::move(parts[(size_t)t]);
        struct PutBack { std::string& from; std::string& to; ~PutBack() { to = std::move(from); } } put_back{s, parts[(size_t)t]};
        std::vector<PHit> hits, srt;
        std::vector<std::vector<PHit>> blocks;
        for (uint64_t r = b; r < e; r++) {
            const uint32_t nr = m->nruns[r];
            if (!nr) continue;
            const uint32_t base = m->hit_off[r];
            for (uint32_t i = 0; i < nr; i++) {
                const ntl_run run = m->runs[base + i];
                hits.clear();
                for (uint32_t h = 0; h < run.count; h++) {
                    const ntl_hit hit = m->hits[base + run.start + h];
                    PHit p;
                    p.cpos = hit.ctg_pos_strand & NTL_POS_MASK; p.cfw = (hit.ctg_pos_strand & NTL_STRAND_BIT) != 0;
                    p.rpos = hit.read_pos_strand & NTL_POS_MASK; p.rfw = (hit.read_pos_strand & NTL_STRAND_BIT) != 0;
                    hits.push_back(p);
                }
                srt = hits;
                std::stable_sort(srt.begin(), srt.end(), [](const PHit& a, const PHit& c) {
                    return a.cpos != c.cpos ? a.cpos < c.cpos : a.rpos < c.rpos;
                });
                auto same = [](const PHit& a, const PHit& c) { return a.cpos == c.cpos && a.rpos == c.rpos; };
                // paf:95-101: detailed check only if the read order is neither the sorted order nor its reverse
                bool fwd_eq = true, rev_eq = true;
                for (size_t q = 0; q < hits.size(); q++) {
                    if (!same(hits[q], srt[q])) fwd_eq = false;
                    if (!same(hits[q], srt[hits.size() - 1 - q])) rev_eq = false;
                }
                blocks.clear();
                if (fwd_eq || rev_eq) blocks.push_back(srt);
                else if (!mapped_blocks(srt, blocks)) continue;
                for (const auto& blk : blocks) {
                    size_t same_strand = 0;
                    for (const auto& p : blk) same_strand += (p.cfw == p.rfw) ? 1 : 0;
                    const char strand = (2 * same_strand >= blk.size()) ? '+' : '-';
                    const PHit& f = blk.front(); const PHit& l = blk.back();
                    const uint64_t ts = std::min(f.cpos, l.cpos), te = (uint64_t)std::max(f.cpos, l.cpos) + (uint64_t)k;
                    const uint64_t qs = std::min(f.rpos, l.rpos), qe = (uint64_t)std::max(f.rpos, l.rpos) + (uint64_t)k;
                    if (!(qs < qe) || qe > read_len[r]) { failed[(size_t)t] = 1; return; }   // paf:127-129
                    s.append(read_names + read_name_off[r], (size_t)(read_name_off[r + 1] - read_name_off[r]));
                    s.push_back('\t'); put_u64(s, read_len[r]);
                    s.push_back('\t'); put_u64(s, qs);
                    s.push_back('\t'); put_u64(s, qe);
                    s.push_back('\t'); s.push_back(strand);
                    s.push_back('\t');
                    s.append(ctg_names + ctg_name_off[run.ctg], (size_t)(ctg_name_off[run.ctg + 1] - ctg_name_off[run.ctg]));
                    s.push_back('\t'); put_u64(s, ctg_len[run.ctg]);
                    s.push_back('\t'); put_u64(s, ts);
                    s.push_back('\t'); put_u64(s, te);
                    s.push_back('\t'); put_u64(s, blk.size());
                    s.push_back('\t'); put_u64(s, te - ts);
                    s.append("\t255\n");
                }
            }
        }
    }, &nch);
    for (int f : failed) if (f) { give_parts(parts); return NTL_ERR_ASSERT; }
    return join_out(parts, out_buf);
}

void ntl_buf_free(char* buf) { release_out(buf); }
void ntl_free(void* p) { if (!ntl_pool_release(p)) free(p); }

}  // extern "C"

// ------------------------------------------------------------------------------------------- reader
// Block reader: the file is consumed in 8 MiB blocks (read(2) for plain files, gzread for gzip, chosen by the magic
// bytes), lines are located with memchr inside the block and sequence lines are appended straight into the output
// buffer (no per-line strings). A line that straddles two blocks is the only case that is assembled separately.
namespace {

struct BigPool {
    std::mutex mu;
    std::unordered_map<void*, size_t> live;          // big buffers handed out (ptr -> capacity)
    std::vector<std::pair<void*, size_t>> parked;    // released, ready for reuse
    static constexpr size_t MAX_PARKED = 2, MAX_PARKED_BYTES = (size_t)6 << 30;
};
BigPool& big_pool() { static BigPool* p = new BigPool(); return *p; }     // never destroyed: buffers may outlive static teardown

void* pool_take(size_t want, size_t& cap_out) {
    BigPool& P = big_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    for (size_t i = 0; i < P.parked.size(); i++) {
        if (P.parked[i].second >= want && P.parked[i].second <= want * 2 + (64u << 20)) {
            void* p = P.parked[i].first;
            cap_out = P.parked[i].second;
            P.parked.erase(P.parked.begin() + (long)i);
            P.live[p] = cap_out;
            return p;
        }
    }
    return nullptr;
}
void pool_register(void* p, size_t cap) {
    BigPool& P = big_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = cap;
}

struct GrowBuf {                 // malloc'ed, geometrically growing byte buffer that is handed to the caller
    char* p = nullptr;
    size_t n = 0, cap = 0;
    bool reserve(size_t want) {
        if (want <= cap) return true;
        size_t c = cap ? cap : (1u << 20);
        while (c < want) c += c / 2 + (1u << 20);
        char* q = nullptr;
        if (!p && c >= (8u << 20)) {
            // first (hinted) allocation of a large batch: a recycled buffer if one is parked, else 2 MiB aligned memory
            // with transparent huge pages, so that filling it does not take one page fault per 4 KiB
            size_t pc = 0;
            void* a = pool_take(c, pc);
            if (a) { q = (char*)a; c = pc; }
            else if (posix_memalign(&a, 2u << 20, c) == 0) {
#ifdef MADV_HUGEPAGE
                madvise(a, c, MADV_HUGEPAGE);
#endif
                pool_register(a, c);
                q = (char*)a;
            }
        }
        if (!q) {
            bool pooled = false;
            if (p) { BigPool& P = big_pool(); std::lock_guard<std::mutex> lk(P.mu); pooled = P.live.erase(p) > 0; }
            (void)pooled;                                    // a pooled buffer that has to grow simply leaves the pool
            q = (char*)realloc(p, c);
        }
        if (!q) return false;
        p = q; cap = c;
        return true;
    }
    bool append(const char* src, size_t len) {
        if (!reserve(n + len + 64)) return false;
        memcpy(p + n, src, len);
        n += len;
        return true;
    }
};

}  // namespace

bool ntl_pool_release(void* p) {
    if (!p) return false;
    BigPool& P = big_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) return false;
    const size_t cap = it->second;
    P.live.erase(it);
    size_t bytes = cap;
    for (auto& b : P.parked) bytes += b.second;
    if (P.parked.size() < BigPool::MAX_PARKED && bytes <= BigPool::MAX_PARKED_BYTES) { P.parked.emplace_back(p, cap); return true; }
    free(p);
    return true;
}


// ---- BGZF (bgzip) input: independent gzip members of <= 64 KiB whose compressed size is in the header, so the members
// of the next stretch of the file are located by walking the headers and inflated by several threads straight into a
// window of decompressed text that the parallel FASTA / FASTQ parser then treats like a mapped plain file.
// Ordinary single-stream gzip has no such structure and stays on zlib's sequential gzread.
struct BgzfBlock { uint64_t coff; uint32_t bsize, hdr, isize; uint64_t out; };
struct BgzfSource {
    int fd = -1;
    const unsigned char* cmap = nullptr;   // the compressed file, mapped read-only
    size_t csize = 0, cpos = 0;            // file size, first block not walked yet
    char* w = nullptr;                     // window of decompressed bytes [wbase, wbase + wlen)
    size_t wcap = 0, wlen = 0;
    uint64_t wbase = 0;
    bool done = false, bad = false;
    int threads = 1;

    ~BgzfSource() {
        free(w);
        if (cmap) munmap((void*)cmap, csize);
        if (fd >= 0) ::close(fd);
    }
    // header of the member at p (n bytes available): total member size and payload offset; false = not a BGZF member
    static bool parse_header(const unsigned char* p, size_t n, uint32_t& bsize, uint32_t& hdr) {
        if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
        const uint32_t xlen = p[10] | (uint32_t)p[11] << 8;
        if (12 + (size_t)xlen > n) return false;
        bsize = 0;
        for (uint32_t at = 12; at + 4 <= 12 + xlen;) {
            const uint32_t slen = p[at + 2] | (uint32_t)p[at + 3] << 8;
            if (p[at] == 'B' && p[at + 1] == 'C' && slen == 2 && at + 6 <= 12 + xlen) bsize = (p[at + 4] | (uint32_t)p[at + 5] << 8) + 1;
            at += 4 + slen;
        }
        if (!bsize) return false;
        size_t h = 12 + xlen;
        if (p[3] & 8) { while (h < n && p[h]) h++; h++; }          // FNAME
        if (p[3] & 16) { while (h < n && p[h]) h++; h++; }         // FCOMMENT
        if (p[3] & 2) h += 2;                                      // FHCRC
        if (h + 8 > bsize) return false;
        hdr = (uint32_t)h;
        return true;
    }
    bool open(int file, size_t size, int nthreads) {
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, file, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, size, MADV_SEQUENTIAL);
        fd = file; cmap = (const unsigned char*)m; csize = size; threads = std::max(1, nthreads);
        return true;
    }
    // make [start, start + need) of the decompressed stream available at w (wbase becomes start); fewer bytes only at
    // the end of the file (done). false = corrupt input or out of memory.
    bool ensure(uint64_t start, uint64_t need);
};

struct ntl_seqfile {
    BgzfSource* bg = nullptr;    // BGZF input, inflated in parallel
    std::string path;
    gzFile gz = nullptr;         // gzip input (or stdin)
    int fd = -1;                 // plain file
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    bool eof = false, io_error = false;
    std::string carry;           // assembled line that straddled two blocks
    std::string pending;         // header line read ahead
    bool have_pending = false;
    // parallel mode (plain FASTA files): the reader works on byte ranges of the file, ppos = next unread record
    bool parallel = false;
    off_t ppos = 0, fsize = 0;
    int threads = 1;
    bool fastq = false;          // parallel mode on a file of 4-line FASTQ records
    const char* map = nullptr;   // the whole file, mapped read-only

    bool fill() {
        if (eof) return false;
        if (buf.empty()) buf.resize(8u << 20);
        long n;
        if (fd >= 0) {
            do { n = (long)::read(fd, buf.data(), buf.size()); } while (n < 0 && errno == EINTR);
        } else {
            n = gzread(gz, buf.data(), (unsigned)buf.size());
            if (n <= 0) {                          // a damaged or truncated gzip stream is an error, not the end of the file
                int errnum = Z_OK;
                gzerror(gz, &errnum);
                if (n < 0 || (errnum != Z_OK && errnum != Z_STREAM_END)) io_error = true;
            }
        }
        if (n < 0 && fd >= 0) io_error = true;
        if (n <= 0) { eof = true; pos = len = 0; return false; }
        pos = 0; len = (size_t)n;
        return true;
    }
    // next line without its terminator, as a view (valid until the next call); false at end of file
    bool getline(const char*& lp, size_t& ll) {
        if (pos >= len && !fill()) return false;
        const char* p = buf.data() + pos;
        const char* nl = (const char*)memchr(p, '\n', len - pos);
        if (nl) {
            lp = p; ll = (size_t)(nl - p);
            pos = (size_t)(nl - buf.data()) + 1;
        } else {                                   // the line continues in the next block(s)
            carry.assign(p, len - pos);
            pos = len;
            for (;;) {
                if (!fill()) break;
                const char* q = buf.data();
                const char* e = (const char*)memchr(q, '\n', len);
                if (e) { carry.append(q, (size_t)(e - q)); pos = (size_t)(e - q) + 1; break; }
                carry.append(q, len);
                pos = len;
            }
            lp = carry.data(); ll = carry.size();
        }
        if (ll && lp[ll - 1] == '\r') ll--;
        return true;
    }
};

namespace {

struct ParRec { uint64_t byte_start; uint64_t name_at; uint32_t name_len; uint64_t seq_len; uint64_t body_start; uint64_t body_end; };
struct ParSeg {
    GrowBuf names;
    std::vector<ParRec> recs;
    bool not_fasta = false, ok = true;
    ParSeg() = default;
    ParSeg(const ParSeg&) = delete;
    ~ParSeg() { ntl_free(names.p); }
};

// pass 1: the records of raw[b, e) (b is a record start or the first byte of the range): names and sequence lengths
void scan_segment(const char* raw, size_t b, size_t e, ParSeg& g) {
    size_t p = b;
    bool in_rec = false;
    g.ok = g.names.reserve(1u << 12);
    while (g.ok && p < e) {
        const char* nl = (const char*)memchr(raw + p, '\n', e - p);
        const size_t le = nl ? (size_t)(nl - raw) : e;          // end of the line (exclusive)
        size_t ll = le - p;
        if (ll && raw[p + ll - 1] == '\r') ll--;
        const char c0 = ll ? raw[p] : 0;
        if (c0 == '>') {
            size_t t = 1;
            while (t < ll && raw[p + t] != ' ' && raw[p + t] != '\t' && raw[p + t] != '\v' && raw[p + t] != '\f' && raw[p + t] != '\r') t++;
            ParRec r; r.byte_start = p; r.name_at = g.names.n; r.name_len = (uint32_t)(t - 1); r.seq_len = 0; r.body_start = le + 1; r.body_end = 0;
            g.ok = g.names.append(raw + p + 1, t - 1);
            g.recs.push_back(r);
            in_rec = true;
        } else if (c0 == '@' || c0 == '+') {
            g.not_fasta = true;                                   // FASTQ-like content: the sequential reader decides
            return;
        } else if (in_rec) {
            g.recs.back().seq_len += ll;
        }
        p = le + 1;
    }
}

// pass 2: the sequence lines of one record (raw[body, end)) without their terminators -> dst
void copy_body(const char* raw, size_t body, size_t end, char* dst) {
    size_t p = body;
    while (p < end) {
        const char* nl = (const char*)memchr(raw + p, '\n', end - p);
        const size_t le = nl ? (size_t)(nl - raw) : end;
        size_t ll = le - p;
        if (ll && raw[p + ll - 1] == '\r') ll--;
        memcpy(dst, raw + p, ll);
        dst += ll;
        p = le + 1;
    }
}

// next record start ("\n>" + 1) at or after position p, or e
size_t next_record_start(const char* raw, size_t p, size_t e) {
    if (p == 0) return 0;
    while (p < e) {
        const char* nl = (const char*)memchr(raw + p - 1, '\n', e - (p - 1));
        if (!nl) return e;
        const size_t q = (size_t)(nl - raw) + 1;
        if (q >= e) return e;
        if (raw[q] == '>') return q;
        p = q + 1;
    }
    return e;
}


// ---- FASTQ (4 lines per record) -------------------------------------------------------------------------------------
// line [p, le) of raw[.., e): le = index of its '\n' (or e); returns the length without a trailing '\r'
inline size_t line_at(const char* raw, size_t p, size_t e, size_t& le) {
    const char* nl = (const char*)memchr(raw + p, '\n', e - p);
    le = nl ? (size_t)(nl - raw) : e;
    size_t ll = le - p;
    if (ll && raw[p + ll - 1] == '\r') ll--;
    return ll;
}

// Is the line starting at q the header of a well-formed 4-line record ('@', sequence, '+', as many quality characters)?
// A quality line may start with '@' too, so the structure is what identifies a header. Fills the record on success.
bool fastq_record_at(const char* raw, size_t q, size_t e, ParRec& r, size_t& next) {
    if (q >= e || raw[q] != '@') return false;
    size_t l1e, l2e, l3e, l4e;
    const size_t l1 = line_at(raw, q, e, l1e);
    if (l1e >= e) return false;
    const size_t l2 = line_at(raw, l1e + 1, e, l2e);
    if (l2e >= e) return false;
    const size_t l3 = line_at(raw, l2e + 1, e, l3e);
    if (l3 == 0 || raw[l2e + 1] != '+' || l3e >= e) return false;
    const size_t l4 = line_at(raw, l3e + 1, e, l4e);
    if (l4 != l2) return false;
    if (l2 && (raw[l1e + 1] == '@' || raw[l1e + 1] == '+' || raw[l1e + 1] == '>')) return false;   // the sequential reader would see a header / separator
    size_t t = 1;
    while (t < l1 && raw[q + t] != ' ' && raw[q + t] != '\t' && raw[q + t] != '\v' && raw[q + t] != '\f' && raw[q + t] != '\r') t++;
    r.byte_start = q; r.name_at = 0; r.name_len = (uint32_t)(t - 1); r.seq_len = l2; r.body_start = l1e + 1; r.body_end = l2e + 1;
    next = l4e < e ? l4e + 1 : e;
    return true;
}

// first record header at a line start >= p
size_t next_fastq_start(const char* raw, size_t p, size_t e) {
    size_t q = p;
    if (q > 0 && raw[q - 1] != '\n') {                               // move to the next line start
        const char* nl = (const char*)memchr(raw + q, '\n', e - q);
        if (!nl) return e;
        q = (size_t)(nl - raw) + 1;
    }
    ParRec r;
    size_t next;
    while (q < e) {
        if (fastq_record_at(raw, q, e, r, next)) return q;
        const char* nl = (const char*)memchr(raw + q, '\n', e - q);
        if (!nl) return e;
        q = (size_t)(nl - raw) + 1;
    }
    return e;
}

// pass 1 for FASTQ: strictly 4-line records from b to e; anything else hands the file over to the sequential reader
void scan_segment_fastq(const char* raw, size_t b, size_t e, size_t file_end, ParSeg& g) {
    g.ok = g.names.reserve(1u << 12);
    size_t p = b;
    while (g.ok && p < e) {
        ParRec r;
        size_t next;
        if (!fastq_record_at(raw, p, file_end, r, next)) {
            // blank lines between records are fine (the sequential reader skips them); anything else is not
            size_t le;
            if (line_at(raw, p, e, le) == 0) { p = le + 1; continue; }
            g.not_fasta = true;
            return;
        }
        r.name_at = g.names.n;
        g.ok = g.names.append(raw + p + 1, r.name_len);
        g.recs.push_back(r);
        p = next;
    }
}

template <class F>
void run_threads(size_t n, F&& body) {
    std::vector<std::thread> th;
    for (size_t g = 1; g < n; g++) th.emplace_back([&, g]() { body(g); });
    if (n) body(0);
    for (auto& x : th) x.join();
}

}  // namespace

bool BgzfSource::ensure(uint64_t start, uint64_t need) {
    if (bad || start < wbase || start > wbase + wlen) return false;
    if (start > wbase) {                                           // drop what the parser has consumed
        const size_t drop = (size_t)(start - wbase);
        memmove(w, w + drop, wlen - drop);
        wlen -= drop; wbase = start;
    }
    while (!done && wlen < need) {
        // walk the headers of the next stretch: enough members for what is missing, at least 32 MiB of output per pass
        const uint64_t goal = std::max<uint64_t>(need - wlen, 32u << 20);
        std::vector<BgzfBlock> blk;
        uint64_t out = 0;
        while (cpos < csize && out < goal) {
            BgzfBlock b;
            if (!parse_header(cmap + cpos, csize - cpos, b.bsize, b.hdr) || cpos + b.bsize > csize) { bad = true; return false; }
            const unsigned char* tail = cmap + cpos + b.bsize - 4;
            b.isize = tail[0] | (uint32_t)tail[1] << 8 | (uint32_t)tail[2] << 16 | (uint32_t)tail[3] << 24;
            b.coff = cpos; b.out = out;
            out += b.isize;
            cpos += b.bsize;
            blk.push_back(b);
        }
        if (cpos >= csize) done = true;
        if (wlen + out + 64 > wcap) {
            size_t c = wcap ? wcap : (64u << 20);
            while (c < wlen + out + 64) c += c / 2;
            char* q = (char*)realloc(w, c);
            if (!q) { bad = true; return false; }
            w = q; wcap = c;
        }
        const size_t nt = (size_t)std::min<size_t>((size_t)threads, std::max<size_t>(1, blk.size() / 16));
        std::vector<int> failed(nt, 0);
        char* dst = w + wlen;
        run_threads(nt, [&](size_t t) {
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { failed[t] = 1; return; }
            const size_t b0 = blk.size() * t / nt, b1 = blk.size() * (t + 1) / nt;
            for (size_t i = b0; i < b1 && !failed[t]; i++) {
                const BgzfBlock& b = blk[i];
                const unsigned char* src = cmap + b.coff;
                zs.next_in = (Bytef*)(src + b.hdr);
                zs.avail_in = b.bsize - b.hdr - 8;
                zs.next_out = (Bytef*)(dst + b.out);
                zs.avail_out = b.isize;
                const int rc = inflate(&zs, Z_FINISH);
                const unsigned char* c = src + b.bsize - 8;
                const uint32_t crc = c[0] | (uint32_t)c[1] << 8 | (uint32_t)c[2] << 16 | (uint32_t)c[3] << 24;
                if (rc != Z_STREAM_END || zs.avail_out != 0 || (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)(dst + b.out), b.isize) != crc)
                    failed[t] = 1;
                inflateReset(&zs);
            }
            inflateEnd(&zs);
        });
        for (int x : failed) if (x) { bad = true; return false; }
        wlen += (size_t)out;
    }
    return true;
}

namespace {

// Parallel FASTA batch straight from the mapped file: pass 1 finds the records of every segment, pass 2 copies the
// sequence lines to their final place. Returns 1 = batch produced, 0 = not applicable (the caller uses the sequential
// reader from f->ppos), -1 = out of memory.
int read_parallel(ntl_seqfile* f, uint64_t max_bases, GrowBuf& seq, GrowBuf& names, std::vector<uint64_t>& offs,
                  std::vector<uint64_t>& noffs) {
    const off_t start = f->ppos;
    const bool bz = f->bg != nullptr;
    const uint64_t margin = 32u << 20;                             // look-ahead past the batch: the record that crosses its end
    if (!bz) {
        if (start >= f->fsize) return 1;                           // end of file: empty batch
        if (!f->map) {
            void* m = mmap(nullptr, (size_t)f->fsize, PROT_READ, MAP_PRIVATE, f->fd, 0);
            if (m == MAP_FAILED) { f->parallel = false; return 0; }
            f->map = (const char*)m;
            madvise(m, (size_t)f->fsize, MADV_SEQUENTIAL);
        }
    }
    // the text the parser sees: raw[0, limit); final = nothing follows it
    const char* raw = nullptr;
    uint64_t limit = 0;
    bool final = true;
    auto view = [&](uint64_t need) -> bool {
        if (!bz) { raw = f->map + start; limit = (uint64_t)(f->fsize - start); return true; }
        if (!f->bg->ensure((uint64_t)start, std::min<uint64_t>(need, 1ull << 60) + margin)) return false;
        raw = f->bg->w; limit = f->bg->wlen; final = f->bg->done;
        return true;
    };
    uint64_t want = max_bases ? max_bases + max_bases / 16 + (16u << 20) : (bz ? 1ull << 60 : (uint64_t)(f->fsize - start));
    if (!view(want)) return -1;
    if (limit == 0 && final) return 1;
    size_t n = 0, cut = 0;
    for (;;) {
        n = (size_t)std::min<uint64_t>(want, limit);
        if (final && n >= limit) { cut = n; break; }                // the rest of the file: every record is complete
        // cut at the last record start in the range; a range without one (a record larger than the range) grows
        size_t q = n;
        cut = 0;
        if (f->fastq) {
            const size_t total = (size_t)limit;
            for (size_t window = 4u << 20; cut == 0; window *= 4) {
                size_t p = n > window ? n - window : 0;
                p = next_fastq_start(raw, p, total);
                while (p < n) {                                       // walk the records of the window; keep the last start below n
                    ParRec r;
                    size_t next;
                    if (!fastq_record_at(raw, p, total, r, next)) { p = next_fastq_start(raw, p + 1, total); continue; }
                    if (p > 0) cut = p;
                    p = next;
                }
                if (n <= window) break;
            }
            if (cut > 0) break;
            want *= 2;
            if (!view(want)) return -1;
            continue;
        }
        while (q > 1) {
            const char* gt = (const char*)memrchr(raw, '>', q);
            if (!gt) break;
            const size_t at = (size_t)(gt - raw);
            if (at > 0 && raw[at - 1] == '\n') { cut = at; break; }
            q = at;
        }
        if (cut > 0) break;
        want *= 2;
        if (!view(want)) return -1;
    }
    // segments at record starts
    std::vector<size_t> bounds;
    bounds.push_back(0);
    const size_t file_end = (size_t)limit;
    for (int t = 1; t < f->threads; t++) {
        const size_t at = (size_t)((uint64_t)cut * t / f->threads);
        const size_t b = f->fastq ? next_fastq_start(raw, at, cut) : next_record_start(raw, at, cut);
        if (b > bounds.back() && b < cut) bounds.push_back(b);
    }
    bounds.push_back(cut);
    const size_t nseg = bounds.size() - 1;
    std::unique_ptr<ParSeg[]> segs(new ParSeg[nseg]);
    run_threads(nseg, [&](size_t g) {
        if (f->fastq) scan_segment_fastq(raw, bounds[g], bounds[g + 1], file_end, segs[g]);
        else scan_segment(raw, bounds[g], bounds[g + 1], segs[g]);
    });
    for (size_t g = 0; g < nseg; g++) {
        if (!segs[g].ok) return -1;
        if (segs[g].not_fasta) { f->parallel = false; return 0; }
    }
    // how many records make the batch (bin/read_fasta.py semantics: whole records until max_bases is reached)
    uint64_t total = 0, names_total = 0;
    off_t next_pos = start + (off_t)cut;
    size_t last_seg = nseg, last_cnt = 0;
    bool full = false;
    for (size_t g = 0; g < nseg && !full; g++) {
        for (size_t r = 0; r < segs[g].recs.size(); r++) {
            if (max_bases && total >= max_bases) {
                next_pos = start + (off_t)segs[g].recs[r].byte_start;
                last_seg = g; last_cnt = r; full = true;
                break;
            }
            total += segs[g].recs[r].seq_len; names_total += segs[g].recs[r].name_len;
        }
    }
    if (!seq.reserve(total + 64) || !names.reserve(names_total + 1)) return -1;
    // offsets and names, then the sequence bytes of every segment in parallel, straight to their final place
    std::vector<uint64_t> seg_seq_at(nseg, 0);
    std::vector<size_t> seg_cnt(nseg, 0);
    uint64_t so = 0, no = 0;
    for (size_t g = 0; g < nseg; g++) {
        const size_t cnt = full ? (g < last_seg ? segs[g].recs.size() : g == last_seg ? last_cnt : 0) : segs[g].recs.size();
        seg_seq_at[g] = so; seg_cnt[g] = cnt;
        for (size_t r = 0; r < cnt; r++) {
            const ParRec& rec = segs[g].recs[r];
            memcpy(names.p + no, segs[g].names.p + rec.name_at, rec.name_len);
            no += rec.name_len;
            noffs.push_back(no);
            so += rec.seq_len;
            offs.push_back(so);
        }
    }
    names.n = no;
    run_threads(nseg, [&](size_t g) {
        uint64_t at = seg_seq_at[g];
        for (size_t r = 0; r < seg_cnt[g]; r++) {
            const ParRec& rec = segs[g].recs[r];
            const size_t end = rec.body_end ? (size_t)rec.body_end
                                            : r + 1 < segs[g].recs.size() ? (size_t)segs[g].recs[r + 1].byte_start : bounds[g + 1];
            copy_body(raw, (size_t)rec.body_start, end, seq.p + at);
            at += rec.seq_len;
        }
    });
    seq.n = so;
    f->ppos = next_pos;
    return 1;
}

}  // namespace

extern "C" {

int ntl_seqfile_open(const char* path, ntl_seqfile** out) {
    if (!path || !out) return NTL_ERR_ARG;
    ntl_seqfile* f = new ntl_seqfile();
    if (strcmp(path, "-") == 0) {
        f->gz = gzdopen(0, "rb");                  // zlib passes plain data through
    } else {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) { delete f; return NTL_ERR_ARG; }
        unsigned char magic[2] = {0, 0};
        const long got = (long)::pread(fd, magic, 2, 0);
        const char* env = getenv("NTL_READER_THREADS");
        const unsigned hw = std::thread::hardware_concurrency();
        const int nthreads = env ? atoi(env) : (int)std::max(1u, std::min(8u, hw / 2));
        f->path = path;
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            // bgzip output (BGZF) is inflated by several threads; any other gzip goes through zlib's sequential reader
            unsigned char head[64];
            const long hn = (long)::pread(fd, head, sizeof head, 0);
            uint32_t bsize = 0, hdr = 0;
            struct stat sb;
            if (nthreads > 1 && hn >= 18 && BgzfSource::parse_header(head, (size_t)hn, bsize, hdr) && fstat(fd, &sb) == 0 &&
                S_ISREG(sb.st_mode) && sb.st_size > 0) {
                f->bg = new BgzfSource();
                if (f->bg->open(fd, (size_t)sb.st_size, env ? nthreads : (int)std::max(2u, std::min(16u, hw / 2))) && f->bg->ensure(0, 1) && f->bg->wlen > 0 &&
                    (f->bg->w[0] == '>' || f->bg->w[0] == '@')) {
                    f->parallel = true; f->threads = nthreads; f->ppos = 0; f->fastq = f->bg->w[0] == '@';
                } else {                                   // not something the parallel parser takes: sequential gzip
                    if (f->bg->fd < 0) ::close(fd);
                    delete f->bg;
                    f->bg = nullptr;
                    f->gz = gzopen(path, "rb");
                }
            } else {
                ::close(fd);
                f->gz = gzopen(path, "rb");
            }
        } else {
            f->fd = fd;
#ifdef POSIX_FADV_SEQUENTIAL
            posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
            // a regular file that starts like FASTA is parsed by several threads (ntl_seqfile_read falls back to the
            // sequential reader the moment it meets something that is not plain FASTA)
            struct stat sb;
            f->threads = nthreads;
            if (f->threads > 1 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && got >= 1 && (magic[0] == '>' || magic[0] == '@')) {
                f->parallel = true; f->fsize = sb.st_size; f->ppos = 0; f->fastq = magic[0] == '@';
            }
        }
    }
    if (!f->gz && f->fd < 0 && !f->bg) { delete f; return NTL_ERR_ARG; }
    if (f->gz) gzbuffer(f->gz, 1u << 20);
    *out = f;
    return NTL_OK;
}

int ntl_seqfile_read(ntl_seqfile* f, uint64_t max_bases, char** seq_out, uint64_t** offsets_out, char** names_out,
                     uint64_t** name_off_out, uint32_t* nseq_out) {
    if (!f || !seq_out || !offsets_out || !names_out || !name_off_out || !nseq_out) return NTL_ERR_ARG;
    GrowBuf seq, names;
    std::vector<uint64_t> offs(1, 0), noffs(1, 0);
    const char* lp = nullptr;
    size_t ll = 0;
    if (f->parallel) {
        const int pr = read_parallel(f, max_bases, seq, names, offs, noffs);
        if (pr < 0) { ntl_free(seq.p); ntl_free(names.p); return NTL_ERR_ARG; }
        if (pr == 1) {
            const uint32_t nseq_p = (uint32_t)(offs.size() - 1);
            uint64_t* o = (uint64_t*)malloc(offs.size() * 8);
            uint64_t* no = (uint64_t*)malloc(noffs.size() * 8);
            if (!o || !no || !seq.reserve(seq.n + 64) || !names.reserve(names.n + 1)) { ntl_free(seq.p); ntl_free(names.p); free(o); free(no); return NTL_ERR_ARG; }
            memset(seq.p + seq.n, 'N', 64);
            memcpy(o, offs.data(), offs.size() * 8);
            memcpy(no, noffs.data(), noffs.size() * 8);
            *seq_out = seq.p; *offsets_out = o; *names_out = names.p; *name_off_out = no; *nseq_out = nseq_p;
            return NTL_OK;
        }
        // not plain FASTA after all: continue sequentially from the first unread record
        ntl_free(seq.p); ntl_free(names.p); seq = GrowBuf(); names = GrowBuf();
        offs.assign(1, 0); noffs.assign(1, 0);
        if (f->bg) {                                   // BGZF text that is not plain FASTA / 4-line FASTQ: zlib reads on from there
            delete f->bg;
            f->bg = nullptr;
            f->gz = gzopen(f->path.c_str(), "rb");
            if (f->gz) gzbuffer(f->gz, 1u << 20);
            if (!f->gz || gzseek(f->gz, (z_off_t)f->ppos, SEEK_SET) < 0) return NTL_ERR_ARG;
        } else {
            lseek(f->fd, f->ppos, SEEK_SET);
        }
        f->pos = f->len = 0; f->eof = false; f->have_pending = false;
    }
    // plain files: what is left of the file bounds the batch, so the buffer never has to grow
    size_t hint = 1u << 20;
    if (f->fd >= 0) {
        struct stat sb;
        const off_t at = lseek(f->fd, 0, SEEK_CUR);
        if (fstat(f->fd, &sb) == 0 && at >= 0 && sb.st_size > at) hint = (size_t)(sb.st_size - at) + (f->len - f->pos) + 4096;
        if (max_bases && hint > max_bases + (64u << 20)) hint = (size_t)max_bases + (64u << 20);
    }
    bool ok = seq.reserve(hint) && names.reserve(1u << 12);
    while (ok) {
        if (max_bases && seq.n >= max_bases) break;
        // find the next header (bin/read_fasta.py:10-16)
        if (!f->have_pending) {
            bool found = false;
            while (f->getline(lp, ll)) {
                if (ll && (lp[0] == '>' || lp[0] == '@')) { found = true; break; }
            }
            if (!found) break;
            f->pending.assign(lp, ll);
        }
        f->have_pending = false;
        const std::string& hdr = f->pending;
        size_t e = 1;
        while (e < hdr.size() && hdr[e] != ' ' && hdr[e] != '\t' && hdr[e] != '\v' && hdr[e] != '\f' && hdr[e] != '\r') e++;
        ok = names.append(hdr.data() + 1, e - 1);
        noffs.push_back(names.n);
        const size_t start = seq.n;
        bool plus = false;
        while (ok && f->getline(lp, ll)) {
            if (ll && (lp[0] == '>' || lp[0] == '@')) { f->pending.assign(lp, ll); f->have_pending = true; break; }
            if (ll && lp[0] == '+') { plus = true; break; }
            ok = seq.append(lp, ll);
        }
        if (plus) {                           // FASTQ: as many quality characters as bases (read_fasta.py:36-43)
            size_t q = 0;
            const size_t need = seq.n - start;
            while (q < need && f->getline(lp, ll)) q += ll;
        }
        offs.push_back(seq.n);
    }
    const uint32_t nseq = (uint32_t)(offs.size() - 1);
    uint64_t* o = (uint64_t*)malloc(offs.size() * 8);
    uint64_t* no = (uint64_t*)malloc(noffs.size() * 8);
    if (!ok || f->io_error || !o || !no || !seq.reserve(seq.n + 64) || !names.reserve(names.n + 1)) {
        ntl_free(seq.p); ntl_free(names.p); free(o); free(no);
        return NTL_ERR_ARG;
    }
    memset(seq.p + seq.n, 'N', 64);
    memcpy(o, offs.data(), offs.size() * 8);
    memcpy(no, noffs.data(), noffs.size() * 8);
    *seq_out = seq.p; *offsets_out = o; *names_out = names.p; *name_off_out = no; *nseq_out = nseq;
    return NTL_OK;
}

void ntl_seqfile_close(ntl_seqfile* f) {
    if (!f) return;
    if (f->gz) gzclose(f->gz);
    delete f->bg;
    if (f->map) munmap((void*)f->map, (size_t)f->fsize);
    if (f->fd >= 0) ::close(f->fd);
    delete f;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------- verbose_mapping.tsv parser
// The checkpoint path and the liftover both start from a verbose_mapping.tsv (bin/ntlink_pair.py:437-488,
// bin/ntlink_liftover_mappings.py:128-147), which is gigabytes of text at genome scale; the reference splits it line by
// line in Python. This is the same parse as ntlink_b200/pair.py::parse_verbose_mappings, natively.
#include <unordered_map>

struct ntl_verbose_file {
    ntl_seqfile* src = nullptr;                       // reuses the block reader (plain or gzip)
    std::unordered_map<std::string, uint32_t> ctg_id;
    std::vector<std::string> ctg_names;
    bool fixed_table = false;                         // ids come from the caller's table: unknown names are an error
    std::string pend_line;                            // first line of the next read block
    bool have_pend = false, done = false;
    std::string err;
};

namespace {

template <class T> T* vec_to_malloc(const std::vector<T>& v) {
    T* p = (T*)malloc((v.size() + 1) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

// one token "ctgpos:ctgstrand_readpos:readstrand" -> (cposf, rposf); false on malformed input
bool parse_token(const char* p, const char* e, uint32_t& cposf, uint32_t& rposf) {
    uint64_t a = 0, b = 0;
    const char* q = p;
    if (q >= e || *q < '0' || *q > '9') return false;
    while (q < e && *q >= '0' && *q <= '9') a = a * 10 + (uint64_t)(*q++ - '0');
    if (q + 3 >= e || q[0] != ':' || (q[1] != '+' && q[1] != '-') || q[2] != '_') return false;
    const bool cs = q[1] == '+';
    q += 3;
    if (q >= e || *q < '0' || *q > '9') return false;
    while (q < e && *q >= '0' && *q <= '9') b = b * 10 + (uint64_t)(*q++ - '0');
    if (q + 2 != e || q[0] != ':' || (q[1] != '+' && q[1] != '-')) return false;
    const bool rs = q[1] == '+';
    if (a > 0x7FFFFFFFull || b > 0x7FFFFFFFull) return false;
    cposf = (uint32_t)a | (cs ? 0x80000000u : 0u);
    rposf = (uint32_t)b | (rs ? 0x80000000u : 0u);
    return true;
}

}  // namespace

extern "C" {

int ntl_verbose_open(const char* path, const char* ctg_names, const uint64_t* ctg_name_off, uint32_t ncontig, ntl_verbose_file** out) {
    if (!path || !out || (ncontig && (!ctg_names || !ctg_name_off))) return NTL_ERR_ARG;
    ntl_verbose_file* f = new ntl_verbose_file();
    if (ntl_seqfile_open(path, &f->src) != NTL_OK) { delete f; return NTL_ERR_ARG; }
    f->src->parallel = false;                         // line-oriented use of the block reader
    if (f->src->fd >= 0) lseek(f->src->fd, 0, SEEK_SET);
    if (ncontig) {
        f->fixed_table = true;
        for (uint32_t i = 0; i < ncontig; i++) {
            std::string n(ctg_names + ctg_name_off[i], (size_t)(ctg_name_off[i + 1] - ctg_name_off[i]));
            f->ctg_id[n] = i;                         // a repeated name: the last one wins, like a dict built from a FASTA
            f->ctg_names.push_back(n);
        }
    }
    *out = f;
    return NTL_OK;
}

const char* ntl_verbose_error(ntl_verbose_file* f) { return f ? f->err.c_str() : ""; }

// Reads whole read blocks until about max_hits hits are collected (0 = to the end of the file).
int ntl_verbose_read(ntl_verbose_file* f, uint64_t max_hits, int share_repeated, ntl_mappings_out* out) {
    if (!f || !out) return NTL_ERR_ARG;
    memset(out, 0, sizeof *out);
    std::vector<uint32_t> hit_off(1, 0), nruns, read_len, runs, hits;           // runs/hits: 3 words per entry
    std::vector<uint64_t> rname_off(1, 0);
    std::string rnames;
    struct Line { uint32_t ctg; uint32_t start, count; };
    std::vector<Line> block;                                                   // lines of the current read
    std::string cur_read;
    bool have_read = false;
    uint32_t base = 0;

    auto flush = [&]() {
        if (!have_read) return;
        // share_repeated (pair:470-472): a contig listed twice refers both times to its LAST listing
        std::vector<Line> eff = block;
        if (share_repeated) {
            std::unordered_map<uint32_t, size_t> last;
            for (size_t i = 0; i < block.size(); i++) last[block[i].ctg] = i;
            for (size_t i = 0; i < block.size(); i++) eff[i] = block[last[block[i].ctg]];
        }
        uint32_t maxpos = 0;
        for (const Line& l : eff) {
            runs.push_back(l.ctg); runs.push_back(l.start); runs.push_back(l.count);
            const uint32_t first = hits[3 * (size_t)(base + l.start) + 2] & 0x7FFFFFFFu;
            const uint32_t last = hits[3 * (size_t)(base + l.start + l.count - 1) + 2] & 0x7FFFFFFFu;
            maxpos = std::max(maxpos, std::max(first, last));
        }
        const uint32_t nh = (uint32_t)(hits.size() / 3) - base, nr = (uint32_t)block.size();
        const uint32_t width = std::max(nh, nr);                               // holey layout: both arrays get `width` slots
        hits.resize(3 * (size_t)(base + width), 0);
        runs.resize(3 * (size_t)(base + width), 0);
        nruns.push_back(nr);
        read_len.push_back(maxpos);
        hit_off.push_back(base + width);
        rnames += cur_read;
        rname_off.push_back(rnames.size());
        base += width;
        block.clear();
        have_read = false;
    };

    const char* lp = nullptr;
    size_t ll = 0;
    std::string line;
    while (!f->done) {
        if (f->have_pend) { line.swap(f->pend_line); f->have_pend = false; }
        else if (f->src->getline(lp, ll)) line.assign(lp, ll);
        else if (f->src->io_error) { f->err = "damaged or truncated gzip input"; return NTL_ERR_ARG; }
        else { f->done = true; break; }
        // line.strip().split('\t') -> exactly 4 fields (pair:449)
        size_t b = 0, e = line.size();
        while (b < e && (unsigned char)line[b] <= ' ') b++;
        while (e > b && (unsigned char)line[e - 1] <= ' ') e--;
        if (b == e) { f->err = "empty line in the mappings file"; return NTL_ERR_ARG; }
        const char* s = line.data();
        const char* t1 = (const char*)memchr(s + b, '\t', e - b);
        const char* t2 = t1 ? (const char*)memchr(t1 + 1, '\t', (size_t)(s + e - (t1 + 1))) : nullptr;
        const char* t3 = t2 ? (const char*)memchr(t2 + 1, '\t', (size_t)(s + e - (t2 + 1))) : nullptr;
        if (!t3 || memchr(t3 + 1, '\t', (size_t)(s + e - (t3 + 1)))) { f->err = "a mappings line must have 4 tab-separated fields"; return NTL_ERR_ARG; }
        const size_t rl = (size_t)(t1 - (s + b));
        const bool same = have_read && cur_read.size() == rl && memcmp(cur_read.data(), s + b, rl) == 0;
        if (!same) {
            if (have_read && max_hits && hits.size() / 3 >= max_hits) {        // batch is full: this line starts the next call
                f->pend_line = line; f->have_pend = true;
                break;
            }
            flush();
            cur_read.assign(s + b, rl);
            have_read = true;
        }
        std::string cname(t1 + 1, (size_t)(t2 - (t1 + 1)));
        uint32_t cid;
        auto it = f->ctg_id.find(cname);
        if (it != f->ctg_id.end()) cid = it->second;
        else if (f->fixed_table) { f->err = "contig " + cname + " of the mappings file is not in the target"; return NTL_ERR_ARG; }
        else { cid = (uint32_t)f->ctg_names.size(); f->ctg_id.emplace(cname, cid); f->ctg_names.push_back(cname); }
        Line l; l.ctg = cid; l.start = (uint32_t)(hits.size() / 3) - base; l.count = 0;
        const char* p = t3 + 1;
        const char* end = s + e;
        while (p < end) {
            const char* sp = (const char*)memchr(p, ' ', (size_t)(end - p));
            const char* te = sp ? sp : end;
            uint32_t cposf, rposf;
            if (!parse_token(p, te, cposf, rposf)) { f->err = "malformed minimizer position in the mappings file: " + std::string(p, (size_t)(te - p)); return NTL_ERR_ARG; }
            hits.push_back(cid); hits.push_back(cposf); hits.push_back(rposf);
            l.count++;
            p = te + 1;
        }
        if (l.count == 0) { f->err = "a mappings line without minimizer positions"; return NTL_ERR_ARG; }
        block.push_back(l);
    }
    flush();
    out->n_reads = (uint32_t)nruns.size();
    out->n_slots = hit_off.back();
    out->hit_off = vec_to_malloc(hit_off); out->nruns = vec_to_malloc(nruns); out->read_len = vec_to_malloc(read_len);
    out->runs = vec_to_malloc(runs); out->hits = vec_to_malloc(hits);
    out->read_names = (char*)malloc(rnames.size() + 1);
    if (out->read_names) memcpy(out->read_names, rnames.data(), rnames.size());
    out->read_name_off = vec_to_malloc(rname_off);
    // the contig table seen so far (all of it: ids are stable across calls)
    std::string cn;
    std::vector<uint64_t> co(1, 0);
    for (const std::string& n : f->ctg_names) { cn += n; co.push_back(cn.size()); }
    out->n_contigs = (uint32_t)f->ctg_names.size();
    out->ctg_names = (char*)malloc(cn.size() + 1);
    if (out->ctg_names) memcpy(out->ctg_names, cn.data(), cn.size());
    out->ctg_name_off = vec_to_malloc(co);
    if (!out->hit_off || !out->nruns || !out->read_len || !out->runs || !out->hits || !out->read_names || !out->read_name_off ||
        !out->ctg_names || !out->ctg_name_off) { f->err = "out of memory"; return NTL_ERR_ARG; }
    return NTL_OK;
}

void ntl_verbose_close(ntl_verbose_file* f) {
    if (!f) return;
    ntl_seqfile_close(f->src);
    delete f;
}

// ------------------------------------------------------------------------------------------- indexlr TSV parser
// The reference's own read interface is text: `indexlr --long --pos --strand [--len] ... | ntlink_pair.py ... -`
// (ntLink:221-225); ntlink_pair.py splits every line and every token in Python (bin/ntlink_pair.py:196-203,355-366), which
// is where its time goes. Same parse, natively, in bounded batches (a 30x human read set is ~10^9 minimizers of text).
//   target TSV:  name \t tok( tok)*          reads TSV:  name \t length \t tok( tok)*      tok = hash:pos:strand
// Lines are strip()ped and split at tabs like the reference does; records with fewer fields are kept as empty sketches
// (the reference skips them, pair:200,357 -- they cannot hit anything either way).
struct ntl_tsv_file {
    ntl_seqfile* src = nullptr;
    bool with_len = false, done = false;
    std::string err;
};

int ntl_tsv_open(const char* path, int with_len, ntl_tsv_file** out) {
    if (!path || !out) return NTL_ERR_ARG;
    ntl_tsv_file* f = new ntl_tsv_file();
    if (ntl_seqfile_open(path, &f->src) != NTL_OK) { delete f; return NTL_ERR_ARG; }
    f->src->parallel = false;
    if (f->src->fd >= 0) lseek(f->src->fd, 0, SEEK_SET);
    f->with_len = with_len != 0;
    *out = f;
    return NTL_OK;
}
const char* ntl_tsv_error(ntl_tsv_file* f) { return f ? f->err.c_str() : ""; }

int ntl_tsv_read(ntl_tsv_file* f, uint64_t max_mx, ntl_tsv_out* out) {
    if (!f || !out) return NTL_ERR_ARG;
    memset(out, 0, sizeof *out);
    std::vector<uint64_t> hash, mx_off(1, 0), name_off(1, 0);
    std::vector<uint32_t> posf, lens;
    std::string names;
    const char* lp = nullptr;
    size_t ll = 0;
    while (!f->done && !(max_mx && hash.size() >= max_mx)) {
        if (!f->src->getline(lp, ll)) {
            if (f->src->io_error) { f->err = "damaged or truncated gzip input"; return NTL_ERR_ARG; }
            f->done = true;
            break;
        }
        const char* b = lp;
        const char* e = lp + ll;
        while (b < e && (unsigned char)*b <= ' ') b++;
        while (e > b && (unsigned char)e[-1] <= ' ') e--;
        if (b == e) continue;
        const char* t1 = (const char*)memchr(b, '\t', (size_t)(e - b));
        const char* name_end = t1 ? t1 : e;
        names.append(b, (size_t)(name_end - b));
        name_off.push_back(names.size());
        const char* p = t1 ? t1 + 1 : e;
        if (f->with_len) {
            uint64_t L = 0;
            const char* t2 = p < e ? (const char*)memchr(p, '\t', (size_t)(e - p)) : nullptr;
            const char* le = t2 ? t2 : e;
            for (const char* q = p; q < le; q++) {
                if (*q < '0' || *q > '9') { f->err = "malformed length column in the sketch TSV"; return NTL_ERR_ARG; }
                L = L * 10 + (uint64_t)(*q - '0');
            }
            if (L > 0xFFFFFFFFull) { f->err = "sequence length does not fit 32 bits"; return NTL_ERR_ARG; }
            lens.push_back((uint32_t)L);
            p = t2 ? t2 + 1 : e;
        }
        // tokens up to the next tab (further columns are ignored, like the reference's line[1] / line[2])
        const char* col_end = p < e ? (const char*)memchr(p, '\t', (size_t)(e - p)) : nullptr;
        if (!col_end) col_end = e;
        while (p < col_end) {
            while (p < col_end && *p == ' ') p++;
            if (p >= col_end) break;
            uint64_t h = 0, pos = 0;
            const char* q = p;
            if (*q < '0' || *q > '9') { f->err = "malformed minimizer in the sketch TSV"; return NTL_ERR_ARG; }
            while (q < col_end && *q >= '0' && *q <= '9') h = h * 10 + (uint64_t)(*q++ - '0');
            if (q >= col_end || *q != ':') { f->err = "malformed minimizer in the sketch TSV (hash:pos:strand expected)"; return NTL_ERR_ARG; }
            q++;
            if (q >= col_end || *q < '0' || *q > '9') { f->err = "malformed minimizer position in the sketch TSV"; return NTL_ERR_ARG; }
            while (q < col_end && *q >= '0' && *q <= '9') pos = pos * 10 + (uint64_t)(*q++ - '0');
            uint32_t fw = 0;
            if (q < col_end && *q == ':') {                       // strand column (absent with `indexlr --pos` only)
                if (q + 1 >= col_end || (q[1] != '+' && q[1] != '-')) { f->err = "malformed minimizer strand in the sketch TSV"; return NTL_ERR_ARG; }
                fw = q[1] == '+' ? 0x80000000u : 0u;
                q += 2;
            }
            if (q < col_end && *q != ' ') { f->err = "malformed minimizer in the sketch TSV"; return NTL_ERR_ARG; }
            if (pos > 0x7FFFFFFFull) { f->err = "minimizer position does not fit 31 bits"; return NTL_ERR_ARG; }
            hash.push_back(h); posf.push_back((uint32_t)pos | fw);
            p = q;
        }
        mx_off.push_back(hash.size());
    }
    out->n_seq = (uint32_t)(mx_off.size() - 1);
    out->n_mx = hash.size();
    out->hash = vec_to_malloc(hash); out->pos_strand = vec_to_malloc(posf); out->mx_off = vec_to_malloc(mx_off);
    out->seq_len = f->with_len ? vec_to_malloc(lens) : nullptr;
    out->names = (char*)malloc(names.size() + 1);
    if (out->names) memcpy(out->names, names.data(), names.size());
    out->name_off = vec_to_malloc(name_off);
    if (!out->hash || !out->pos_strand || !out->mx_off || (f->with_len && !out->seq_len) || !out->names || !out->name_off) { f->err = "out of memory"; return NTL_ERR_ARG; }
    return NTL_OK;
}

void ntl_tsv_close(ntl_tsv_file* f) {
    if (!f) return;
    ntl_seqfile_close(f->src);
    delete f;
}

}  // extern "C"

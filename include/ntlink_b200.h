/*
 * ntlink_b200.h -- C ABI of libntlink_b200.so: the B200-native (sm_100a) implementation of ntLink's
 * data-parallel hot path (minimizer sketching + minimizer mapping).
 *
 * Plain C, pointers and sizes only; every entry point names the reference interface it replaces
 * (paths relative to bcgsc/ntLink v1.3.11). The reference has no FFI of its own: its seams are two
 * executables (`indexlr` from btllib, `bin/ntlink_pair.py`) and the Python functions inside the latter.
 * INTEGRATION.md shows the ctypes binding a maintainer adds to bin/ntlink_pair.py.
 *
 * Conventions
 *   - one ntl_ctx per GPU, not thread-safe per ctx; ctypes releases the GIL during calls;
 *   - all calls return NTL_OK (0) or a negative NTL_ERR_* code; ntl_last_error() gives the text;
 *   - inputs are caller-owned HOST pointers; outputs point into ctx-owned pinned host memory that stays
 *     valid until the next call producing the same output struct on that ctx (or ntl_destroy);
 *   - positions and strands are packed as  pos | (strand == '+' ? 1u << 31 : 0)  ("pos_strand");
 *   - there is no CPU fallback: without a CUDA device ntl_init fails with NTL_ERR_CUDA.
 */
#ifndef NTLINK_B200_H
#define NTLINK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTL_OK 0
#define NTL_ERR_CUDA (-1)       /* CUDA runtime error (no device, out of memory, launch failure) */
#define NTL_ERR_ARG (-2)        /* invalid argument */
#define NTL_ERR_WORKSPACE (-3)  /* a device-side capacity was exceeded even after growing it */
#define NTL_ERR_STATE (-4)      /* call order violated, e.g. mapping before an index was built */
#define NTL_ERR_ASSERT (-5)     /* an assertion of the reference would have failed (e.g. paf:127-129) */

#define NTL_STRAND_BIT 0x80000000u
#define NTL_POS_MASK 0x7FFFFFFFu

typedef struct ntl_ctx ntl_ctx;

/* ---- lifecycle ---------------------------------------------------------------------------------- */
int ntl_init(int device, ntl_ctx** out);
void ntl_destroy(ntl_ctx* ctx);
const char* ntl_last_error(const ntl_ctx* ctx);
int ntl_version(void);
/* tuning knobs: "strip_len" (k-mer positions per thread, multiple of 8; 0 = chosen per batch, the default), "tile" (1: the
 * single-pass tile kernel where the window size allows it), "small" (dense-mode kernels for windows of w <= 16: 0 off,
 * 1 automatic (default), 2 tile form, 3 streaming form), "cand_c" (candidate threshold,
 * expected candidates per window), "batch_bases" (bases per device batch), "pipeline_min_bases" (4x the chunk size of
 * the pipelined ntl_map_reads), "async" (1: ntl_map_reads / ntl_map_resident enqueue the whole call without waiting for
 * the device and synchronise once; 0: the step-by-step path the sync-free one falls back to), "graph" (1: each chunk of
 * a sync-free call is launched as one CUDA graph), "resident_chunk_bases" (bases per device batch of ntl_map_resident), "copy_threads" (host threads that move pageable caller memory into
 * pinned bounce buffers; -1 auto, 0 leaves the staging to the driver) */
int ntl_set_option(ntl_ctx* ctx, const char* name, double value);

/* ---- sketch -------------------------------------------------------------------------------------
 * Replaces the external sketcher   indexlr --long --pos --strand [--len] -k K -w W -t T FILE
 * (btllib <= 1.6.2; invoked at ntLink:198-199 for the target and ntLink:221-225 for the reads): ntHash
 * canonical hash, windowed minimizers over valid k-mers, second hash printed. `seq` holds nseq sequences
 * back to back (ASCII, any case, any non-ACGT byte is an invalid base), offsets[nseq+1] delimit them.
 * Output order = input order; minimizers of sequence i are [seq_off[i], seq_off[i+1]). */
typedef struct {
    uint64_t n_mx;
    uint32_t nseq;
    uint32_t reserved;
    const uint64_t* hash;        /* [n_mx]  the hash indexlr prints */
    const uint32_t* pos_strand;  /* [n_mx]  k-mer start | strand bit */
    const uint64_t* seq_off;     /* [nseq+1] */
} ntl_sketch_out;
int ntl_sketch(ntl_ctx* ctx, const char* seq, const uint64_t* offsets, uint32_t nseq, int k, int w,
               ntl_sketch_out* out);

/* ---- target index -------------------------------------------------------------------------------
 * Replaces NtLink.read_minimizers (bin/ntlink_pair.py:189-211): hash -> (contig, position, strand); a hash
 * that occurs two or more times anywhere in the target is dropped entirely. contig[] are ids 0..ncontig-1
 * in target FASTA order; contig_len[] their lengths (bin/ntlink_utils.py:65-73); name_rank[] the rank of
 * each contig NAME under Python string comparison (normalize_pair, bin/ntlink_pair.py:213-219). */
int ntl_index_build(ntl_ctx* ctx, const uint64_t* hash, const uint32_t* contig, const uint32_t* pos_strand,
                    uint64_t n, const uint32_t* contig_len, const uint32_t* name_rank, uint32_t ncontig);
/* Same, but sketches the target on the device first (ntLink:198-199 + the above in one call, the sketch never
 * leaves the GPU). If sketch_out is not NULL the target sketch is also returned (for writing target.tsv). */
int ntl_index_build_from_sequences(ntl_ctx* ctx, const char* seq, const uint64_t* offsets, uint32_t ncontig,
                                   int k, int w, const uint32_t* name_rank, ntl_sketch_out* sketch_out);
int ntl_index_stats(ntl_ctx* ctx, uint64_t* n_inserted, uint64_t* n_unique, uint64_t* table_slots);
/* Replicated index for multi-GPU runs: raw target minimizer triples of this rank (device pointers, for an NCCL
 * all-gather by the caller) -- see ntlink_b200/dist.py. */
int ntl_device_sketch_arrays(ntl_ctx* ctx, uint64_t* n_mx, void** d_hash, void** d_pos_strand, void** d_seq_off);
int ntl_index_build_device(ntl_ctx* ctx, const void* d_hash, const void* d_contig, const void* d_pos_strand,
                           uint64_t n, const uint32_t* contig_len, const uint32_t* name_rank, uint32_t ncontig);

/* ---- mapping ------------------------------------------------------------------------------------
 * Replaces the read loop of NtLink.find_scaffold_pairs (bin/ntlink_pair.py:336-414) with
 * ntlink_utils.get_accepted_anchor_contigs (bin/ntlink_utils.py:200-294) and tally_pairs_from_mappings /
 * add_pair / calculate_pair_info / calculate_gap_size (bin/ntlink_pair.py:157-187,213-239,315-334,416-435). */
typedef struct {
    int32_t k, w;
    int32_t z;              /* -z  minimum contig length */
    int32_t f;              /* -f  max contigs in a run for full transitive edges */
    double x;               /* -x  fudge factor (0 = off) */
    int32_t sensitive;      /* --sensitive */
    int32_t repeat_filter;  /* --repeat-filter */
} ntl_params;

typedef struct {            /* one accepted minimizer hit, 12 bytes */
    uint32_t ctg;           /* contig id */
    uint32_t ctg_pos_strand;
    uint32_t read_pos_strand;
} ntl_hit;

typedef struct {            /* one accepted contig run of a read (a verbose_mapping.tsv line), 12 bytes */
    uint32_t ctg;
    uint32_t start;         /* first hit, relative to the read's hit region */
    uint32_t count;         /* hit_count */
} ntl_run;

typedef struct {            /* one contig-pair observation (an add_pair call that was accepted), 24 bytes */
    uint32_t read;          /* global read ordinal */
    uint32_t ord;           /* insertion order within the read */
    uint32_t src, tgt;      /* contig ids, normalised: lexicographically smaller name first */
    int32_t gap;
    uint32_t flags;         /* bit0 src '+', bit1 tgt '+', bit2 both runs have hit_count > 1 (anchor) */
} ntl_event;

typedef struct {
    uint32_t n_reads;
    uint32_t reserved;
    uint64_t n_mx;            /* read minimizers sketched / received */
    uint64_t n_hits;          /* minimizers found in the index */
    uint64_t n_runs;          /* total accepted runs */
    uint64_t n_events;        /* total pair observations */
    /* per read r: its hit region is [hit_off[r], hit_off[r+1]); after chaining the first nruns[r] entries of
     * runs[hit_off[r] ..] are its accepted runs in output order and hits[hit_off[r] + run.start ..] their hits */
    const uint32_t* hit_off;  /* [n_reads+1] */
    const uint32_t* nruns;    /* [n_reads] */
    const ntl_run* runs;      /* [n_hits] */
    const ntl_hit* hits;      /* [n_hits] */
    /* per read r: events [ev_off[r], ev_off[r] + ev_cnt[r]) in reference insertion order */
    const uint32_t* ev_off;   /* [n_reads+1] */
    const uint32_t* ev_cnt;   /* [n_reads] */
    const ntl_event* events;  /* [ev_off[n_reads]] */
} ntl_map_out;

/* sketch + map nreads reads given as sequences (the fused path: gzip -cd reads | indexlr ... | ntlink_pair.py).
 * first_read_ordinal numbers the reads globally (multi-batch / multi-GPU order). Events are also appended to the
 * ctx's device-side event log for ntl_pairs_finish. */
int ntl_map_reads(ntl_ctx* ctx, const char* seq, const uint64_t* offsets, uint32_t nreads,
                  uint64_t first_read_ordinal, const ntl_params* prm, ntl_map_out* out);
/* map reads given as an indexlr sketch (ntlink_pair.py reading the TSV of `indexlr --len`): mx_off[nreads+1] */
int ntl_map_sketch(ntl_ctx* ctx, const uint64_t* hash, const uint32_t* pos_strand, const uint64_t* mx_off,
                   const uint32_t* read_len, uint32_t nreads, uint64_t first_read_ordinal,
                   const ntl_params* prm, ntl_map_out* out);

/* Checkpoint path: the pair events of reads whose accepted runs/hits are already known (parsed from a
 * verbose_mapping.tsv). Replaces NtLink.find_scaffold_pairs_checkpoints / parse_verbose_entries
 * (bin/ntlink_pair.py:437-488): arrays in the layout of ntl_map_out, read_len[r] = the reference's substitute read
 * length (largest first/last read position of the read's runs, :487). Needs the contig lengths / name ranks of an
 * index (ntl_index_build with n = 0 is enough). Events are appended to the ctx's event log. read_len = NULL: the
 * substitute read length is computed on the device. hit_off = NULL: use the mappings ntl_liftover_mappings left on
 * the device (nreads must match). */
int ntl_tally_mappings(ntl_ctx* ctx, const uint32_t* hit_off, const uint32_t* nruns, const ntl_run* runs, const ntl_hit* hits,
                       const uint32_t* read_len, uint32_t nreads, uint64_t first_read_ordinal, const ntl_params* prm,
                       uint64_t* n_events_out);

/* Grouped mapping for gap filling. Replaces, for ALL gaps at once, the per-gap body of map_long_reads
 * (bin/ntlink_patch_gaps.py:412-442): read_btllib_minimizers (:397-410, a hash seen twice among the group's targets is
 * dropped), the membership filter (:437-438) and ntlink_utils.get_accepted_anchor_contigs (bin/ntlink_utils.py:200-294).
 * Group g = target sequences [group_t_off[g], group_t_off[g+1]) (the two scaffold ends around a gap) + read g. Sketches
 * in the ntl_sketch_out layout (hash, pos|strand<<31, per-sequence offsets), t_len = the lengths the z filter uses (the
 * full scaffolds, :206), r_len = read lengths. Output: ntl_map_out with one "read" per group, contig ids = target
 * sequence indices, hit_off = the read's minimizer offsets (holey regions), no events. */
int ntl_map_groups(ntl_ctx* ctx, const uint64_t* t_hash, const uint32_t* t_pos_strand, const uint64_t* t_mx_off, const uint32_t* t_len,
                   uint32_t ntargets, const uint32_t* group_t_off, const uint64_t* r_hash, const uint32_t* r_pos_strand,
                   const uint64_t* r_mx_off, const uint32_t* r_len, uint32_t ngroups, const ntl_params* prm, ntl_map_out* out);

/* Mapping liftover between rounds. Replaces bin/ntlink_liftover_mappings.py (liftover_ctg_mappings :61-88 and
 * print_adjusted_mappings :90-124; called from ntLink_rounds:122-125): the accepted runs of every read are rewritten
 * from contig coordinates to the coordinates of the scaffolds they were joined into. One ntl_agp_row per contig id of
 * the input mappings (built from the AGP by the caller, read_agp :40-50). Output in the ntl_map_out layout with the
 * same hit_off as the input (regions keep their size, unused slots are holes), contig ids in the new namespace; `out`
 * may be NULL. The lifted arrays also stay on the device: ntl_tally_mappings with hit_off = NULL tallies them without
 * another transfer (build the index of the new namespace in between; read_len = NULL selects the checkpoint read length). */
typedef struct ntl_agp_row {
    uint32_t new_id;      /* id of the path, or of the contig's own name if it has no AGP entry, in the new namespace */
    uint32_t flags;       /* NTL_AGP_* */
    uint32_t scaf_start;  /* AGP column 2 (1-based) */
    uint32_t ctg_start;   /* AGP column 7 (1-based) */
    uint32_t ctg_end;     /* AGP column 8 (1-based) */
} ntl_agp_row;
enum { NTL_AGP_IN = 1 /* has an entry */, NTL_AGP_MINUS = 2 /* orientation '-' */,
       NTL_AGP_KEEP = 4 /* path id == contig id, or orientation not +/-: hits pass through untouched (:84-85) */ };
int ntl_liftover_mappings(ntl_ctx* ctx, const uint32_t* hit_off, const uint32_t* nruns, const ntl_run* runs, const ntl_hit* hits,
                          uint32_t nreads, const ntl_agp_row* agp, uint32_t ncontig, int k, ntl_map_out* out);

/* ---- pair tally ---------------------------------------------------------------------------------
 * Replaces the `pairs` accumulator of find_scaffold_pairs (bin/ntlink_pair.py:327-332): per normalised pair the
 * gap estimates in read order, the anchor count, and the order in which pairs were first seen. */
typedef struct {
    uint32_t src, tgt;
    uint32_t flags;           /* bit0 src '+', bit1 tgt '+' */
    uint32_t n;               /* supporting reads = len(gap_estimates) */
    uint32_t anchor;
    uint32_t reserved;
    uint64_t gap_off;         /* gaps[gap_off .. gap_off+n) in read order */
    uint64_t first_key;       /* (read << 24 | ord) of the first observation; pairs are returned sorted by it */
} ntl_pair;
typedef struct {
    uint64_t n_pairs;
    uint64_t n_gaps;
    const ntl_pair* pairs;
    const int32_t* gaps;
} ntl_pairs_out;
int ntl_events_reset(ntl_ctx* ctx);
int ntl_events_append(ntl_ctx* ctx, const ntl_event* events, uint64_t n);    /* e.g. gathered from other ranks */
int ntl_events_count(ntl_ctx* ctx, uint64_t* n);
int ntl_events_device(ntl_ctx* ctx, uint64_t* n, void** d_events);            /* device pointer for NCCL */
int ntl_events_append_device(ntl_ctx* ctx, const void* d_events, uint64_t n);
/* multi-GPU exchange of the event logs with ONE fixed-size collective (ntlink_b200/dist.py): export writes a header
 * row {count,0,0,0,0,0} followed by at most cap_events events (rows of 6 int32 = ntl_event) into a caller-owned
 * DEVICE buffer of (cap_events+1) rows; import_gathered appends the events of `world` such buffers laid out back to
 * back (the all-gather result), in rank order = global read order. *overflow is set when some rank had more than
 * cap_events events (nothing is imported then: grow the buffers and repeat). */
int ntl_events_export(ntl_ctx* ctx, void* d_dst, uint64_t cap_events, uint64_t* n_out);
int ntl_events_import_gathered(ntl_ctx* ctx, const void* d_src, uint32_t world, uint64_t cap_events, int* overflow);
/* The same exchange without host synchronisation inside the library: ntl_events_export_async only enqueues on the
 * context's stream (ntl_stream; make the collective's stream wait for it), ntl_events_import_counts takes the per-rank
 * counts the caller already read from the gathered headers and only enqueues the copies. */
int ntl_stream(ntl_ctx* ctx, void** cuda_stream_out);
int ntl_events_export_async(ntl_ctx* ctx, void* d_dst, uint64_t cap_events, uint64_t* n_out);
int ntl_events_import_counts(ntl_ctx* ctx, const void* d_src, uint32_t world, uint64_t cap_events, const uint32_t* counts);
/* ... and without the host ever reading the counts: the gathered headers are interpreted on the device, the event log
 * is sized by the bound world * cap_events until ntl_pairs_finish reads the exact count back together with the pair
 * table; a rank that sent more than cap_events makes ntl_pairs_finish fail with NTL_ERR_WORKSPACE. */
int ntl_events_import_device(ntl_ctx* ctx, const void* d_src, uint32_t world, uint64_t cap_events);
int ntl_pairs_finish(ntl_ctx* ctx, ntl_pairs_out* out);

/* ---- host text emitters (byte-identical to the reference's files) -------------------------------
 * Each returns the number of bytes written to a malloc'ed buffer (*out_buf, release with ntl_buf_free) or a
 * negative error. names: concatenated NUL-free names with name_off[n+1]. */
int64_t ntl_format_sketch_tsv(const ntl_sketch_out* sk, const char* names, const uint64_t* name_off,
                              const uint64_t* seq_len /* NULL = no --len column */, int with_pos, int with_strand,
                              int threads, char** out_buf);
/* verbose_mapping.tsv lines (bin/ntlink_pair.py:382-388) */
int64_t ntl_format_verbose(const ntl_map_out* m, const char* read_names, const uint64_t* read_name_off,
                           const char* ctg_names, const uint64_t* ctg_name_off, int threads, char** out_buf);
/* PAF-like lines (bin/ntlink_paf_output.py:103-135); returns NTL_ERR_ASSERT where the reference asserts */
int64_t ntl_format_paf(const ntl_map_out* m, const char* read_names, const uint64_t* read_name_off,
                       const uint32_t* read_len, const char* ctg_names, const uint64_t* ctg_name_off,
                       const uint32_t* ctg_len, int k, int threads, char** out_buf);
void ntl_buf_free(char* buf);
/* FASTA/FASTQ (plain or gzip, multi-line) reader: id = header up to the first whitespace (bin/read_fasta.py).
 * Returns malloc'ed arrays; release with ntl_seqfile_free. max_bases = 0 reads everything, otherwise stops after
 * the record that crosses max_bases (call again with the same handle to continue). */
typedef struct ntl_seqfile ntl_seqfile;
int ntl_seqfile_open(const char* path, ntl_seqfile** out);
int ntl_seqfile_read(ntl_seqfile* f, uint64_t max_bases, char** seq, uint64_t** offsets, char** names,
                     uint64_t** name_off, uint32_t* nseq);
void ntl_seqfile_close(ntl_seqfile* f);
void ntl_free(void* p);

/* verbose_mapping.tsv (plain or gzip) -> the arrays ntl_tally_mappings / ntl_liftover_mappings take. Replaces the Python
 * line loops of find_scaffold_pairs_checkpoints / parse_verbose_entries (bin/ntlink_pair.py:437-488) and liftover_mappings
 * (bin/ntlink_liftover_mappings.py:128-147): reads = maximal blocks of consecutive lines with one read id, one run per
 * line, read_len = the largest first/last read position of the read's runs (:487). With a contig table (names as in the
 * target FASTA) unknown contigs are an error, without one ids are assigned in order of appearance; the table seen so far
 * is returned with every batch. share_repeated = 1 keeps the reference's dict quirk of the checkpoint path (a contig
 * listed twice in a read refers both times to its last listing, :470-472); the liftover passes 0. max_hits = 0 reads the
 * whole file, otherwise whole read blocks until about that many hits. All arrays are malloc'ed: release with ntl_free. */
typedef struct ntl_verbose_file ntl_verbose_file;
typedef struct ntl_mappings_out {
    uint32_t n_reads, n_slots, n_contigs, reserved;
    uint32_t* hit_off;        /* [n_reads + 1] */
    uint32_t* nruns;          /* [n_reads] */
    uint32_t* read_len;       /* [n_reads] */
    uint32_t* runs;           /* [n_slots * 3] {ctg, start, count} */
    uint32_t* hits;           /* [n_slots * 3] {ctg, ctg_pos|strand<<31, read_pos|strand<<31} */
    char* read_names;  uint64_t* read_name_off;     /* [n_reads + 1] */
    char* ctg_names;   uint64_t* ctg_name_off;      /* [n_contigs + 1] */
} ntl_mappings_out;
int ntl_verbose_open(const char* path, const char* ctg_names, const uint64_t* ctg_name_off, uint32_t ncontig, ntl_verbose_file** out);
int ntl_verbose_read(ntl_verbose_file* f, uint64_t max_hits, int share_repeated, ntl_mappings_out* out);
const char* ntl_verbose_error(ntl_verbose_file* f);
void ntl_verbose_close(ntl_verbose_file* f);

/* indexlr TSV (plain or gzip, "-" = stdin) -> sketch arrays in bounded batches. Replaces the Python line / token splitting
 * of NtLink.read_minimizers and find_scaffold_pairs (bin/ntlink_pair.py:196-203,355-366) for the reference's own text
 * interface `indexlr ... | ntlink_pair.py ... -` (ntLink:221-225). with_len = 1 for `indexlr --len` output (reads). Each
 * call returns whole records until about max_mx minimizers (0 = the rest of the file); n_seq == 0 at the end. Arrays are
 * malloc'ed: release with ntl_free. */
typedef struct ntl_tsv_file ntl_tsv_file;
typedef struct ntl_tsv_out {
    uint32_t n_seq, reserved;
    uint64_t n_mx;
    uint64_t* hash;           /* [n_mx] */
    uint32_t* pos_strand;     /* [n_mx] */
    uint64_t* mx_off;         /* [n_seq + 1] */
    uint32_t* seq_len;        /* [n_seq] or NULL without --len */
    char* names;  uint64_t* name_off;   /* [n_seq + 1] */
} ntl_tsv_out;
int ntl_tsv_open(const char* path, int with_len, ntl_tsv_file** out);
int ntl_tsv_read(ntl_tsv_file* f, uint64_t max_mx, ntl_tsv_out* out);
const char* ntl_tsv_error(ntl_tsv_file* f);
void ntl_tsv_close(ntl_tsv_file* f);

/* ---- device-resident / timing interface (bench.py) ----------------------------------------------- */
/* copy a read batch to the device once; ntl_map_resident then runs the whole hot path on it without touching
 * the host (results stay on the device; only the counters come back). */
int ntl_reads_upload(ntl_ctx* ctx, const char* seq, const uint64_t* offsets, uint32_t nreads);
int ntl_map_resident(ntl_ctx* ctx, uint64_t first_read_ordinal, const ntl_params* prm, ntl_map_out* counts_only);
/* target kept resident the same way */
int ntl_target_upload(ntl_ctx* ctx, const char* seq, const uint64_t* offsets, uint32_t ncontig,
                      const uint32_t* name_rank);
int ntl_index_build_resident(ntl_ctx* ctx, int k, int w);
/* multi-GPU: sketch contigs [first, first + count) of the resident target; the triples (hash, GLOBAL
 * contig id, pos|strand) stay on the device for the caller's all-gather, ntl_index_build_device then builds the replicated
 * index from the gathered arrays with the lengths / name ranks ntl_target_resident_meta returns ([ncontig] each). */
int ntl_target_sketch_resident(ntl_ctx* ctx, uint32_t first, uint32_t count, int k, int w, uint64_t* n_mx, void** d_hash,
                               void** d_contig, void** d_pos_strand);
int ntl_target_resident_meta(ntl_ctx* ctx, uint32_t* contig_len, uint32_t* name_rank);
/* device timings (ms, CUDA events on the library's stream) of the last call, indexed by NTL_T_*; and the
 * number of kernels this library launched since ntl_timing_reset */
enum { NTL_T_PACK = 0, NTL_T_DENSE, NTL_T_SELECT, NTL_T_GAP, NTL_T_EMIT, NTL_T_LOOKUP, NTL_T_CHAIN, NTL_T_TALLY,
       NTL_T_INDEX, NTL_T_TOTAL, NTL_T_NUM };
int ntl_timing_reset(ntl_ctx* ctx);
/* counters since ntl_init: "async_calls" (calls that took the sync-free path), "async_fallbacks" (of those, how many had
 * to be repeated on the synchronous path because a capacity bound was too small), "graph_launches", "graph_failures" (a
 * CUDA-graph step failed: graphs are then switched off for the context and plain launches are used) */
int ntl_get_stat(ntl_ctx* ctx, const char* name, double* value);
int ntl_timing(ntl_ctx* ctx, double* ms_accum /* [NTL_T_NUM] */, uint64_t* launches, uint64_t* dense_launches,
               uint64_t* dense_bases);
/* the dominant kernel (k_dense) alone, restricted to launches over at least min 16 Mbp (the read batches): accumulated
 * device time, launches and bases since ntl_timing_reset -- what bench.py's roofline block is computed from */
int ntl_timing_dense(ntl_ctx* ctx, double* ms_accum, uint64_t* launches, uint64_t* bases);
int ntl_device_sync(ntl_ctx* ctx);
/* CUDA-event stopwatch on the library's stream: ntl_mark(ctx, 0) ... work ... ntl_mark(ctx, 1); after a sync
 * ntl_mark_elapsed gives the device time between the two marks in milliseconds */
int ntl_mark(ntl_ctx* ctx, int which);
int ntl_mark_elapsed(ntl_ctx* ctx, double* ms);
/* device-to-device copy on the library's stream (synchronous on return); lets the caller move the library's device
 * arrays into buffers it owns (e.g. torch tensors handed to NCCL) */
int ntl_copy_device(ntl_ctx* ctx, void* d_dst, const void* d_src, uint64_t bytes);

/* ---- synthetic benchmark / test inputs ------------------------------------------------
 * Deterministic, counter-based: the device entry points fill the resident target / reads (the multi-Gbp configurations of
 * BASELINE.json never exist on the host), the host entry points produce the same bytes for the tests and for the CPU
 * reference arm. A genome base is a function of (seed, position); a contig / read is a genome slice, optionally
 * reverse-complemented; contigs may carry one run of N; reads get iid substitutions / deletions / insertions drawn per
 * source position with thresholds out of 65536 (ONT-like: 4 % / 3 % / 3 % = 2621 / 1966 / 1966). */
typedef struct { uint64_t start; uint32_t len, flip, n_start, n_len; uint32_t pad[2]; } ntl_synth_contig;
typedef struct { uint64_t start; uint32_t len, flip; uint64_t id; } ntl_synth_read;
int ntl_synth_target_resident(ntl_ctx* ctx, uint64_t seed, const ntl_synth_contig* contigs, uint32_t ncontig, const uint32_t* name_rank);
int ntl_synth_reads_resident(ntl_ctx* ctx, uint64_t seed, const ntl_synth_read* reads, uint32_t nreads, uint32_t sub16, uint32_t del16,
                             uint32_t ins16, uint64_t* total_bases);
/* which: 0 = resident target, 1 = resident reads */
int ntl_resident_info(ntl_ctx* ctx, int which, uint32_t* nseq, uint64_t* nbases);
/* sequences [first, first + count) of the resident target / reads -> host: off_out[count + 1] (rebased to 0), seq_out may be NULL */
int ntl_resident_download(ntl_ctx* ctx, int which, uint32_t first, uint32_t count, char* seq_out, uint64_t* off_out);
/* host generators: off_out[n + 1] is always written; with seq_out == NULL only the offsets (sizes) are computed */
int ntl_synth_host_contigs(uint64_t seed, const ntl_synth_contig* contigs, uint32_t ncontig, uint64_t* off_out, char* seq_out, int threads);
int ntl_synth_host_reads(uint64_t seed, const ntl_synth_read* reads, uint32_t nreads, uint32_t sub16, uint32_t del16, uint32_t ins16,
                         uint64_t* off_out, char* seq_out, int threads);

#ifdef __cplusplus
}
#endif
#endif /* NTLINK_B200_H */

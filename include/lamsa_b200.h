/*
 * lamsa_b200.h -- C ABI of liblamsa_b200.so, the B200 (sm_100a) replacement for
 * the banded-DP hot path of LAMSA.
 *
 * Two layers are exported:
 *
 *  1. DROP-IN entry points.  Same names, argument meaning, ownership rules and
 *     error behaviour as the reference prototypes, so that the reference's
 *     callers (frag_check.c, split_mapping.c, bwt_aln.c) link against this
 *     library instead of their own ksw.o without source changes.  Each
 *     prototype cites the reference declaration it replaces (paths relative to
 *     the reference tree).  Every DP cell is evaluated by the CUDA kernels in
 *     lamsa_b200/csrc; there is NO CPU implementation behind these symbols and
 *     they abort with a message on stderr when no sm_100 device is usable.
 *
 *  2. BATCH interface (lb2_*).  The producer side of "thousands of independent
 *     DP tasks per launch": the caller describes tasks with plain pointers and
 *     sizes, the library packs them, runs them on one GPU and hands back
 *     scores, end points and CIGARs.  The drop-in symbols are thin wrappers
 *     that submit a batch of one.
 *
 * Only plain C types cross this boundary (no torch / C++ types).
 */
#ifndef LAMSA_B200_H
#define LAMSA_B200_H

#include <stdint.h>
#include <stdio.h>
#include <pthread.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ types -- */

#ifndef cigar32_t
#define cigar32_t int32_t          /* src/lamsa_aln.h:210  (len<<4 | op)        */
#endif

/* BAM-style CIGAR operators: src/lamsa_aln.h:183-204 */
enum { LB2_CMATCH = 0, LB2_CINS = 1, LB2_CDEL = 2, LB2_CREF_SKIP = 3,
       LB2_CSOFT_CLIP = 4, LB2_CHARD_CLIP = 5 };

/* Layout-identical restatement of `lamsa_aln_para` (src/lamsa_aln.h:384-432;
 * 224 bytes on x86-64, offsets in SURVEY.md appendix B).  The layout is checked
 * against the reference header by tests/test_abi.py through oracle/ref_shim.c.
 * When the reference's own headers are in scope (a real drop-in link, where
 * __LAMSA_ALN_H__-style guards already defined the type) define
 * LAMSA_B200_NO_PARA_TYPE before including this header. */
#ifndef LAMSA_B200_NO_PARA_TYPE
typedef struct {
    int n_thread;
    int seed_len, seed_step, seed_inv;
    int per_aln_m;
    int first_loci_thd;
    int SV_len_thd;
    int ske_max;
    float ovlp_rat;
    int bwt_seed_len, bwt_max_len, bwt_min_len;
    int fastest;
    int split_len;
    int split_pen;
    int res_mul_max;
    int hash_len, hash_key_len, hash_step, hash_size;
    uint8_t supp_soft, comm;
    FILE *outp;
    int match_dis, mismatch_thd;
    int del_thd; int ins_thd;
    int *frag_score_table;
    int ins_gapo, ins_gape, del_gapo, del_gape;      /* penalties of the global fill   */
    int ins_ext_o, ins_ext_e, del_ext_o, del_ext_e;  /* penalties of the extension     */
    int match, mis;
    int8_t sc_mat[25];
    int band_w, end_bonus, zdrop;
    float ed_rate, mis_rate, id_rate, mat_rate;
    int read_type;
    uint8_t aln_mode;
} lamsa_aln_para;
#endif

/* ------------------------------------------------- 1. drop-in entry points -- */

/* replaces src/ksw.h:83 / src/ksw.c:543.  *cigar is malloc'd, caller frees.
 * qlen<0 || tlen<0 -> message on stderr and exit(-1) (src/ksw.c:547-548). */
int ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int *n_cigar, cigar32_t **cigar);
/* replaces src/ksw.h:82 / src/ksw.c:655 */
int ksw_global(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
               int m, const int8_t *mat, int gapo, int gape,
               int w, int *n_cigar, cigar32_t **cigar);
/* replaces src/ksw.h:107 / src/ksw.c:387 (score-only extension) */
int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off);
/* replaces src/ksw.h:106 / src/ksw.c:492 */
int ksw_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
               int m, const int8_t *mat, int gapo, int gape,
               int w, int end_bonus, int zdrop, int h0,
               int *qle, int *tle, int *gtle, int *gscore, int *max_off);
/* replaces src/ksw.h:110 / src/ksw.c:667.  Reads ins_ext_o/e, del_ext_o/e,
 * end_bonus, zdrop from AP.  Outputs are written only when n_cigar_ && cigar_
 * (src/ksw.c:781); *cigar_ stays NULL with *n_cigar_==0 when nothing aligned. */
int ksw_extend_core(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                    int m, const int8_t *mat, int w, int h0, lamsa_aln_para *AP,
                    int *_qle, int *_tle, cigar32_t **cigar_, int *n_cigar_, int *m_cigar_);
/* replaces src/ksw.h:114 / src/ksw.c:809: 0 query end reached, 1 target end, 2 neither */
int ksw_extend_c(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                 int m, const int8_t *mat, int w, int h0, lamsa_aln_para *AP,
                 int *_qle, int *_tle, cigar32_t **cigar_, int *n_cigar_, int *m_cigar_);
/* replaces src/ksw.h:117 / src/ksw.c:820 (extension of the reversed pair) */
int ksw_extend_r(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                 int m, const int8_t *mat, int w, int h0, lamsa_aln_para *AP,
                 int *_qre, int *_tre, cigar32_t **cigar_, int *n_cigar_, int *m_cigar_);
/* replaces src/ksw.h:123 / src/ksw.c:862 */
int ksw_bi_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                  int m, const int8_t *mat, int lh0, int rh0, lamsa_aln_para *AP,
                  cigar32_t **cigar_, int *n_cigar_, int *m_cigar_);
/* replaces src/ksw.c:841 (extern-declared by src/frag_check.c:526) */
void sw_mid_fix(cigar32_t **cigar, int *cigar_n, int *cigar_m,
                cigar32_t *lcigar, int ln_cigar, cigar32_t *rcigar, int rn_cigar,
                const uint8_t *query, int qlen, int lqe, int rqe,
                const uint8_t *target, int tlen, int lte, int rte,
                lamsa_aln_para *AP, int m, const int8_t *mat);

/* ---- sparse-DP chaining entry points (src/lamsa_dp_con.h:8-15).  The struct types are the
 * reference's (src/lamsa_aln.h:230-245 map_msg, :368-377 lamsa_aln_per_para, :282-295 aln_reg,
 * :300-303 line_node, :345-366 frag_dp_node; src/frag_check.h:13-43 frag_msg; kseq_t); this
 * header only names them -- lamsa_b200/csrc/ref_abi.h holds the layout-identical restatements,
 * checked against the reference headers by tests/test_abi.py.  The scratch arguments (line,
 * line_start_len, line_rank, ..., f_node) are accepted and not touched: no caller reads them
 * (only lamsa_dp_con.c itself did); f_node serves as the per-worker key that ties a read's
 * frag_line_remain call to its frag_line_BCC call.  *f_msg is malloc'd exactly like
 * frag_init_msg/frag_set_msg would (src/frag_check.c:19-93) and freed by frag_check's
 * frag_free_msg (:959); frag_line_remain sorts and merges a_reg in place like get_remain_reg
 * does (src/lamsa_aln.c:558). */
#ifndef LAMSA_B200_NO_PARA_TYPE   /* with the reference's headers in scope its own prototypes stand */
struct lb2_ref_map_msg; struct lb2_ref_frag_msg; struct lb2_ref_per_para; struct lb2_ref_aln_reg;
struct lb2_ref_line_node; struct lb2_ref_kseq;
/* replaces src/lamsa_dp_con.h:12 / src/lamsa_dp_con.c:1305 */
int frag_line_BCC(struct lb2_ref_map_msg *m_msg, struct lb2_ref_frag_msg **f_msg,
                  struct lb2_ref_per_para *APP, lamsa_aln_para *AP, struct lb2_ref_kseq *seqs,
                  struct lb2_ref_line_node *line, int *line_start_len, int *line_rank, int *line_select_rank,
                  void ***f_node, struct lb2_ref_line_node *_line, int line_n_max);
/* replaces src/lamsa_dp_con.h:8 / src/lamsa_dp_con.c:1252 */
int frag_line_remain(struct lb2_ref_aln_reg *a_reg, struct lb2_ref_map_msg *m_msg, struct lb2_ref_frag_msg **f_msg,
                     struct lb2_ref_per_para *APP, lamsa_aln_para *AP, struct lb2_ref_kseq *seqs,
                     struct lb2_ref_line_node *line, int *line_start_len, int *line_rank, int *line_select_rank,
                     void ***f_node, struct lb2_ref_line_node *_line, int *_line_start_len, int *_line_rank,
                     int line_n_max);
/* helpers of the same translation units that other reference files bind by `extern`
 * (src/bwt_aln.c:103-105,148; src/lamsa_aln.c:609) or through src/lamsa_heap.h:5-17.
 * node_score = src/lamsa_aln.h:325-332 (restated in ref_abi.h as lb2_ref_node_score). */
struct lb2_ref_node_score;
struct lb2_ref_node_score *node_init_score(int n);                          /* src/lamsa_dp_con.c:29  */
void  node_free_score(struct lb2_ref_node_score *ns);                       /* src/lamsa_dp_con.c:39  */
float cover_rate(int s1, int e1, int s2, int e2);                           /* src/lamsa_dp_con.c:61  */
void  build_node_max_heap(struct lb2_ref_node_score *ns);                   /* src/lamsa_heap.c:35    */
void  build_node_min_heap(struct lb2_ref_node_score *ns);                   /* src/lamsa_heap.c:171   */
void  build_node_minpos_heap(struct lb2_ref_node_score *ns);                /* src/lamsa_heap.c:121   */
#endif
/* heap_add_node (src/lamsa_dp_con.c:44), node_pop (src/lamsa_heap.c:5), node_heap_extract_max (:42),
 * node_heap_extract_minpos (:128) and node_heap_update_min (:191) pass or return a line_node BY VALUE
 * ({int x, y}, 8 bytes, in one integer register on x86-64); they are exported with that calling
 * convention and declared in ref_abi.h. */

/* ------------------------------------------------------ 2. batch interface -- */

typedef struct lb2_ctx lb2_ctx;        /* one per (process, GPU) */
typedef struct lb2_batch lb2_batch;    /* packed, device-resident task set */

enum { LB2_KIND_GLOBAL = 0, LB2_KIND_EXTEND = 1 };
enum { LB2_FLAG_CIGAR = 1,             /* produce the traceback / CIGAR                                   */
       LB2_FLAG_TARGET_PAC = 2,        /* target = window [target_pac, +tlen) of the resident reference   */
       LB2_FLAG_TARGET_REV = 4 };      /* ... read back to front (what ksw_extend_r does, src/ksw.c:829)  */

/* One DP task.  `w` is the band as the CALLER would pass it to the reference;
 * the adjustments of src/ksw.c:549 and :696-704 are applied by the library. */
typedef struct {
    int32_t kind;                 /* LB2_KIND_*                                  */
    int32_t flags;                /* LB2_FLAG_*                                  */
    int32_t qlen, tlen;
    const uint8_t *query;         /* qlen codes 0..m-1, host memory              */
    const uint8_t *target;        /* tlen codes 0..m-1, host memory              */
    int32_t w;
    int32_t h0;                   /* extension only, must be > 0                 */
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t end_bonus, zdrop;     /* extension only                              */
    int32_t m;                    /* alphabet size, 1..8                         */
    const int8_t *mat;            /* m*m scores, host memory                     */
    int64_t target_pac;           /* LB2_FLAG_TARGET_PAC: pac coordinate of the window's first base  */
} lb2_task;

typedef struct {
    int32_t score;                /* global: H(tlen-1,qlen-1); extend: best      */
    int32_t qle, tle;             /* extend: end point chosen by src/ksw.c:785-791
                                     (with LB2_FLAG_CIGAR) or max_j+1,max_i+1    */
    int32_t gtle, gscore, max_off;/* extend: src/ksw.c:484-488                   */
    int32_t n_cigar;              /* number of CIGAR words                       */
    int32_t reserved;
    int64_t cigar_off;            /* first word inside the batch's CIGAR pool    */
    int64_t cells;                /* inner-loop bodies evaluated     */
} lb2_result;

/* Keep the reference's 2-bit packed FORWARD sequence resident on the GPU (the .pac
 * image `lamsa aln` loads at src/lamsa_aln.c:1238-1239; base k is
 * pac[k>>2] >> ((~k&3)<<1) & 3, src/bntseq.c:242).  Tasks flagged LB2_FLAG_TARGET_PAC
 * then carry (target_pac, tlen) instead of unpacked bytes: replaces the per-call
 * pac2fa_core unpack (src/bntseq.c:465-477) and the H2D copy of target bytes. */
int  lb2_ctx_set_reference(lb2_ctx *ctx, const uint8_t *pac, int64_t l_pac);

/* All functions return 0 on success, non-zero on error (message via lb2_last_error). */
int  lb2_ctx_create(int device, lb2_ctx **out);
void lb2_ctx_destroy(lb2_ctx *ctx);
const char *lb2_last_error(void);
/* cap on device scratch used for direction bits per launch wave (bytes) */
int  lb2_ctx_set_scratch_limit(lb2_ctx *ctx, uint64_t bytes);

/* One-shot: host tasks in, host results out (pack + H2D + kernels + D2H).
 * `*cigar_pool` is malloc'd (caller frees with lb2_free); result i owns words
 * [cigar_off, cigar_off+n_cigar). */
int  lb2_dp_run(lb2_ctx *ctx, int64_t n, const lb2_task *tasks, lb2_result *results,
                cigar32_t **cigar_pool, int64_t *cigar_pool_n);

/* The same with every sequence of `tasks` lying inside ONE caller-owned host pool [pool, pool+pool_bytes)
 * (page-locked memory from lb2_host_alloc makes the copy a plain DMA; pageable memory works, slower).  The
 * library then does not copy sequences on the host at all: per chunk it uploads the range of the pool the
 * chunk's tasks use, as it lies, and a device kernel lays the sequences out for the fill kernels (32-byte
 * aligned, zero padded).  Host work per task is the classification only.  Results as lb2_dp_run. */
int  lb2_dp_run_pool(lb2_ctx *ctx, const uint8_t *pool, int64_t pool_bytes, int64_t n, const lb2_task *tasks,
                     lb2_result *results, cigar32_t **cigar_pool, int64_t *cigar_pool_n);
/* helper: copy the sequences of `tasks` into `pool` in task order and re-point the records at the copies
 * (*used = bytes taken; fails when pool_bytes is too small -- sum of (qlen+tlen rounded up to 16) is enough) */
int  lb2_pool_pack(int64_t n, lb2_task *tasks, uint8_t *pool, int64_t pool_bytes, int64_t *used);
int  lb2_host_alloc(size_t bytes, void **out);    /* cudaHostAlloc, portable */
void lb2_host_free(void *p);

/* lb2_dp_run cuts batches into chunks of about this many tasks and pipelines host packing,
 * H2D, kernels and read-backs across them (default 131072; at most 16 chunks) */
int  lb2_ctx_set_chunk_tasks(lb2_ctx *ctx, int64_t tasks);
/* byte / launch counters of the last lb2_dp_run on this context */
int  lb2_ctx_last_run_stats(const lb2_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes, int64_t *launches);
/* CUDA-event time of the fill and traceback kernels of the last lb2_dp_run on this context */
int  lb2_ctx_last_run_kernel_ms(const lb2_ctx *ctx, float *fill_ms, float *trace_ms);

/* Staged form used by the benchmark and by pipelined producers. */
int  lb2_batch_create(lb2_ctx *ctx, int64_t n, const lb2_task *tasks, lb2_batch **out); /* pack to pinned host */
/* staged form of lb2_dp_run_pool: no host copy of the sequences; upload = DMA of the pool range + device re-layout */
int  lb2_batch_create_pool(lb2_ctx *ctx, const uint8_t *pool, int64_t pool_bytes, int64_t n, const lb2_task *tasks,
                           lb2_batch **out);
int  lb2_batch_upload(lb2_batch *b);                       /* H2D, async on the ctx stream */
int  lb2_batch_compute(lb2_batch *b, float *kernel_ms);    /* fill + traceback kernels; CUDA-event ms or NULL */
/* lb2_batch_compute in two halves: _async enqueues the kernels and returns, _done polls (1 = finished),
 * _wait blocks and reads the timings.  One batch per context may be between _async and _wait. */
int  lb2_batch_compute_async(lb2_batch *b);
int  lb2_batch_compute_done(lb2_batch *b);
int  lb2_batch_compute_wait(lb2_batch *b, float *kernel_ms);
int  lb2_batch_download(lb2_batch *b, lb2_result *results,
                        cigar32_t **cigar_pool, int64_t *cigar_pool_n);
/* as lb2_batch_download, but the CIGAR pool is a view into pinned staging owned by
 * the batch (valid until lb2_batch_destroy or the next download; do not free) */
int  lb2_batch_download_view(lb2_batch *b, lb2_result *results,
                             const cigar32_t **cigar_pool, int64_t *cigar_pool_n);
int  lb2_batch_stats(const lb2_batch *b, int64_t *h2d_bytes, int64_t *d2h_bytes,
                     int64_t *launches, float *fill_ms, float *trace_ms);
/* Roofline bookkeeping: with class timing on, lb2_batch_compute launches the batch's kernel classes one after the
 * other (instead of concurrently) and times each with its own CUDA events; after a download, lb2_batch_class_stats
 * reports per class the kernel, its tasks, the DP cells it evaluated and its time.  Returns the number of classes. */
typedef struct {
    int32_t class_id, kind, variant, window_slots;
    int64_t tasks, cells;
    float ms;
    char kernel[44];
} lb2_class_stat;
int  lb2_batch_set_class_timing(lb2_batch *b, int on);
int  lb2_batch_class_stats(const lb2_batch *b, lb2_class_stat *out, int cap);
void lb2_batch_destroy(lb2_batch *b);
void lb2_free(void *p);

/* Integer-pipe issue-rate microbenchmark: returns measured G(int16 lane-ops)/s
 * for the s16x2 DPX max/add chain the fill kernels are built from. */
int  lb2_int_peak(lb2_ctx *ctx, double *gops_s16x2, double *gops_s32, int *sm_count, int *clock_khz);

/* ------------------------------------------- 3. sparse-DP anchor chaining -- */
/*
 * SDP = the skeleton search of src/lamsa_dp_con.c: frag_line_BCC (:1305, round 1
 * over all seed hits of a read) and frag_line_remain (:1252, round 2 inside the
 * read regions round 1 left unaligned).  A read is described by its seeds that
 * have hits (src/lamsa_aln.c:943-951: seed_out of seed_all), each with map_n
 * hits; a hit is the part of map_t (src/lamsa_aln.h:230-239) the chaining reads.
 * One warp per read; the predecessor scans of frag_dp_update (:713-751) run across
 * the lanes, node order and every tie rule of the reference are kept.
 */
typedef struct {
    int64_t offset;      /* map_t.offset: 1-based reference coordinate of the hit      */
    int32_t nchr;        /* map_t.nchr                                                 */
    int32_t NM;          /* map_t.NM                                                   */
    int32_t len_dif;     /* map_t.len_dif                                              */
    int32_t nstrand;     /* map_t.nstrand: +1 / -1                                     */
} lb2_sdp_hit;           /* 24 bytes */

/* the fields of lamsa_aln_para the chaining reads */
typedef struct {
    int32_t seed_len, seed_step, seed_inv;
    int32_t per_aln_m, first_loci_thd, SV_len_thd, ske_max;
    float   ovlp_rat;
    int32_t split_len, match_dis, mismatch_thd, aln_mode, bwt_seed_len;
    int32_t frag_score_table[10];   /* src/lamsa_aln.c:177-188 */
} lb2_sdp_para;

/* one aligned record of round 1 = one push_reg_res entry (src/lamsa_aln.c:580-606) */
typedef struct {
    int32_t beg, end;           /* read interval, 1-based inclusive */
    int32_t chr, is_rev;
    int64_t ref_beg, ref_end;
} lb2_sdp_reg;           /* 32 bytes */

typedef struct {
    int32_t seed_out;    /* seeds with hits (lamsa_aln_per_para.seed_out)              */
    int32_t seed_all;    /* all seeds of the read (lamsa_aln_per_para.seed_all)        */
    int32_t read_len;
    int32_t n_reg;       /* aligned records handed to the remain stage                 */
    int64_t seed_first;  /* first entry in seed_id[] / map_n[]                         */
    int64_t hit_first;   /* first entry in hits[]; hits of a read are seed-major       */
    int64_t reg_first;   /* first entry in regs[]                                      */
} lb2_sdp_read;          /* 40 bytes */

/*
 * Result of one stage for one read = the chosen skeletons as an int32 stream
 * (what frag_dp_path, src/lamsa_dp_con.c:1152-1250, turns into frag_msg):
 *   n_lines, then per line:  line_score, frag_num,
 *                            then per fragment: seed_num, seed_num x (seed_i, aln_i)
 * fragments and seeds in the order frag_set_msg (src/frag_check.c:55) receives them;
 * frag_left_bound is 0 and frag_right_bound is seed_all+1 for every line (:1191,:1228).
 */
typedef struct lb2_sdp_batch lb2_sdp_batch;

/* Packs and uploads the reads (hits stay resident for both stages). */
int  lb2_sdp_create(lb2_ctx *ctx, const lb2_sdp_para *para, int64_t n_reads, const lb2_sdp_read *reads,
                    const int32_t *seed_id, const int32_t *map_n, const lb2_sdp_hit *hits,
                    lb2_sdp_batch **out);
/* Re-load an existing batch object with another set of reads (device buffers are grow-only). */
int  lb2_sdp_reset(lb2_sdp_batch *b, const lb2_sdp_para *para, int64_t n_reads, const lb2_sdp_read *reads,
                   const int32_t *seed_id, const int32_t *map_n, const lb2_sdp_hit *hits);
/* Stage 1 (frag_line_BCC) for every read.  stream/offsets are owned by the batch and stay
 * valid until the next stage call or lb2_sdp_destroy; read r owns words
 * [off[r], off[r+1]). */
int  lb2_sdp_run_bcc(lb2_sdp_batch *b, const int32_t **stream, const int64_t **off, float *kernel_ms);
/* Stage 2 (frag_line_remain): regs = the aligned records of stage 1 per read
 * (reads[r].reg_first / n_reg of the `reads` passed here, which may differ from create's
 * only in those two fields). */
int  lb2_sdp_run_remain(lb2_sdp_batch *b, const lb2_sdp_read *reads, const lb2_sdp_reg *regs,
                        const int32_t **stream, const int64_t **off, float *kernel_ms);
/* The state stage 2 needs from stage 1 is one flag per hit: "on a stage-1 skeleton" (TRACKED_FLAG,
 * src/lamsa_dp_con.c:948).  get_tracked after lb2_sdp_run_bcc copies the flags out (one byte per hit,
 * hits in batch order); set_tracked loads them into a batch object that never ran stage 1, so that
 * a producer may regroup reads between the stages (reads reach stage 2 at different times). */
int  lb2_sdp_get_tracked(lb2_sdp_batch *b, uint8_t *flags);
int  lb2_sdp_set_tracked(lb2_sdp_batch *b, const uint8_t *flags);
/* predecessor pairs evaluated by get_fseed_dis inside frag_dp_update in the last stage
 * (the unit of work of the chaining) */
int  lb2_sdp_stats(const lb2_sdp_batch *b, int64_t *pairs, int64_t *h2d_bytes, int64_t *d2h_bytes);
void lb2_sdp_destroy(lb2_sdp_batch *b);
/* ------------------------------------------------------ 4. batch producer -- */
/*
 * The reference aligns reads on `-t N` pthreads, one read per thread at a time, every DP call
 * blocking (src/lamsa_aln.c:825-891, :1151-1162).  These functions have the signatures of
 * pthread_create / pthread_join and run the worker functions as user-level fibers on one OS thread per
 * host core instead (LB2_HOST_THREADS): a worker that reaches one of the drop-in entry points above
 * parks its request and yields; ONE submitter per GPU gathers the parked requests of all threads into
 * batches, and the results come back to the fibers' home threads (producer.cu).  With tens of thousands
 * of workers -- one per read in flight, lamsa_b200/host/aln_core.c -- a launch carries thousands of DP
 * tasks.  Workers start when the first of them is joined.
 */
int lb2_worker_spawn(pthread_t *id, const pthread_attr_t *attr, void *(*fn)(void *), void *arg);
int lb2_worker_join(pthread_t id, void **ret);
/* Called by a worker: lets the other workers of its thread run (a worker waiting for a sibling's progress). */
void lb2_worker_yield(void);
/* Called by a worker: seconds it has spent parked on DP / chaining requests so far (statistics). */
double lb2_worker_parked_seconds(void);
/* Self test of the worker scheduler, its context switch and the request round trip through another thread,
 * without a GPU: n workers yielding `yields` times each on `threads` scheduler threads; returns the number of
 * workers with a wrong result (0 = pass). */
int lb2_fiber_selftest(int n, int yields, int threads);
/* Opens the batch producer's GPUs (contexts, batch slots, device threads) from a helper thread, so that CUDA
 * start-up overlaps the caller's own start-up (index loading).  Optional. */
void lb2_dropin_warmup(void);

/* ------------------------------------------- 5. alignment record statistics -- */
/*
 * The reference-touching half of lamsa_res_aux (src/frag_check.c:793-853): a record's CIGAR walked
 * against its read and the resident 2-bit reference (lb2_ctx_set_reference), giving the counts from which NM and AS
 * follow (:835-838: NM = mismatches + inserted + deleted bases; AS = matches*match - mismatches*mis - gap costs).
 * CIGAR operators other than M / I / D / S are reported as ref_used = -(op+1) (the reference exits there).
 */
typedef struct {
    const cigar32_t *cigar; int32_t n_cigar;
    int32_t read_len;
    const uint8_t *read;          /* codes 0..4, in the orientation the record is reported in             */
    int64_t ref_pac;              /* forward pac coordinate of the record's first reference base           */
} lb2_aux_task;
typedef struct {
    int32_t n_match, n_mismatch, n_ins_open, n_ins_ext, n_del_open, n_del_ext;
    int32_t read_used, ref_used;  /* read / reference bases the CIGAR consumed                             */
} lb2_aux_result;
int  lb2_aux_run(lb2_ctx *ctx, int64_t n, const lb2_aux_task *tasks, lb2_aux_result *results);
/* The same for callers inside the batch producer's worker fibers: parks the request; all parked requests of all
 * workers are served as one launch (producer.cu).  lb2_producer_set_reference hands the producer the reference
 * once (the pointer must stay valid; it is uploaded to the GPUs that serve these requests). */
int  lb2_worker_aux_counts(int64_t n, const lb2_aux_task *tasks, lb2_aux_result *results);
void lb2_producer_set_reference(const uint8_t *pac, int64_t l_pac);

/* field offsets / sizes of the library's restatements of the reference structs (ref_abi.h), in the
 * order of oracle/sdp_ref_shim.c:ref_sdp_offsets / ref_sdp_sizes; returns the count */
int  lb2_ref_abi_offsets(int *out);
int  lb2_ref_abi_sizes(int *out);

/* ------------------------------------------- 7. local Smith-Waterman (ksw_align) -- */
/* src/ksw.h:15-20, :62-63; src/ksw.c:68 (ksw_qinit), :116 (ksw_u8), :237 (ksw_i16), :344 (ksw_align2), :373 (ksw_align).
 * No caller inside LAMSA; exported because the reference's header declares them.  kswq_t stays opaque
 * (callers pass it back and free() it).  Results equal the SSE2 code's, including its dependence on the profile
 * width (16 byte lanes / 8 word lanes), the padding columns in the row maxima, and score 255 on byte overflow. */
#define LB2_KSW_XBYTE  0x10000
#define LB2_KSW_XSTOP  0x20000
#define LB2_KSW_XSUBO  0x40000
#define LB2_KSW_XSTART 0x80000
#ifndef LAMSA_B200_NO_PARA_TYPE
struct _kswq_t;
typedef struct _kswq_t kswq_t;
typedef struct { int score; int te, qe; int score2, te2; int tb, qb; } kswr_t;
kswq_t *ksw_qinit(int size, int qlen, const uint8_t *query, int m, const int8_t *mat);
kswr_t ksw_u8(kswq_t *q, int tlen, const uint8_t *target, int o_del, int e_del, int o_ins, int e_ins, int xtra);
kswr_t ksw_i16(kswq_t *q, int tlen, const uint8_t *target, int o_del, int e_del, int o_ins, int e_ins, int xtra);
kswr_t ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                  int o_del, int e_del, int o_ins, int e_ins, int xtra, kswq_t **qry);
kswr_t ksw_align(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                 int gapo, int gape, int xtra, kswq_t **qry);
#endif
/* batch form: one warp per pair (sw_local.cuh); size 1 = byte profile (ksw_u8), 2 = word profile (ksw_i16);
 * xtra as in src/ksw.h (KSW_XSTART is the caller's second pass, see ksw_align2); tb/qb come back -1 */
typedef struct {
    const uint8_t *query; int32_t qlen;
    const uint8_t *target; int32_t tlen;
    int32_t m; const int8_t *mat;
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t xtra, size;
} lb2_sw_task;
typedef struct { int32_t score, te, qe, score2, te2, tb, qb; } lb2_sw_result;
int lb2_sw_run(lb2_ctx *ctx, int64_t n, const lb2_sw_task *tasks, lb2_sw_result *results);

/* ---------------------------------- 6. local split mapping: seeds and line -- */
/*
 * The seed-and-chain half of the reference's hash_split_map (src/split_mapping.c:634-686): k-mer index of a
 * reference window (init_hash :181-208), look-up of the read's k-mers (:654-675, at most 50 hits per k-mer) and
 * the chaining of the hits into one line (hash_main_line :492-602).  One warp per request (hash_line.cuh).
 * The line comes back as 3 ints per node: read position, diagonal (window position - read position), relation to
 * the previous node (F_MATCH 0, F_MISMATCH 2, F_LONG_MISMATCH 3, F_INSERT 4, F_DELETE 5; src/lamsa_aln.h:101-113).
 * The stitching of the line with DP calls (:688-821) is the drop-in `hash_split_map` below.
 */
typedef struct {
    const uint8_t *ref;  int32_t ref_len;       /* reference window, codes 0..3 (4 hashes as 2, src/bntseq.c:78) */
    const uint8_t *read; int32_t read_len;      /* read piece, codes 0..4                                        */
    int32_t ref_offset;                          /* hash_split_map's ref_offset (> 0: duplication window)         */
    int32_t hash_len, hash_step, split_len;      /* lamsa_aln_para: hash_len (<= 15), hash_step, split_len        */
    int32_t head, tail;                          /* _head, _tail                                                  */
    int32_t *line; int32_t line_cap;             /* out: 3 ints per node; capacity in nodes, at least
                                                    (read_len - hash_len) / hash_step + 1                         */
    int32_t m_len;                               /* out: nodes on the line (hash_main_line's return value)        */
    int32_t n_hits;                              /* out: k-mer hits that entered the chaining                     */
} lb2_hash_task;
int lb2_hash_line_run(lb2_ctx *ctx, int64_t n, lb2_hash_task *tasks);

/* drop-in replacements of src/split_mapping.c:181 and :634 (same signatures).  init_hash no longer builds the host
 * index (it only leaves the arrays its callers free in a freeable state); hash_split_map sends the pair to the GPU
 * for the line and stitches it with this library's ksw_global2 / ksw_bi_extend / ksw_extend_core.  The reference's
 * own definitions are marked weak when split_mapping.c is compiled. */
int init_hash(uint8_t *ref_seq, int ref_len, int hash_len, uint32_t **hash_num, uint64_t ***hash_node,
              int ***hash_node_num, int32_t **hash_pos, int key_len, int hash_size);
int hash_split_map(cigar32_t **split_cigar, int *split_clen, int *split_m,
                   uint8_t *ref_seq, int ref_len, int ref_offset, uint8_t *read_seq, int read_len,
                   lamsa_aln_para *AP, uint32_t *hash_num, uint64_t **hash_node, int **hash_node_num,
                   int32_t *hash_pos, int _head, int _tail);
/* the same two under library names, for a program that keeps its own (weak) init_hash / hash_split_map and forwards
 * to these from one of its own objects (lamsa_b200/host/split_map.c; oracle/Makefile `producer_hash`) */
int lb2_init_hash(uint8_t *ref_seq, int ref_len, int hash_len, uint32_t **hash_num, uint64_t ***hash_node,
                  int ***hash_node_num, int32_t **hash_pos, int key_len, int hash_size);
int lb2_hash_split_map(cigar32_t **split_cigar, int *split_clen, int *split_m,
                       uint8_t *ref_seq, int ref_len, int ref_offset, uint8_t *read_seq, int read_len,
                       lamsa_aln_para *AP, uint32_t *hash_num, uint64_t **hash_node, int **hash_node_num,
                       int32_t *hash_pos, int _head, int _tail);

#ifdef __cplusplus
}
#endif
#endif /* LAMSA_B200_H */

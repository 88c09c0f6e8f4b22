/*
 * aln_core.c -- this repo's `lamsa_aln_core`: the alignment stage of `lamsa aln` restructured into the
 * batch producer's read pipeline.  It takes the place of the reference's driver loop
 * (src/lamsa_aln.c:1116-1177: read a chunk of 128 reads -> `-t` pthreads, one read at a time, every
 * DP call blocking -> join -> write the chunk) and of its worker (lamsa_main_aln, :825-891):
 *
 *   - tens of thousands of reads are in flight at once, each on a worker fiber of liblamsa_b200
 *     (lb2_worker_spawn / _join, producer.cu); every banded-DP and chaining call a read makes is
 *     parked and served in GPU batches gathered over ALL reads in flight;
 *   - no chunk barrier: a worker that has finished a read takes the next one from the input at once
 *     (the input is read under a lock, one read at a time, by whichever worker needs one);
 *   - SAM records are formatted by the workers into memory and written in input order as soon as
 *     all earlier reads are out (the reference's output order, src/lamsa_aln.c:1102-1110).
 *
 * Everything a read goes through is the reference's own code, called in the reference's order
 * (gem_map_msg, map_cal_msg, frag_line_BCC, frag_check, get_reg, frag_line_remain, frag_check,
 * bwt_aln_remain, get_cov_f, rearr_aln_res, aln_res_output); this file only owns the loop around it.
 * It is compiled against the reference's headers and linked with the reference's other translation
 * units (its lamsa_aln.c included: that file's own lamsa_aln_core is marked weak at compile time so
 * that this definition is the one `lamsa_aln_c` calls -- oracle/Makefile `producer`).
 */
#include <malloc.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#include "lamsa_aln.h"
#include "lamsa_dp_con.h"
#include "bwt.h"
#include "bntseq.h"
#include "frag_check.h"
#include "split_mapping.h"
#include "bwt_aln.h"
#include "gem_parse.h"

/* liblamsa_b200 (include/lamsa_b200.h section 4); declared here because that header restates types the
 * reference's headers above already define */
extern int lb2_worker_spawn(pthread_t *id, const pthread_attr_t *attr, void *(*fn)(void *), void *arg);
extern int lb2_worker_join(pthread_t id, void **ret);
extern void lb2_worker_yield(void);
extern double lb2_worker_parked_seconds(void);
extern void lb2_dropin_warmup(void);
#ifdef LB2_CORE_RES_AUX
extern void lb2_producer_set_reference(const uint8_t *pac, int64_t l_pac);   /* res_aux.c: record statistics on the GPU */
#endif

/* non-static parts of the reference's src/lamsa_aln.c that its header does not declare */
typedef struct { lamsa_aln_per_para *APP; map_msg *m_msg; aln_res *a_res; } lamsa_seq_t;   /* src/lamsa_aln.c:782-794 */
extern aln_res *aln_init_res(int l_m, int n, int XA_max);                                    /* :369  */
extern void aln_reset_res(aln_res *a_res, int n, int read_len);                              /* :396  */
extern void aln_res_free(aln_res *res, int n);                                               /* :438  */
extern void get_reg(aln_res *res, aln_reg *reg);                                             /* :597  */
extern float get_cov_f(aln_res *res, aln_reg *reg);                                          /* :639  */
extern void rearr_aln_res(aln_res *res, int n, float ovlp_r);                                /* :654  */
extern void map_cal_msg(map_msg *m_msg, bntseq_t *bns);                                      /* :767  */
extern void init_aln_per_para(lamsa_aln_per_para *APP, seed_msg *s_msg, int read_n);          /* :776  */
extern void aln_res_output(lamsa_aln_para AP, aln_res *res, int res_n, char *name, char *seq, char *qual,
                           bntseq_t *bns);                                                   /* :1001 */
#define LB2_LINE_SIZE 65536                                                                   /* LINE_SIZE, :21 */

/* CUDA start-up overlaps index loading: the GPUs are opened from a helper thread as soon as the program starts */
__attribute__((constructor)) static void lb2_core_warm(void) { lb2_dropin_warmup(); }

typedef struct { char *buf; size_t len; int ready; } out_slot;

typedef struct {
	lamsa_aln_para *AP; bwt_t *bwt; bntseq_t *bns; uint8_t *pac; seed_msg *s_msg;
	/* input: one read at a time, under read_mu */
	pthread_mutex_t read_mu;
	kstream_t *fs; FILE *seed_mapfp; char *gem_line;
	long next_seq; int eof;
	/* output: records of read k go out when reads 0..k-1 are out */
	pthread_mutex_t out_mu;
	out_slot *ring; long ring_n; long next_out;
	long done;
	double read_hold_s, out_hold_s;   /* time inside the two locks (LB2_FIBER_STATS) */
	pthread_mutex_t stat_mu; double ph[6];  /* summed over workers: input-lock wait, parse, align (host part), format, emit, set-up */
	double t_begin; float *t_start, *t_end;   /* LB2_READ_TRACE: per read, seconds since the stage began */
} pipeline_t;

/* The chaining entry points of liblamsa_b200 keep their node tables on the GPU and never look at the per-thread
 * scratch the reference allocates in aux_dp_init (src/lamsa_aln.c:969-980: frag_dp_node table, line buffers);
 * f_node only identifies the worker.  LB2_CORE_REF_CHAINING builds this file for the reference's own CPU
 * lamsa_dp_con.c instead (oracle/Makefile `pipeline_cpu`, a test of this pipeline on a box without a GPU), which
 * needs that scratch for real. */
typedef struct {
	frag_dp_node ***f_node;
	line_node *line, *_line;
	int *line_start_len, *_line_start_len, *line_rank, *_line_rank, *line_select_rank;
} chain_scratch;
#ifdef LB2_CORE_REF_CHAINING
extern frag_dp_node ***fnode_alloc(int seed_m, int per_aln_m);                              /* src/lamsa_aln.c:726 */
extern void fnode_free(frag_dp_node ***f_node, int seed_m, int per_aln_m);                   /* :755 */
#endif

typedef struct {
	pipeline_t *P;
	lamsa_seq_t slot;          /* APP, m_msg, a_res of the read being aligned */
	kseq_t seq;                /* its name / bases / qualities (shares the pipeline's stream) */
	char *lines; size_t lines_n, lines_m;   /* the read's lines of the seed map, as read: text, NUL after each line */
	chain_scratch cs;          /* cs.f_node identifies this worker to frag_line_BCC / frag_line_remain */
} worker_t;

/* hand the records of read `seq` to the writer; writes every read that has become due */
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void emit(pipeline_t *P, long seq, char *buf, size_t len)
{
	pthread_mutex_lock(&P->out_mu);
	const double t0 = now_s();
	out_slot *s = &P->ring[seq % P->ring_n];
	s->buf = buf; s->len = len; s->ready = 1;
	for (;;) {
		out_slot *o = &P->ring[P->next_out % P->ring_n];
		if (!o->ready) break;
		if (o->len) fwrite(o->buf, 1, o->len, P->AP->outp);
		free(o->buf);
		o->buf = NULL; o->len = 0; o->ready = 0;
		__atomic_store_n(&P->next_out, P->next_out + 1, __ATOMIC_RELEASE);
	}
	P->out_hold_s += now_s() - t0;
	pthread_mutex_unlock(&P->out_mu);
}

/* The input step of the reference (lamsa_read_seq, src/lamsa_aln.c:927-956) in two halves.  Only what has to be
 * sequential stays under the input lock: the next FASTA/FASTQ record and the read's seed_all lines of the seed map,
 * taken as text.  Returns 0 at the end of the input. */
static int take_next_read(pipeline_t *P, worker_t *w)
{
	if (kseq_read(&w->seq) < 0) return 0;
	++(P->s_msg->read_count);
	init_aln_per_para(w->slot.APP, P->s_msg, P->s_msg->read_count);
	const int seed_all = w->slot.APP->seed_all;
	int k;
	w->lines_n = 0;
	for (k = 0; k < seed_all; ++k) {
		if (fgets(P->gem_line, LB2_LINE_SIZE, P->seed_mapfp) == NULL) {
			fprintf(stderr, "[lamsa_read_seq] Seeds' GEM map-result do NOT match.\n"); exit(1);      /* src/gem_parse.c:213-214 */
		}
		const size_t l = strlen(P->gem_line) + 1;
		if (w->lines_n + l > w->lines_m) {
			w->lines_m = (w->lines_n + l) * 2 + 1024;
			w->lines = (char *)realloc(w->lines, w->lines_m);
		}
		memcpy(w->lines + w->lines_n, P->gem_line, l);
		w->lines_n += l;
	}
	return 1;
}
/* ... and everything else a worker does for itself: the map_msg table and, per line, what gem_map_read does after
 * its fgets (src/gem_parse.c:216-227: drop the newline, skip four tab-separated fields, keep the hit list unless
 * it is "-") with lamsa_read_seq's seed numbering (:945-952). */
static void parse_seed_lines(worker_t *w)
{
	lamsa_seq_t *p = &w->slot;
	const int seed_all = p->APP->seed_all;
	int seed_n, seed_out = 0;
	char *line = w->lines;
	p->m_msg = map_init_msg(seed_all);
	for (seed_n = 0; seed_n < seed_all; ++seed_n) {
		size_t l = strlen(line), i; int ct = 0;
		char *next = line + l + 1;
		if (l) line[--l] = 0;
		for (i = 0; i < l; ++i) {
			if (line[i] == '\t') { if (ct == 3) break; else ct++; }
		}
		if (line[i + 1] != '-') {
			p->m_msg[seed_out].map_str = strdup(line + i + 1);
			p->m_msg[seed_out].seed_id = seed_n + 1;
			++seed_out;
		}
		line = next;
	}
	p->APP->seed_out = seed_out;
}

/* one read through the reference's stages, in the order of src/lamsa_aln.c:846-880 */
static void align_read(worker_t *w, frag_msg **f_msg, uint32_t **hash_num, uint64_t ***hash_node)
{
	pipeline_t *P = w->P;
	lamsa_aln_para *AP = P->AP; bntseq_t *bns = P->bns; uint8_t *pac = P->pac;
	lamsa_seq_t *la = &w->slot; lamsa_aln_per_para *APP = la->APP; kseq_t *seqs = &w->seq;
	chain_scratch *cs = &w->cs;
	const int line_n_max = P->s_msg->seed_max * AP->per_aln_m;
	int k, line_n;

	for (k = 0; k < APP->seed_out; ++k) {
		gem_map_msg(la->m_msg + k, AP->per_aln_m);
		map_cal_msg(la->m_msg + k, bns);
	}
	aln_reset_res(la->a_res, 3, seqs->seq.l);
	aln_reg *a_reg = aln_init_reg(seqs->seq.l);
	line_n = frag_line_BCC(la->m_msg, f_msg, APP, AP, seqs, cs->line, cs->line_start_len, cs->line_rank, cs->line_select_rank,
	                       cs->f_node, cs->_line, line_n_max);

	uint8_t *bseq = (uint8_t *)malloc(seqs->seq.l + 1);
	for (k = 0; k < (int)seqs->seq.l; ++k) bseq[k] = nst_nt4_table[(int)(seqs->seq.s[k])];
	uint8_t *rbseq = NULL;
	if (line_n > 0) {
		frag_check(la->m_msg, f_msg, la->a_res, bns, pac, bseq, &rbseq, APP, AP, seqs, line_n, hash_num, hash_node);
		get_reg(la->a_res, a_reg);
	}
	line_n = frag_line_remain(a_reg, la->m_msg, f_msg, APP, AP, seqs, cs->line, cs->line_start_len, cs->line_rank, cs->line_select_rank,
	                          cs->f_node, cs->_line, cs->_line_start_len, cs->_line_rank, line_n_max);
	if (line_n > 0) {
		frag_check(la->m_msg, f_msg, la->a_res + 1, bns, pac, bseq, &rbseq, APP, AP, seqs, line_n, hash_num, hash_node);
		get_reg(la->a_res + 1, a_reg);
	}
	bwt_aln_remain(a_reg, la->a_res + 2, P->bwt, bns, pac, bseq, &rbseq, AP, seqs);
	get_reg(la->a_res + 2, a_reg);
	la->a_res->cov_f = get_cov_f(la->a_res, a_reg);
	rearr_aln_res(la->a_res, 3, AP->ovlp_rat);
	aln_free_reg(a_reg);
	if (rbseq) free(rbseq);
	free(bseq);
}

static void *read_worker(void *arg)
{
	worker_t *w = (worker_t *)arg;
	pipeline_t *P = w->P;
	lamsa_aln_para *AP = P->AP;
	const int n_key = (int)pow(NT_N, AP->hash_key_len);
	int k;
	/* per-worker state, as lamsa_seq_init (src/lamsa_aln.c:912-923) and aux_dp_init (:982-983) set it up */
	double ph[6] = {0, 0, 0, 0, 0, 0};
	double tp = now_s();
	w->slot.a_res = aln_init_res(1, 3, AP->res_mul_max);
	w->slot.APP = (lamsa_aln_per_para *)malloc(sizeof(lamsa_aln_per_para));
	uint32_t *hash_num = (uint32_t *)calloc(n_key, sizeof(uint32_t));
	uint64_t **hash_node = (uint64_t **)calloc(n_key, sizeof(uint64_t *));
	frag_msg **f_msg = (frag_msg **)malloc(sizeof(frag_msg *));
#ifdef LB2_CORE_REF_CHAINING
	{
		const int line_m = P->s_msg->seed_max * AP->per_aln_m, node_m = line_m * (1 + L_EXTRA);
		w->cs.f_node = fnode_alloc(P->s_msg->seed_max + 2, AP->per_aln_m);
		w->cs.line = (line_node *)malloc(node_m * sizeof(line_node)); w->cs._line = (line_node *)malloc(node_m * sizeof(line_node));
		w->cs.line_start_len = (int *)malloc(line_m * 2 * sizeof(int)); w->cs._line_start_len = (int *)malloc(line_m * 2 * sizeof(int));
		w->cs.line_rank = (int *)malloc(line_m * sizeof(int)); w->cs._line_rank = (int *)malloc(line_m * sizeof(int));
		w->cs.line_select_rank = (int *)malloc(line_m * sizeof(int));
	}
#else
	w->cs.f_node = (frag_dp_node ***)&w->cs;       /* a unique address per worker; never dereferenced */
#endif
	ph[5] += now_s() - tp;
	for (;;) {
		tp = now_s();
		pthread_mutex_lock(&P->read_mu);
		/* keep the writer's window: a read far behind must not let the others run ahead without bound */
		while (!P->eof && P->next_seq - __atomic_load_n(&P->next_out, __ATOMIC_ACQUIRE) >= P->ring_n) {
			pthread_mutex_unlock(&P->read_mu);
			lb2_worker_yield();
			pthread_mutex_lock(&P->read_mu);
		}
		const double tr0 = now_s();
		if (P->eof || !take_next_read(P, w)) {
			P->eof = 1;
			pthread_mutex_unlock(&P->read_mu);
			break;
		}
		const long seq = P->next_seq++;
		P->read_hold_s += now_s() - tr0;
		if (P->t_start) P->t_start[seq] = (float)(now_s() - P->t_begin);
		pthread_mutex_unlock(&P->read_mu);
		{ const double t = now_s(); ph[0] += t - tp; tp = t; }

		parse_seed_lines(w);
		{ const double t = now_s(); ph[1] += t - tp; tp = t; }
		const double parked0 = lb2_worker_parked_seconds();
		align_read(w, f_msg, &hash_num, &hash_node);
		{ const double t = now_s(); ph[2] += t - tp - (lb2_worker_parked_seconds() - parked0); tp = t; }

		/* the read's SAM records, formatted here and written by whoever completes the input order */
		char *buf = NULL; size_t len = 0;
		FILE *mem = open_memstream(&buf, &len);
		if (!mem) { fprintf(stderr, "[lamsa_b200] open_memstream failed\n"); exit(1); }
		lamsa_aln_para ap = *AP;
		ap.outp = mem;
		aln_res_output(ap, w->slot.a_res, 3, w->seq.name.s, w->seq.seq.s, w->seq.qual.s, P->bns);
		fclose(mem);
		map_free_msg(w->slot.m_msg, w->slot.APP->seed_all);
		{ const double t = now_s(); ph[3] += t - tp; tp = t; }
		emit(P, seq, buf, len);
		ph[4] += now_s() - tp;
		if (P->t_end) P->t_end[seq] = (float)(now_s() - P->t_begin);
		const long done = __atomic_add_fetch(&P->done, 1, __ATOMIC_RELAXED);
		if (done % 100000 == 0) fprintf(stderr, "%16ld reads have been aligned.\n", done);
	}
	pthread_mutex_lock(&P->stat_mu);
	for (k = 0; k < 6; ++k) P->ph[k] += ph[k];
	pthread_mutex_unlock(&P->stat_mu);
	free(f_msg);
#ifdef LB2_CORE_REF_CHAINING
	fnode_free(w->cs.f_node, P->s_msg->seed_max + 2, AP->per_aln_m);
	free(w->cs.line); free(w->cs._line); free(w->cs.line_start_len); free(w->cs._line_start_len);
	free(w->cs.line_rank); free(w->cs._line_rank); free(w->cs.line_select_rank);
#endif
	for (k = 0; k < n_key; ++k) free(hash_node[k]);
	free(hash_node); free(hash_num);
	free(w->slot.APP); aln_res_free(w->slot.a_res, 3);
	free(w->seq.name.s); free(w->seq.comment.s); free(w->seq.seq.s); free(w->seq.qual.s); free(w->lines);
	return NULL;
}

static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e && *e ? atoi(e) : dflt; }

int lamsa_aln_core(const char *read_prefix, char *seed_result, seed_msg *s_msg,
                   bwt_t *bwt, bntseq_t *bns, uint8_t *pac, lamsa_aln_para *AP)
{
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	/* Thousands of reads in flight allocate and free gigabytes in small pieces (map_msg tables, CIGARs, sequences).
	 * glibc's defaults hand freed memory back to the kernel and take it again page fault by page fault; keep it. */
	if (!getenv("LB2_DEFAULT_MALLOC")) {
		mallopt(M_TRIM_THRESHOLD, 1 << 30);
		mallopt(M_MMAP_THRESHOLD, 32 << 20);
		mallopt(M_TOP_PAD, 64 << 20);
	}
	pipeline_t P;
	memset(&P, 0, sizeof P);
	P.AP = AP; P.bwt = bwt; P.bns = bns; P.pac = pac; P.s_msg = s_msg;
#ifdef LB2_CORE_RES_AUX
	lb2_producer_set_reference(pac, bns->l_pac);
#endif
	gzFile readfp;
	if ((P.seed_mapfp = fopen(seed_result, "r")) == NULL) { fprintf(stderr, "\n[lamsa_aln_core] Can't open seed-result file %s.\n", seed_result); exit(1); }
	if ((readfp = gzopen(read_prefix, "r")) == NULL) { fprintf(stderr, "\n[lamsa_aln_core] Can't open read file %s.\n", read_prefix); exit(1); }
	P.fs = ks_init(readfp);
	P.gem_line = (char *)malloc(LB2_LINE_SIZE);
	s_msg->read_count = 0;
	pthread_mutex_init(&P.read_mu, NULL);
	pthread_mutex_init(&P.out_mu, NULL);
	pthread_mutex_init(&P.stat_mu, NULL);

	/* reads in flight: `-t` counts host threads in the reference; here the host threads are the library's
	 * (one per core, LB2_HOST_THREADS) and the number that matters is how many reads are open at once */
	long n_workers = env_int("LB2_READS_IN_FLIGHT", 8192);
	if (AP->n_thread > n_workers) n_workers = AP->n_thread;
	if (s_msg->read_all > 0 && n_workers > s_msg->read_all) n_workers = s_msg->read_all;
	if (n_workers < 1) n_workers = 1;
	/* a worker's stack and its guard page are two mappings; stay well below vm.max_map_count (65530 by default) */
	if (n_workers > 24576) n_workers = 24576;
	P.ring_n = 4 * n_workers;
	P.ring = (out_slot *)calloc(P.ring_n, sizeof(out_slot));

	const char *read_trace = getenv("LB2_READ_TRACE");
	if (read_trace && *read_trace && s_msg->read_all > 0) {
		P.t_start = (float *)calloc(s_msg->read_all + 1, sizeof(float)); P.t_end = (float *)calloc(s_msg->read_all + 1, sizeof(float));
		P.t_begin = now_s();
	}
	worker_t *ws = (worker_t *)calloc(n_workers, sizeof(worker_t));
	pthread_t *ids = (pthread_t *)calloc(n_workers, sizeof(pthread_t));
	long i;
	for (i = 0; i < n_workers; ++i) {
		ws[i].P = &P;
		ws[i].seq.f = P.fs;
		lb2_worker_spawn(&ids[i], NULL, read_worker, ws + i);
	}
	for (i = 0; i < n_workers; ++i) lb2_worker_join(ids[i], NULL);

	if (P.t_start) {
		FILE *tf = fopen(read_trace, "w");
		if (tf) { for (i = 0; i < P.done; ++i) fprintf(tf, "%ld %.6f %.6f\n", i, P.t_start[i], P.t_end[i]); fclose(tf); }
		free(P.t_start); free(P.t_end);
	}
	free(ids); free(ws); free(P.ring); free(P.gem_line);
	fclose(P.seed_mapfp); ks_destroy(P.fs); gzclose(readfp);
	pthread_mutex_destroy(&P.read_mu); pthread_mutex_destroy(&P.out_mu);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (getenv("LB2_FIBER_STATS"))
		fprintf(stderr, "[lamsa_b200] alignment stage: %ld reads, %ld in flight, %.3f s wall (%.3f s inside the input lock, %.3f s inside the output lock)\n",
		        P.done, n_workers, (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), P.read_hold_s, P.out_hold_s);
	if (getenv("LB2_FIBER_STATS") && P.done)
		fprintf(stderr, "[lamsa_b200] host time per read, summed over the workers (ms): taking the read (input lock incl. waiting) %.3f, seed-map lines %.3f, "
		        "alignment stages without the time parked on the GPU %.3f, SAM formatting %.3f, handing to the writer %.3f, worker set-up %.3f\n",
		        1e3 * P.ph[0] / P.done, 1e3 * P.ph[1] / P.done, 1e3 * P.ph[2] / P.done, 1e3 * P.ph[3] / P.done, 1e3 * P.ph[4] / P.done, 1e3 * P.ph[5] / P.done);
	return 0;
}

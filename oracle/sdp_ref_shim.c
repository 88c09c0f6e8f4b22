/*
 * sdp_ref_shim.c -- drives the UNMODIFIED reference sparse-DP chaining
 * (frag_line_BCC / frag_line_remain, /root/reference/src/lamsa_dp_con.c) from the
 * flat read description of include/lamsa_b200.h section 3.  Compiled ONLY into
 *   oracle/_ref/liblamsa_ref.so   (all reference translation units but main.c + this file)
 *   oracle/_ref/lamsa_rec         (-DSDP_RECORDER: the whole reference program with
 *                                  frag_line_BCC/_remain wrapped by the linker so that
 *                                  every call's inputs and outputs are written to $SDP_REC)
 * It includes the reference's own headers; nothing of the reference is copied.
 * Test infrastructure; never linked into the product.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "lamsa_aln.h"
#include "frag_check.h"
#include "lamsa_dp_con.h"

#define LAMSA_B200_NO_PARA_TYPE
#include "../include/lamsa_b200.h"

/* skeleton stream of one stage (format: include/lamsa_b200.h, "Result of one stage") */
static int64_t emit_fmsg(const frag_msg *fm, int line_n, int32_t *out, int64_t cap)
{
	int64_t n = 0;
#define PUT(v) do { if (n < cap) out[n] = (int32_t)(v); ++n; } while (0)
	PUT(line_n);
	for (int l = 0; l < line_n; ++l) {
		PUT(fm[l].line_score);
		PUT(fm[l].frag_num);
		for (int f = 0; f < fm[l].frag_num; ++f) {
			const frag_aln_msg *fa = fm[l].fa_msg + f;
			PUT(fa->seed_num);
			for (int s = 0; s < fa->seed_num; ++s) { PUT(fa->seed_i[s]); PUT(fa->seed_aln_i[s]); }
		}
	}
#undef PUT
	return n;
}

static void para_to_flat(const lamsa_aln_para *AP, lb2_sdp_para *P)
{
	memset(P, 0, sizeof *P);
	P->seed_len = AP->seed_len; P->seed_step = AP->seed_step; P->seed_inv = AP->seed_inv;
	P->per_aln_m = AP->per_aln_m; P->first_loci_thd = AP->first_loci_thd; P->SV_len_thd = AP->SV_len_thd;
	P->ske_max = AP->ske_max; P->ovlp_rat = AP->ovlp_rat; P->split_len = AP->split_len;
	P->match_dis = AP->match_dis; P->mismatch_thd = AP->mismatch_thd; P->aln_mode = AP->aln_mode;
	P->bwt_seed_len = AP->bwt_seed_len;
	for (int i = 0; i < 10; ++i) P->frag_score_table[i] = AP->frag_score_table[i];
}

#ifndef SDP_RECORDER
/* ------------------------------------------------------------ library form -- */
char lamsa_pg[1024];          /* defined by main.c in the program; the library form leaves main.c out */
extern int f_BCC_score_table[10];
extern frag_dp_node ***fnode_alloc(int seed_m, int per_aln_m);
extern void fnode_free(frag_dp_node ***f_node, int seed_m, int per_aln_m);
extern void push_reg(aln_reg *reg, int beg, int end, int beg_n, reg_b ref_beg[], int end_n, reg_b ref_end[]);
extern void frag_free_msg(frag_msg *f_msg, int line_num);

int ref_sdp_sizes(int *o)
{
	int n = 0;
	o[n++] = (int)sizeof(map_t); o[n++] = (int)sizeof(map_msg); o[n++] = (int)sizeof(frag_dp_node);
	o[n++] = (int)sizeof(frag_msg); o[n++] = (int)sizeof(frag_aln_msg); o[n++] = (int)sizeof(lamsa_aln_per_para);
	o[n++] = (int)sizeof(aln_reg); o[n++] = (int)sizeof(reg_t); o[n++] = (int)sizeof(reg_b);
	o[n++] = (int)sizeof(kseq_t); o[n++] = (int)sizeof(line_node); o[n++] = (int)sizeof(node_score);
	return n;
}

/* field offsets the product's layout restatements (lamsa_b200/csrc/ref_abi.h) are checked against;
 * order must match tests/test_abi.py:SDP_OFFSETS */
#include <stddef.h>
int ref_sdp_offsets(int *o)
{
	int n = 0;
	o[n++] = offsetof(map_t, nstrand); o[n++] = offsetof(map_t, nchr); o[n++] = offsetof(map_t, offset);
	o[n++] = offsetof(map_t, NM); o[n++] = offsetof(map_t, len_dif);
	o[n++] = offsetof(map_msg, map); o[n++] = offsetof(map_msg, map_n); o[n++] = offsetof(map_msg, seed_id);
	o[n++] = offsetof(frag_msg, frag_max); o[n++] = offsetof(frag_msg, frag_num); o[n++] = offsetof(frag_msg, fa_msg);
	o[n++] = offsetof(frag_msg, line_score); o[n++] = offsetof(frag_msg, frag_left_bound); o[n++] = offsetof(frag_msg, frag_right_bound);
	o[n++] = offsetof(frag_aln_msg, chr); o[n++] = offsetof(frag_aln_msg, strand); o[n++] = offsetof(frag_aln_msg, cigar);
	o[n++] = offsetof(frag_aln_msg, cigar_len); o[n++] = offsetof(frag_aln_msg, cigar_max); o[n++] = offsetof(frag_aln_msg, flag);
	o[n++] = offsetof(frag_aln_msg, seed_max); o[n++] = offsetof(frag_aln_msg, seed_num); o[n++] = offsetof(frag_aln_msg, seed_i);
	o[n++] = offsetof(frag_aln_msg, seed_aln_i);
	o[n++] = offsetof(lamsa_aln_per_para, seed_all); o[n++] = offsetof(lamsa_aln_per_para, seed_out);
	o[n++] = offsetof(aln_reg, reg); o[n++] = offsetof(aln_reg, reg_n); o[n++] = offsetof(aln_reg, reg_m); o[n++] = offsetof(aln_reg, read_len);
	o[n++] = offsetof(reg_t, ref_beg); o[n++] = offsetof(reg_t, ref_end); o[n++] = offsetof(reg_t, beg_n); o[n++] = offsetof(reg_t, end_n);
	o[n++] = offsetof(reg_t, beg_m); o[n++] = offsetof(reg_t, end_m); o[n++] = offsetof(reg_t, beg); o[n++] = offsetof(reg_t, end);
	o[n++] = offsetof(reg_b, is_rev); o[n++] = offsetof(reg_b, chr); o[n++] = offsetof(reg_b, ref_pos);
	o[n++] = offsetof(kseq_t, seq) + offsetof(kstring_t, l);
	o[n++] = offsetof(node_score, node); o[n++] = offsetof(node_score, score); o[n++] = offsetof(node_score, NM);
	o[n++] = offsetof(node_score, min_score_thd); o[n++] = offsetof(node_score, max_n); o[n++] = offsetof(node_score, node_n);
	o[n++] = CIGAR_LEN_M; o[n++] = UNCOVERED;
	return n;
}

/* stages: bit 0 = frag_line_BCC, bit 1 = frag_line_remain (needs bit 0).  off1/off2 have
 * n_reads+1 entries.  Returns 0, or -1 when an output buffer was too small. */
int ref_sdp_run_batch(const lb2_sdp_para *P, int64_t n_reads, const lb2_sdp_read *reads,
                      const int32_t *seed_id, const int32_t *map_n, const lb2_sdp_hit *hits,
                      const lb2_sdp_reg *regs, int stages,
                      int32_t *out1, int64_t cap1, int64_t *off1,
                      int32_t *out2, int64_t cap2, int64_t *off2)
{
	lamsa_aln_para AP;
	memset(&AP, 0, sizeof AP);
	AP.seed_len = P->seed_len; AP.seed_step = P->seed_step; AP.seed_inv = P->seed_inv;
	AP.per_aln_m = P->per_aln_m; AP.first_loci_thd = P->first_loci_thd; AP.SV_len_thd = P->SV_len_thd;
	AP.ske_max = P->ske_max; AP.ovlp_rat = P->ovlp_rat; AP.split_len = P->split_len;
	AP.match_dis = P->match_dis; AP.mismatch_thd = P->mismatch_thd; AP.aln_mode = (uint8_t)P->aln_mode;
	AP.bwt_seed_len = P->bwt_seed_len;
	int table[10];
	for (int i = 0; i < 10; ++i) table[i] = P->frag_score_table[i], f_BCC_score_table[i] = P->frag_score_table[i];
	AP.frag_score_table = table;

	int seed_max = 1;
	for (int64_t r = 0; r < n_reads; ++r) if (reads[r].seed_out > seed_max) seed_max = reads[r].seed_out;
	/* scratch exactly as aux_dp_init lays it out (src/lamsa_aln.c:969-982) */
	frag_dp_node ***f_node = fnode_alloc(seed_max + 2, AP.per_aln_m);
	int line_m = seed_max * AP.per_aln_m, line_node_m = line_m * (1 + L_EXTRA);
	line_node *line = malloc(line_node_m * sizeof(line_node)), *_line = malloc(line_node_m * sizeof(line_node));
	int *lsl = malloc(line_m * 2 * sizeof(int)), *_lsl = malloc(line_m * 2 * sizeof(int));
	int *line_rank = malloc(line_m * sizeof(int)), *_line_rank = malloc(line_m * sizeof(int));
	int *line_select_rank = malloc(line_m * sizeof(int));
	kseq_t *seqs = calloc(1, sizeof(kseq_t));
	seqs->name.s = strdup("read");
	frag_msg **f_msg = malloc(sizeof(frag_msg*));
	int64_t n1 = 0, n2 = 0;
	int rc = 0;

	for (int64_t r = 0; r < n_reads; ++r) {
		const lb2_sdp_read *rd = reads + r;
		lamsa_aln_per_para APP = { 0, rd->seed_all, rd->seed_out };
		map_msg *m_msg = calloc(rd->seed_out > 0 ? rd->seed_out : 1, sizeof(map_msg));
		int64_t h = rd->hit_first;
		for (int i = 0; i < rd->seed_out; ++i) {
			int mn = map_n[rd->seed_first + i];
			m_msg[i].seed_id = seed_id[rd->seed_first + i];
			m_msg[i].map_n = m_msg[i].map_m = mn;
			m_msg[i].map = calloc(mn > 0 ? mn : 1, sizeof(map_t));
			for (int j = 0; j < mn; ++j, ++h) {
				map_t *m = m_msg[i].map + j;
				m->nstrand = (int8_t)hits[h].nstrand; m->strand = hits[h].nstrand == 1 ? '+' : '-';
				m->nchr = hits[h].nchr; m->offset = hits[h].offset; m->NM = hits[h].NM; m->len_dif = hits[h].len_dif;
			}
		}
		seqs->seq.l = rd->read_len;
		off1[r] = n1; off2[r] = n2;
		if (stages & 1) {
			int line_n = frag_line_BCC(m_msg, f_msg, &APP, &AP, seqs, line, lsl, line_rank, line_select_rank,
			                           f_node, _line, line_m);
			n1 += emit_fmsg(line_n > 0 ? *f_msg : 0, line_n, out1 + (n1 < cap1 ? n1 : 0), n1 < cap1 ? cap1 - n1 : 0);
			if (line_n > 0) frag_free_msg(*f_msg, line_n);
		}
		if (stages & 2) {
			aln_reg *a_reg = aln_init_reg(rd->read_len);
			for (int k = 0; k < rd->n_reg; ++k) {
				const lb2_sdp_reg *g = regs + rd->reg_first + k;
				reg_b rb = { (int8_t)g->is_rev, g->chr, g->ref_beg }, re = { (int8_t)g->is_rev, g->chr, g->ref_end };
				push_reg(a_reg, g->beg, g->end, 1, &rb, 1, &re);
			}
			int line_n = frag_line_remain(a_reg, m_msg, f_msg, &APP, &AP, seqs, line, lsl, line_rank, line_select_rank,
			                              f_node, _line, _lsl, _line_rank, line_m);
			n2 += emit_fmsg(line_n > 0 ? *f_msg : 0, line_n, out2 + (n2 < cap2 ? n2 : 0), n2 < cap2 ? cap2 - n2 : 0);
			if (line_n > 0) frag_free_msg(*f_msg, line_n);
			aln_free_reg(a_reg);
		}
		for (int i = 0; i < rd->seed_out; ++i) free(m_msg[i].map);
		free(m_msg);
	}
	off1[n_reads] = n1; off2[n_reads] = n2;
	if (n1 > cap1 || n2 > cap2) rc = -1;
	fnode_free(f_node, seed_max + 2, AP.per_aln_m);
	free(line); free(_line); free(lsl); free(_lsl); free(line_rank); free(_line_rank); free(line_select_rank);
	free(seqs->name.s); free(seqs); free(f_msg);
	return rc;
}

#else
/* ------------------------------------------------------------ recorder form -- */
/* Linked with -Wl,--wrap=frag_line_BCC -Wl,--wrap=frag_line_remain.  Record layout
 * (int32 words, little endian), one record per call:
 *   BCC   : 1, nwords, para(sizeof/4 words), seed_out, seed_all, read_len,
 *           seed_out x (seed_id, map_n), hits (6 words each), n_out, out stream
 *   remain: 2, nwords, n_reg, n_reg x reg (8 words each; before the call sorts/merges them),
 *           n_out, out stream
 * `lamsa aln -t 1` keeps the two calls of a read adjacent. */
static FILE *rec_fp(void)
{
	static FILE *fp;
	if (!fp) {
		const char *p = getenv("SDP_REC");
		fp = fopen(p ? p : "sdp.rec", "wb");
		if (!fp) { perror("SDP_REC"); exit(1); }
	}
	return fp;
}
typedef struct { int32_t *w; int64_t n, m; } wbuf;
static void wput(wbuf *b, const void *src, int64_t nwords)
{
	if (b->n + nwords > b->m) { b->m = (b->n + nwords) * 2 + 64; b->w = realloc(b->w, b->m * 4); }
	memcpy(b->w + b->n, src, nwords * 4); b->n += nwords;
}
static void wput1(wbuf *b, int32_t v) { wput(b, &v, 1); }
static void wflush(wbuf *b, int tag)
{
	FILE *fp = rec_fp();
	int32_t hd[2] = { tag, (int32_t)b->n };
	fwrite(hd, 4, 2, fp); fwrite(b->w, 4, b->n, fp); fflush(fp);
	free(b->w);
}
static void wput_out(wbuf *b, const frag_msg *fm, int line_n)
{
	int64_t n = emit_fmsg(fm, line_n, 0, 0);
	int32_t *t = malloc(n * 4 + 4);
	emit_fmsg(fm, line_n, t, n);
	wput1(b, (int32_t)n); wput(b, t, n); free(t);
}

int __real_frag_line_BCC(map_msg *m_msg, frag_msg **f_msg, lamsa_aln_per_para *APP, lamsa_aln_para *AP, kseq_t *seqs,
        line_node *line, int *lsl, int *line_rank, int *line_select_rank, frag_dp_node ***f_node,
        line_node *_line, int line_n_max);
int __wrap_frag_line_BCC(map_msg *m_msg, frag_msg **f_msg, lamsa_aln_per_para *APP, lamsa_aln_para *AP, kseq_t *seqs,
        line_node *line, int *lsl, int *line_rank, int *line_select_rank, frag_dp_node ***f_node,
        line_node *_line, int line_n_max)
{
	wbuf b = { 0, 0, 0 };
	lb2_sdp_para P; para_to_flat(AP, &P);
	wput(&b, &P, sizeof P / 4);
	wput1(&b, APP->seed_out); wput1(&b, APP->seed_all); wput1(&b, (int32_t)seqs->seq.l);
	for (int i = 0; i < APP->seed_out; ++i) { wput1(&b, m_msg[i].seed_id); wput1(&b, m_msg[i].map_n); }
	for (int i = 0; i < APP->seed_out; ++i)
		for (int j = 0; j < m_msg[i].map_n; ++j) {
			const map_t *m = m_msg[i].map + j;
			lb2_sdp_hit h = { m->offset, m->nchr, m->NM, m->len_dif, m->nstrand };
			wput(&b, &h, sizeof h / 4);
		}
	int line_n = __real_frag_line_BCC(m_msg, f_msg, APP, AP, seqs, line, lsl, line_rank, line_select_rank, f_node, _line, line_n_max);
	wput_out(&b, line_n > 0 ? *f_msg : 0, line_n);
	wflush(&b, 1);
	return line_n;
}

int __real_frag_line_remain(aln_reg *a_reg, map_msg *m_msg, frag_msg **f_msg, lamsa_aln_per_para *APP, lamsa_aln_para *AP,
        kseq_t *seqs, line_node *line, int *lsl, int *line_rank, int *line_select_rank, frag_dp_node ***f_node,
        line_node *_line, int *_lsl, int *_line_rank, int line_n_max);
int __wrap_frag_line_remain(aln_reg *a_reg, map_msg *m_msg, frag_msg **f_msg, lamsa_aln_per_para *APP, lamsa_aln_para *AP,
        kseq_t *seqs, line_node *line, int *lsl, int *line_rank, int *line_select_rank, frag_dp_node ***f_node,
        line_node *_line, int *_lsl, int *_line_rank, int line_n_max)
{
	wbuf b = { 0, 0, 0 };
	wput1(&b, a_reg->reg_n);
	for (int k = 0; k < a_reg->reg_n; ++k) {
		const reg_t *g = a_reg->reg + k;
		if (g->beg_n != 1 || g->end_n != 1) { fprintf(stderr, "[sdp recorder] unexpected region lists\n"); exit(1); }
		lb2_sdp_reg fr = { g->beg, g->end, g->ref_beg[0].chr, g->ref_beg[0].is_rev, g->ref_beg[0].ref_pos, g->ref_end[0].ref_pos };
		wput(&b, &fr, sizeof fr / 4);
	}
	int line_n = __real_frag_line_remain(a_reg, m_msg, f_msg, APP, AP, seqs, line, lsl, line_rank, line_select_rank, f_node,
	                                     _line, _lsl, _line_rank, line_n_max);
	wput_out(&b, line_n > 0 ? *f_msg : 0, line_n);
	wflush(&b, 2);
	return line_n;
}
#endif

/*
 * res_aux.c -- `lamsa_res_aux` (reference src/frag_check.c:793-853: NM / AS of every record of an alignment line,
 * records with a negative score dropped) with the part that touches the reference -- walking each CIGAR against
 * the read and the unpacked reference window -- done on the GPU over the RESIDENT 2-bit reference
 * (liblamsa_b200: lb2_worker_aux_counts, aux_scan.cuh) instead of pac2fa_core + a byte loop per record.
 * The records of a line go out as one parked request; all workers' requests of a moment are one launch.
 * What stays here is the reference's bookkeeping, restated in its order: the window checks of pac2fa_core
 * (src/bntseq.c:465-477), the length check (:834-835), NM / AS (:837-838), the removal of negative-score records
 * (:839-845) and the line totals (:846-851).
 *
 * OPT-IN build (oracle/Makefile `producer_aux`): frag_check.c is compiled unmodified with its own lamsa_res_aux
 * marked weak (`#pragma weak`, force-included), like lamsa_aln_core in aln_core.c.  In the worker-fiber model this
 * trades a linear host scan for one more GPU round trip per alignment line, so it is not the default link.
 * lamsa_res_split (:712-776) only rewrites the record's CIGAR list and never touches the reference: host code.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <zlib.h>

#include "lamsa_aln.h"
#include "bntseq.h"
#include "frag_check.h"

/* include/lamsa_b200.h section 5 (declared here: that header restates types the reference's headers define) */
typedef struct { const cigar32_t *cigar; int32_t n_cigar; int32_t read_len; const uint8_t *read; int64_t ref_pac; } lb2_aux_task;
typedef struct { int32_t n_match, n_mismatch, n_ins_open, n_ins_ext, n_del_open, n_del_ext, read_used, ref_used; } lb2_aux_result;
extern int lb2_worker_aux_counts(int64_t n, const lb2_aux_task *tasks, lb2_aux_result *results);
extern const char *lb2_last_error(void);
extern void copy_res(res_t *f, res_t *t);                        /* src/frag_check.c:778 */
extern char READ_NAME[];                                          /* src/frag_check.c:17 */

void lamsa_res_aux(line_aln_res *la, bntseq_t *bns, uint8_t *pac, uint8_t *read_bseq, int read_len, lamsa_aln_para *AP, kseq_t *seqs)
{
	const int n = la->cur_res_n + 1;
	int m, i;
	(void)pac;
	if (n > 0) {
		lb2_aux_task *tasks = (lb2_aux_task *)malloc(n * sizeof(lb2_aux_task));
		lb2_aux_result *res = (lb2_aux_result *)malloc(n * sizeof(lb2_aux_result));
		int *ref_len = (int *)malloc(n * sizeof(int)), *idx = (int *)malloc(n * sizeof(int));
		for (m = 0; m < n; ++m) {
			res_t *r = la->res + m;
			const int64_t start = r->offset - 1;                     /* 0-based, :808 */
			const int clen = bns->anns[r->chr - 1].len;
			ref_len[m] = refInCigar(r->cigar, r->cigar_len);
			if (start > clen || start < 0) {                         /* src/bntseq.c:470-472 */
				fprintf(stderr, "\n[bntseq] Error: Coor is longger than sequence lenth.(%lld > %d)\n", (long long)start, clen); exit(1);
			}
			if (start + ref_len[m] > clen) ref_len[m] = clen - (int)start;      /* :474 */
			tasks[m].cigar = r->cigar; tasks[m].n_cigar = r->cigar_len;
			tasks[m].read = read_bseq; tasks[m].read_len = read_len;
			tasks[m].ref_pac = bns->anns[r->chr - 1].offset + start;
			idx[m] = m;
		}
		if (lb2_worker_aux_counts(n, tasks, res)) { fprintf(stderr, "[lamsa_b200] %s\n", lb2_last_error()); exit(1); }
		for (m = 0; m <= la->cur_res_n; ++m) {
			res_t *r = la->res + m;
			const lb2_aux_result *c = res + idx[m];
			if (c->ref_used < 0) {                                   /* :827-829 */
				fprintf(stderr, "\n%s\n[lamsa_gen_aux] Error: Unexpected cigar operation: %d.\n", READ_NAME, -c->ref_used - 1);
				printcigar(stderr, r->cigar, r->cigar_len); exit(1);
			}
			if (c->read_used != read_len || c->ref_used != ref_len[idx[m]]) {        /* :834-835 */
				fprintf(stderr, "[lamsa_gen_aux] Error: %s Unmatched length: read: %d %d.\tref: %d %d\n", seqs->name.s, c->read_used, read_len, c->ref_used, ref_len[idx[m]]);
				printcigar(stderr, r->cigar, r->cigar_len); fprintf(stderr, "\n"); exit(1);
			}
			r->NM = c->n_mismatch + c->n_ins_ext + c->n_del_ext;
			r->score = c->n_match * AP->match - c->n_mismatch * AP->mis - c->n_ins_open * AP->ins_gapo - c->n_ins_ext * AP->ins_gape
			           - c->n_del_open * AP->del_gapo - c->n_del_ext * AP->del_gape;
			if (r->score < 0) {
				for (i = m + 1; i <= la->cur_res_n; ++i) { copy_res(la->res + i, la->res + i - 1); idx[i - 1] = idx[i]; }
				m--;
				la->cur_res_n--;
			} else {
				la->tol_score += r->score;
				la->tol_NM += r->NM;
			}
		}
		free(tasks); free(res); free(ref_len); free(idx);
	}
	if (la->cur_res_n < 0) la->tol_score = -1;
	else la->tol_score -= (la->cur_res_n * AP->split_pen);
}

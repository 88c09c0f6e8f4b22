/*
 * dp_oracle.c -- CPU restatement of LAMSA's banded affine-gap DP.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lamsa_b200/ may link, import or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check the CUDA path.  Parity of this
 * restatement is pinned against oracle/_ref/libksw_ref.so (the unmodified
 * reference ksw.c compiled by oracle/Makefile) in tests/test_oracle.py and tests/test_wrappers_oracle.py
 * and against the committed fixtures under tests/golden/.
 *
 * Reference being restated (all paths relative to /root/reference):
 *   banded global alignment + traceback ........ src/ksw.c:543-653 (ksw_global2)
 *   banded extension, z-drop, adaptive band .... src/ksw.c:667-807 (ksw_extend_core)
 *   score-only extension ........................ src/ksw.c:387-490 (ksw_extend2)
 *   forward / reverse wrappers .................. src/ksw.c:809-836
 *   two-sided extension + seam repair ........... src/ksw.c:841-926
 *   CIGAR append helpers ........................ src/frag_check.h:139-198
 *
 * The reference keeps one array of {h,e} pairs that is updated in place; here
 * the same state is held in two "slot" arrays hs[]/es[] where, on entry to row
 * i, hs[j] is H(i-1,j-1) and es[j] is E(i,j).  Slots that a row does not visit
 * keep whatever they held before (the reference behaves the same way and the
 * adaptive band of the extension kernel can read such slots later), so the
 * slot arrays, not a clean recurrence, are the specification.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dp_oracle.h"

#define ORC_NEG_INF (-0x40000000)   /* src/ksw.c:504 */

static __thread uint64_t g_cells;   /* inner-loop bodies executed (SURVEY 8d "cell") */

uint64_t orc_cells_get(void) { return g_cells; }
void orc_cells_reset(void) { g_cells = 0; }

/* ---------------------------------------------------------------- CIGAR -- */

typedef struct { int32_t *v; int n, cap; } oplist;

/* run-length append, capacity 4 then doubling: src/ksw.c:506-516 */
static void ops_add(oplist *l, int op, int len)
{
	if (l->n > 0 && (l->v[l->n - 1] & 0xf) == op) { l->v[l->n - 1] += len << 4; return; }
	if (l->n == l->cap) {
		l->cap = l->cap ? l->cap * 2 : 4;
		l->v = (int32_t *)realloc(l->v, (size_t)l->cap * 4);
	}
	l->v[l->n++] = len << 4 | op;
}

static void ops_flip(int32_t *v, int n)
{
	int a = 0, b = n - 1;
	while (a < b) { int32_t t = v[a]; v[a] = v[b]; v[b] = t; ++a; --b; }
}

/* Walk the direction bytes from (i,k) to the origin: src/ksw.c:636-649, 792-801.
 * dir byte layout: bits0-1 source of H (0 diag, 1 E, 2 F); bit2 E was extended;
 * bit5 F was extended.  `state` selects which 2-bit field is consulted next. */
static void walk_back(const uint8_t *z, int n_col, int w, int i, int k, oplist *out)
{
	int state = 0;
	while (i >= 0 && k >= 0) {
		int off = i > w ? i - w : 0;
		state = z[(long)i * n_col + (k - off)] >> (state << 1) & 3;
		if (state == 0) { ops_add(out, 0, 1); --i; --k; }
		else if (state == 1) { ops_add(out, 2, 1); --i; }
		else { ops_add(out, 1, 1); --k; }
	}
	if (i >= 0) ops_add(out, 2, i + 1);
	if (k >= 0) ops_add(out, 1, k + 1);
	ops_flip(out->v, out->n);
}

/* --------------------------------------------------------------- global -- */

int orc_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int *n_cigar_, int32_t **cigar_)
{
	if (qlen < 0 || tlen < 0) {                       /* src/ksw.c:547-548 */
		fprintf(stderr, "[orc_global2] Error: qlen: %d tlen: %d\n", qlen, tlen);
		exit(-1);
	}
	int dl = qlen > tlen ? qlen - tlen : tlen - qlen;
	if (w < dl + 3) w = dl + 3;                       /* :549 */
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	const int want_path = n_cigar_ && cigar_;
	const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;   /* :559 */
	if (n_cigar_) *n_cigar_ = 0;

	uint8_t *z = want_path ? (uint8_t *)malloc((size_t)n_col * tlen + 1) : 0;
	int32_t *hs = (int32_t *)malloc(((size_t)qlen + 1) * 4);
	int32_t *es = (int32_t *)malloc(((size_t)qlen + 1) * 4);

	/* row -1: :569-572 */
	hs[0] = 0; es[0] = ORC_NEG_INF;
	for (int j = 1; j <= qlen; ++j) {
		hs[j] = j <= w ? -(o_ins + e_ins * j) : ORC_NEG_INF;
		es[j] = ORC_NEG_INF;
	}
	for (int i = 0; i < tlen; ++i) {                  /* :574-633 */
		const int8_t *srow = mat + (size_t)target[i] * m;
		int lo = i > w ? i - w : 0;
		int hi = i + w + 1 < qlen ? i + w + 1 : qlen;
		int32_t left = lo == 0 ? -(o_del + e_del * (i + 1)) : ORC_NEG_INF;
		int32_t f = ORC_NEG_INF;
		uint8_t *zr = z ? z + (size_t)i * n_col : 0;
		for (int j = lo; j < hi; ++j) {
			int32_t diag = hs[j] + srow[query[j]];
			int32_t e = es[j];
			hs[j] = left;
			uint8_t d = 0;
			int32_t h = diag;
			if (!(diag >= e)) { d = 1; h = e; }        /* ties keep the diagonal */
			if (!(h >= f)) { d = 2; h = f; }
			left = h;
			int32_t open = diag - oe_del;
			e -= e_del;
			if (e > open) d |= 1 << 2; else e = open;  /* ties re-open */
			es[j] = e;
			open = diag - oe_ins;
			f -= e_ins;
			if (f > open) d |= 2 << 4; else f = open;
			if (zr) zr[j - lo] = d;
			++g_cells;
		}
		hs[hi] = left; es[hi] = ORC_NEG_INF;          /* :632 */
	}
	int score = hs[qlen];                             /* :634 */
	if (want_path) {
		oplist ops = {0, 0, 0};
		int i = tlen - 1;
		int k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;   /* :638 */
		walk_back(z, n_col, w, i, k, &ops);
		*n_cigar_ = ops.n; *cigar_ = ops.v;
	}
	free(hs); free(es); free(z);
	return score;
}

/* --------------------------------------------------------------- extend -- */

/* The band clamp of src/ksw.c:696-704 (double division, truncation). */
static int clamp_band(int w, int qlen, int m, const int8_t *mat, int end_bonus,
                      int o_del, int e_del, int o_ins, int e_ins)
{
	int best = 0;
	for (int a = 0; a < m * m; ++a) if (mat[a] > best) best = mat[a];
	int lim = (int)((double)(qlen * best + end_bonus - o_ins) / e_ins + 1.);
	if (lim < 1) lim = 1;
	if (w > lim) w = lim;
	lim = (int)((double)(qlen * best + end_bonus - o_del) / e_del + 1.);
	if (lim < 1) lim = 1;
	if (w > lim) w = lim;
	return w;
}

typedef struct {
	int best, best_i, best_j, row_end, gscore, max_off, w, n_col;
} ext_state;

/* Shared body of the two extension variants.  z==NULL => score only
 * (src/ksw.c:387-490); otherwise direction bytes are kept (:667-780). */
static void extend_fill(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                        int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                        int w, int end_bonus, int zdrop, int h0, uint8_t *z, int n_col,
                        ext_state *st)
{
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int32_t *hs = (int32_t *)calloc((size_t)qlen + 2, 4);
	int32_t *es = (int32_t *)calloc((size_t)qlen + 2, 4);
	/* row -1: :692-694 */
	hs[0] = h0;
	hs[1] = h0 > oe_ins ? h0 - oe_ins : 0;   /* slot 1 exists even when qlen==0 (calloc qlen+2) */
	for (int j = 2; j <= qlen && hs[j - 1] > e_ins; ++j) hs[j] = hs[j - 1] - e_ins;

	int best = h0, best_i = -1, best_j = -1, row_end = -1, gscore = -1, max_off = 0;
	int lo = 0, hi = qlen;
	for (int i = 0; i < tlen; ++i) {
		const int8_t *srow = mat + (size_t)target[i] * m;
		int rowmax = 0, rowarg = -1;
		int32_t f = 0, left;
		if (lo < i - w) lo = i - w;                   /* :718-720 */
		if (hi > i + w + 1) hi = i + w + 1;
		if (hi > qlen) hi = qlen;
		if (lo == 0) { left = h0 - (o_del + e_del * (i + 1)); if (left < 0) left = 0; }
		else left = 0;
		const int zoff = i > w ? i - w : 0;           /* static band start, :735 */
		uint8_t *zr = z ? z + (size_t)i * n_col : 0;
		int j;
		for (j = lo; j < hi; ++j) {
			int32_t diag = hs[j], e = es[j];
			hs[j] = left;
			diag = diag ? diag + srow[query[j]] : 0;   /* :737 */
			uint8_t d = 0;
			int32_t h = diag;
			if (!(diag > e)) { d = 1; h = e; }         /* ties prefer E, then F */
			if (!(h > f)) { d = 2; h = f; }
			left = h;
			if (!(rowmax > h)) { rowarg = j; rowmax = h; }   /* last argmax */
			int32_t open = diag - oe_del; if (open < 0) open = 0;
			e -= e_del;
			if (e > open) d |= 1 << 2; else e = open;
			es[j] = e;
			open = diag - oe_ins; if (open < 0) open = 0;
			f -= e_ins;
			if (f > open) d |= 2 << 4; else f = open;
			if (zr) zr[j - zoff] = d;
			++g_cells;
		}
		hs[hi] = left; es[hi] = 0;                    /* :758 */
		if (j == qlen) {                              /* :759-762 */
			if (!(gscore > left)) row_end = i;
			if (left > gscore) gscore = left;
		}
		if (rowmax == 0) break;                       /* :763 */
		if (rowmax > best) {
			best = rowmax; best_i = i; best_j = rowarg;
			int off = rowarg - i; if (off < 0) off = -off;
			if (off > max_off) max_off = off;
		} else if (zdrop > 0) {                       /* :767-773 */
			int di = i - best_i, dj = rowarg - best_j;
			if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
			else         { if (best - rowmax - (dj - di) * e_ins > zdrop) break; }
		}
		/* band trim: :775-778 */
		for (j = lo; j < hi && hs[j] == 0 && es[j] == 0; ++j) ;
		lo = j;
		for (j = hi; j >= lo && hs[j] == 0 && es[j] == 0; --j) ;
		hi = j + 2 < qlen ? j + 2 : qlen;
	}
	free(hs); free(es);
	st->best = best; st->best_i = best_i; st->best_j = best_j;
	st->row_end = row_end; st->gscore = gscore; st->max_off = max_off;
}

int orc_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off)
{
	if (h0 <= 0) { fprintf(stderr, "[orc_extend2] h0 must be positive\n"); abort(); }  /* :395 */
	ext_state st;
	w = clamp_band(w, qlen, m, mat, end_bonus, o_del, e_del, o_ins, e_ins);
	extend_fill(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins,
	            w, end_bonus, zdrop, h0, 0, 0, &st);
	if (qle) *qle = st.best_j + 1;
	if (tle) *tle = st.best_i + 1;
	if (gtle) *gtle = st.row_end + 1;
	if (gscore) *gscore = st.gscore;
	if (max_off) *max_off = st.max_off;
	return st.best;
}

int orc_extend_core(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                    int m, const int8_t *mat, int w, int h0, const orc_ext_par *P,
                    int *qle, int *tle, int32_t **cigar_, int *n_cigar_, int *m_cigar_)
{
	if (qlen < 0 || tlen < 0) {                       /* src/ksw.c:672-673 */
		fprintf(stderr, "[orc_extend_core] Error: qlen: %d tlen: %d\n", qlen, tlen);
		exit(-1);
	}
	if (h0 <= 0) { fprintf(stderr, "[orc_extend_core] h0 must be positive\n"); abort(); } /* :682 */
	ext_state st;
	w = clamp_band(w, qlen, m, mat, P->end_bonus, P->o_del, P->e_del, P->o_ins, P->e_ins);
	int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;  /* :705 */
	size_t zsz = (size_t)n_col * tlen;
	uint8_t *z = (uint8_t *)malloc(zsz + 1);
	memset(z, 255, zsz);                              /* :707 */
	extend_fill(qlen, query, tlen, target, m, mat, P->o_del, P->e_del, P->o_ins, P->e_ins,
	            w, P->end_bonus, P->zdrop, h0, z, n_col, &st);
	if (n_cigar_ && cigar_) {                         /* :781-803 */
		int i, k;
		if (st.gscore <= 0 || st.gscore <= st.best - P->end_bonus) { i = st.best_i; k = st.best_j; }
		else { i = st.row_end; k = qlen - 1; }
		if (qle) *qle = k + 1;
		if (tle) *tle = i + 1;
		oplist ops = {0, 0, 0};
		walk_back(z, n_col, w, i, k, &ops);
		*n_cigar_ = ops.n; *cigar_ = ops.v; *m_cigar_ = ops.cap;
	}
	free(z);
	return st.best;
}

/* :809-818 */
int orc_extend_c(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                 int m, const int8_t *mat, int w, int h0, const orc_ext_par *P,
                 int *qle, int *tle, int32_t **cigar_, int *n_cigar_, int *m_cigar_)
{
	*n_cigar_ = *m_cigar_ = 0;
	orc_extend_core(qlen, query, tlen, target, m, mat, w, h0, P, qle, tle, cigar_, n_cigar_, m_cigar_);
	return *qle == qlen ? 0 : *tle == tlen ? 1 : 2;
}

/* :820-836 */
int orc_extend_r(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                 int m, const int8_t *mat, int w, int h0, const orc_ext_par *P,
                 int *qre, int *tre, int32_t **cigar_, int *n_cigar_, int *m_cigar_)
{
	uint8_t *rq = (uint8_t *)malloc(qlen > 0 ? qlen : 1), *rt = (uint8_t *)malloc(tlen > 0 ? tlen : 1);
	for (int i = 0; i < qlen; ++i) rq[i] = query[qlen - 1 - i];
	for (int i = 0; i < tlen; ++i) rt[i] = target[tlen - 1 - i];
	*n_cigar_ = *m_cigar_ = 0;
	orc_extend_core(qlen, rq, tlen, rt, m, mat, w, h0, P, qre, tre, cigar_, n_cigar_, m_cigar_);
	free(rq); free(rt);
	return *qre == qlen ? 0 : *tre == tlen ? 1 : 2;
}

/* ------------------------------------------------ CIGAR list helpers ----- */

/* src/frag_check.h:139-154 */
static void list_push_one(int32_t **c, int *n, int *cap, int32_t op)
{
	if (*n > 0 && (((*c)[*n - 1] ^ op) & 0xf) == 0) { (*c)[*n - 1] += (op >> 4) << 4; return; }
	if (*n == *cap) {
		*cap = *cap ? *cap << 1 : 4;
		*c = (int32_t *)realloc(*c, (size_t)*cap * 4);
	}
	(*c)[(*n)++] = op;
}
/* src/frag_check.h:156-159 */
static void list_push_nonempty(int32_t **c, int *n, int *cap, int32_t op)
{
	if (op >> 4) list_push_one(c, n, cap, op);
}
/* src/frag_check.h:161-188: joins runs, and folds I next to S into S */
static void list_push_many(int32_t **c, int *n, int *cap, const int32_t *src, int cnt)
{
	if (cnt == 0) return;
	int i = *n, j = 0;
	if (i > 0) {
		int a = (*c)[i - 1] & 0xf, b = src[0] & 0xf;
		if (a == b) { (*c)[i - 1] += (src[0] >> 4) << 4; j = 1; }
		else if ((a == 1 && b == 4) || (a == 4 && b == 1)) {
			(*c)[i - 1] = ((((*c)[i - 1] >> 4) + (src[0] >> 4)) << 4) | 4; j = 1;
		}
	}
	for (; j < cnt; ++i, ++j) {
		if (i == *cap) {
			*cap = *cap ? *cap << 1 : 4;
			*c = (int32_t *)realloc(*c, (size_t)*cap * 4);
		}
		(*c)[i] = src[j];
	}
	*n = i;
}

/* src/ksw.c:841-860 */
void orc_mid_fix(int32_t **cigar, int *cigar_n, int *cigar_m,
                 int32_t *lc, int ln, int32_t *rc, int rn,
                 const uint8_t *query, int qlen, int lqe, int rqe,
                 const uint8_t *target, int tlen, int lte, int rte,
                 const orc_bi_par *P, int m, const int8_t *mat)
{
	int Sn = qlen - lqe - rqe, Hn = tlen - lte - rte, half = P->split_len / 2;
	if (abs(Sn) >= half || abs(Hn) >= half || abs(Sn - Hn) >= half || tlen < 0 || qlen < 0) {
		list_push_many(cigar, cigar_n, cigar_m, lc, ln);
		list_push_one(cigar, cigar_n, cigar_m, (Sn << 4) | 4);
		list_push_one(cigar, cigar_n, cigar_m, Hn << 4 | 5);
		list_push_many(cigar, cigar_n, cigar_m, rc, rn);
	} else {
		int32_t *g = 0; int gn = 0;
		orc_global2(qlen, query, tlen, target, m, mat, P->del_gapo, P->del_gape,
		            P->ins_gapo, P->ins_gape, P->band_w, &gn, &g);
		list_push_many(cigar, cigar_n, cigar_m, g, gn);
		free(g);
	}
}

/* src/ksw.c:862-926 */
int orc_bi_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m,
                  const int8_t *mat, int lh0, int rh0, const orc_bi_par *P,
                  int32_t **cigar_, int *n_cigar_, int *m_cigar_)
{
	if (*n_cigar_) *n_cigar_ = 0;
	int dl = abs(qlen - tlen);
	int w = dl + 3 > P->band_w ? dl + 3 : P->band_w;                       /* :873 */
	int lqe, lte, ln, lm; int32_t *lc = 0;
	int res = orc_extend_c(qlen, query, tlen, target, m, mat, w, lh0, &P->ext, &lqe, &lte, &lc, &ln, &lm);
	/* the reference compares an int with a float expression: :881 */
	int near = dl < P->split_len + tlen * P->id_rate * (P->aln_mode & 2);
	if (res < 2) {
		*cigar_ = lc; *n_cigar_ = ln; *m_cigar_ = lm;
		list_push_nonempty(cigar_, n_cigar_, m_cigar_,
		                   res == 0 ? (((tlen - lte) << 4) | 2) : (((qlen - lqe) << 4) | 1));
		return 0;
	} else if (near && ((lqe << 1 > qlen) || (lte << 1 > tlen))) {
		if (lc) free(lc);
		orc_global2(qlen, query, tlen, target, m, mat, P->del_gapo, P->del_gape,
		            P->ins_gapo, P->ins_gape, P->band_w, n_cigar_, cigar_);
		*m_cigar_ = *n_cigar_;
		return 0;
	}
	int rqe, rte, rn, rm; int32_t *rc = 0;
	res = orc_extend_r(qlen, query, tlen, target, m, mat, w, rh0, &P->ext, &rqe, &rte, &rc, &rn, &rm);
	if (res < 2) {
		list_push_nonempty(&rc, &rn, &rm,
		                   res == 0 ? (((tlen - rte) << 4) | 2) : (((qlen - rqe) << 4) | 1));
		ops_flip(rc, rn);
		*cigar_ = rc; *n_cigar_ = rn; *m_cigar_ = rm;
		free(lc);
		return 0;
	} else if (near && ((rqe << 1 > qlen) || (rte << 1 > tlen))) {
		if (lc) free(lc);
		if (rc) free(rc);
		orc_global2(qlen, query, tlen, target, m, mat, P->del_gapo, P->del_gape,
		            P->ins_gapo, P->ins_gape, P->band_w, n_cigar_, cigar_);
		*m_cigar_ = *n_cigar_;
		return 0;
	}
	if (rn > 1) ops_flip(rc, rn);
	int32_t *out = (int32_t *)malloc(10 * 4); int on = 0, om = 10;      /* :911-912 */
	int Sn = qlen - lqe - rqe;
	orc_mid_fix(&out, &on, &om, lc, ln, rc, rn, query, qlen, lqe, rqe, target, tlen, lte, rte, P, m, mat);
	*cigar_ = out; *n_cigar_ = on; *m_cigar_ = om;
	if (lc) free(lc);
	if (rc) free(rc);
	return Sn >= P->split_len ? 1 : 0;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------- batch driver ----- */
/* Runs lb2_task records (include/lamsa_b200.h) through the restatement above
 * on `nthreads` host threads; used by the tests and by bench.py's cpu_baseline. */
#define LAMSA_B200_NO_PARA_TYPE
typedef struct lamsa_aln_para_opaque lamsa_aln_para;   /* only named in prototypes */
#include "../include/lamsa_b200.h"

static int bd_ext(const lb2_task *t, lb2_result *r, int32_t **c)
{
	orc_ext_par P = { t->o_del, t->e_del, t->o_ins, t->e_ins, t->end_bonus, t->zdrop };
	int qle = 0, tle = 0, nc = 0, mc = 0;
	*c = 0;
	r->score = orc_extend_core(t->qlen, t->query, t->tlen, t->target, t->m, t->mat, t->w, t->h0, &P,
	                           &qle, &tle, c, &nc, &mc);
	r->qle = qle; r->tle = tle; r->n_cigar = nc; r->reserved = mc;
	return 0;
}
static int bd_ext2(const lb2_task *t, lb2_result *r)
{
	int qle, tle, gtle, gs, mo;
	r->score = orc_extend2(t->qlen, t->query, t->tlen, t->target, t->m, t->mat, t->o_del, t->e_del,
	                       t->o_ins, t->e_ins, t->w, t->end_bonus, t->zdrop, t->h0, &qle, &tle, &gtle, &gs, &mo);
	r->qle = qle; r->tle = tle; r->gtle = gtle; r->gscore = gs; r->max_off = mo;
	return 0;
}
#define BD_NAME orc_run_batch
#define BD_GLOBAL(t, nc, c) orc_global2((t)->qlen, (t)->query, (t)->tlen, (t)->target, (t)->m, (t)->mat, \
                                        (t)->o_del, (t)->e_del, (t)->o_ins, (t)->e_ins, (t)->w, nc, c)
#define BD_EXTEND(t, r, c) bd_ext(t, r, c)
#define BD_EXTEND2(t, r) bd_ext2(t, r)
#define BD_CELLS_RESET orc_cells_reset()
#define BD_CELLS_GET orc_cells_get()
#include "batch_driver.h"

/*
 * sw_oracle.c -- CPU restatement of the reference's local Smith-Waterman entry points
 * (`ksw_qinit` /root/reference/src/ksw.c:68-114, `ksw_u8` :116-235, `ksw_i16` :237-335, `ksw_align2` :344-371): the
 * striped SSE2 kernels restated as the scalar recurrence they evaluate.
 *
 * TEST INFRASTRUCTURE ONLY (checker of lamsa_b200/csrc/sw_local.cuh and of the ksw_align* drop-ins).  Parity is
 * PINNED: tests/test_sw_oracle.py compares it with the unmodified reference (oracle/_ref/libksw_ref.so) on seeded pairs.
 *
 * What the striped code computes, column j of the query, row i of the target (all values >= 0, saturating at 0):
 *   T(i,j)   = max(0, H(i-1,j-1) + s(i,j), E(i,j))
 *   Fseg     = the insertion chain restarted at every multiple of slen = ceil(qlen / lanes) (the main loop, :147-168,
 *              carries f only inside a SIMD lane's segment); Ffull = the chain over the whole row (what the lazy-F
 *              loop :170-181 completes)
 *   H(i,j)   = max(T, Ffull)                  -- stored for the next row and for the end-point scan
 * The row has slen * lanes columns: the padding columns behind the query score 0 against everything (:101, :110), so
 * they carry H(i-1, qlen-1) along the diagonal and take part in the row maximum (and hence in the 2nd-best list).
 *   E(i+1,j) = max(0, E(i,j) - e_del, max(T, Fseg)(i,j) - o_del - e_del)   -- E is fed by the value BEFORE the lazy-F
 *              pass (:166 "we do not need to set E(i,j) ..."), so the result depends on the lane count (16 for bytes,
 *              8 for words); this is reproduced, not repaired.
 * Byte mode: a row whose maximum reaches 255 - shift ends the run with score 255 (:197, :202).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XBYTE 0x10000
#define XSTOP 0x20000
#define XSUBO 0x40000
#define XSTART 0x80000

typedef struct { int score, te, qe, score2, te2, tb, qb; } swr_t;

static int imax2(int a, int b) { return a > b ? a : b; }

static swr_t sw_core(int size, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                     int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	swr_t r = {0, -1, -1, -1, -1, -1, -1};
	const int lanes = size == 1 ? 16 : 8;
	const int slen = (qlen + lanes - 1) / lanes, W = slen * lanes;
	int minv = 127, maxv = 0, a, i, j;
	for (a = 0; a < m * m; ++a) { if (mat[a] < minv) minv = mat[a]; if (mat[a] > maxv) maxv = mat[a]; }
	const int shift = (uint8_t)(256 - (uint8_t)minv);             /* :88-96 */
	const int minsc = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000, endsc = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
	int *H = calloc((size_t)W + 1, sizeof(int)), *E = calloc((size_t)W + 1, sizeof(int)), *Hn = calloc((size_t)W + 1, sizeof(int));
	int *Hbest = calloc((size_t)W + 1, sizeof(int));
	int64_t *b = NULL; int n_b = 0, m_b = 0;                      /* (row maximum, row) of rows reaching minsc, runs merged */
	int gmax = 0, te = -1, overflow = 0;
	if (qlen <= 0) { free(H); free(E); free(Hn); free(Hbest); return r; }   /* slen 0: the reference reads H0[-1] (:147) */
	for (i = 0; i < tlen; ++i) {
		const int8_t *row = mat + (int)target[i] * m;
		int fseg = 0, ffull = 0, rmax = 0;
		for (j = 0; j < W; ++j) {
			int t, hs, h;
			if (j % slen == 0) fseg = 0;
			t = imax2(imax2(0, (j ? H[j - 1] : 0) + (j < qlen ? row[query[j]] : 0)), E[j]);
			hs = imax2(t, fseg);
			h = imax2(t, ffull);
			rmax = imax2(rmax, h);
			Hn[j] = h;
			E[j] = imax2(0, imax2(E[j] - e_del, hs - o_del - e_del));
			fseg = imax2(0, imax2(fseg - e_ins, hs - o_ins - e_ins));
			ffull = imax2(0, imax2(ffull - e_ins, h - o_ins - e_ins));
		}
		if (size == 1 && rmax >= 255 - shift) rmax = 255 - shift;
		if (rmax >= minsc) {                                      /* :186-194 */
			if (n_b == 0 || (int)(b[n_b - 1] & 0xffffffff) + 1 != i) {
				if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; b = realloc(b, sizeof(int64_t) * (size_t)m_b); }
				b[n_b++] = (int64_t)rmax << 32 | i;
			} else if ((int)(b[n_b - 1] >> 32) < rmax) b[n_b - 1] = (int64_t)rmax << 32 | i;
		}
		if (rmax > gmax) {
			gmax = rmax; te = i;
			memcpy(Hbest, Hn, sizeof(int) * (size_t)W);
			if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) { overflow = size == 1 && gmax + shift >= 255; break; }
		}
		{ int *sw = H; H = Hn; Hn = sw; }
	}
	r.score = (size == 1 && gmax + shift >= 255) ? 255 : gmax;
	r.te = te;
	if (!(size == 1 && r.score == 255)) {
		int best = -1;
		r.qe = -1;
		for (j = 0; j < W; ++j) if (Hbest[j] > best) { best = Hbest[j]; r.qe = j; }      /* smallest column among the maxima */
		if (b) {
			const int x = (r.score + maxv - 1) / maxv, low = te - x, high = te + x;
			for (i = 0; i < n_b; ++i) {
				const int e = (int)(b[i] & 0xffffffff), v = (int)(b[i] >> 32);
				if ((e < low || e > high) && v > r.score2) { r.score2 = v; r.te2 = e; }
			}
		}
	}
	(void)overflow;
	free(H); free(E); free(Hn); free(Hbest); free(b);
	return r;
}

/* size: 1 = ksw_u8, 2 = ksw_i16 (what ksw_qinit was called with) */
void orc_sw_core(int size, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                 int o_del, int e_del, int o_ins, int e_ins, int xtra, int *out7)
{
	const swr_t r = sw_core(size, qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra);
	memcpy(out7, &r, sizeof r);
}

/* ksw_align2 (src/ksw.c:344-371) without a cached profile */
void orc_sw_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                   int o_del, int e_del, int o_ins, int e_ins, int xtra, int *out7)
{
	const int size = (xtra & XBYTE) ? 1 : 2;
	swr_t r = sw_core(size, qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra), rr;
	if ((xtra & XSTART) && !((xtra & XSUBO) && r.score < (xtra & 0xffff)) && r.qe >= 0) {      /* qe < 0: byte overflow; the reference then builds an empty profile and reads before it */
		uint8_t *rq = malloc((size_t)r.qe + 1), *rt = malloc((size_t)tlen + 1);
		int k;
		for (k = 0; k <= r.qe; ++k) rq[k] = query[r.qe - k];
		memcpy(rt, target, (size_t)tlen);
		for (k = 0; k <= r.te; ++k) rt[k] = target[r.te - k];
		rr = sw_core(size, r.qe + 1, rq, tlen, rt, m, mat, o_del, e_del, o_ins, e_ins, XSTOP | r.score);
		if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
		free(rq); free(rt);
	}
	memcpy(out7, &r, sizeof r);
}

/*
 * sdp_oracle.c -- CPU restatement of LAMSA's sparse-DP anchor chaining, the checker the
 * CUDA path (lamsa_b200/csrc/sdp_*.cu*) is compared against.  TEST INFRASTRUCTURE: only
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it; the product never does.
 *
 * It restates, on the flat read description of include/lamsa_b200.h section 3,
 *   frag_line_BCC      /root/reference/src/lamsa_dp_con.c:1305-1445   (stage 1)
 *   frag_line_remain   :1252-1302 + get_remain_reg src/lamsa_aln.c:548-572  (stage 2)
 * and everything below them (edge classification :596, node init :636-681/:766, predecessor
 * scan :701, tree pruning :786-920, gap refill :1068, region DP :923, skeleton clustering
 * :12-494, overlap filter :568, path -> fragments :1152, heaps src/lamsa_heap.c).
 * Parity is PINNED: tests/test_sdp_oracle.py runs it against the unmodified reference
 * (oracle/_ref/liblamsa_ref.so) and against recorded call streams of whole `lamsa aln` runs
 * (tests/golden/sdp_golden.npz).
 *
 * Data layout is this file's own (one flat node array per read, indexable by (seed, hit));
 * the order of evaluation, every tie rule and every side effect follow the reference.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/lamsa_b200.h"

/* edge kinds, src/lamsa_aln.h:103-121 */
enum { E_MATCH = 0, E_MISMATCH = 2, E_MATCH_THD = 2, E_LONG_MISMATCH = 3, E_INSERT = 4, E_DELETE = 5,
       E_CHR_DIF = 6, E_UNCONNECT = 8, E_INIT = 20 };
/* node states, src/split_mapping.h:60-67 */
enum { ST_MIN = 1, ST_MULTI = 2, ST_WHOLE = 4, ST_TRACKED = 5 };
/* skeleton flags and trailer slots, src/lamsa_aln.h:124-165 */
enum { SK_MERGB = 0, SK_NMERG = 1, SK_MERGH = 2, SK_INTER = 4, SK_DUMP = 8, SK_EXTRA = 5 };

typedef struct { int x, y; } pair_t;
static const pair_t NONE = { -1, 0 };            /* START_NODE, src/lamsa_aln.h:98 */

typedef struct {                                 /* frag_dp_node, src/lamsa_aln.h:345-366 */
	int son_flag;
	pair_t from;
	int in_de, son_n, son_max;
	pair_t *son;
	int max_score, max_NM;
	pair_t max_node;
	int score, tol_NM;
	int match_flag;
	int dp_flag;
	int node_n;
} node_t;

typedef struct { pair_t *node; int *score, *NM; int thd, max_n, n; } heap_t;   /* node_score */
typedef struct { pair_t n1, n2; } trig_t;

typedef struct {
	const lb2_sdp_para *P;
	int seed_out, seed_all, read_len;
	const int32_t *seed_id, *map_n;
	const lb2_sdp_hit *hit;
	int64_t *hoff;                               /* seed -> first node */
	node_t *nd;
	int64_t n_nodes;
	pair_t *line, *tline, *mini;                 /* line / _line of the reference + gap refill list */
	int *lsl, *tlsl, *rank, *trank, *srank;
	int line_n_max;
	int64_t pairs;                               /* edge classifications inside predecessor scans */
} sdp_t;

#define ND(s, p)   ((s)->nd[(s)->hoff[(p).x] + (p).y])
#define NDXY(s, a, b) ((s)->nd[(s)->hoff[a] + (b)])
#define HIT(s, a, b)  ((s)->hit[(s)->hoff[a] + (b)])
/* skeleton trailer accessors (slot layout of src/lamsa_aln.h:138-150) */
#define SK_NODE(s, L, sl, i) ((L) + (sl)[(i) << 1])
#define SK_LEN(sl, i)        ((sl)[((i) << 1) + 1])
#define T_LB(l, n) ((l)[(n)].x)
#define T_RB(l, n) ((l)[(n)].y)
#define T_ELB(l, n) ((l)[(n) + 1].x)
#define T_ERB(l, n) ((l)[(n) + 1].y)
#define T_MF(l, n) ((l)[(n) + 2].x)
#define T_MH(l, n) ((l)[(n) + 2].y)
#define T_LS(l, n) ((l)[(n) + 3].x)
#define T_BS(l, n) ((l)[(n) + 3].y)
#define T_NM(l, n) ((l)[(n) + 4].x)

static int iabs(int v) { return v < 0 ? -v : v; }

/* ---------------------------------------------------------------- heaps ---- */
/* src/lamsa_dp_con.c:29-42 */
static heap_t *heap_new(int n)
{
	heap_t *h = malloc(sizeof *h);
	h->max_n = n; h->n = 0; h->thd = 0;
	h->node = malloc((n > 0 ? n : 1) * sizeof(pair_t));
	h->score = malloc((n > 0 ? n : 1) * sizeof(int));
	h->NM = malloc((n > 0 ? n : 1) * sizeof(int));
	return h;
}
static void heap_free(heap_t *h) { free(h->node); free(h->score); free(h->NM); free(h); }
/* src/lamsa_heap.c:5-13: a stack pop, not a heap extraction */
static pair_t heap_pop_last(heap_t *h, int *score, int *NM)
{
	if (h->n < 1) return NONE;
	--h->n;
	*score = h->score[h->n]; *NM = h->NM[h->n];
	return h->node[h->n];
}
static void heap_swap(heap_t *h, int a, int b)
{
	pair_t t = h->node[a]; h->node[a] = h->node[b]; h->node[b] = t;
	int v = h->score[a]; h->score[a] = h->score[b]; h->score[b] = v;
	v = h->NM[a]; h->NM[a] = h->NM[b]; h->NM[b] = v;
}
/* src/lamsa_heap.c:152-170: min by score, then larger NM */
static void heap_sift_min(heap_t *h, int i)
{
	for (;;) {
		int l = 2 * i + 1, r = 2 * i + 2, m = i;
		if (l < h->n && (h->score[l] < h->score[i] || (h->score[l] == h->score[i] && h->NM[l] > h->NM[i]))) m = l;
		if (r < h->n && (h->score[r] < h->score[m] || (h->score[r] == h->score[m] && h->NM[r] > h->NM[m]))) m = r;
		if (m == i) return;
		heap_swap(h, i, m); i = m;
	}
}
static void heap_build_min(heap_t *h) { for (int i = (h->n - 1) / 2; i >= 0; --i) heap_sift_min(h, i); }
/* src/lamsa_heap.c:102-119: min by skeleton id (.x) */
static void heap_sift_minpos(heap_t *h, int i)
{
	for (;;) {
		int l = 2 * i + 1, r = 2 * i + 2, m = i;
		if (l < h->n && h->node[l].x < h->node[i].x) m = l;
		if (r < h->n && h->node[r].x < h->node[m].x) m = r;
		if (m == i) return;
		heap_swap(h, i, m); i = m;
	}
}
static void heap_build_minpos(heap_t *h) { for (int i = (h->n - 1) / 2; i >= 0; --i) heap_sift_minpos(h, i); }
static pair_t heap_take_minpos(heap_t *h)
{
	if (h->n < 1) return NONE;
	pair_t top = h->node[0];
	--h->n;
	h->node[0] = h->node[h->n]; h->score[0] = h->score[h->n]; h->NM[0] = h->NM[h->n];
	heap_sift_minpos(h, 0);
	return top;
}
/* src/lamsa_dp_con.c:44-59 + src/lamsa_heap.c:191-201: bounded top-k; -1 kept, -2 rejected,
 * otherwise the skeleton id that was pushed out */
static int heap_offer(heap_t *h, pair_t node, int score, int NM)
{
	if (h->n < h->max_n) {
		h->score[h->n] = score; h->NM[h->n] = NM; h->node[h->n++] = node;
		if (h->n == h->max_n) heap_build_min(h);
		return -1;
	}
	if (h->score[0] < score || (h->score[0] == score && h->NM[0] > NM)) {
		int out = h->node[0].x;
		h->score[0] = score; h->NM[0] = NM; h->node[0] = node;
		heap_sift_min(h, 0);
		return out;
	}
	return -2;
}

/* ---------------------------------------------------- edge classification -- */
/* get_fseed_dis, src/lamsa_dp_con.c:596-634 */
static int edge_kind(const sdp_t *s, int pre, int pre_a, int i, int j)
{
	const lb2_sdp_para *P = s->P;
	if (pre == -1 || i == -1) return E_MATCH;
	if (pre == i) return pre_a == j ? E_MATCH : E_UNCONNECT;
	const lb2_sdp_hit *hp = &HIT(s, pre, pre_a), *hi = &HIT(s, i, j);
	if (hi->nchr != hp->nchr || hi->nstrand != hp->nstrand) return E_CHR_DIF;
	int sp = s->seed_id[pre], si = s->seed_id[i], ds = iabs(sp - si);
	if (ds * P->seed_step < P->seed_len) return E_UNCONNECT;
	int64_t exp = hp->offset + hp->nstrand * (si - sp) * P->seed_step;
	int64_t act = hi->offset;
	int dis = (int)(hp->nstrand * ((sp < si) ? (act - exp) : (exp - act))
	                - ((hp->nstrand * (sp - si) < 0) ? hp->len_dif : hi->len_dif));
	int mat_dis = P->match_dis * ((P->aln_mode & 2) ? ds : 1);
	if (dis <= mat_dis && dis >= -mat_dis) {
		if (ds == 1) return E_MATCH;
		if (ds <= 3 * P->mismatch_thd) return E_MISMATCH;
		return E_LONG_MISMATCH;
	}
	if (dis > mat_dis && dis < P->SV_len_thd) return E_DELETE;
	if ((dis < -mat_dis && dis >= 0 - (ds * P->seed_step - P->seed_len))
	    || (dis < -(P->split_len / 2) && dis >= -P->SV_len_thd)) return E_INSERT;
	return E_UNCONNECT;
}

/* ------------------------------------------------------------ node setup --- */
/* fnode_set, :636-655 */
static void node_set(node_t *n, int x, int y, pair_t from, int score, int NM, int match_flag, int dp_flag)
{
	n->son_flag = E_INIT; n->from = from; n->score = score; n->tol_NM = NM;
	n->match_flag = match_flag; n->dp_flag = dp_flag;
	n->node_n = 1; n->in_de = 0; n->son_n = 0;
	n->max_score = score; n->max_NM = NM; n->max_node.x = x; n->max_node.y = y;
}
/* frag_dp_per_init, :766-784 (frag_dp_init :657-681 is the same per hit of a seed) */
static void node_init(sdp_t *s, int x, int y, pair_t from, int dp_flag)
{
	node_t *n = &NDXY(s, x, y);
	if (from.x == NONE.x) { node_set(n, x, y, from, 1, HIT(s, x, y).NM, E_MATCH, dp_flag); return; }
	int k = edge_kind(s, from.x, from.y, x, y);
	if (k != E_UNCONNECT && k != E_CHR_DIF)
		node_set(n, x, y, from, 2 + s->P->frag_score_table[k], HIT(s, x, y).NM + HIT(s, from.x, from.y).NM, k, dp_flag);
	else n->dp_flag = 0 - dp_flag;
}
/* fnode_add_son, :683-698 */
static void node_add_son(sdp_t *s, pair_t fa, pair_t son)
{
	node_t *f = &ND(s, fa);
	++f->in_de;
	if (f->son_n == f->son_max) { f->son_max <<= 1; f->son = realloc(f->son, f->son_max * sizeof(pair_t)); }
	f->son[f->son_n++] = son;
}

/* ------------------------------------------------------ predecessor scan --- */
/* frag_dp_update, :701-764 */
static void node_update(sdp_t *s, int x, int y, int start, int dp_flag)
{
	node_t *me = &NDXY(s, x, y);
	const int *tbl = s->P->frag_score_table;
	pair_t best = me->from;
	int best_score = me->score, best_NM = me->tol_NM, best_flag = me->dp_flag;
	for (int i = x - 1; i >= start; --i) {
		for (int j = 0; j < s->map_n[i]; ++j) {
			node_t *p = &NDXY(s, i, j);
			if (p->dp_flag != dp_flag) continue;
			int strand = HIT(s, i, j).nstrand;
			if (strand == 1 && p->son_flag <= E_MATCH_THD) continue;       /* already has a match-like successor */
			int k = edge_kind(s, i, j, x, y);
			++s->pairs;
			if (k == E_UNCONNECT || k == E_CHR_DIF) continue;
			if (strand == -1 && k <= E_MATCH_THD) {                          /* first match-like predecessor wins */
				best.x = i; best.y = j;
				best_score = p->score + 1 + tbl[k]; best_flag = k; best_NM = p->tol_NM + me->tol_NM;
				goto done;
			}
			int sc = p->score + 1 + tbl[k];
			if (sc > best_score || (sc == best_score && me->tol_NM + p->tol_NM < best_NM)) {
				best.x = i; best.y = j;
				best_score = sc; best_flag = k; best_NM = p->tol_NM + me->tol_NM;
			}
		}
	}
done:
	if (best.x != me->from.x || best.y != me->from.y) {
		ND(s, best).son_flag = best_flag;
		me->from = best; me->score = best_score; me->tol_NM = best_NM; me->match_flag = best_flag;
		me->node_n = ND(s, best).node_n + 1;
		pair_t self = { x, y };
		node_add_son(s, best, self);
	}
}

/* ----------------------------------------------------------- tree pruning -- */
/* node_add_score, :786-802 */
static void path_push(sdp_t *s, int score, int NM, pair_t node, heap_t *h)
{
	if (score < h->thd) return;
	if (h->n > h->max_n - 1) { fprintf(stderr, "[sdp_oracle] path list overflow (%d %d)\n", h->n, h->max_n); exit(1); }
	h->score[h->n] = score; h->NM[h->n] = NM; h->node[h->n++] = node;
	ND(s, node).dp_flag = ST_TRACKED;
	for (pair_t t = ND(s, node).from; t.x != -1; t = ND(s, t).from) ND(s, t).dp_flag = ST_TRACKED;
}
/* detach `son` from its parent: its subtree becomes a path of its own (:842-847, :851-857, :894-899) */
static void detach(sdp_t *s, pair_t son, pair_t max_node, heap_t *h)
{
	node_t *c = &ND(s, son);
	c->from = NONE;
	c->max_score -= c->score - 1;
	c->max_NM -= c->tol_NM - HIT(s, son.x, son.y).NM;
	ND(s, max_node).node_n -= c->node_n - 1;
	path_push(s, c->max_score, c->max_NM, max_node, h);
}
/* get_max_son, :808-829 */
static pair_t best_son(sdp_t *s, int x, int y)
{
	node_t *f = &NDXY(s, x, y);
	int max_score = 0, max_NM = 0, max_dis = 0, flag_thd = E_INIT;
	pair_t best = { 0, 0 };
	for (int i = 0; i < f->son_n; ++i) {
		pair_t c = f->son[i];
		node_t *n = &ND(s, c);
		if (n->match_flag <= flag_thd && (n->max_score > max_score ||
		        (n->max_score == max_score && (c.x - x < max_dis || n->max_NM < max_NM)))) {
			best = c; max_score = n->max_score; max_NM = n->max_NM; max_dis = c.x - x;
			if (n->match_flag <= E_MATCH_THD) flag_thd = E_MATCH_THD;
		}
	}
	return best;
}
/* cut_branch, :831-870 */
static void cut_branch(sdp_t *s, int x, int y, heap_t *h)
{
	node_t *f = &NDXY(s, x, y);
	pair_t keep = best_son(s, x, y);
	for (int i = 0; i < f->son_n; ++i) {
		pair_t c = f->son[i];
		if (c.x == keep.x && c.y == keep.y) continue;
		detach(s, c, ND(s, c).max_node, h);
	}
	node_t *k = &ND(s, keep);
	if (f->score > k->max_score) {               /* negative edge */
		k->in_de = -1;
		detach(s, keep, k->max_node, h);
		f->son_n = 0;
		f->max_node.x = x; f->max_node.y = y; f->max_score = f->score; f->max_NM = f->tol_NM;
	} else {
		f->son_n = 1; f->son[0] = keep;
		f->max_node = k->max_node; f->max_score = k->max_score; f->max_NM = k->max_NM;
	}
	f->in_de = 0;
}
/* branch_track_new, :873-920 */
static void track_from_leaf(sdp_t *s, int x, int y, heap_t *h)
{
	node_t *n = &NDXY(s, x, y);
	int max_score, max_NM; pair_t max_node;
	n->in_de = -1;
	if (n->son_n == 0) {
		max_node.x = x; max_node.y = y; n->max_node = max_node;
		max_score = n->max_score = n->score;
		max_NM = n->max_NM = n->tol_NM;
	} else { max_node = n->max_node; max_score = n->max_score; max_NM = n->max_NM; }
	pair_t fa = n->from;
	while (fa.x != NONE.x) {
		node_t *f = &ND(s, fa);
		if (f->son_n != 1) {                     /* a branching node: wait for its last son */
			if (--f->in_de == 0) cut_branch(s, fa.x, fa.y, h);
			return;
		}
		if (f->score > max_score) {              /* negative edge: cut below fa */
			pair_t c = f->son[0];
			ND(s, c).in_de = -1;
			detach(s, c, max_node, h);
			f->son_n = 0;
			max_score = f->score; max_NM = f->tol_NM; max_node = fa;
		}
		f->max_score = max_score; f->max_NM = max_NM; f->max_node = max_node; f->in_de = -1;
		fa = f->from;
	}
	path_push(s, max_score, max_NM, max_node, h);
}

/* --------------------------------------------------------------- gap refill -- */
/* frag_mini_dp_line, :1068-1150: chain of multi-hit seeds strictly between `left` and `right` */
static int gap_refill(sdp_t *s, pair_t left, pair_t right, pair_t *out, int *d_score, int *d_NM, int use_head, int has_tail)
{
	const int *tbl = s->P->frag_score_table;          /* == f_BCC_score_table (:1074) */
	pair_t head = use_head ? left : NONE;
	int old_score, old_NM, left_NM = (left.x == NONE.x) ? 0 : HIT(s, left.x, left.y).NM;
	if (!has_tail) { old_score = 1; old_NM = left_NM; }
	else { old_score = 2 + tbl[ND(s, right).match_flag]; old_NM = left_NM + HIT(s, right.x, right.y).NM; }
	const int dp_flag = ST_MULTI;
	for (int i = left.x + 1; i < right.x; ++i)
		for (int j = 0; j < s->map_n[i]; ++j) {
			int f = NDXY(s, i, j).dp_flag;
			if (f == dp_flag || f == 0 - dp_flag) node_init(s, i, j, head, dp_flag);
		}
	for (int i = left.x + 2; i < right.x; ++i)
		for (int j = 0; j < s->map_n[i]; ++j)
			if (NDXY(s, i, j).dp_flag == dp_flag) node_update(s, i, j, left.x + 1, dp_flag);
	int max_score, max_NM = 0, max_n = 0;
	pair_t max_node = head;
	if (!has_tail) {
		max_score = old_score;
		for (int i = right.x - 1; i > left.x; --i)
			for (int j = 0; j < s->map_n[i]; ++j) {
				node_t *n = &NDXY(s, i, j);
				if (n->dp_flag != dp_flag) continue;
				if (n->score > max_score || (n->score == max_score && n->tol_NM < max_NM)) {
					max_score = n->score; max_NM = n->tol_NM; max_node.x = i; max_node.y = j; max_n = n->node_n;
				}
			}
	} else {
		node_t *r = &ND(s, right);
		r->from = head; r->score = old_score; r->tol_NM = old_NM; r->node_n = 1;
		node_update(s, right.x, right.y, left.x + 1, dp_flag);
		max_score = r->score; max_NM = r->tol_NM; max_node = r->from; max_n = r->node_n - 1;
	}
	int k = max_n - 1;
	for (pair_t t = max_node; t.x != head.x; t = ND(s, t).from) {
		if (k < 0) { fprintf(stderr, "[sdp_oracle] gap refill: path longer than its count\n"); exit(1); }
		out[k--] = t;
	}
	if (k >= 0) { fprintf(stderr, "[sdp_oracle] gap refill: path shorter than its count\n"); exit(1); }
	*d_score += max_score - old_score;
	*d_NM += max_NM - old_NM;
	return max_n;
}

/* frag_min_extend, :1031-1066 */
static void promote_colinear(sdp_t *s, int x, int y, int aln_min, int dp_flag)
{
	for (int i = x - 1; i >= 0; --i) {
		if (s->map_n[i] <= aln_min) continue;
		for (int j = 0; j < s->map_n[i]; ++j) {
			int k = edge_kind(s, i, j, x, y);
			if (k == E_MATCH || k == E_MISMATCH || k == E_LONG_MISMATCH) { NDXY(s, i, j).dp_flag = dp_flag; break; }
		}
	}
	for (int i = x + 1; i < s->seed_out; ++i) {
		if (s->map_n[i] <= aln_min) continue;
		for (int j = 0; j < s->map_n[i]; ++j) {
			int k = edge_kind(s, x, y, i, j);
			if (k == E_MATCH || k == E_MISMATCH || k == E_LONG_MISMATCH) { NDXY(s, i, j).dp_flag = dp_flag; break; }
		}
	}
}

/* ------------------------------------------------------ skeleton clustering -- */
/* line_sort_endpos, :10-27: descending last-seed index, equal keys keep their order
 * (glibc qsort is a stable merge sort) */
static void sort_by_end(sdp_t *s, pair_t *L, int *sl, int *rank, int *srank, int ls, int len)
{
	int *pos = malloc((len > 0 ? len : 1) * sizeof(int)), *id = malloc((len > 0 ? len : 1) * sizeof(int));
	for (int i = 0; i < len; ++i) { id[i] = ls + i; pos[i] = SK_NODE(s, L, sl, ls + i)[SK_LEN(sl, ls + i) - 1].x; }
	for (int i = 1; i < len; ++i) {                  /* stable insertion sort, descending */
		int p = pos[i], v = id[i], k = i - 1;
		while (k >= 0 && pos[k] < p) { pos[k + 1] = pos[k]; id[k + 1] = id[k]; --k; }
		pos[k + 1] = p; id[k + 1] = v;
	}
	for (int i = 0; i < len; ++i) { rank[ls + i] = id[i]; srank[id[i]] = ls + i; }
	free(pos); free(id);
}
/* line_merge, :69-112 */
static int merge_pair(sdp_t *s, int a, int b, pair_t *L, int *sl, float ovlp_r)
{
	pair_t *na = SK_NODE(s, L, sl, a), *nb = SK_NODE(s, L, sl, b);
	int la = SK_LEN(sl, a), lb = SK_LEN(sl, b);
	int s1, e1, s2 = na[0].x, e2 = na[la - 1].x, hi, lhi;
	pair_t *nhi;
	if (T_MF(nb, lb) & SK_NMERG) { hi = b; nhi = nb; lhi = lb; s1 = nb[0].x; e1 = nb[lb - 1].x; }
	else { hi = T_MH(nb, lb); nhi = SK_NODE(s, L, sl, hi); lhi = SK_LEN(sl, hi); s1 = T_LB(nhi, lhi); e1 = T_RB(nhi, lhi); }
	int st = s2 > s1 ? s2 : s1, en = e2 < e1 ? e2 : e1;
	float r1 = (en - st + 1 + 0.0) / (e1 - s1 + 1 + 0.0), r2 = (en - st + 1 + 0.0) / (e2 - s2 + 1 + 0.0);
	if (r1 < ovlp_r && r2 < ovlp_r) { T_MF(na, la) = SK_NMERG; return 0; }
	if (T_LS(na, la) <= T_LS(nb, lb) / 2 || T_LS(na, la) <= T_BS(nb, lb) / 2) {
		T_LB(nhi, lhi) = s1; T_RB(nhi, lhi) = e1;
		T_MF(nhi, lhi) = SK_MERGH;
		T_MF(na, la) = SK_MERGB; T_MH(na, la) = hi;
		T_MF(na, la) |= SK_DUMP;
		return 1;
	}
	T_LB(nhi, lhi) = s1 + s2 - st; T_RB(nhi, lhi) = e1 + e2 - en;
	T_MF(nhi, lhi) = SK_MERGH;
	T_MF(na, la) = SK_MERGB; T_MH(na, la) = hi;
	if (T_BS(nb, lb) > T_BS(na, la)) T_BS(na, la) = T_BS(nb, lb);
	return 1;
}

#define MFI(i) T_MF(SK_NODE(s, L, sl, i), SK_LEN(sl, i))
#define MHI(i) T_MH(SK_NODE(s, L, sl, i), SK_LEN(sl, i))
#define LSI(i) T_LS(SK_NODE(s, L, sl, i), SK_LEN(sl, i))
#define NMI(i) T_NM(SK_NODE(s, L, sl, i), SK_LEN(sl, i))

/* Picks the kept skeletons of one cluster (shared part of line_filter :160-235 and line_filter1
 * :345-400).  cl = members (skeleton id, score, NM) in rank order; kept[] receives
 * [0]=best, [1..]=head then bodies; tri_n (may be NULL) is zeroed for dumped members. */
typedef struct { int x, y, z; } tri_t;
static int pick_in_cluster(sdp_t *s, pair_t *L, int *sl, const tri_t *cl, int cn, heap_t *h, int *kept, int *tri_n)
{
	int b_score = 0, s_score = 0, kn = 1;
	for (int j = 0; j < cn; ++j) {
		if (cl[j].y > b_score) { s_score = b_score; b_score = cl[j].y; }
		else if (cl[j].y > s_score) s_score = cl[j].y;
	}
	if (s_score >= b_score / 2) {
		h->n = 0;
		for (int j = 0; j < cn; ++j) {
			if (cl[j].y >= b_score / 2) {
				pair_t nd = { cl[j].x, -1 };
				int ret = heap_offer(h, nd, cl[j].y, cl[j].z);
				if (ret == -2) { MFI(cl[j].x) |= SK_DUMP; if (tri_n) tri_n[cl[j].x] = 0; }
				else if (ret != -1) { MFI(ret) |= SK_DUMP; if (tri_n) tri_n[ret] = 0; }
			} else { MFI(cl[j].x) |= SK_DUMP; if (tri_n) tri_n[cl[j].x] = 0; }
		}
		heap_build_minpos(h);
		int head = heap_take_minpos(h).x;
		pair_t *hn = SK_NODE(s, L, sl, head); int hl = SK_LEN(sl, head);
		T_MF(hn, hl) = SK_MERGH;
		int min_l = hn[0].x, max_r = hn[hl - 1].x, body;
		if (T_LS(hn, hl) == b_score) kept[0] = head;
		kept[kn++] = head;
		while ((body = heap_take_minpos(h).x) != -1) {
			pair_t *bn = SK_NODE(s, L, sl, body); int bl = SK_LEN(sl, body);
			T_MF(bn, bl) = SK_MERGB; T_MH(bn, bl) = head;
			if (bn[0].x < min_l) min_l = bn[0].x;
			if (bn[bl - 1].x > max_r) max_r = bn[bl - 1].x;
			if (T_LS(bn, bl) == b_score) kept[0] = body;
			kept[kn++] = body;
		}
		T_LB(hn, hl) = hn[0].x < min_l ? hn[0].x : min_l;
		T_RB(hn, hl) = hn[hl - 1].x > max_r ? hn[hl - 1].x : max_r;
	} else {
		for (int j = 0; j < cn; ++j) {
			if (cl[j].y == b_score) { MFI(cl[j].x) = SK_NMERG; kept[0] = cl[j].x; kept[kn++] = cl[j].x; }
			else { MFI(cl[j].x) |= SK_DUMP; if (tri_n) tri_n[cl[j].x] = 0; }
		}
	}
	return kn;
}

/* line_filter, :122-319 */
static void filter_stage1(sdp_t *s, pair_t *L, int *sl, int *rank, int *srank, int ls, int len,
                          trig_t **trg, int *tri_n, int per_max)
{
	tri_t **cl = malloc(len * sizeof(tri_t *));
	int **kept = malloc(len * sizeof(int *));
	int *cn = malloc(len * sizeof(int)), *kn = malloc(len * sizeof(int));
	for (int i = 0; i < len; ++i) { cl[i] = malloc(len * sizeof(tri_t)); kept[i] = malloc((len + 1) * sizeof(int)); }
	int m_i = -1;
	for (int _i = ls; _i < ls + len; ++_i) {
		int i = rank[_i], mf = MFI(i);
		if (mf & SK_DUMP) continue;
		if (mf & SK_NMERG) { ++m_i; cl[m_i][0].x = i; cl[m_i][0].y = -2; cn[m_i] = 1; }
		else if (mf & SK_MERGH) { ++m_i; cl[m_i][0].x = i; cl[m_i][0].y = LSI(i); cl[m_i][0].z = NMI(i); cn[m_i] = 1; }
		else { cl[m_i][cn[m_i]].x = i; cl[m_i][cn[m_i]].y = LSI(i); cl[m_i][cn[m_i]].z = NMI(i); cn[m_i]++; }
	}
	heap_t *h = heap_new(per_max);
	for (int i = 0; i <= m_i; ++i) {
		if (cl[i][0].y == -2) { kept[i][0] = cl[i][0].x; kn[i] = 1; continue; }
		kn[i] = pick_in_cluster(s, L, sl, cl[i], cn[i], h, kept[i], tri_n);
		/* inter-skeleton candidates inside the gaps ("triggers") of the kept ones, :236-274 */
		for (int ii = 1; ii < kn[i]; ++ii) {
			int j = kept[i][ii], _j = srank[j];
			for (int k = 0; k < tri_n[j]; ++k) {
				int head = -1;
				pair_t t1 = trg[j][k].n1, t2 = trg[j][k].n2;
				for (int _l = _j + 1; _l < ls + len; ++_l) {
					int l = rank[_l];
					pair_t *nl = SK_NODE(s, L, sl, l); int ll = SK_LEN(sl, l);
					if ((T_MF(nl, ll) & 0x3) != 0) break;
					if (!(nl[0].x > t1.x && nl[ll - 1].x < t2.x)) continue;
					int mfk = ND(s, t2).match_flag;
					if (mfk != E_MISMATCH && mfk != E_LONG_MISMATCH) continue;
					const lb2_sdp_hit *hs = &HIT(s, nl[0].x, nl[0].y), *he = &HIT(s, nl[ll - 1].x, nl[ll - 1].y);
					const lb2_sdp_hit *h1 = &HIT(s, t1.x, t1.y), *h2 = &HIT(s, t2.x, t2.y);
					int st = hs->nstrand;
					if (st == h1->nstrand || hs->nchr != h1->nchr ||
					    st * hs->offset < st * h2->offset || st * he->offset > st * h1->offset) continue;
					T_ELB(nl, ll) = t1.x; T_ERB(nl, ll) = t2.x;
					T_MF(nl, ll) = SK_INTER;
					if (head == -1) { T_MF(nl, ll) |= SK_NMERG; head = l; }
					else { T_MF(nl, ll) |= SK_MERGB; T_MH(nl, ll) = head; MFI(head) = SK_INTER | SK_MERGH; }
				}
			}
		}
	}
	heap_free(h);
	/* a one/two-seed cluster at either end next to a real one is dropped, :279-316 */
	if (m_i > 0) {
		for (int side = 0; side < 2; ++side) {
			int a = side == 0 ? 0 : m_i, b = side == 0 ? 1 : m_i - 1;
			pair_t *na = SK_NODE(s, L, sl, kept[a][0]); int la = SK_LEN(sl, kept[a][0]);
			pair_t *nb = SK_NODE(s, L, sl, kept[b][0]); int lb = SK_LEN(sl, kept[b][0]);
			int span_a = na[la - 1].x - na[0].x, span_b = nb[lb - 1].x - nb[0].x;
			if (span_a < 2 && span_b >= 2) {
				for (int i = 1; i < kn[a]; ++i) {
					MFI(kept[a][i]) = SK_DUMP;
					for (int _j = ls; _j < ls + len; ++_j) {
						int j = rank[_j], mf = MFI(j);
						if (!(mf & SK_NMERG) && !(mf & SK_MERGH) && !(mf & SK_DUMP) && MHI(j) == kept[a][i]) MFI(j) = SK_DUMP;
					}
				}
			}
		}
	}
	for (int i = 0; i < len; ++i) { free(cl[i]); free(kept[i]); }
	free(cl); free(kept); free(cn); free(kn);
}
/* line_filter1, :321-404 */
static void filter_stage2(sdp_t *s, pair_t *L, int *sl, int *rank, int ls, int len, int per_max)
{
	tri_t **cl = malloc(len * sizeof(tri_t *));
	int *cn = malloc(len * sizeof(int)), *kept = malloc((len + 2) * sizeof(int));
	for (int i = 0; i < len; ++i) cl[i] = malloc(len * sizeof(tri_t));
	int m_i = -1;
	for (int _i = ls; _i < ls + len; ++_i) {
		int i = rank[_i], mf = MFI(i);
		if ((mf & SK_DUMP) || (mf & SK_NMERG)) continue;
		if (mf & SK_MERGH) { ++m_i; cl[m_i][0].x = i; cl[m_i][0].y = LSI(i); cl[m_i][0].z = NMI(i); cn[m_i] = 1; }
		else { cl[m_i][cn[m_i]].x = i; cl[m_i][cn[m_i]].y = LSI(i); cl[m_i][cn[m_i]].z = NMI(i); cn[m_i]++; }
	}
	heap_t *h = heap_new(per_max);
	for (int i = 0; i <= m_i; ++i) pick_in_cluster(s, L, sl, cl[i], cn[i], h, kept, 0);
	heap_free(h);
	for (int i = 0; i < len; ++i) free(cl[i]);
	free(cl); free(cn); free(kept);
}
/* line_remove, :406-423 */
static int drop_dumped(sdp_t *s, pair_t *L, int *sl, int *rank, int ls, int len)
{
	int cur = ls;
	for (int _l = ls; _l < ls + len; ++_l) {
		int l = rank[_l];
		if (!(MFI(l) & SK_DUMP)) rank[cur++] = l;
	}
	return cur - ls;
}
/* extend boundaries of the kept skeletons, :443-493 (identical in line_set_bound1 :513-565) */
static void set_extents(sdp_t *s, pair_t *L, int *sl, int *rank, int ls, int len, int left, int right)
{
	int _i, i = 0, l, r = right, m, st;
	pair_t *ni = 0; int li = 0;
	for (_i = ls; _i < ls + len; ++_i) {
		i = rank[_i]; ni = SK_NODE(s, L, sl, i); li = SK_LEN(sl, i);
		if (!(T_MF(ni, li) & SK_DUMP)) { T_ERB(ni, li) = r; break; }
	}
	if (!ni) return;
	st = (T_MF(ni, li) & SK_NMERG) ? ni[0].x : T_LB(ni, li);
	m = 0;
	for (++_i; _i < ls + len; ++_i) {
		i = rank[_i]; ni = SK_NODE(s, L, sl, i); li = SK_LEN(sl, i);
		if (T_MF(ni, li) & (SK_DUMP | SK_INTER)) continue;
		if (T_MF(ni, li) & (SK_NMERG | SK_MERGH)) {
			r = st; l = (T_MF(ni, li) & SK_NMERG) ? ni[li - 1].x : T_RB(ni, li);
			for (int _j = _i - 1; _j >= m; --_j) {
				int j = rank[_j];
				if (MFI(j) & (SK_DUMP | SK_INTER)) continue;
				T_ELB(SK_NODE(s, L, sl, j), SK_LEN(sl, j)) = l;
			}
			st = (T_MF(ni, li) & SK_NMERG) ? ni[0].x : T_LB(ni, li);
			m = _i;
		}
		T_ERB(ni, li) = r;
	}
	for (int _j = _i - 1; _j >= m; --_j) {
		int j = rank[_j];
		if (MFI(j) & (SK_DUMP | SK_INTER)) continue;
		T_ELB(SK_NODE(s, L, sl, j), SK_LEN(sl, j)) = left;
	}
}
/* line_set_bound :425-494 (stage 1, trg != NULL) / line_set_bound1 :496-566 (stage 2) */
static void cluster_skeletons(sdp_t *s, pair_t *L, int *sl, int *rank, int *srank, int ls, int *o_len, int left, int right,
                              trig_t **trg, int *tri_n, int stage1)
{
	if (*o_len <= 0) return;
	int len = *o_len;
	sort_by_end(s, L, sl, rank, srank, ls, len);
	MFI(rank[ls]) = SK_NMERG;
	for (int i = 1; i < len; ++i) merge_pair(s, rank[ls + i], rank[ls + i - 1], L, sl, s->P->ovlp_rat);
	if (stage1) filter_stage1(s, L, sl, rank, srank, ls, len, trg, tri_n, s->P->ske_max);
	else filter_stage2(s, L, sl, rank, ls, len, s->P->ske_max);
	*o_len = len = drop_dumped(s, L, sl, rank, ls, len);
	set_extents(s, L, sl, rank, ls, len, left, right);
}

/* line_filter_overlap, :568-594 */
static void drop_ref_overlaps(sdp_t *s, pair_t *L, int *sl, int *rank, int line_n)
{
	int seed_len = s->P->seed_len;
	for (int _i = 0; _i < line_n; ++_i) {
		int i = rank[_i];
		pair_t *ni = SK_NODE(s, L, sl, i); int li = SK_LEN(sl, i), last = 0;
		for (int j = 1; j < li; ++j) {
			int is_tail = (j == li - 1);
			if (is_tail && li - 1 == last) break;
			pair_t c = ni[j], p = ni[last];
			const lb2_sdp_hit *hc = &HIT(s, c.x, c.y), *hp = &HIT(s, p.x, p.y);
			int ovl = seed_len + ((hc->nstrand == 1) ? hp->len_dif : hc->len_dif) > hc->nstrand * (hc->offset - hp->offset)
			          && ND(s, c).match_flag != E_INSERT;
			if (is_tail) { if (ovl) ni[last].x = -1; }
			else if (ovl) ni[j].x = -1;
			else last = j;
		}
	}
}

/* frag_dp_path, :1152-1250 -> skeleton stream */
typedef struct { int32_t *w; int64_t n, cap; } stream_t;
static void put(stream_t *o, int v) { if (o->n < o->cap) o->w[o->n] = v; ++o->n; }
static void emit_skeletons(sdp_t *s, pair_t *L, int *sl, int *rank, int line_n, stream_t *o)
{
	put(o, line_n);
	if (line_n == 0) return;
	if (s->P->aln_mode & 1) drop_ref_overlaps(s, L, sl, rank, line_n);
	for (int _l = 0; _l < line_n; ++_l) {
		int l = rank[_l];
		pair_t *ln = SK_NODE(s, L, sl, l); int ll = SK_LEN(sl, l);
		put(o, T_LS(ln, ll));
		int64_t frag_num_at = o->n; put(o, 0);
		int frag_num = 0;
		int64_t seed_num_at = o->n; put(o, 1);           /* FRAG_END opens a fragment with its first seed */
		int seed_num = 1;
		pair_t pre = ln[ll - 1], cur;
		put(o, pre.x); put(o, pre.y);
		for (int i = ll - 1; i > 0; --i) {
			cur = pre;
			if (ln[i - 1].x < 0) continue;
			pre = ln[i - 1];
			int mf = ND(s, cur).match_flag;
			if (mf == E_INSERT || mf == E_DELETE || mf == E_MISMATCH || mf == E_LONG_MISMATCH) {
				if (seed_num_at < o->cap) o->w[seed_num_at] = seed_num;
				++frag_num;
				seed_num_at = o->n; put(o, 1); seed_num = 1;
				put(o, pre.x); put(o, pre.y);
			} else if (mf == E_MATCH) {
				put(o, pre.x); put(o, pre.y); ++seed_num;
			} else { fprintf(stderr, "[sdp_oracle] unknown edge kind %d on a skeleton\n", mf); exit(1); }
		}
		if (seed_num_at < o->cap) o->w[seed_num_at] = seed_num;
		if (frag_num_at < o->cap) o->w[frag_num_at] = frag_num + 1;
	}
}

/* ------------------------------------------------------------------ stage 1 -- */
/* frag_line_BCC, :1305-1445 */
static void stage_bcc(sdp_t *s, stream_t *o)
{
	const lb2_sdp_para *P = s->P;
	int min_n = P->first_loci_thd, min_exist = 0, min_num = 0;
	for (int i = 0; i < s->seed_out; ++i) {
		int fl = ST_MULTI;
		if (s->map_n[i] <= min_n) { fl = ST_MIN; min_exist = 1; ++min_num; }
		for (int j = 0; j < s->map_n[i]; ++j) node_init(s, i, j, NONE, fl);
	}
	if (!min_exist || min_num * 3 < s->seed_out) {
		for (int i = 0; i < s->seed_out; ++i)
			for (int j = 0; j < s->map_n[i]; ++j) NDXY(s, i, j).dp_flag = ST_MIN;
		min_n = P->per_aln_m;
	}
	if (min_n != P->per_aln_m)
		for (int i = 0; i < s->seed_out; ++i)
			if (s->map_n[i] <= min_n)
				for (int j = 0; j < s->map_n[i]; ++j) promote_colinear(s, i, j, min_n, ST_MIN);
	for (int i = 1; i < s->seed_out; ++i)
		for (int j = 0; j < s->map_n[i]; ++j)
			if (NDXY(s, i, j).dp_flag == ST_MIN) node_update(s, i, j, 0, ST_MIN);

	heap_t *h = heap_new(s->line_n_max);
	h->thd = 2;
	for (int i = s->seed_out - 1; i >= 0; --i)
		for (int j = 0; j < s->map_n[i]; ++j)
			if (NDXY(s, i, j).dp_flag == ST_MIN && NDXY(s, i, j).in_de == 0) track_from_leaf(s, i, j, h);

	int l_i = 0, min_l = h->n, o_l = h->n, next_start = 0;
	trig_t **trg = malloc((o_l > 0 ? o_l : 1) * sizeof(trig_t *));
	int *tri_n = malloc((o_l > 0 ? o_l : 1) * sizeof(int));
	for (int i = 0; i < o_l; ++i) trg[i] = malloc((s->seed_out > 0 ? s->seed_out : 1) * sizeof(trig_t));
	pair_t *L = s->line; int *sl = s->lsl;
	for (;;) {
		int line_score, line_NM;
		pair_t max_node = heap_pop_last(h, &line_score, &line_NM);
		if (max_node.x == -1) break;
		int node_i = 0, mini_len;
		tri_n[l_i] = 0; sl[l_i << 1] = next_start;
		pair_t *ln = L + next_start, last_n, right, left;
		if (max_node.x < s->seed_out - 1) {                 /* refill to the right of the path end */
			pair_t bound = { s->seed_out, 0 };
			mini_len = gap_refill(s, max_node, bound, s->mini, &line_score, &line_NM, 1, 0);
			for (int k = mini_len - 1; k >= 0; --k) { ln[node_i++] = s->mini[k]; ND(s, s->mini[k]).dp_flag = ST_TRACKED; }
			ln[node_i] = max_node;
			last_n = ln[0];
			for (int k = mini_len - 1; k >= 0; --k) {
				if (last_n.x - ln[node_i - k].x > 2) { trg[l_i][tri_n[l_i]].n1 = ln[node_i - k]; trg[l_i][tri_n[l_i]].n2 = last_n; ++tri_n[l_i]; }
				last_n = ln[node_i - k];
			}
		}
		right = max_node;
		while (right.x != NONE.x) {
			ln[node_i++] = right;
			left = ND(s, right).from;
			if (left.x < right.x - 1) {                      /* refill every gap of the path */
				mini_len = gap_refill(s, left, right, s->mini, &line_score, &line_NM, 1, 1);
				for (int k = mini_len - 1; k >= 0; --k) { ln[node_i++] = s->mini[k]; ND(s, s->mini[k]).dp_flag = ST_TRACKED; }
				ln[node_i] = left;
				last_n = right;
				for (int k = mini_len; k >= 0; --k) {
					if (last_n.x - ln[node_i - k].x > 2) {
						if (ln[node_i - k].x == NONE.x) continue;
						trg[l_i][tri_n[l_i]].n1 = ln[node_i - k]; trg[l_i][tri_n[l_i]].n2 = last_n; ++tri_n[l_i];
					}
					last_n = ln[node_i - k];
				}
			}
			right = left;
		}
		for (int k = 0; k < node_i / 2; ++k) { pair_t t = ln[k]; ln[k] = ln[node_i - k - 1]; ln[node_i - k - 1] = t; }
		sl[(l_i << 1) + 1] = node_i;
		T_LS(ln, node_i) = line_score; T_BS(ln, node_i) = line_score; T_NM(ln, node_i) = line_NM;
		++l_i; next_start += node_i + SK_EXTRA;
	}
	cluster_skeletons(s, L, sl, s->rank, s->srank, 0, &min_l, -1, s->seed_out, trg, tri_n, 1);
	for (int i = 0; i < o_l; ++i) free(trg[i]);
	free(trg); free(tri_n); heap_free(h);
	emit_skeletons(s, L, sl, s->rank, min_l, o);
}

/* ------------------------------------------------------------------ stage 2 -- */
typedef struct { int chr; int64_t pos; } refpt_t;
typedef struct { int beg, end; refpt_t *rb, *re; int bn, en, bm, em; } region_t;
static void pts_push(refpt_t **a, int *n, int *m, const refpt_t *src, int cnt)
{
	for (int i = 0; i < cnt; ++i) {
		if (*n == *m) { *m = *m ? *m * 2 : 4; *a = realloc(*a, *m * sizeof(refpt_t)); }
		(*a)[(*n)++] = src[i];
	}
}
/* get_remain_reg, src/lamsa_aln.c:548-572 (+ aln_sort_reg :477, aln_merg_reg :498-519, push_reg :533-546):
 * the read intervals stage 1 left unaligned, each with the reference ends of its flanks */
static int remaining_regions(const sdp_t *s, const lb2_sdp_reg *regs, int n_reg, region_t **out)
{
	int read_len = s->read_len, lo = s->P->seed_len, hi = read_len, n_out = 0;
	region_t *R = calloc(n_reg + 2, sizeof(region_t));
	*out = R;
	if (n_reg == 0) {
		if (lo < read_len && read_len <= hi) { R[0].beg = 1; R[0].end = read_len; return 1; }
		return 0;
	}
	region_t *A = calloc(n_reg, sizeof(region_t));
	int *ord = malloc(n_reg * sizeof(int));
	for (int i = 0; i < n_reg; ++i) ord[i] = i;
	for (int i = 1; i < n_reg; ++i) {                /* stable, ascending beg */
		int v = ord[i], k = i - 1;
		while (k >= 0 && regs[ord[k]].beg > regs[v].beg) { ord[k + 1] = ord[k]; --k; }
		ord[k + 1] = v;
	}
	for (int i = 0; i < n_reg; ++i) {
		const lb2_sdp_reg *g = regs + ord[i];
		refpt_t b = { g->chr, g->ref_beg }, e = { g->chr, g->ref_end };
		A[i].beg = g->beg; A[i].end = g->end;
		pts_push(&A[i].rb, &A[i].bn, &A[i].bm, &b, 1);
		pts_push(&A[i].re, &A[i].en, &A[i].em, &e, 1);
	}
	int cur = 0;
	for (int i = 1; i < n_reg; ++i) {                /* merge neighbours closer than bwt_seed_len */
		if (A[i].beg - A[cur].end - 1 < s->P->bwt_seed_len) {
			if (A[i].end > A[cur].end) A[cur].end = A[i].end;
			pts_push(&A[cur].rb, &A[cur].bn, &A[cur].bm, A[i].rb, A[i].bn);
			pts_push(&A[cur].re, &A[cur].en, &A[cur].em, A[i].re, A[i].en);
		} else {
			++cur;
			if (cur != i) {
				A[cur].beg = A[i].beg; A[cur].end = A[i].end; A[cur].bn = A[cur].en = 0;
				pts_push(&A[cur].rb, &A[cur].bn, &A[cur].bm, A[i].rb, A[i].bn);
				pts_push(&A[cur].re, &A[cur].en, &A[cur].em, A[i].re, A[i].en);
			}
		}
	}
	int n = cur + 1, i;
#define EMIT(b_, e_, pb, pbn, pe, pen) do { region_t *q = R + n_out++; q->beg = (b_); q->end = (e_); \
		pts_push(&q->rb, &q->bn, &q->bm, (pb), (pbn)); pts_push(&q->re, &q->en, &q->em, (pe), (pen)); } while (0)
	if (A[0].beg > lo && A[0].beg - 1 <= hi) EMIT(1, A[0].beg - 1, 0, 0, A[0].rb, A[0].bn);
	for (i = 1; i < n; ++i)
		if (A[i].beg - A[i - 1].end > lo && A[i].beg - 1 - A[i - 1].end <= hi)
			EMIT(A[i - 1].end + 1, A[i].beg - 1, A[i - 1].re, A[i - 1].en, A[i].rb, A[i].bn);
	if (read_len - A[i - 1].end > lo && read_len - A[i - 1].end <= hi)
		EMIT(A[i - 1].end + 1, read_len, A[i - 1].re, A[i - 1].en, 0, 0);
#undef EMIT
	for (i = 0; i < n_reg; ++i) { free(A[i].rb); free(A[i].re); }
	free(A); free(ord);
	return n_out;
}

/* frag_mini_dp_multi_line, :923-1017: chaining restricted to seeds strictly inside (left_b, right_b) */
static int region_dp(sdp_t *s, int left_b, int right_b, const region_t *rg, pair_t *L, int *sl)
{
	if (left_b + 1 >= right_b) return 0;
	const lb2_sdp_para *P = s->P;
	int start = left_b + 1, end = right_b - 1, dp_flag = ST_WHOLE;
	for (int i = start; i <= end; ++i)
		for (int j = 0; j < s->map_n[i]; ++j)
			if (NDXY(s, i, j).dp_flag != ST_TRACKED) node_init(s, i, j, NONE, dp_flag);
	for (int i = start + 1; i <= end; ++i)
		for (int j = 0; j < s->map_n[i]; ++j)
			if (NDXY(s, i, j).dp_flag == dp_flag) node_update(s, i, j, start, dp_flag);
	heap_t *h = heap_new(s->line_n_max);
	h->thd = 0;
	for (int i = end; i >= start; --i)
		for (int j = 0; j < s->map_n[i]; ++j)
			if (NDXY(s, i, j).dp_flag == dp_flag && NDXY(s, i, j).in_de == 0) track_from_leaf(s, i, j, h);
	int l_i = 0, next_start = 0;
	for (;;) {
		int score, NM;
		pair_t right = heap_pop_last(h, &score, &NM);
		if (right.x == NONE.x) break;
		int node_i = ND(s, right).node_n - 1;
		sl[l_i << 1] = next_start; sl[(l_i << 1) + 1] = node_i + 1;
		next_start += node_i + 1 + SK_EXTRA;
		const lb2_sdp_hit *hr = &HIT(s, right.x, right.y);
		int near = 0;                                  /* bonus when the path end is near an aligned flank, :979-996 */
		for (int i = 0; i < rg->bn && !near; ++i)
			if (hr->nchr == rg->rb[i].chr &&
			    labs((long)((hr->offset - rg->rb[i].pos) - (right.x - left_b) * P->seed_step)) < P->SV_len_thd) near = 1;
		for (int i = 0; i < rg->en && !near; ++i)
			if (hr->nchr == rg->re[i].chr &&
			    labs((long)((hr->offset - rg->re[i].pos) - (right.x - left_b) * P->seed_step)) < P->SV_len_thd) near = 1;
		if (near) { if (score > 1) score += score / 2; else score++; }
		pair_t *node = SK_NODE(s, L, sl, l_i); int len = SK_LEN(sl, l_i);
		T_LS(node, len) = score; T_BS(node, len) = score; T_NM(node, len) = NM;
		for (pair_t t = right; t.x != NONE.x; t = ND(s, t).from) {
			if (node_i < 0) { fprintf(stderr, "[sdp_oracle] region dp: path longer than its count\n"); exit(1); }
			node[node_i--] = t;
		}
		if (node_i >= 0) { fprintf(stderr, "[sdp_oracle] region dp: path shorter than its count\n"); exit(1); }
		++l_i;
	}
	heap_free(h);
	return l_i;
}

/* frag_line_remain, :1252-1302 */
static void stage_remain(sdp_t *s, const lb2_sdp_reg *regs, int n_reg, stream_t *o)
{
	const lb2_sdp_para *P = s->P;
	region_t *R;
	int nR = remaining_regions(s, regs, n_reg, &R);
	int l_n = 0, next_start = 0;
	for (int i = 0; i < nR; ++i) {
		int left_id = (R[i].beg + P->seed_inv - 1) / P->seed_step + 1;
		int right_id = (R[i].end - 1) / P->seed_step + 1;
		if (right_id > s->seed_all) right_id -= 1;
		int left = -2, right = -2;
		for (int j = 0; j < s->seed_out; ++j) if (s->seed_id[j] >= left_id) { left = j - 1; break; }
		if (left == -2) continue;
		for (int j = s->seed_out - 1; j >= 0; --j) if (s->seed_id[j] <= right_id) { right = j + 1; break; }
		if (right == -2) continue;
		int l = region_dp(s, left, right, R + i, s->tline, s->tlsl);
		cluster_skeletons(s, s->tline, s->tlsl, s->trank, s->srank, 0, &l, left, right, 0, 0, 0);
		for (int _j = 0; _j < l; ++_j) {
			int j = s->trank[_j], len = s->tlsl[(j << 1) + 1];
			s->lsl[(l_n + _j) << 1] = next_start; s->lsl[((l_n + _j) << 1) + 1] = len;
			for (int k = 0; k < len + SK_EXTRA; ++k) s->line[next_start + k] = s->tline[s->tlsl[j << 1] + k];
			next_start += len + SK_EXTRA;
			s->rank[l_n + _j] = l_n + _j;
		}
		l_n += l;
	}
	for (int i = 0; i < nR; ++i) { free(R[i].rb); free(R[i].re); }
	free(R);
	emit_skeletons(s, s->line, s->lsl, s->rank, l_n, o);
}

/* ------------------------------------------------------------------- driver -- */
/* Same contract as ref_sdp_run_batch (oracle/sdp_ref_shim.c).  pairs (may be NULL) receives the
 * number of edge classifications done inside predecessor scans, per stage [bcc, remain]. */
int orc_sdp_run_batch(const lb2_sdp_para *P, int64_t n_reads, const lb2_sdp_read *reads,
                      const int32_t *seed_id, const int32_t *map_n, const lb2_sdp_hit *hits,
                      const lb2_sdp_reg *regs, int stages,
                      int32_t *out1, int64_t cap1, int64_t *off1,
                      int32_t *out2, int64_t cap2, int64_t *off2, int64_t *pairs)
{
	stream_t o1 = { out1, 0, cap1 }, o2 = { out2, 0, cap2 };
	int64_t pr[2] = { 0, 0 };
	for (int64_t r = 0; r < n_reads; ++r) {
		const lb2_sdp_read *rd = reads + r;
		sdp_t s;
		memset(&s, 0, sizeof s);
		s.P = P; s.seed_out = rd->seed_out; s.seed_all = rd->seed_all; s.read_len = rd->read_len;
		s.seed_id = seed_id + rd->seed_first; s.map_n = map_n + rd->seed_first; s.hit = hits + rd->hit_first;
		s.hoff = malloc((rd->seed_out + 2) * sizeof(int64_t));
		int64_t H = 0;
		for (int i = 0; i < rd->seed_out; ++i) { s.hoff[i] = H; H += s.map_n[i]; }
		s.hoff[rd->seed_out] = s.hoff[rd->seed_out + 1] = H;
		s.n_nodes = H;
		s.nd = calloc(H + 1, sizeof(node_t));
		for (int64_t k = 0; k < H; ++k) { s.nd[k].son_max = 4; s.nd[k].son = calloc(4, sizeof(pair_t)); }
		int64_t cap = 6 * H + 16 + rd->seed_out;
		s.line = malloc(cap * sizeof(pair_t)); s.tline = malloc(cap * sizeof(pair_t)); s.mini = malloc((rd->seed_out + 2) * sizeof(pair_t));
		s.lsl = malloc((2 * H + 2) * sizeof(int)); s.tlsl = malloc((2 * H + 2) * sizeof(int));
		s.rank = malloc((H + 1) * sizeof(int)); s.trank = malloc((H + 1) * sizeof(int)); s.srank = malloc((H + 1) * sizeof(int));
		s.line_n_max = (int)(H + 1);
		off1[r] = o1.n; off2[r] = o2.n;
		if (stages & 1) { stage_bcc(&s, &o1); pr[0] += s.pairs; s.pairs = 0; }
		if (stages & 2) { stage_remain(&s, regs + rd->reg_first, rd->n_reg, &o2); pr[1] += s.pairs; }
		for (int64_t k = 0; k < H; ++k) free(s.nd[k].son);
		free(s.nd); free(s.hoff); free(s.line); free(s.tline); free(s.mini);
		free(s.lsl); free(s.tlsl); free(s.rank); free(s.trank); free(s.srank);
	}
	off1[n_reads] = o1.n; off2[n_reads] = o2.n;
	if (pairs) { pairs[0] = pr[0]; pairs[1] = pr[1]; }
	return (o1.n > cap1 || o2.n > cap2) ? -1 : 0;
}

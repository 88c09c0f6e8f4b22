/*
 * hash_oracle.c -- CPU restatement of the seed-and-chain half of the reference's local split mapping
 * (`hash_split_map`, /root/reference/src/split_mapping.c:634-686): k-mer index of the reference window
 * (init_hash :181-208), look-up of the read's k-mers (:654-675) and the chaining of the hits into one line
 * (hash_main_line :492-602 with hash_main_dis :218-261, hash_dp_init :264-310, hash_min_extend :312-338,
 * hash_dp_update :341-397, hash_mini_dp_init :399-441, mini_hash_main_line :444-488).
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else) as the checker of the CUDA kernel
 * lamsa_b200/csrc/hash_line.cuh.  Parity is PINNED: tests/test_hash_oracle.py runs it against the unmodified
 * reference functions (oracle/_ref/liblamsa_ref.so through oracle/hash_ref_shim.c) on seeded SV-shaped inputs.
 *
 * Flat formulation shared with the kernel: nodes are numbered head = 0, the hits seed-major 1..N, tail = N+1;
 * `from` is a node number (-1: none); the hits of one k-mer are the window positions in ascending order
 * (init_hash_core stores them in scan order, :202-206); a k-mer with more than 50 positions has no hits (:669).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { HF_MATCH = 0, HF_MISMATCH = 2, HF_MATCH_THD = 2, HF_LONG_MISMATCH = 3, HF_INSERT = 4, HF_DELETE = 5,
       HF_UNCONNECT = 8, HF_UNMATCH = 9 };                       /* src/lamsa_aln.h:101-113 */
enum { FLAG_MIN = 1, FLAG_MULTI = 2, FLAG_UNLIMITED = 3 };       /* src/split_mapping.h:60-64 */
enum { SV_PEN = 2, MAX_HITS = 50 };                              /* src/split_mapping.h:75, src/split_mapping.c:669 */

typedef struct {
	int hash_len, hash_step, split_len;     /* lamsa_aln_para: hash_len, hash_step, split_len */
	int ref_len, read_len, ref_offset;
} hpar;

typedef struct { int x, read_i, offset, from, score, node_n, mflag, dflag; } hnode;

static int iabs(int v) { return v < 0 ? -v : v; }

/* relation of two hits (a: read position a_i, diagonal a_off; b likewise); src/split_mapping.c:218-261 */
static int relation(const hpar *P, int a_i, int a_off, int b_i, int b_off)
{
	const int L = P->hash_len, st = P->hash_step;
	const int d = a_i > b_i ? a_off - b_off : b_off - a_off;
	const int gap = iabs(b_i - a_i);
	if (d == 0) return gap < L + 2 * st ? HF_MATCH : gap < L + 6 * st ? HF_MISMATCH : HF_LONG_MISMATCH;
	if (d > 0) return HF_DELETE;
	if (d >= -(gap - L)) return HF_INSERT;
	if (d > -(P->split_len / 2)) return HF_UNCONNECT;
	/* overlapping insertion of at least split_len/2 */
	{
		const int lo_i = b_i > a_i ? a_i : b_i, lo_off = b_i > a_i ? a_off : b_off;
		const int hi_i = b_i > a_i ? b_i : a_i, hi_off = b_i > a_i ? b_off : a_off;
		int ok;
		if (P->ref_offset > 0) ok = P->read_len - P->ref_len + hi_off >= -(lo_i + L - 1) && P->read_len - lo_off >= hi_i;
		else ok = hi_off >= -(lo_i - 1) && P->ref_len - lo_off >= hi_i;
		return ok ? HF_INSERT : HF_UNCONNECT;
	}
}

static int edge_pen(int rel) { return rel <= HF_MATCH_THD ? 0 : SV_PEN; }

/* (re)initialise one hit against `head`: src/split_mapping.c:264-310 / :399-441 */
static void init_node(const hpar *P, hnode *nd, int id, int head, int flag)
{
	hnode *v = nd + id;
	if (nd[head].dflag == FLAG_UNLIMITED) { v->from = head; v->score = 1; v->node_n = 1; v->mflag = HF_MATCH; v->dflag = flag; return; }
	{
		const int rel = relation(P, nd[head].read_i, nd[head].offset, v->read_i, v->offset);
		if (rel == HF_UNCONNECT) { v->from = -1; v->score = 0; v->node_n = 0; v->mflag = rel; v->dflag = -flag; }
		else { v->from = head; v->score = 2 - edge_pen(rel); v->node_n = 1; v->mflag = rel; v->dflag = flag; }
	}
}

/* best predecessor of node `id` among the nodes of seeds [first_x, x-1] carrying `flag`, seeds descending, hits of
 * a seed ascending, a strictly better score wins; src/split_mapping.c:341-397 */
static void update_node(const hpar *P, hnode *nd, const int *start, const int *len, int id, int first_x, int flag)
{
	hnode *v = nd + id;
	int best = v->score, from = v->from, best_rel = 0, x, y;
	if (v->dflag != FLAG_UNLIMITED) {
		for (x = v->x - 1; x >= first_x; --x)
			for (y = 0; y < len[x]; ++y) {
				const hnode *c = nd + start[x] + y;
				int rel, sc;
				if (c->dflag != flag) continue;
				rel = relation(P, c->read_i, c->offset, v->read_i, v->offset);
				if (rel == HF_UNCONNECT) continue;
				sc = c->score + 1 - edge_pen(rel);
				if (sc > best) { best = sc; from = start[x] + y; best_rel = rel; }
			}
		if (from != v->from) {
			v->score = best; v->from = from; v->mflag = best_rel;
			if (best_rel == HF_MATCH) nd[from].dflag = -flag;          /* a matched predecessor serves once */
			v->node_n += nd[from].node_n;
		}
	} else {
		for (x = v->x - 1; x >= first_x; --x)
			for (y = 0; y < len[x]; ++y) {
				const hnode *c = nd + start[x] + y;
				if (c->dflag == flag && c->score > best) { best = c->score; from = start[x] + y; }
			}
		if (best > 0) { v->from = from; v->read_i = -1; v->offset = -1; v->score = best; v->node_n = nd[from].node_n; v->mflag = HF_MATCH; v->dflag = FLAG_UNLIMITED; }
	}
}

/* chain of multi-hit nodes between two consecutive nodes of the main line; writes the chain forward into `out`,
 * returns its length; src/split_mapping.c:444-488 */
static int mini_line(const hpar *P, hnode *nd, const int *start, const int *len, int left, int right, int *out)
{
	const int lx = nd[left].x, rx = nd[right].x;
	int x, y, k, id;
	for (x = lx + 1; x < rx; ++x) for (y = 0; y < len[x]; ++y) init_node(P, nd, start[x] + y, left, FLAG_MULTI);
	nd[right].from = left; nd[right].score = 0; nd[right].node_n = 0; nd[right].dflag = FLAG_MULTI;
	for (x = lx + 2; x < rx; ++x)
		for (y = 0; y < len[x]; ++y) if (nd[start[x] + y].dflag == FLAG_MULTI) update_node(P, nd, start, len, start[x] + y, lx + 1, FLAG_MULTI);
	update_node(P, nd, start, len, right, lx + 1, FLAG_MULTI);
	k = nd[right].node_n - 1;
	for (id = nd[right].from; nd[id].x != lx; id = nd[id].from) { if (k >= 0) out[k] = id; --k; }
	return nd[right].node_n;
}

/*
 * The line of one (reference window, read) pair.  out_*: per line node its read position, diagonal (window position -
 * read position) and relation to its predecessor on the line (the stitching of :717-783 reads exactly these).
 * Returns the number of line nodes (what hash_main_line returns), or -1 when `cap` is too small.
 */
int orc_hash_line(const uint8_t *ref, int ref_len, const uint8_t *read, int read_len, int ref_offset,
                  int hash_len, int hash_step, int split_len, int head_on, int tail_on,
                  int32_t *out_read_i, int32_t *out_offset, int32_t *out_flag, int cap)
{
	static const int nt4[5] = {0, 1, 2, 3, 2};                    /* src/bntseq.c:78: N hashes as G */
	hpar P = {hash_len, hash_step, split_len, ref_len, read_len, ref_offset};
	const int S = read_len >= hash_len ? (read_len - hash_len) / hash_step + 1 : 0;
	const int R = ref_len >= hash_len ? ref_len - hash_len + 1 : 0;
	uint64_t *rcode = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(R + 1));
	int *start = (int *)calloc((size_t)S + 2, sizeof(int)), *len = (int *)calloc((size_t)S + 2, sizeof(int));
	int *hits = NULL, n_hits = 0, cap_hits = 0;
	hnode *nd;
	int *line, *mini, n_line = 0, N, s, p, i, have_single = 0, head = 0, tail, rc;

	for (p = 0; p < R; ++p) { uint64_t c = 0; for (i = 0; i < hash_len; ++i) c = c << 2 | (uint64_t)nt4[ref[p + i] > 4 ? 4 : ref[p + i]]; rcode[p] = c; }
	/* hits of every read k-mer: window positions in ascending order, none when there are more than 50 */
	len[0] = 1;
	for (s = 1; s <= S; ++s) {
		uint64_t c = 0; int cnt = 0;
		const uint8_t *q = read + (size_t)(s - 1) * hash_step;
		for (i = 0; i < hash_len; ++i) c = c << 2 | (uint64_t)nt4[q[i] > 4 ? 4 : q[i]];
		for (p = 0; p < R; ++p) cnt += rcode[p] == c;
		if (cnt > MAX_HITS) cnt = 0;
		if (n_hits + cnt > cap_hits) { cap_hits = (n_hits + cnt) * 2 + 64; hits = (int *)realloc(hits, sizeof(int) * (size_t)cap_hits); }
		start[s] = 1 + n_hits; len[s] = cnt;
		if (cnt) for (p = 0; p < R; ++p) if (rcode[p] == c) hits[n_hits++] = p;
	}
	N = n_hits; tail = N + 1;
	start[0] = 0; start[S + 1] = tail; len[S + 1] = 1;
	nd = (hnode *)calloc((size_t)N + 2, sizeof(hnode));
	line = (int *)malloc(sizeof(int) * (size_t)(N + 2)); mini = (int *)malloc(sizeof(int) * (size_t)(N + 2));

	/* head and tail: src/split_mapping.c:505-513 */
	nd[head] = (hnode){0, head_on ? -hash_len : -1, head_on ? 0 : -1, -1, 0, 0, HF_MATCH, head_on ? FLAG_MIN : FLAG_UNLIMITED};
	if (tail_on) nd[tail] = (hnode){S + 1, read_len, ref_len - read_len, head, 0, 0, head_on ? HF_UNMATCH : HF_MATCH, FLAG_MIN};
	else nd[tail] = (hnode){S + 1, -1, -1, head, 0, 0, HF_MATCH, FLAG_UNLIMITED};
	for (s = 1; s <= S; ++s)
		for (i = 0; i < len[s]; ++i) {
			hnode *v = nd + start[s] + i;
			v->x = s; v->read_i = (s - 1) * hash_step; v->offset = hits[start[s] - 1 + i] - v->read_i;
			init_node(&P, nd, start[s] + i, head, len[s] == 1 ? FLAG_MIN : FLAG_MULTI);
			if (len[s] == 1) have_single = 1;
		}

	if (have_single) {
		int right, left;
		/* a multi-hit node on the diagonal of ANY single-hit node (head and tail count: their lists have one
		 * entry) joins the single-hit pass; src/split_mapping.c:312-338, :529-532 */
		for (s = 1; s <= S; ++s) {
			if (len[s] <= 1) continue;
			for (i = 0; i < len[s]; ++i) {
				hnode *v = nd + start[s] + i;
				int x;
				if (v->dflag < 0) continue;
				for (x = 0; x <= S + 1; ++x) if (len[x] == 1 && nd[start[x]].offset == v->offset) { v->dflag = FLAG_MIN; break; }
			}
		}
		for (s = 2; s <= S; ++s)
			for (i = 0; i < len[s]; ++i) if (nd[start[s] + i].dflag == FLAG_MIN) update_node(&P, nd, start, len, start[s] + i, 1, FLAG_MIN);
		update_node(&P, nd, start, len, tail, 1, FLAG_MIN);
		/* walk back from the tail; gaps after a non-match edge are refilled with multi-hit nodes (:547-561) */
		right = tail; left = nd[tail].from;
		for (;;) {
			if (nd[right].mflag != HF_MATCH && nd[left].x < nd[right].x - 1) {
				const int m = mini_line(&P, nd, start, len, left, right, mini);
				for (i = m - 1; i >= 0; --i) line[n_line++] = mini[i];
			}
			if (nd[left].x == 0) break;
			line[n_line++] = left;
			right = left; left = nd[right].from;
		}
		for (i = 0; i < n_line / 2; ++i) { const int t = line[i]; line[i] = line[n_line - 1 - i]; line[n_line - 1 - i] = t; }
	} else {
		int k, id;
		for (s = 2; s <= S; ++s)
			for (i = 0; i < len[s]; ++i) if (nd[start[s] + i].dflag == FLAG_MULTI) update_node(&P, nd, start, len, start[s] + i, 1, FLAG_MULTI);
		update_node(&P, nd, start, len, tail, 1, FLAG_MULTI);
		n_line = nd[tail].node_n;
		k = n_line - 1;
		for (id = nd[tail].from; nd[id].x != 0; id = nd[id].from) { if (k >= 0) line[k] = id; --k; }
	}
	rc = n_line;
	if (n_line > cap) rc = -1;
	else for (i = 0; i < n_line; ++i) { out_read_i[i] = nd[line[i]].read_i; out_offset[i] = nd[line[i]].offset; out_flag[i] = nd[line[i]].mflag; }
	free(rcode); free(start); free(len); free(hits); free(nd); free(line); free(mini);
	return rc;
}

/*
 * hash_ref_shim.c -- drives the UNMODIFIED reference local split mapping (/root/reference/src/split_mapping.c) from
 * plain arrays.  Compiled ONLY into oracle/_ref/liblamsa_ref.so (all reference translation units but main.c).
 * It includes the reference's own headers and calls its own functions; nothing of the reference is copied.
 * Test infrastructure; never linked into the product.
 *
 *   ref_hash_line       init_hash + the k-mer look-up + hash_main_line -> the chained line (pins oracle/hash_oracle.c)
 *   ref_hash_split_map  the whole hash_split_map (index, line, DP stitching) -> CIGAR (golden vectors of the drop-in)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lamsa_aln.h"
#include "frag_check.h"
#include "split_mapping.h"

extern int init_hash(uint8_t *ref_seq, int ref_len, int hash_len, uint32_t **hash_num, uint64_t ***hash_node, int ***hash_node_num,
                     int32_t **hash_pos, int key_len, int hash_size);
extern int hash_calcu(int *key_int, int *kmer_int, uint8_t *seed, int hash_len, int key_len);
extern int hash_hit(uint32_t *hash_num, uint64_t **hash_node, int *node_i, int key_int, int kmer_int);
extern int hash_main_line(int *hash_pos, int *start_a, int *len_a, int ref_len, int read_len, int ref_offset, int hash_seed_n,
                          lamsa_aln_para *AP, hash_dp_node **h_node, line_node *line, int _head, int _tail);
extern int hash_split_map(cigar32_t **split_cigar, int *split_clen, int *split_m, uint8_t *ref_seq, int ref_len, int ref_offset,
                          uint8_t *read_seq, int read_len, lamsa_aln_para *AP, uint32_t *hash_num, uint64_t **hash_node,
                          int **hash_node_num, int32_t *hash_pos, int _head, int _tail);

int ref_hash_line(const uint8_t *ref, int ref_len, const uint8_t *read, int read_len, int ref_offset,
                  int hash_len, int hash_step, int split_len, int head_on, int tail_on,
                  int32_t *out_read_i, int32_t *out_offset, int32_t *out_flag, int cap)
{
	lamsa_aln_para AP;
	memset(&AP, 0, sizeof AP);
	AP.hash_len = hash_len; AP.hash_step = hash_step; AP.hash_key_len = HASH_KEY; AP.hash_size = (int)pow(NT_N, HASH_KEY);
	AP.split_len = split_len;
	uint32_t *hash_num = calloc(AP.hash_size, sizeof(uint32_t));
	uint64_t **hash_node = calloc(AP.hash_size, sizeof(uint64_t *));
	int32_t *hash_pos = malloc((ref_len > 0 ? ref_len : 1) * sizeof(int32_t));
	int **hash_node_num;
	init_hash((uint8_t *)ref, ref_len, hash_len, &hash_num, &hash_node, &hash_node_num, &hash_pos, AP.hash_key_len, AP.hash_size);

	/* the look-up of src/split_mapping.c:645-675, through the reference's own hash_calcu / hash_hit */
	const int n = (read_len - hash_len) / hash_step + 1;
	int *start_a = malloc((n + 2) * sizeof(int)), *len_a = malloc((n + 2) * sizeof(int));
	int i, key, kmer, at;
	len_a[0] = 1;
	for (i = 0; i <= read_len - hash_len; i += hash_step) {
		hash_calcu(&key, &kmer, (uint8_t *)read + i, hash_len, AP.hash_key_len);
		if (hash_hit(hash_num, hash_node, &at, key, kmer) == 1) {
			start_a[i / hash_step + 1] = (int)(hash_node[key][at] & 0xffffffff);
			if ((len_a[i / hash_step + 1] = hash_node_num[key][at]) > 50) len_a[i / hash_step + 1] = 0;
		} else len_a[i / hash_step + 1] = 0;
	}
	len_a[i / hash_step + 1] = 1;
	line_node *line = malloc((n > 0 ? n : 1) * sizeof(line_node));
	hash_dp_node **h_node = malloc((n + 2) * sizeof(hash_dp_node *));
	for (i = 0; i < n + 2; ++i) h_node[i] = malloc((len_a[i] > 0 ? len_a[i] : 1) * sizeof(hash_dp_node));
	const int m = hash_main_line(hash_pos, start_a, len_a, ref_len, read_len, ref_offset, n, &AP, h_node, line, head_on, tail_on);
	int rc = m;
	if (m > cap) rc = -1;
	else for (i = 0; i < m; ++i) {
		const hash_dp_node *v = &h_node[line[i].x][line[i].y];
		out_read_i[i] = v->read_i; out_offset[i] = v->offset; out_flag[i] = v->match_flag;
	}
	for (i = 0; i < n + 2; ++i) free(h_node[i]);
	free(h_node); free(line); free(start_a); free(len_a); free(hash_pos);
	for (i = 0; i < AP.hash_size; ++i) { free(hash_node_num[i]); free(hash_node[i]); }
	free(hash_node_num); free(hash_node); free(hash_num);
	return rc;
}

/* AP: a fully set lamsa_aln_para (the caller fills it through its own layout-identical struct).  Returns the number of
 * CIGAR words (copied into out, at most cap; -1 when cap is too small); *res = hash_split_map's return value. */
int ref_hash_split_map(const uint8_t *ref, int ref_len, int ref_offset, const uint8_t *read, int read_len, lamsa_aln_para *AP,
                       int head_on, int tail_on, int32_t *out, int cap, int *res)
{
	uint32_t *hash_num = calloc(AP->hash_size, sizeof(uint32_t));
	uint64_t **hash_node = calloc(AP->hash_size, sizeof(uint64_t *));
	int32_t *hash_pos = malloc((ref_len > 0 ? ref_len : 1) * sizeof(int32_t));
	int **hash_node_num;
	int i;
	init_hash((uint8_t *)ref, ref_len, AP->hash_len, &hash_num, &hash_node, &hash_node_num, &hash_pos, AP->hash_key_len, AP->hash_size);
	int m = CIGAR_LEN_M, n = 0;
	cigar32_t *cigar = malloc(m * sizeof(cigar32_t));
	*res = hash_split_map(&cigar, &n, &m, (uint8_t *)ref, ref_len, ref_offset, (uint8_t *)read, read_len, AP, hash_num, hash_node, hash_node_num, hash_pos, head_on, tail_on);
	int rc = n;
	if (n > cap) rc = -1; else memcpy(out, cigar, n * sizeof(cigar32_t));
	free(cigar); free(hash_pos);
	for (i = 0; i < AP->hash_size; ++i) { free(hash_node_num[i]); free(hash_node[i]); }
	free(hash_node_num); free(hash_node); free(hash_num);
	return rc;
}

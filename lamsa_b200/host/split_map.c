/*
 * split_map.c -- strong `init_hash` / `hash_split_map` (reference src/split_mapping.c:181, :634) for a program that is
 * linked from the reference's UNMODIFIED sources: split_mapping.c is compiled with its own two definitions marked weak
 * (`#pragma weak`, force-included: weak_split_map.h), and these forward to liblamsa_b200 (hash_dropin.cu: k-mer index,
 * look-up and chaining of the hits on the GPU, the reference's stitching around the library's GPU DP calls).
 * OPT-IN build (oracle/Makefile `producer_hash`), like res_aux.c.
 */
#include <stdint.h>
#include <stdio.h>
#include <zlib.h>
#include "lamsa_aln.h"

extern int lb2_init_hash(uint8_t *ref_seq, int ref_len, int hash_len, uint32_t **hash_num, uint64_t ***hash_node,
                         int ***hash_node_num, int32_t **hash_pos, int key_len, int hash_size);
extern int lb2_hash_split_map(cigar32_t **split_cigar, int *split_clen, int *split_m, uint8_t *ref_seq, int ref_len, int ref_offset,
                              uint8_t *read_seq, int read_len, lamsa_aln_para *AP, uint32_t *hash_num, uint64_t **hash_node,
                              int **hash_node_num, int32_t *hash_pos, int _head, int _tail);

int init_hash(uint8_t *ref_seq, int ref_len, int hash_len, uint32_t **hash_num, uint64_t ***hash_node, int ***hash_node_num,
              int32_t **hash_pos, int key_len, int hash_size)
{
	return lb2_init_hash(ref_seq, ref_len, hash_len, hash_num, hash_node, hash_node_num, hash_pos, key_len, hash_size);
}

int hash_split_map(cigar32_t **split_cigar, int *split_clen, int *split_m, uint8_t *ref_seq, int ref_len, int ref_offset,
                   uint8_t *read_seq, int read_len, lamsa_aln_para *AP, uint32_t *hash_num, uint64_t **hash_node,
                   int **hash_node_num, int32_t *hash_pos, int _head, int _tail)
{
	return lb2_hash_split_map(split_cigar, split_clen, split_m, ref_seq, ref_len, ref_offset, read_seq, read_len, AP, hash_num, hash_node,
	                          hash_node_num, hash_pos, _head, _tail);
}

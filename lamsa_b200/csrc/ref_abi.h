// ref_abi.h -- layout-identical restatements of the reference structs that cross the sparse-DP
// drop-in boundary (frag_line_BCC / frag_line_remain and the node_score helpers).  Field order and
// types follow the reference headers cited per struct; sizes and offsets are asserted here (x86-64)
// and checked against the reference's own headers by tests/test_abi.py via oracle/sdp_ref_shim.c.
#pragma once
#include <cstddef>
#include <cstdint>

extern "C" {

typedef struct lb2_ref_line_node { int x, y; } lb2_ref_line_node;                 // src/lamsa_aln.h:300-303

typedef struct lb2_ref_map {                                                      // map_t, src/lamsa_aln.h:230-239
    char strand; int8_t nstrand;
    char chr[1024]; int32_t nchr;
    int64_t offset;
    int NM;
    void* cigar;
    int len_dif, bmax;
} lb2_ref_map;
typedef struct lb2_ref_map_msg {                                                  // map_msg, :240-245
    lb2_ref_map* map;
    int32_t map_n, map_m;
    int32_t seed_id;
    char* map_str;
} lb2_ref_map_msg;

typedef struct lb2_ref_frag_aln_msg {                                             // frag_aln_msg, src/frag_check.h:13-34
    int chr, strand;
    int64_t cigar_ref_start, cigar_ref_end;
    int cigar_read_start, cigar_read_end;
    int32_t* cigar;
    int cigar_len, cigar_max;
    int len_dif;
    int per_n, flag, b_f;
    int seed_max, seed_num;
    int* seed_i;
    int* seed_aln_i;
} lb2_ref_frag_aln_msg;
typedef struct lb2_ref_frag_msg {                                                 // frag_msg, src/frag_check.h:36-43
    int frag_max, frag_num;
    lb2_ref_frag_aln_msg* fa_msg;
    int line_score;
    int frag_left_bound, frag_right_bound;
} lb2_ref_frag_msg;

typedef struct lb2_ref_per_para { int last_len, seed_all, seed_out; } lb2_ref_per_para;   // src/lamsa_aln.h:368-377

typedef struct lb2_ref_reg_b { int8_t is_rev; int chr; int64_t ref_pos; } lb2_ref_reg_b;  // src/lamsa_aln.h:276-280
typedef struct lb2_ref_reg {                                                      // reg_t, :282-286
    lb2_ref_reg_b *ref_beg, *ref_end;
    int beg_n, end_n, beg_m, end_m;
    int beg, end;
} lb2_ref_reg;
typedef struct lb2_ref_aln_reg { lb2_ref_reg* reg; int reg_n, reg_m; int read_len; } lb2_ref_aln_reg;  // :291-295

typedef struct lb2_ref_kstring { size_t l, m; char* s; } lb2_ref_kstring;         // src/kstring.h:39-42
typedef struct lb2_ref_kseq { lb2_ref_kstring name, comment, seq, qual; void* f; } lb2_ref_kseq;  // src/kseq.h:221-225

typedef struct lb2_ref_node_score {                                               // node_score, src/lamsa_aln.h:325-332
    lb2_ref_line_node* node;
    int *score, *NM;
    int min_score_thd;
    int max_n;
    int node_n;
} lb2_ref_node_score;

// by-value line_node helpers (see the note in include/lamsa_b200.h)
int heap_add_node(lb2_ref_node_score* ns, lb2_ref_line_node node, int score, int NM);
lb2_ref_line_node node_pop(lb2_ref_node_score* ns, int* score, int* NM);
lb2_ref_line_node node_heap_extract_max(lb2_ref_node_score* ns, int* score);
lb2_ref_line_node node_heap_extract_minpos(lb2_ref_node_score* ns);
int node_heap_update_min(lb2_ref_node_score* ns, lb2_ref_line_node node, int score, int NM);

}  // extern "C"

static_assert(sizeof(lb2_ref_map) == 1064 && offsetof(lb2_ref_map, nchr) == 1028 && offsetof(lb2_ref_map, offset) == 1032 &&
              offsetof(lb2_ref_map, NM) == 1040 && offsetof(lb2_ref_map, len_dif) == 1056, "map_t layout");
static_assert(sizeof(lb2_ref_map_msg) == 32 && offsetof(lb2_ref_map_msg, seed_id) == 16, "map_msg layout");
static_assert(sizeof(lb2_ref_frag_aln_msg) == 88 && offsetof(lb2_ref_frag_aln_msg, cigar) == 32 &&
              offsetof(lb2_ref_frag_aln_msg, flag) == 56 && offsetof(lb2_ref_frag_aln_msg, seed_i) == 72, "frag_aln_msg layout");
static_assert(sizeof(lb2_ref_frag_msg) == 32 && offsetof(lb2_ref_frag_msg, line_score) == 16, "frag_msg layout");
static_assert(sizeof(lb2_ref_per_para) == 12, "lamsa_aln_per_para layout");
static_assert(sizeof(lb2_ref_reg_b) == 16 && sizeof(lb2_ref_reg) == 40 && sizeof(lb2_ref_aln_reg) == 24, "aln_reg layout");
static_assert(sizeof(lb2_ref_kseq) == 104 && offsetof(lb2_ref_kseq, seq) == 48, "kseq_t layout");
static_assert(sizeof(lb2_ref_node_score) == 40 && offsetof(lb2_ref_node_score, node_n) == 32, "node_score layout");

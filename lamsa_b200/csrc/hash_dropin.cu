// hash_dropin.cu -- drop-in `init_hash` and `hash_split_map` (reference src/split_mapping.c:181-208, :634-825).
// The seed-and-chain half (k-mer index of the window, look-up of the read's k-mers, hash_main_line) runs on the
// GPU (hash_line.cuh through lb2_hash_line_run; a caller inside a worker fiber of the batch producer parks and all
// parked requests are one launch).  What stays here is the reference's host control flow around its DP calls --
// the stitching of the line (:688-821) and make_indel_cigar (:606-632) -- restated in the reference's order, with
// this library's ksw_global2 / ksw_bi_extend / ksw_extend_core (GPU) underneath.  No CPU path: without a GPU the
// calls below exit like every other entry point of the library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/lamsa_b200.h"
#include "dropin_internal.h"
#include "cigar_list.h"

using namespace lb2::cigar_list;

namespace {
std::mutex g_hash_mu;

// src/split_mapping.c:606-632: the indel between two adjacent line nodes; returns their overlap
int indel_word(int ref_left, int read_left, int ref_right, int read_right, int* clen, cigar32_t* word, int split_len, int* split_flag) {
    const int dlen = ref_left - ref_right + 1, ilen = read_left - read_right + 1;
    if (dlen < 0 && ilen < 0) { fprintf(stderr, "[make_indel_cigar] Error: dlen: %d, ilen: %d.\n", dlen, ilen); exit(1); }
    const int len = ilen - dlen;
    *clen = len != 0;
    if (len > 0) { *word = (len << 4) + LB2_CDEL; if (len >= split_len) *split_flag |= 2; }
    else if (len < 0) { *word = ((-len) << 4) + LB2_CINS; if (-len >= split_len) *split_flag |= 2; }
    return dlen > ilen ? dlen : ilen;
}

// global fill below split_len, two-sided extension above (:702-707, :736-742, :792-797, :810-816)
int fill_blank(int qlen, const uint8_t* q, int tlen, const uint8_t* t, lamsa_aln_para* AP, cigar32_t** out, int* n, int* m) {
    cigar32_t* c = nullptr; int cn = 0, cm = 0, res = 0;
    if (tlen < AP->split_len && qlen < AP->split_len)
        ksw_global2(qlen, q, tlen, t, 5, AP->sc_mat, AP->del_gapo, AP->del_gape, AP->ins_gapo, AP->ins_gape, AP->band_w, &cn, &c);
    else
        res = ksw_bi_extend(qlen, q, tlen, t, 5, AP->sc_mat, AP->hash_len * AP->match, AP->hash_len * AP->match, AP, &c, &cn, &cm);
    list_append(out, n, m, c, cn);
    free(c);
    return res;
}
}  // namespace

static int init_hash_impl(uint8_t* ref_seq, int ref_len, int hash_len, uint32_t** hash_num, uint64_t*** hash_node,
                          int*** hash_node_num, int32_t** hash_pos, int key_len, int hash_size)
{
    // the index lives on the GPU, built per request by hash_split_map below; leave what the callers free
    // (src/split_mapping.c:843-846: hash_pos, hash_node_num[i], hash_node_num) in a freeable state
    (void)ref_seq; (void)ref_len; (void)hash_len; (void)hash_node; (void)hash_pos; (void)key_len;
    for (int i = 0; i < hash_size; ++i) (*hash_num)[i] = 0;
    *hash_node_num = (int**)malloc(sizeof(int*) * (size_t)(hash_size > 0 ? hash_size : 1));
    for (int i = 0; i < hash_size; ++i) (*hash_node_num)[i] = (int*)calloc(1, sizeof(int));
    return 0;
}

static int hash_split_map_impl(cigar32_t** split_cigar, int* split_clen, int* split_m,
                               uint8_t* ref_seq, int ref_len, int ref_offset, uint8_t* read_seq, int read_len,
                               lamsa_aln_para* AP, uint32_t* hash_num, uint64_t** hash_node, int** hash_node_num,
                               int32_t* hash_pos, int _head, int _tail)
{
    (void)hash_num; (void)hash_node; (void)hash_node_num; (void)hash_pos;
    const int hash_len = AP->hash_len, hash_step = AP->hash_step;
    const int split_len = AP->split_pen;            // sic: src/split_mapping.c:640
    int res = 0;
    *split_clen = 0;

    // ---- the line, on the GPU
    const int cap = read_len >= hash_len ? (read_len - hash_len) / hash_step + 1 : 1;
    std::vector<int32_t> line((size_t)cap * 3);
    lb2_hash_task ht; memset(&ht, 0, sizeof ht);
    ht.ref = ref_seq; ht.ref_len = ref_len; ht.read = read_seq; ht.read_len = read_len; ht.ref_offset = ref_offset;
    ht.hash_len = hash_len; ht.hash_step = hash_step; ht.split_len = AP->split_len; ht.head = _head; ht.tail = _tail;
    ht.line = line.data(); ht.line_cap = cap;
    int rc;
    if (lb2::fiber_active()) rc = lb2::worker_hash_line(&ht);
    else { std::lock_guard<std::mutex> lk(g_hash_mu); rc = lb2_hash_line_run(lb2::dropin_ctx(), 1, &ht); }
    if (rc) { fprintf(stderr, "[lamsa_b200] hash_split_map: %s\n", lb2_last_error()); exit(1); }
    const int m_len = ht.m_len;
    auto RI = [&](int k) { return line[(size_t)k * 3]; };
    auto OFF = [&](int k) { return line[(size_t)k * 3 + 1]; };
    auto FLAG = [&](int k) { return line[(size_t)k * 3 + 2]; };

    // ---- stitching (src/split_mapping.c:688-821)
    const int tail_in = hash_len / 2, head_in = (hash_len + 1) / 2;
    if (m_len > 0) {
        cigar32_t g = 0; int gn = 0;
        {   // 1. left bound .. first node
            const int refi = RI(0) + OFF(0), readi = RI(0);
            if (_head) {
                if (readi != 0 && refi != 0) res |= fill_blank(readi + tail_in, read_seq, refi + tail_in, ref_seq, AP, split_cigar, split_clen, split_m);
                else {
                    indel_word(-1, -1, refi, readi, &gn, &g, split_len, &res);
                    list_append(split_cigar, split_clen, split_m, &g, gn);
                    list_add_nonempty(split_cigar, split_clen, split_m, (tail_in << 4) | LB2_CMATCH);
                }
            }
        }
        // 2. between the nodes: runs of match-like nodes are one M; the seams are filled, overlapped or indels
        int start_i = 0, overlap = 0;
        for (int i = 0; i < m_len; ++i) {
            if (!(i == m_len - 1 || FLAG(i + 1) >= 2 /* F_MATCH_THD */)) continue;
            list_add_nonempty(split_cigar, split_clen, split_m, ((RI(i) - RI(start_i) + hash_len - tail_in - head_in - overlap) << 4) | LB2_CMATCH);
            if (i == m_len - 1) break;
            const int l_readi = RI(i) + hash_len - 1, r_readi = RI(i + 1);
            const int l_refi = RI(i) + hash_len + OFF(i) - 1, r_refi = RI(i + 1) + OFF(i + 1);
            if (l_readi + 1 < r_readi && l_refi + 1 < r_refi) {             // a blank on both sequences
                const int ql = r_readi - (l_readi + 1) + head_in + tail_in, tl = ql + OFF(i + 1) - OFF(i);
                res |= fill_blank(ql, read_seq + l_readi + 1 - head_in, tl, ref_seq + l_refi + 1 - head_in, AP, split_cigar, split_clen, split_m);
                overlap = 0;
            } else if (l_refi >= r_refi) {                                  // the nodes overlap on the reference
                const int extra = ref_offset > 0 ? hash_len : 0;
                int ql = r_readi - (l_readi + 1) + head_in, tl = ql + extra;
                int lqe = 0, lte = 0, rqe = 0, rte = 0, cn = 0, cm = 0; cigar32_t* c = nullptr;
                ksw_extend_core(ql, read_seq + l_readi + 1 - head_in, tl, ref_seq + l_refi + 1 - head_in, 5, AP->sc_mat, AP->band_w,
                                hash_len * AP->match, AP, &lqe, &lte, &c, &cn, &cm);
                list_append(split_cigar, split_clen, split_m, c, cn);
                free(c); c = nullptr;
                const int ql_left = ql;
                ql = r_readi - (l_readi + 1) + tail_in; tl = ql + extra;
                if (r_readi + tail_in - ql < 0 || r_refi + tail_in - tl < -ref_offset - extra) { fprintf(stderr, "[hash_split_map] BUG.\n"); exit(1); }
                std::vector<uint8_t> rq((size_t)(ql > 0 ? ql : 1)), rt((size_t)(tl > 0 ? tl : 1));
                for (int j = 0; j < ql; ++j) rq[(size_t)j] = read_seq[r_readi + tail_in - 1 - j];
                for (int j = 0; j < tl; ++j) rt[(size_t)j] = ref_seq[r_refi + tail_in - 1 - j];
                cn = 0; cm = 0;
                ksw_extend_core(ql, rq.data(), tl, rt.data(), 5, AP->sc_mat, AP->band_w, hash_len * AP->match, AP, &rqe, &rte, &c, &cn, &cm);
                list_reverse(c, cn);
                (void)ql_left;
                const int Sn = ql + head_in - lqe - rqe, Hn = r_refi + head_in + tail_in - l_refi - 1 - lte - rte;
                list_add(split_cigar, split_clen, split_m, (Sn << 4) | LB2_CSOFT_CLIP);
                list_add(split_cigar, split_clen, split_m, (Hn << 4) | LB2_CHARD_CLIP);
                list_append(split_cigar, split_clen, split_m, c, cn);
                overlap = 0;
                free(c);
            } else {                                                        // adjacent: a plain indel
                list_add_nonempty(split_cigar, split_clen, split_m, (head_in << 4) | LB2_CMATCH);
                overlap = indel_word(l_refi, l_readi, r_refi, r_readi, &gn, &g, split_len, &res);
                list_append(split_cigar, split_clen, split_m, &g, gn);
                list_add_nonempty(split_cigar, split_clen, split_m, (tail_in << 4) | LB2_CMATCH);
            }
            start_i = i + 1;
        }
        {   // 3. last node .. right bound
            const int readi = RI(m_len - 1) + hash_len - 1, refi = RI(m_len - 1) + OFF(m_len - 1) + hash_len - 1;
            if (_tail) {
                if (readi + 1 < read_len && refi + 1 < ref_len)
                    res |= fill_blank(read_len - (readi + 1) + head_in, read_seq + readi + 1 - head_in, ref_len - (refi + 1) + head_in,
                                      ref_seq + refi + 1 - head_in, AP, split_cigar, split_clen, split_m);
                else {
                    list_add_nonempty(split_cigar, split_clen, split_m, (head_in << 4) | LB2_CMATCH);
                    indel_word(refi, readi, ref_len, read_len, &gn, &g, split_len, &res);
                    list_append(split_cigar, split_clen, split_m, &g, gn);
                }
            }
        }
    } else if (_head && _tail) {
        res |= fill_blank(read_len, read_seq, ref_len, ref_seq, AP, split_cigar, split_clen, split_m);
    }
    return res;
}

// Exported twice: under the reference's own names (a maintainer deletes the two functions from split_mapping.c and
// links this library), and under lb2_ names for a link that keeps split_mapping.c unmodified with its definitions
// marked weak -- there the strong definitions must sit in an object of the program itself (lamsa_b200/host/split_map.c
// forwards to these), because a definition in a shared library never overrides one in the program's own objects.
extern "C" {
int init_hash(uint8_t* ref_seq, int ref_len, int hash_len, uint32_t** hash_num, uint64_t*** hash_node, int*** hash_node_num,
              int32_t** hash_pos, int key_len, int hash_size) {
    return init_hash_impl(ref_seq, ref_len, hash_len, hash_num, hash_node, hash_node_num, hash_pos, key_len, hash_size);
}
int lb2_init_hash(uint8_t* ref_seq, int ref_len, int hash_len, uint32_t** hash_num, uint64_t*** hash_node, int*** hash_node_num,
                  int32_t** hash_pos, int key_len, int hash_size) {
    return init_hash_impl(ref_seq, ref_len, hash_len, hash_num, hash_node, hash_node_num, hash_pos, key_len, hash_size);
}
int hash_split_map(cigar32_t** split_cigar, int* split_clen, int* split_m, uint8_t* ref_seq, int ref_len, int ref_offset,
                   uint8_t* read_seq, int read_len, lamsa_aln_para* AP, uint32_t* hash_num, uint64_t** hash_node,
                   int** hash_node_num, int32_t* hash_pos, int _head, int _tail) {
    return hash_split_map_impl(split_cigar, split_clen, split_m, ref_seq, ref_len, ref_offset, read_seq, read_len, AP, hash_num, hash_node, hash_node_num, hash_pos, _head, _tail);
}
int lb2_hash_split_map(cigar32_t** split_cigar, int* split_clen, int* split_m, uint8_t* ref_seq, int ref_len, int ref_offset,
                       uint8_t* read_seq, int read_len, lamsa_aln_para* AP, uint32_t* hash_num, uint64_t** hash_node,
                       int** hash_node_num, int32_t* hash_pos, int _head, int _tail) {
    return hash_split_map_impl(split_cigar, split_clen, split_m, ref_seq, ref_len, ref_offset, read_seq, read_len, AP, hash_num, hash_node, hash_node_num, hash_pos, _head, _tail);
}
}

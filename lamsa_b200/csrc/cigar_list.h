// cigar_list.h -- the reference's CIGAR list helpers (src/frag_check.h:139-188: _push_cigar0 / _push_cigar1 /
// _push_cigar / _invert_cigar), shared by the drop-in translation units (ksw_dropin.cu, hash_dropin.cu).
#pragma once
#include <cstdio>
#include <cstdlib>
#include "../../include/lamsa_b200.h"

namespace lb2 { namespace cigar_list {

inline void list_add(cigar32_t** c, int* n, int* cap, cigar32_t op) {            // _push_cigar0
    if (*n > 0 && (((*c)[*n - 1] ^ op) & 0xf) == 0) { (*c)[*n - 1] += (op >> 4) << 4; return; }
    if (*n == *cap) {
        *cap = *cap ? *cap << 1 : 4;
        *c = (cigar32_t*)realloc(*c, sizeof(cigar32_t) * (size_t)*cap);
        if (!*c) { fprintf(stderr, "\n[lamsa_b200] out of memory.\n"); exit(1); }
    }
    (*c)[(*n)++] = op;
}
inline void list_add_nonempty(cigar32_t** c, int* n, int* cap, cigar32_t op) {   // _push_cigar1
    if (op >> 4) list_add(c, n, cap, op);
}
inline void list_append(cigar32_t** c, int* n, int* cap, const cigar32_t* src, int cnt) {   // _push_cigar
    if (cnt == 0) return;
    int i = *n, j = 0;
    if (i > 0) {
        const int a = (*c)[i - 1] & 0xf, b = src[0] & 0xf;
        if (a == b) { (*c)[i - 1] += (src[0] >> 4) << 4; j = 1; }
        else if ((a == LB2_CINS && b == LB2_CSOFT_CLIP) || (a == LB2_CSOFT_CLIP && b == LB2_CINS)) {
            (*c)[i - 1] = ((((*c)[i - 1] >> 4) + (src[0] >> 4)) << 4) | LB2_CSOFT_CLIP; j = 1;
        }
    }
    for (; j < cnt; ++i, ++j) {
        if (i == *cap) {
            *cap = *cap ? *cap << 1 : 4;
            *c = (cigar32_t*)realloc(*c, sizeof(cigar32_t) * (size_t)*cap);
            if (!*c) { fprintf(stderr, "\n[lamsa_b200] out of memory.\n"); exit(1); }
        }
        (*c)[i] = src[j];
    }
    *n = i;
}
inline void list_reverse(cigar32_t* c, int n) {                                  // _invert_cigar
    for (int a = 0, b = n - 1; a < b; ++a, --b) { cigar32_t t = c[a]; c[a] = c[b]; c[b] = t; }
}


}}  // namespace lb2::cigar_list

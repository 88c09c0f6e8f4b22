/*
 * chunk_wrapper.c -- compiled INSTEAD of the reference's src/lamsa_aln.c in the
 * "wide" drop-in build (make -C oracle dropin_wide).  It is the reference's own
 * translation unit with one macro changed at compile time: the read chunk
 * (src/lamsa_aln.h:9, 128 reads) becomes LAMSA_CHUNK reads, so that `lamsa aln
 * -t N` keeps up to N reads -- hence N DP tasks per combined GPU launch -- in
 * flight (the chunk size does not influence results: per-read state only,
 * output in input order, src/lamsa_aln.c:1102-1110).  No reference code is
 * copied: the source is included from where it lies.
 */
#include <stdio.h>
#include <stdint.h>
#include <zlib.h>
#include "kseq.h"
#include "lamsa_aln.h"          /* sets the include guard and CHUNK_READ_N 128 */
#undef CHUNK_READ_N
#ifndef LAMSA_CHUNK
#define LAMSA_CHUNK 8192
#endif
#define CHUNK_READ_N LAMSA_CHUNK
#include "lamsa_aln.c"

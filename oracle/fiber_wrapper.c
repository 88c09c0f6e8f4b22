/*
 * fiber_wrapper.c -- compiled INSTEAD of the reference's src/lamsa_aln.c in the batch-producer
 * build (make -C oracle dropin_fiber).  It is the reference's own translation unit, included from
 * where it lies, with three names changed at compile time:
 *   pthread_create / pthread_join  -> lb2_worker_spawn / lb2_worker_join (include/lamsa_b200.h
 *                                     section 4): the `-t N` workers of src/lamsa_aln.c:1151-1162
 *                                     become fibers whose DP / chaining calls are batched;
 *   CHUNK_READ_N (src/lamsa_aln.h:9, 128 reads) -> LAMSA_CHUNK, so that N reads really are in
 *                                     flight (chunk size does not influence results: per-read
 *                                     state only, output in input order, :1102-1110).
 * No reference code is copied.
 */
#include <pthread.h>            /* before the renames, so the system declarations stay as they are */
#include <stdio.h>
#include <stdint.h>
#include <zlib.h>
#include "kseq.h"
#include "lamsa_aln.h"          /* sets the include guard and CHUNK_READ_N 128 */
#undef CHUNK_READ_N
#ifndef LAMSA_CHUNK
#define LAMSA_CHUNK 16384
#endif
#define CHUNK_READ_N LAMSA_CHUNK
extern int lb2_worker_spawn(pthread_t *id, const pthread_attr_t *attr, void *(*fn)(void *), void *arg);
extern int lb2_worker_join(pthread_t id, void **ret);
#define pthread_create lb2_worker_spawn
#define pthread_join lb2_worker_join
#include "lamsa_aln.c"

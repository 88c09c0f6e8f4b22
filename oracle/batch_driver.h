/*
 * batch_driver.h -- pthread driver that pushes an array of lb2_task records
 * (include/lamsa_b200.h) through a scalar CPU implementation and fills
 * lb2_result records + one concatenated CIGAR pool.  TEST INFRASTRUCTURE ONLY
 * (see dp_oracle.c).  Included twice: by dp_oracle.c (the restatement) and by
 * ref_shim.c (the unmodified reference ksw.c), which define
 *   BD_NAME             exported function name
 *   BD_GLOBAL(t,nc,c)   run a global task, return score
 *   BD_EXTEND(t,r,c)    run an extension task with CIGAR; fills r->score/qle/tle/n_cigar/reserved
 *   BD_EXTEND2(t,r)     score-only extension
 *   BD_CELLS_RESET / BD_CELLS_GET   optional per-thread cell counter (or 0)
 * Work distribution mirrors the reference's own worker loop: a shared counter
 * handed out under a lock (src/lamsa_aln.c:838-842), here an atomic.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
	int64_t n;
	const lb2_task *tasks;
	lb2_result *res;
	int32_t **cig;            /* per-task malloc'd CIGAR */
	int64_t *next;
} bd_shared;

static void *bd_worker(void *arg)
{
	bd_shared *S = (bd_shared *)arg;
	for (;;) {
		int64_t a = __atomic_fetch_add(S->next, 64, __ATOMIC_RELAXED);
		if (a >= S->n) break;
		int64_t e = a + 64 < S->n ? a + 64 : S->n;
		for (int64_t i = a; i < e; ++i) {
			const lb2_task *t = &S->tasks[i];
			lb2_result *r = &S->res[i];
			memset(r, 0, sizeof *r);
			S->cig[i] = 0;
			BD_CELLS_RESET;
			if (t->kind == LB2_KIND_GLOBAL) {
				if (t->flags & LB2_FLAG_CIGAR) { int nc = 0; r->score = BD_GLOBAL(t, &nc, &S->cig[i]); r->n_cigar = nc; }
				else r->score = BD_GLOBAL(t, 0, 0);
				r->qle = t->qlen; r->tle = t->tlen;
			} else if (t->flags & LB2_FLAG_CIGAR) {
				BD_EXTEND(t, r, &S->cig[i]);
			} else {
				BD_EXTEND2(t, r);
			}
			r->cells = (int64_t)(BD_CELLS_GET);
		}
	}
	return 0;
}

/* returns 0; *seconds = wall time of the parallel section */
int BD_NAME(int64_t n, const lb2_task *tasks, lb2_result *res, int32_t **pool, int64_t *pool_n,
            int nthreads, double *seconds)
{
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	bd_shared S;
	int64_t next = 0;
	S.n = n; S.tasks = tasks; S.res = res; S.next = &next;
	S.cig = (int32_t **)calloc(n > 0 ? n : 1, sizeof(int32_t *));
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	pthread_t th[256];
	for (int k = 1; k < nthreads; ++k) pthread_create(&th[k], 0, bd_worker, &S);
	bd_worker(&S);
	for (int k = 1; k < nthreads; ++k) pthread_join(th[k], 0);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	int64_t tot = 0;
	for (int64_t i = 0; i < n; ++i) { res[i].cigar_off = tot; tot += res[i].n_cigar; }
	if (pool) {
		int32_t *p = (int32_t *)malloc((tot > 0 ? tot : 1) * sizeof(int32_t));
		for (int64_t i = 0; i < n; ++i)
			if (res[i].n_cigar) memcpy(p + res[i].cigar_off, S.cig[i], (size_t)res[i].n_cigar * 4);
		*pool = p;
	}
	if (pool_n) *pool_n = tot;
	for (int64_t i = 0; i < n; ++i) free(S.cig[i]);
	free(S.cig);
	return 0;
}

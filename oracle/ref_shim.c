/*
 * ref_shim.c -- compiled ONLY into oracle/_ref/libksw_ref.so, next to the
 * unmodified /root/reference/src/ksw.c (see oracle/Makefile).  It includes the
 * reference's own headers so that tests can (1) learn the true layout of
 * lamsa_aln_para and check include/lamsa_b200.h against it, and (2) fill a
 * genuine lamsa_aln_para to drive the reference's ksw_extend_core/ksw_bi_extend.
 * Test infrastructure; never linked into the product.
 */
#include <stddef.h>
#include <stdio.h>
#include <zlib.h>
#include <string.h>
#include "lamsa_aln.h"
#include "ksw.h"
#include "frag_check.h"

#define OFF(f) (int)offsetof(lamsa_aln_para, f)

int ref_para_sizeof(void) { return (int)sizeof(lamsa_aln_para); }

/* order must match lamsa_b200/_abi.py:PARA_FIELDS */
int ref_para_offsets(int *o)
{
	int n = 0;
	o[n++] = OFF(n_thread);   o[n++] = OFF(seed_len);  o[n++] = OFF(seed_step); o[n++] = OFF(seed_inv);
	o[n++] = OFF(per_aln_m);  o[n++] = OFF(first_loci_thd); o[n++] = OFF(SV_len_thd); o[n++] = OFF(ske_max);
	o[n++] = OFF(ovlp_rat);   o[n++] = OFF(split_len); o[n++] = OFF(split_pen); o[n++] = OFF(res_mul_max);
	o[n++] = OFF(match_dis);  o[n++] = OFF(mismatch_thd); o[n++] = OFF(frag_score_table);
	o[n++] = OFF(ins_gapo);   o[n++] = OFF(ins_gape);  o[n++] = OFF(del_gapo);  o[n++] = OFF(del_gape);
	o[n++] = OFF(ins_ext_o);  o[n++] = OFF(ins_ext_e); o[n++] = OFF(del_ext_o); o[n++] = OFF(del_ext_e);
	o[n++] = OFF(match);      o[n++] = OFF(mis);       o[n++] = OFF(sc_mat);
	o[n++] = OFF(band_w);     o[n++] = OFF(end_bonus); o[n++] = OFF(zdrop);
	o[n++] = OFF(ed_rate);    o[n++] = OFF(mis_rate);  o[n++] = OFF(id_rate);   o[n++] = OFF(mat_rate);
	o[n++] = OFF(read_type);  o[n++] = OFF(aln_mode);
	return n;
}

int ref_struct_sizes(int *o)
{
	int n = 0;
	o[n++] = (int)sizeof(kswr_t);    o[n++] = (int)sizeof(line_node); o[n++] = (int)sizeof(frag_dp_node);
	o[n++] = (int)sizeof(map_t);     o[n++] = (int)sizeof(map_msg);   o[n++] = (int)sizeof(frag_msg);
	o[n++] = (int)sizeof(frag_aln_msg); o[n++] = (int)sizeof(node_score); o[n++] = (int)sizeof(aln_reg);
	o[n++] = (int)sizeof(reg_t);     o[n++] = (int)sizeof(lamsa_aln_per_para); o[n++] = (int)sizeof(cigar_t);
	return n;
}

/* ---- batch driver over the UNMODIFIED reference entry points ------------- */
#define LAMSA_B200_NO_PARA_TYPE
#include "../include/lamsa_b200.h"

static int bd_ext(const lb2_task *t, lb2_result *r, int32_t **c)
{
	lamsa_aln_para AP;
	memset(&AP, 0, sizeof AP);
	AP.del_ext_o = t->o_del; AP.del_ext_e = t->e_del; AP.ins_ext_o = t->o_ins; AP.ins_ext_e = t->e_ins;
	AP.end_bonus = t->end_bonus; AP.zdrop = t->zdrop;
	int qle = 0, tle = 0, nc = 0, mc = 0;
	*c = 0;
	r->score = ksw_extend_core(t->qlen, t->query, t->tlen, t->target, t->m, t->mat, t->w, t->h0, &AP,
	                           &qle, &tle, c, &nc, &mc);
	r->qle = qle; r->tle = tle; r->n_cigar = nc; r->reserved = mc;
	return 0;
}
static int bd_ext2(const lb2_task *t, lb2_result *r)
{
	int qle, tle, gtle, gs, mo;
	r->score = ksw_extend2(t->qlen, t->query, t->tlen, t->target, t->m, t->mat, t->o_del, t->e_del,
	                       t->o_ins, t->e_ins, t->w, t->end_bonus, t->zdrop, t->h0, &qle, &tle, &gtle, &gs, &mo);
	r->qle = qle; r->tle = tle; r->gtle = gtle; r->gscore = gs; r->max_off = mo;
	return 0;
}
#define BD_NAME ref_run_batch
#define BD_GLOBAL(t, nc, c) ksw_global2((t)->qlen, (t)->query, (t)->tlen, (t)->target, (t)->m, (t)->mat, \
                                        (t)->o_del, (t)->e_del, (t)->o_ins, (t)->e_ins, (t)->w, nc, c)
#define BD_EXTEND(t, r, c) bd_ext(t, r, c)
#define BD_EXTEND2(t, r) bd_ext2(t, r)
#define BD_CELLS_RESET ((void)0)
#define BD_CELLS_GET 0
#include "batch_driver.h"

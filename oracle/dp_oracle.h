/* dp_oracle.h -- interface of the CPU restatement (test infrastructure only;
 * see the header of dp_oracle.c).  Plain C ABI so tests can bind it with ctypes. */
#ifndef LAMSA_B200_DP_ORACLE_H
#define LAMSA_B200_DP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* extension penalties the reference reads from lamsa_aln_para (src/ksw.c:680) */
typedef struct {
	int o_del, e_del, o_ins, e_ins;   /* del_ext_o, del_ext_e, ins_ext_o, ins_ext_e */
	int end_bonus, zdrop;
} orc_ext_par;

/* everything ksw_bi_extend / sw_mid_fix read from lamsa_aln_para (src/ksw.c:841-926) */
typedef struct {
	orc_ext_par ext;
	int del_gapo, del_gape, ins_gapo, ins_gape;
	int band_w, split_len, aln_mode;
	float id_rate;
} orc_bi_par;

int orc_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int *n_cigar, int32_t **cigar);
int orc_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off);
int orc_extend_core(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                    int m, const int8_t *mat, int w, int h0, const orc_ext_par *P,
                    int *qle, int *tle, int32_t **cigar, int *n_cigar, int *m_cigar);
int orc_extend_c(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                 int m, const int8_t *mat, int w, int h0, const orc_ext_par *P,
                 int *qle, int *tle, int32_t **cigar, int *n_cigar, int *m_cigar);
int orc_extend_r(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                 int m, const int8_t *mat, int w, int h0, const orc_ext_par *P,
                 int *qre, int *tre, int32_t **cigar, int *n_cigar, int *m_cigar);
void orc_mid_fix(int32_t **cigar, int *cigar_n, int *cigar_m,
                 int32_t *lc, int ln, int32_t *rc, int rn,
                 const uint8_t *query, int qlen, int lqe, int rqe,
                 const uint8_t *target, int tlen, int lte, int rte,
                 const orc_bi_par *P, int m, const int8_t *mat);
int orc_bi_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m,
                  const int8_t *mat, int lh0, int rh0, const orc_bi_par *P,
                  int32_t **cigar, int *n_cigar, int *m_cigar);
uint64_t orc_cells_get(void);
void orc_cells_reset(void);
void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif

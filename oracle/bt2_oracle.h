/* oracle/bt2_oracle.h -- TEST INFRASTRUCTURE ONLY (see bt2_oracle.c header). */
#ifndef BT2_ORACLE_H_
#define BT2_ORACLE_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int       off_size;      /* 4 (.bt2) or 8 (.bt2l) */
	uint64_t  len, bwt_len;
	int32_t   line_rate, off_rate, ftab_chars, flags;
	uint64_t  off_mask_rate; /* = off_rate */
	uint32_t  side_sz, side_bwt_sz, side_bwt_len;
	uint64_t  num_sides, ebwt_tot_len, offs_len, ftab_len, eftab_len;
	uint64_t  n_pat, n_frag;
	uint64_t *plen;          /* [n_pat] */
	uint64_t *rstarts;       /* [3*n_frag] (NULL for mirror) */
	uint8_t  *ebwt;          /* [ebwt_tot_len] raw sides */
	uint64_t  z_off;
	uint64_t  fchr[5];
	uint64_t *ftab;          /* [ftab_len], widened */
	uint64_t *eftab;         /* [eftab_len], widened */
	void     *offs;          /* raw u32/u64 SA sample, NULL for mirror */
	int       is_fw;         /* forward index? (mirror: 0) */
} bt2o_ebwt;

typedef struct {
	uint64_t  n_recs, n_refs;
	uint64_t *rec_off, *rec_len; uint8_t *rec_first;
	uint64_t *ref_rec_offs;  /* [n_refs+1] */
	uint64_t *ref_offs;      /* [n_refs+1] unambiguous chars preceding */
	uint64_t *ref_lens;      /* [n_refs] */
	uint8_t  *buf;           /* 2-bit packed */
	uint64_t  buf_sz;        /* # bases */
} bt2o_ref;

typedef struct {
	bt2o_ebwt fw, bw;
	bt2o_ref  ref;
	int has_bw, has_ref;
} bt2o_index;

/* loading */
bt2o_index *bt2o_open(const char *base, int load_mirror, int load_ref);
void        bt2o_close(bt2o_index *ix);
uint64_t    bt2o_scalar(const bt2o_index *ix, int mirror, int which);

/* FM primitives; "mirror" selects the .rev index */
void     bt2o_rank4(const bt2o_index *ix, int mirror, uint64_t row, uint64_t out4[4]);
uint64_t bt2o_rank1(const bt2o_index *ix, int mirror, uint64_t row, int c);
int      bt2o_rowL(const bt2o_index *ix, int mirror, uint64_t row);
uint64_t bt2o_maplf1(const bt2o_index *ix, int mirror, uint64_t row, int c);
void     bt2o_maplf_range(const bt2o_index *ix, int mirror, uint64_t top, uint64_t num, uint64_t upto[4], uint64_t in[4], uint8_t *chars);
void     bt2o_ftab_lohi(const bt2o_index *ix, int mirror, uint64_t i, uint64_t *top, uint64_t *bot);
uint64_t bt2o_get_offset(const bt2o_index *ix, uint64_t row);
int      bt2o_joined_to_text(const bt2o_index *ix, uint64_t qlen, uint64_t off, int reject_straddle,
                             uint64_t *tidx, uint64_t *textoff, uint64_t *tlen, int *straddled);
int      bt2o_get_stretch(const bt2o_index *ix, uint64_t tidx, int64_t off, int64_t count, uint8_t *out);

/* seed search */
uint64_t bt2o_exact_sweep(const bt2o_index *ix, const uint8_t *codes, int len, int nofw, int norc,
                          uint64_t mine2[2], uint64_t topbot4[4]);
int      bt2o_seed_search(const bt2o_index *ix, const uint8_t *codes, const uint8_t *quals, int len,
                          int seedlen, int interval, int offset, int nofw, int norc,
                          int max_seeds, uint64_t *out_ranges);

void     bt2o_extend(const bt2o_index *ix, const uint8_t *codes, int len, int fw, uint64_t off, uint64_t seedlen,
                     uint64_t topf, uint64_t botf, uint64_t topb, uint64_t botb, uint64_t nlex_nrex[2]);

/* seed-extension DP (end-to-end) */
typedef struct {
	int match_bonus, mmp_max, mmp_min, n_pen;
	int rdgap_const, rdgap_linear, rfgap_const, rfgap_linear, gapbar, local;
	double nceil_const, nceil_linear;            /* --n-ceil (Scoring::nCeil, scoring.h:61-63: 0, 0.15f): oneMmSearch and ungappedAlign evaluate it themselves */
} bt2o_scoring;
void bt2o_scoring_default(bt2o_scoring *sc, int local);
int  bt2o_one_mm(const bt2o_index *ix, const bt2o_scoring *sc, const uint8_t *codes, const uint8_t *quals, int len,
                 int64_t minsc, int nofw, int norc, int max_hits, int64_t *out, int *out_fw);
int  bt2o_ungapped(const bt2o_index *ix, const bt2o_scoring *sc, const uint8_t *codes, const uint8_t *quals, int len, int fw,
                   uint64_t tidx, int64_t off, int64_t tlen, int ohang, int64_t minsc, int64_t *out6, uint8_t *editmask);
/* log of backtrace attempts of the following bt2o_dp calls: triples [candidate score, alignment index or -1, candidate index] */
void bt2o_dp_attempt_log(int64_t *buf, int cap);
int  bt2o_dp_attempt_count(void);
int  bt2o_dp(const bt2o_index *ix, const bt2o_scoring *sc, const uint8_t *codes, const uint8_t *quals, int len, int fw,
             uint64_t tidx, int64_t refl, int64_t refr, int triml, int corel, int corer, int64_t minsc, int nceil,
             int max_cands, int max_alns, int max_edits,
             int64_t *summary, int64_t *cands, int64_t *alns, int32_t *edits);

#ifdef __cplusplus
}
#endif
#endif

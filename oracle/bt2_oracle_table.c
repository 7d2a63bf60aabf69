/* bt2_oracle_table.c -- TEST INFRASTRUCTURE (like everything under oracle/): the entry points of include/bt2g.h that the
 * exact-policy engine calls (bt2g_policy_backend), answered on the CPU by the plain-C restatement in bt2_oracle.c, in the
 * entry points' own array conventions.  Lets the CPU suite and tools/parity_subset.py drive bt2g_policy_align at C speed.
 * Only tests/ and tools/ load it; the product constructs its table with bt2g_policy_backend_gpu alone. */
#include <stdlib.h>
#include <string.h>
#include "bt2_oracle.h"
#include "../include/bt2g.h"

typedef struct { bt2o_index *ix; bt2o_scoring sc; } table_ctx;

static const uint8_t *rd_seq(const bt2g_reads *r, uint64_t i, int *len) { *len = (int)(r->off[i + 1] - r->off[i]); return r->seq + r->off[i]; }
static const uint8_t *rd_qual(const bt2g_reads *r, uint64_t i) { return r->qual ? r->qual + r->off[i] : NULL; }
static int code_of(int ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; }

static int t_exact_sweep(void *c, const bt2g_reads *reads, int nofw, int norc, uint8_t *mine, uint64_t *ee) {
	table_ctx *t = (table_ctx *)c;
	for(uint64_t i = 0; i < reads->n_reads; i++) {
		int len; const uint8_t *s = rd_seq(reads, i, &len);
		uint64_t m2[2] = {0, 0}, tb[4] = {0, 0, 0, 0};
		bt2o_exact_sweep(t->ix, s, len, nofw, norc, m2, tb);
		mine[2 * i] = (uint8_t)m2[0]; mine[2 * i + 1] = (uint8_t)m2[1];
		memcpy(ee + 4 * i, tb, sizeof(tb));
	}
	return 0;
}

static int t_seed_search(void *c, const bt2g_reads *reads, const bt2g_seed_plan *plan, uint64_t *out, int32_t *nseeds) {
	table_ctx *t = (table_ctx *)c;
	const size_t per = (size_t)2 * plan->max_seeds * 4;
	uint64_t *tmp = (uint64_t *)calloc(per, sizeof(uint64_t));
	for(uint64_t i = 0; i < reads->n_reads; i++) {
		int len; const uint8_t *s = rd_seq(reads, i, &len);
		memset(tmp, 0, per * sizeof(uint64_t));
		nseeds[i] = bt2o_seed_search(t->ix, s, rd_qual(reads, i), len, plan->seed_len, plan->interval[i], plan->offset[i], plan->nofw, plan->norc,
		                             plan->max_seeds, tmp);
		memcpy(out + i * per, tmp, per * sizeof(uint64_t));
	}
	free(tmp);
	return 0;
}

static int t_one_mm(void *c, const bt2g_reads *reads, const int32_t *minsc, const uint8_t *mask, int32_t max_hits, bt2g_mm_hit *hits, int32_t *counts) {
	table_ctx *t = (table_ctx *)c;
	int64_t *o = (int64_t *)malloc(sizeof(int64_t) * 6 * 512);
	int *fw = (int *)malloc(sizeof(int) * 512);
	for(uint64_t i = 0; i < reads->n_reads; i++) {
		int len; const uint8_t *s = rd_seq(reads, i, &len);
		const int n = bt2o_one_mm(t->ix, &t->sc, s, rd_qual(reads, i), len, minsc[i], !(mask[i] & 1), !(mask[i] & 2), 512, o, fw);
		for(int k = 0; k < 4; k++) counts[4 * i + k] = 0;
		for(int k = 0; k < n && k < 512; k++) {
			const int task = fw[k] ? 0 : 2;
			int32_t *cnt = &counts[4 * i + task];
			if(*cnt >= max_hits) continue;
			bt2g_mm_hit *h = &hits[((size_t)i * 4 + task) * max_hits + *cnt];
			h->top = (uint64_t)o[6 * k]; h->bot = (uint64_t)o[6 * k + 1]; h->pos = (int32_t)o[6 * k + 2];
			h->chr = code_of((int)o[6 * k + 3]); h->qchr = code_of((int)o[6 * k + 4]); h->score = (int32_t)o[6 * k + 5];
			(*cnt)++;
		}
	}
	free(o); free(fw);
	return 0;
}

static int t_extend_exact(void *c, const bt2g_reads *reads, const bt2g_seed_plan *plan, const uint64_t *ranges, uint8_t *out) {
	table_ctx *t = (table_ctx *)c;
	const int ms = plan->max_seeds;
	for(uint64_t i = 0; i < reads->n_reads; i++) {
		int len; const uint8_t *s = rd_seq(reads, i, &len);
		const int L = plan->seed_len < len ? plan->seed_len : len;
		int n = 1;
		if(len - plan->offset[i] > L) n += (len - plan->offset[i] - L) / plan->interval[i];
		for(int st = 0; st < 2; st++) for(int k = 0; k < ms; k++) {
			const uint64_t *rg = ranges + (((size_t)i * 2 + st) * ms + k) * 4;
			uint8_t *o = out + (((size_t)i * 2 + st) * ms + k) * 2;
			o[0] = o[1] = 0;
			if(k >= n || rg[1] <= rg[0]) continue;
			uint64_t lr[2] = {0, 0};
			bt2o_extend(t->ix, s, len, st == 0, (uint64_t)(plan->offset[i] + k * plan->interval[i]), (uint64_t)L, rg[0], rg[1], rg[2], rg[3], lr);
			o[0] = (uint8_t)(lr[0] > 255 ? 255 : lr[0]); o[1] = (uint8_t)(lr[1] > 255 ? 255 : lr[1]);
		}
	}
	return 0;
}

static int t_resolve(void *c, const uint64_t *rows, const uint32_t *hitlen, uint64_t n, int reject, uint64_t *joined, uint64_t *tidx, uint64_t *textoff,
                     uint64_t *tlen, uint8_t *flags) {
	table_ctx *t = (table_ctx *)c;
	for(uint64_t i = 0; i < n; i++) {
		joined[i] = bt2o_get_offset(t->ix, rows[i]);
		uint64_t ti = 0, to = 0, tl = 0; int st = 0;
		const int ok = bt2o_joined_to_text(t->ix, hitlen[i], joined[i], reject, &ti, &to, &tl, &st);
		flags[i] = (uint8_t)((st ? 1 : 0) | (ok ? 0 : 2));
		tidx[i] = ok ? ti : 0; textoff[i] = ok ? to : 0; tlen[i] = ok ? tl : 0;
	}
	return 0;
}

static int t_get_stretch(void *c, const uint64_t *tidx, const int64_t *off, const int32_t *count, uint64_t n, int32_t stride, uint8_t *out) {
	table_ctx *t = (table_ctx *)c;
	for(uint64_t i = 0; i < n; i++) {
		memset(out + i * (size_t)stride, 4, (size_t)stride);
		bt2o_get_stretch(t->ix, tidx[i], off[i], count[i] < stride ? count[i] : stride, out + i * (size_t)stride);
	}
	return 0;
}

static int t_ungapped(void *c, const bt2g_reads *reads, const bt2g_ungapped_problem *probs, uint64_t n, bt2g_ungapped_result *out, uint8_t *mask, uint32_t stride) {
	table_ctx *t = (table_ctx *)c;
	for(uint64_t k = 0; k < n; k++) {
		const bt2g_ungapped_problem *p = &probs[k];
		int len; const uint8_t *s = rd_seq(reads, p->read_idx, &len);
		int64_t o6[6] = {0, 0, 0, 0, 0, 0};
		uint8_t *m = (uint8_t *)calloc((size_t)len + 1, 1);
		const int rc = bt2o_ungapped(t->ix, &t->sc, s, rd_qual(reads, p->read_idx), len, (int)p->fw, p->tidx, p->refoff, (int64_t)p->reflen, p->ohang, p->minsc, o6, m);
		memset(&out[k], 0, sizeof(out[k]));
		out[k].status = rc;
		if(mask) memset(mask + k * (size_t)stride, 0, stride);
		if(rc == 1) {
			out[k].score = (int32_t)o6[0]; out[k].rowi = (int32_t)o6[1]; out[k].rowf = (int32_t)o6[2]; out[k].ns = (int32_t)o6[3];
			out[k].refns = (int32_t)o6[4]; out[k].nedits = (int32_t)o6[5];
			if(mask) memcpy(mask + k * (size_t)stride, m, (size_t)(len < (int)stride ? len : (int)stride));
		}
		free(m);
	}
	return 0;
}

/* alignment (oracle edit list, 5'-relative positions) -> device op string, last read row first (policy_engine.py: aln_to_ops) */
static int aln_ops(const uint8_t *codes, int rdlen, int fw, int trim5, int trim3, const int32_t *ed, int ned, uint8_t *ops, int max_ops) {
	const int ext = rdlen - trim5 - trim3, row0 = fw ? trim5 : trim3;
	/* left-to-right edits */
	int32_t *l = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(ned + 1));
	for(int k = 0; k < ned; k++) {
		const int32_t *s = fw ? ed + 4 * k : ed + 4 * (ned - 1 - k);
		l[4 * k] = fw ? s[0] : ext - s[0] - (s[3] == 1 ? 0 : 1); l[4 * k + 1] = s[1]; l[4 * k + 2] = s[2]; l[4 * k + 3] = s[3];
	}
	uint8_t *fwd = (uint8_t *)malloc((size_t)ext + (size_t)ned + 4);
	int n = 0, k = 0;
	for(int rel = 0; rel < ext; rel++) {
		while(k < ned && l[4 * k] == rel && l[4 * k + 3] == 1) { fwd[n++] = (uint8_t)(BT2G_OP_READGAP | (code_of(l[4 * k + 1]) << 2)); k++; }
		if(k < ned && l[4 * k] == rel) { fwd[n++] = l[4 * k + 3] == 2 ? (uint8_t)BT2G_OP_REFGAP : (uint8_t)(BT2G_OP_MM | (code_of(l[4 * k + 1]) << 2)); k++; }
		else {
			const int row = row0 + rel;
			const int cc = fw ? codes[row] : (codes[rdlen - 1 - row] > 3 ? 4 : 3 - codes[rdlen - 1 - row]);
			fwd[n++] = (uint8_t)(BT2G_OP_MATCH | (cc << 2));
		}
	}
	for(int i = 0; i < n && i < max_ops; i++) ops[i] = fwd[n - 1 - i];
	free(l); free(fwd);
	return n;
}

static int t_dp_extend(void *c, const bt2g_reads *reads, const bt2g_dp_problem *probs, uint64_t n, int32_t max_cands, int32_t max_alns, int32_t max_ops,
                       bt2g_dp_summary *summ, bt2g_dp_cand *cands, bt2g_dp_aln *alns, uint8_t *ops) {
	table_ctx *t = (table_ctx *)c;
	const int MC = 65536, MA = 256, ME = 65536;
	int64_t *ocands = (int64_t *)malloc(sizeof(int64_t) * 3 * MC), *oalns = (int64_t *)malloc(sizeof(int64_t) * 8 * MA);
	int32_t *oed = (int32_t *)malloc(sizeof(int32_t) * 4 * ME);
	int64_t *att = (int64_t *)malloc(sizeof(int64_t) * 3 * MC);
	for(uint64_t k = 0; k < n; k++) {
		const bt2g_dp_problem *p = &probs[k];
		int len; const uint8_t *s = rd_seq(reads, p->read_idx, &len);
		int64_t summary[4] = {0, 0, 0, 0};
		bt2o_dp_attempt_log(att, MC);
		bt2o_dp(t->ix, &t->sc, s, rd_qual(reads, p->read_idx), len, (int)p->fw, p->tidx, p->refl, p->refr, p->triml, p->corel, p->corer, p->minsc, p->nceil,
		        MC, MA, ME, summary, ocands, oalns, oed);
		const int natt = bt2o_dp_attempt_count();
		bt2o_dp_attempt_log(NULL, 0);
		bt2g_dp_summary *sm = &summ[k];
		memset(sm, 0, sizeof(*sm));
		sm->found = (int32_t)summary[0]; sm->best = (int32_t)summary[1]; sm->ncand = (int32_t)summary[2]; sm->naln = (int32_t)summary[3];
		if(!sm->found) continue;
		if(sm->ncand > max_cands) sm->flags |= BT2G_DP_FLAG_CAND_OVERFLOW;
		if(sm->naln > max_alns) sm->flags |= BT2G_DP_FLAG_ALN_OVERFLOW;
		bt2g_dp_cand *cd = cands + k * (size_t)max_cands;
		for(int ci = 0; ci < sm->ncand && ci < max_cands; ci++) { cd[ci].row = (int32_t)ocands[3 * ci]; cd[ci].col = (int32_t)ocands[3 * ci + 1]; cd[ci].score = (int32_t)ocands[3 * ci + 2]; cd[ci].fate = BT2G_CAND_FILT_START; }
		int *cand_of = (int *)calloc((size_t)sm->naln + 1, sizeof(int));
		for(int a = 0; a < natt && a < MC; a++) {
			const int ai = (int)att[3 * a + 1], ci = (int)att[3 * a + 2];
			if(ci < max_cands) cd[ci].fate = ai >= 0 ? BT2G_CAND_SUCCEEDED : BT2G_CAND_FAILED;
			if(ai >= 0 && ai <= sm->naln) cand_of[ai] = ci;
		}
		int e0 = 0;
		for(int ai = 0; ai < sm->naln && ai < MA; ai++) {
			const int64_t *a = oalns + 8 * ai;
			const int ne = (int)a[6], fw = (int)a[7], trim5 = (int)a[4], trim3 = (int)a[5];
			if(ai < max_alns) {
				bt2g_dp_aln *al = &alns[k * (size_t)max_alns + ai];
				memset(al, 0, sizeof(*al));
				const int tl = fw ? trim5 : trim3, ext = len - trim5 - trim3;
				int ins = 0, del = 0;
				for(int q = 0; q < ne; q++) { ins += oed[4 * (e0 + q) + 3] == 1; del += oed[4 * (e0 + q) + 3] == 2; }
				const int rext = ext + ins - del;
				uint8_t *buf = (uint8_t *)malloc((size_t)rext + 1);
				bt2o_get_stretch(t->ix, p->tidx, a[3], rext, buf);
				int refns = 0; for(int q = 0; q < rext; q++) refns += buf[q] > 3;
				free(buf);
				al->cand_idx = cand_of[ai]; al->score = (int32_t)a[0]; al->ns = (int32_t)a[1]; al->gaps = (int32_t)a[2]; al->refns = refns;
				al->row0 = tl; al->col0 = (int32_t)(a[3] - p->refl); al->trim_beg = tl; al->trim_end = len - ext - tl;
				al->nops = aln_ops(s, len, fw, trim5, trim3, oed + 4 * e0, ne, ops + ((size_t)k * max_alns + ai) * max_ops, max_ops);
				if(al->nops > max_ops) sm->flags |= BT2G_DP_FLAG_OPS_OVERFLOW;
			}
			e0 += ne;
		}
		free(cand_of);
	}
	free(ocands); free(oalns); free(oed); free(att);
	return 0;
}

/* fills `be` (a bt2g_policy_backend) with the functions above; returns the context to pass to bt2o_policy_table_free */
void *bt2o_policy_table(bt2o_index *ix, int local, int off_size, bt2g_policy_backend *be) {
	table_ctx *t = (table_ctx *)calloc(1, sizeof(table_ctx));
	t->ix = ix;
	bt2o_scoring_default(&t->sc, local);
	be->ctx = t;
	be->exact_sweep = t_exact_sweep; be->seed_search = t_seed_search; be->one_mm = t_one_mm; be->extend_exact = t_extend_exact;
	be->resolve = t_resolve; be->get_stretch = t_get_stretch; be->ungapped = t_ungapped; be->dp_extend = t_dp_extend;
	be->off_size = off_size; be->reserved = 0;
	return t;
}
/* non-default penalties for the table's calls (--mp / --np / --rdg / --rfg / --ma) */
void bt2o_policy_table_scoring(void *tv, int match_bonus, int mmp_max, int mmp_min, int n_pen, int rdgap_const, int rdgap_linear,
                               int rfgap_const, int rfgap_linear) {
	table_ctx *t = (table_ctx *)tv;
	t->sc.match_bonus = match_bonus; t->sc.mmp_max = mmp_max; t->sc.mmp_min = mmp_min; t->sc.n_pen = n_pen;
	t->sc.rdgap_const = rdgap_const; t->sc.rdgap_linear = rdgap_linear; t->sc.rfgap_const = rfgap_const; t->sc.rfgap_linear = rfgap_linear;
}
void bt2o_policy_table_nceil(void *tv, double nceil_const, double nceil_linear) {
	table_ctx *t = (table_ctx *)tv;
	t->sc.nceil_const = nceil_const; t->sc.nceil_linear = nceil_linear;
}
void bt2o_policy_table_free(void *t) { free(t); }

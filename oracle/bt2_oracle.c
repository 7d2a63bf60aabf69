/* oracle/bt2_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the bowtie2 2.5.5 alignment hot path,
 * written from the reference's documented behaviour; every function cites the reference
 * file:line it follows.  It exists to CHECK the CUDA product path (tests/, smoke(), the
 * cpu_baseline "port" leg of bench.py) and is itself pinned against the real reference
 * compiled into oracle/_ref/libbt2ref_{s,l}.so (tests/test_oracle_vs_reference.py) and
 * against committed golden fixtures in tests/golden/.  Parity status: PINNED.
 *
 * Nothing under bowtie2_b200/ or include/ may include, link or call this file.
 */
#include "bt2_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OFFMASK64 (~(uint64_t)0)

/* ------------------------------------------------------------------------------------ */
/* file reading helpers                                                                   */
/* ------------------------------------------------------------------------------------ */
static int rd_bytes(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n ? 0 : -1; }

static int rd_i32(FILE *f, int32_t *v) { return rd_bytes(f, v, 4); }

static int rd_off(FILE *f, int off_size, uint64_t *v) {
	if(off_size == 4) { uint32_t x; if(rd_bytes(f, &x, 4)) return -1; *v = x; return 0; }
	return rd_bytes(f, v, 8);
}

static uint64_t *rd_off_array(FILE *f, int off_size, uint64_t n) {
	uint64_t *a = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
	if(!a) return NULL;
	for(uint64_t i = 0; i < n; i++) {
		if(rd_off(f, off_size, &a[i])) { free(a); return NULL; }
	}
	return a;
}

/* EbwtParams::init (bt2_idx.h:133-167) */
static void params_init(bt2o_ebwt *e) {
	uint64_t bwt_sz = e->len / 4 + 1;
	e->bwt_len = e->len + 1;
	e->side_sz = 1u << e->line_rate;
	e->side_bwt_sz = e->side_sz - 4 * (uint32_t)e->off_size;
	e->side_bwt_len = e->side_bwt_sz * 4;
	e->num_sides = (bwt_sz + e->side_bwt_sz - 1) / e->side_bwt_sz;
	e->ebwt_tot_len = e->num_sides * e->side_sz;
	e->eftab_len = (uint64_t)e->ftab_chars * 2;
	e->ftab_len = ((uint64_t)1 << (e->ftab_chars * 2)) + 1;
	e->offs_len = (e->bwt_len + ((uint64_t)1 << e->off_rate) - 1) >> e->off_rate;
}

/* Ebwt::readIntoMemory (bt2_io.cpp:131-511), little-endian files only */
static int load_ebwt(bt2o_ebwt *e, const char *base, const char *ext, int off_size, int is_fw, int load_sa) {
	char path[4096];
	memset(e, 0, sizeof(*e));
	e->off_size = off_size;
	e->is_fw = is_fw;
	snprintf(path, sizeof(path), "%s.1.%s", base, ext);
	FILE *f = fopen(path, "rb");
	if(!f) return -1;
	int32_t one, lines_per_side;
	if(rd_i32(f, &one) || one != 1) { fclose(f); return -2; }
	if(rd_off(f, off_size, &e->len)) { fclose(f); return -3; }
	if(rd_i32(f, &e->line_rate) || rd_i32(f, &lines_per_side) || rd_i32(f, &e->off_rate) ||
	   rd_i32(f, &e->ftab_chars) || rd_i32(f, &e->flags)) { fclose(f); return -3; }
	params_init(e);
	if(rd_off(f, off_size, &e->n_pat)) { fclose(f); return -3; }
	e->plen = rd_off_array(f, off_size, e->n_pat);
	if(rd_off(f, off_size, &e->n_frag)) { fclose(f); return -3; }
	e->rstarts = rd_off_array(f, off_size, e->n_frag * 3);
	e->ebwt = (uint8_t *)malloc(e->ebwt_tot_len);
	if(!e->plen || !e->rstarts || !e->ebwt || rd_bytes(f, e->ebwt, e->ebwt_tot_len)) { fclose(f); return -4; }
	if(rd_off(f, off_size, &e->z_off)) { fclose(f); return -3; }
	for(int i = 0; i < 5; i++) if(rd_off(f, off_size, &e->fchr[i])) { fclose(f); return -3; }
	e->ftab = rd_off_array(f, off_size, e->ftab_len);
	e->eftab = rd_off_array(f, off_size, e->eftab_len);
	fclose(f);
	if(!e->ftab || !e->eftab) return -4;
	if(load_sa) {
		snprintf(path, sizeof(path), "%s.2.%s", base, ext);
		f = fopen(path, "rb");
		if(!f) return -5;
		if(rd_i32(f, &one) || one != 1) { fclose(f); return -2; }
		e->offs = malloc(e->offs_len * (uint64_t)off_size);
		if(!e->offs || rd_bytes(f, e->offs, e->offs_len * (uint64_t)off_size)) { fclose(f); return -4; }
		fclose(f);
	}
	return 0;
}

/* BitPairReference ctor (reference.cpp:96-200) */
static int load_ref(bt2o_ref *r, const char *base, const char *ext, int off_size) {
	char path[4096];
	memset(r, 0, sizeof(*r));
	snprintf(path, sizeof(path), "%s.3.%s", base, ext);
	FILE *f = fopen(path, "rb");
	if(!f) return -1;
	int32_t one;
	if(rd_i32(f, &one) || one != 1) { fclose(f); return -2; }
	if(rd_off(f, off_size, &r->n_recs)) { fclose(f); return -3; }
	uint64_t n = r->n_recs;
	r->rec_off = (uint64_t *)malloc(n * 8); r->rec_len = (uint64_t *)malloc(n * 8);
	r->rec_first = (uint8_t *)malloc(n);
	r->ref_rec_offs = (uint64_t *)malloc((n + 1) * 8);
	r->ref_offs = (uint64_t *)malloc((n + 1) * 8);
	r->ref_lens = (uint64_t *)malloc((n + 1) * 8);
	uint64_t cumsz = 0, cumlen = 0, nrefs = 0, nlens = 0;
	for(uint64_t i = 0; i < n; i++) {
		int c;
		if(rd_off(f, off_size, &r->rec_off[i]) || rd_off(f, off_size, &r->rec_len[i]) || (c = fgetc(f)) == EOF) {
			fclose(f); return -3;
		}
		r->rec_first[i] = c ? 1 : 0;
		if(r->rec_first[i]) {
			r->ref_rec_offs[nrefs] = i;
			r->ref_offs[nrefs] = cumsz;
			if(nrefs > 0) r->ref_lens[nlens++] = cumlen;
			cumlen = 0;
			nrefs++;
		}
		cumsz += r->rec_len[i];
		cumlen += r->rec_off[i] + r->rec_len[i];
	}
	fclose(f);
	r->ref_rec_offs[nrefs] = n;
	r->ref_offs[nrefs] = cumsz;
	r->ref_lens[nlens++] = cumlen;
	r->n_refs = nrefs;
	r->buf_sz = cumsz;
	uint64_t bytes = (cumsz + 3) >> 2;
	r->buf = (uint8_t *)malloc(bytes ? bytes : 1);
	snprintf(path, sizeof(path), "%s.4.%s", base, ext);
	f = fopen(path, "rb");
	if(!f) return -1;
	if(rd_bytes(f, r->buf, bytes)) { fclose(f); return -3; }
	fclose(f);
	return 0;
}

static void free_ebwt(bt2o_ebwt *e) {
	free(e->plen); free(e->rstarts); free(e->ebwt); free(e->ftab); free(e->eftab); free(e->offs);
}

bt2o_index *bt2o_open(const char *base, int load_mirror, int load_reference) {
	char path[4096];
	const char *ext = "bt2";
	int off_size = 4;
	snprintf(path, sizeof(path), "%s.1.bt2", base);
	FILE *f = fopen(path, "rb");
	if(!f) {
		snprintf(path, sizeof(path), "%s.1.bt2l", base);
		f = fopen(path, "rb");
		if(!f) return NULL;
		ext = "bt2l"; off_size = 8;
	}
	fclose(f);
	bt2o_index *ix = (bt2o_index *)calloc(1, sizeof(bt2o_index));
	if(load_ebwt(&ix->fw, base, ext, off_size, 1, 1)) { free(ix); return NULL; }
	if(load_mirror) {
		snprintf(path, sizeof(path), "%s.rev", base);
		if(load_ebwt(&ix->bw, path, ext, off_size, 0, 0)) { free_ebwt(&ix->fw); free(ix); return NULL; }
		ix->has_bw = 1;
	}
	if(load_reference) {
		if(load_ref(&ix->ref, base, ext, off_size)) { free_ebwt(&ix->fw); free(ix); return NULL; }
		ix->has_ref = 1;
	}
	return ix;
}

void bt2o_close(bt2o_index *ix) {
	if(!ix) return;
	free_ebwt(&ix->fw);
	if(ix->has_bw) free_ebwt(&ix->bw);
	if(ix->has_ref) {
		bt2o_ref *r = &ix->ref;
		free(r->rec_off); free(r->rec_len); free(r->rec_first); free(r->ref_rec_offs);
		free(r->ref_offs); free(r->ref_lens); free(r->buf);
	}
	free(ix);
}

uint64_t bt2o_scalar(const bt2o_index *ix, int mirror, int which) {
	const bt2o_ebwt *e = mirror ? &ix->bw : &ix->fw;
	switch(which) {
		case 0: return e->len;        case 1: return e->bwt_len;
		case 2: return (uint64_t)e->line_rate; case 3: return (uint64_t)e->off_rate;
		case 4: return (uint64_t)e->ftab_chars; case 5: return e->num_sides;
		case 6: return e->side_sz;    case 7: return e->side_bwt_sz;
		case 8: return e->z_off;      case 9: return e->n_pat;
		case 10: return e->n_frag;    case 11: return e->offs_len;
		case 12: return e->ftab_len;  case 13: return e->eftab_len;
		case 14: return e->ebwt_tot_len;
	}
	return 0;
}

/* ------------------------------------------------------------------------------------ */
/* FM primitives                                                                          */
/* ------------------------------------------------------------------------------------ */
static inline uint64_t side_occ(const bt2o_ebwt *e, const uint8_t *side, int c) {
	const uint8_t *p = side + e->side_bwt_sz + (size_t)c * e->off_size;
	if(e->off_size == 4) { uint32_t v; memcpy(&v, p, 4); return v; }
	uint64_t v; memcpy(&v, p, 8); return v;
}

/* countUpToEx (bt2_idx.h:2038): occurrences of each nucleotide in the first char_off
 * 2-bit characters of the side.  Written as the obvious loop; the reference's word
 * popcount + LUT tail computes the same thing. */
static void count_upto4(const uint8_t *side, uint32_t char_off, uint64_t cnt[4]) {
	cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
	for(uint32_t k = 0; k < char_off; k++) {
		cnt[(side[k >> 2] >> ((k & 3) << 1)) & 3]++;
	}
}

/* countBt2SideEx (bt2_idx.h:1887-1919) with SideLocus::initFromRow (:369-397).
 * The "$" is stored as an A at row z_off; an A-count taken strictly after it within
 * the same side is decremented (:1891-1899). */
void bt2o_rank4(const bt2o_index *ix, int mirror, uint64_t row, uint64_t out4[4]) {
	const bt2o_ebwt *e = mirror ? &ix->bw : &ix->fw;
	uint64_t side_num = row / e->side_bwt_len;
	uint32_t char_off = (uint32_t)(row % e->side_bwt_len);
	const uint8_t *side = e->ebwt + side_num * e->side_sz;
	uint64_t cnt[4];
	count_upto4(side, char_off, cnt);
	uint64_t z_side = e->z_off / e->side_bwt_len;
	uint32_t z_char = (uint32_t)(e->z_off % e->side_bwt_len);
	if(side_num == z_side && char_off > z_char) cnt[0]--;
	for(int c = 0; c < 4; c++) out4[c] = cnt[c] + side_occ(e, side, c) + e->fchr[c];
}

/* countBt2Side / mapLF(l,c) (bt2_idx.h:1758-1793, :2344) */
uint64_t bt2o_rank1(const bt2o_index *ix, int mirror, uint64_t row, int c) {
	uint64_t r[4];
	bt2o_rank4(ix, mirror, row, r);
	return r[c];
}

/* rowL (bt2_idx.h:2247) */
int bt2o_rowL(const bt2o_index *ix, int mirror, uint64_t row) {
	const bt2o_ebwt *e = mirror ? &ix->bw : &ix->fw;
	uint64_t side_num = row / e->side_bwt_len;
	uint32_t k = (uint32_t)(row % e->side_bwt_len);
	const uint8_t *side = e->ebwt + side_num * e->side_sz;
	return (side[k >> 2] >> ((k & 3) << 1)) & 3;
}

/* mapLF1(row, l, c) (bt2_idx.h:2420-2443) */
uint64_t bt2o_maplf1(const bt2o_index *ix, int mirror, uint64_t row, int c) {
	const bt2o_ebwt *e = mirror ? &ix->bw : &ix->fw;
	if(bt2o_rowL(ix, mirror, row) != c || row == e->z_off) return OFFMASK64;
	return bt2o_rank1(ix, mirror, row, c);
}

/* mapLFRange (bt2_idx.h:2268-2305) = countBt2SideRange (:1804-1865) + countBt2SideRange2 (:2177-2239), the
 * GroupWalk step for a range that is still several rows wide (group_walk.h:897).  upto[c] = rank of c at `top`
 * (as countBt2SideEx: the "$" stored as an A is not counted); in[c] = rows of [top, top+num) whose BWT character
 * is c -- the "$" row, if inside, IS tallied as an A there (:2209) -- and chars[j] is that character (the
 * reference's four bool lists masks[c][j] == (chars[j] == c)). */
void bt2o_maplf_range(const bt2o_index *ix, int mirror, uint64_t top, uint64_t num, uint64_t upto[4], uint64_t in[4], uint8_t *chars) {
	bt2o_rank4(ix, mirror, top, upto);
	in[0] = in[1] = in[2] = in[3] = 0;
	for(uint64_t j = 0; j < num; j++) {
		int c = bt2o_rowL(ix, mirror, top + j);   /* nextSide (:1857) == the row's own side */
		in[c]++;
		chars[j] = (uint8_t)c;
	}
}

/* ftabHi / ftabLo / ftabLoHi (bt2_idx.h:1428-1554): entries > len are indirections */
static uint64_t ftab_hi(const bt2o_ebwt *e, uint64_t i) {
	uint64_t v = e->ftab[i];
	if(v <= e->len) return v;
	uint64_t mask = e->off_size == 4 ? 0xffffffffull : OFFMASK64;
	uint64_t ef = (v ^ mask) & mask;
	return e->eftab[ef * 2 + 1];
}
static uint64_t ftab_lo(const bt2o_ebwt *e, uint64_t i) {
	uint64_t v = e->ftab[i];
	if(v <= e->len) return v;
	uint64_t mask = e->off_size == 4 ? 0xffffffffull : OFFMASK64;
	uint64_t ef = (v ^ mask) & mask;
	return e->eftab[ef * 2];
}
void bt2o_ftab_lohi(const bt2o_index *ix, int mirror, uint64_t i, uint64_t *top, uint64_t *bot) {
	const bt2o_ebwt *e = mirror ? &ix->bw : &ix->fw;
	*top = ftab_hi(e, i);
	*bot = ftab_lo(e, i + 1);
}

static inline uint64_t offs_at(const bt2o_ebwt *e, uint64_t k) {
	if(e->off_size == 4) return ((const uint32_t *)e->offs)[k];
	return ((const uint64_t *)e->offs)[k];
}

/* Ebwt::getOffset(row) (bt2_idx.cpp:150-171); == GroupWalk result (group_walk.h:517-520) */
uint64_t bt2o_get_offset(const bt2o_index *ix, uint64_t row) {
	const bt2o_ebwt *e = &ix->fw;
	uint64_t jumps = 0;
	uint64_t rate_mask = ((uint64_t)1 << e->off_rate) - 1;
	for(;;) {
		if(row == e->z_off) return jumps;
		if((row & rate_mask) == 0) return jumps + offs_at(e, row >> e->off_rate);
		int c = bt2o_rowL(ix, 0, row);
		row = bt2o_rank1(ix, 0, row, c);
		jumps++;
	}
}

/* Ebwt::joinedToTextOff (bt2_idx.cpp:54-124), forward index */
int bt2o_joined_to_text(const bt2o_index *ix, uint64_t qlen, uint64_t off, int reject_straddle,
                        uint64_t *tidx, uint64_t *textoff, uint64_t *tlen, int *straddled) {
	const bt2o_ebwt *e = &ix->fw;
	uint64_t top = 0, bot = e->n_frag;
	*straddled = 0;
	for(;;) {
		uint64_t elt = top + ((bot - top) >> 1);
		uint64_t lower = e->rstarts[elt * 3];
		uint64_t upper = (elt == e->n_frag - 1) ? e->len : e->rstarts[(elt + 1) * 3];
		if(lower <= off) {
			if(upper > off) {
				if(off + qlen > upper) {
					*straddled = 1;
					if(reject_straddle) { *tidx = OFFMASK64; *textoff = 0; *tlen = 0; return 0; }
				}
				*tidx = e->rstarts[elt * 3 + 1];
				*textoff = (off - lower) + e->rstarts[elt * 3 + 2];
				break;
			}
			top = elt;
		} else {
			bot = elt;
		}
	}
	*tlen = e->plen[*tidx];
	return 1;
}

/* BitPairReference::getBase (reference.cpp:330-358), with off-end positions read as N (4) */
static int ref_get_base(const bt2o_ref *r, uint64_t tidx, uint64_t toff) {
	uint64_t reci = r->ref_rec_offs[tidx], recf = r->ref_rec_offs[tidx + 1];
	uint64_t buf_off = r->ref_offs[tidx], off = 0;
	for(uint64_t i = reci; i < recf; i++) {
		off += r->rec_off[i];
		if(toff < off) return 4;
		uint64_t rec_end = off + r->rec_len[i];
		if(toff < rec_end) {
			buf_off += toff - off;
			return (r->buf[buf_off >> 2] >> ((buf_off & 3) << 1)) & 3;
		}
		buf_off += r->rec_len[i];
		off = rec_end;
	}
	return 4;
}

int bt2o_get_stretch(const bt2o_index *ix, uint64_t tidx, int64_t off, int64_t count, uint8_t *out) {
	const bt2o_ref *r = &ix->ref;
	int64_t tlen = (int64_t)r->ref_lens[tidx];
	for(int64_t i = 0; i < count; i++) {
		int64_t p = off + i;
		out[i] = (p < 0 || p >= tlen) ? 4 : (uint8_t)ref_get_base(r, tidx, (uint64_t)p);
	}
	return 0;
}

/* ------------------------------------------------------------------------------------ */
/* seed search                                                                            */
/* ------------------------------------------------------------------------------------ */

/* ftabSeqToInt (bt2_idx.h:1373-1398): chars OR-ed in search order. fwex selects L->R. */
static uint64_t ftab_seq_to_int(const uint8_t *seq, int off, int fc, int left_to_right) {
	uint64_t v = 0;
	for(int i = 0; i < fc; i++) {
		int c = left_to_right ? seq[off + i] : seq[off + fc - i - 1];
		if(c > 3) return OFFMASK64;
		v = (v << 2) | (uint64_t)c;
	}
	return v;
}

/* one strand of SeedAligner::exactSweep (aligner_seed.cpp:856-970; helpers :760-850).
 * The reference interleaves the fw and rc sweeps only for prefetching; each strand's
 * state is independent. */
static void exact_sweep_strand(const bt2o_index *ix, const uint8_t *seq, int len, int mine_max,
                               uint64_t *mine, uint64_t *otop, uint64_t *obot, int *finished) {
	const bt2o_ebwt *e = &ix->fw;
	int ftab_len = e->ftab_chars;
	uint64_t top = 0, bot = 0;
	int dep = 0, nedit = 0, do_init = 1, done = 0;
	while(dep < len && !done) {
		if(do_init) {
			/* exactSweepInit (:760-800) */
			int left = len - dep;
			int do_ftab = ftab_len > 1 && left >= ftab_len;
			top = bot = 0;
			if(do_ftab) {
				for(int i = 0; i < ftab_len; i++) if(seq[left - 1 - i] > 3) { do_ftab = 0; break; }
			}
			if(do_ftab) {
				uint64_t fi = ftab_seq_to_int(seq, left - ftab_len, ftab_len, 1);
				top = ftab_hi(e, fi); bot = ftab_lo(e, fi + 1);
				dep += ftab_len;
			} else {
				int c = seq[len - dep - 1];
				if(c < 4) { top = e->fchr[c]; bot = e->fchr[c + 1]; }
				dep++;
			}
			/* exactSweepStep (:826-848) */
			if(bot <= top) {
				nedit++;
				if(nedit >= mine_max) { *mine = (uint64_t)nedit; done = 1; }
				continue;
			}
			do_init = 0;
		}
		if(dep < len) {
			/* exactSweepMapLF (:802-824) */
			int c = seq[len - dep - 1];
			if(c > 3) {
				top = bot = 0;
			} else if(bot - top > 1) {
				uint64_t t = bt2o_rank1(ix, 0, top, c), b = bt2o_rank1(ix, 0, bot, c);
				top = t; bot = b;
			} else {
				uint64_t t = bt2o_maplf1(ix, 0, top, c);
				if(t == OFFMASK64) { top = bot = 0; } else { top = t; bot = t + 1; }
			}
			if(bot <= top) {
				nedit++;
				if(nedit >= mine_max) { *mine = (uint64_t)nedit; done = 1; }
				do_init = 1;
			}
			dep++;
		}
	}
	*finished = 0;
	*otop = *obot = 0;
	if(!done && dep >= len) {
		*mine = (uint64_t)nedit;
		*finished = 1;
		if(nedit == 0 && bot > top) { *otop = top; *obot = bot; }
	}
}

uint64_t bt2o_exact_sweep(const bt2o_index *ix, const uint8_t *codes, int len, int nofw, int norc,
                          uint64_t mine2[2], uint64_t topbot4[4]) {
	uint8_t *rc = (uint8_t *)malloc((size_t)len + 1);
	for(int i = 0; i < len; i++) { int c = codes[len - 1 - i]; rc[i] = (uint8_t)(c > 3 ? 4 : 3 - c); }
	uint64_t nelt = 0;
	mine2[0] = mine2[1] = 0;
	topbot4[0] = topbot4[1] = topbot4[2] = topbot4[3] = 0;
	int fin;
	if(!nofw) {
		exact_sweep_strand(ix, codes, len, 2, &mine2[0], &topbot4[0], &topbot4[1], &fin);
		nelt += topbot4[1] - topbot4[0];
	}
	if(!norc) {
		exact_sweep_strand(ix, rc, len, 2, &mine2[1], &topbot4[2], &topbot4[3], &fin);
		nelt += topbot4[3] - topbot4[2];
	}
	free(rc);
	return nelt;
}

/* One exact seed: Seed::instantiate SEED_TYPE_EXACT (aligner_seed.cpp:252-259, N check
 * :326-352), CacheAndSeed ftab indices (:88-112), startSearchSeedBi (:1637-1718, NDEBUG
 * branch for the mirror range), searchSeedBi exact path (:1858-2037) with mapBiLFEx
 * (bt2_idx.h:2372-2413) and mapLF1 (:2420). seq is the seed as it aligns to the Watson
 * strand.  Returns 1 and fills r[4]=topf,botf,topb,botb if the seed occurs. */
static int exact_seed(const bt2o_index *ix, const uint8_t *seq, int seedlen, uint64_t r[4]) {
	const bt2o_ebwt *fw = &ix->fw;
	const bt2o_ebwt *bw = ix->has_bw ? &ix->bw : NULL;
	int ftab_len = fw->ftab_chars;
	for(int i = 0; i < seedlen; i++) if(seq[i] > 3) return 0; /* exact zone cannot absorb an N */
	uint64_t topf, botf, topb = 0, botb = 0;
	int step;
	if(ftab_len > 1 && ftab_len <= seedlen) {
		int off = seedlen - ftab_len;
		uint64_t fwi0 = ftab_seq_to_int(seq, off, ftab_len, 1);
		topf = ftab_hi(fw, fwi0); botf = ftab_lo(fw, fwi0 + 1);
		if(botf - topf == 0) return 0;
		if(bw) {
			uint64_t bwi0 = ftab_seq_to_int(seq, off, ftab_len, 0);
			topb = ftab_hi(bw, bwi0);
			botb = topb + (botf - topf);
		}
		step = ftab_len;
	} else {
		int c = seq[seedlen - 1];
		topf = topb = fw->fchr[c];
		botf = botb = fw->fchr[c + 1];
		if(botf - topf == 0) return 0;
		step = 1;
	}
	for(; step < seedlen; step++) {
		int c = seq[seedlen - step - 1];
		if(botf - topf > 1) {
			uint64_t t[4], b[4];
			bt2o_rank4(ix, 0, topf, t);
			bt2o_rank4(ix, 0, botf, b);
			/* mirror range by prefix sums of widths (bt2_idx.h:2404-2412) */
			uint64_t tp = topb;
			for(int j = 0; j < c; j++) tp += b[j] - t[j];
			if(b[c] == t[c]) return 0;
			topf = t[c]; botf = b[c];
			topb = tp; botb = tp + (b[c] - t[c]);
		} else {
			uint64_t t = bt2o_maplf1(ix, 0, topf, c);
			if(t == OFFMASK64) return 0;
			topf = t; botf = t + 1;
			/* topb/botb unchanged (tp[]/bp[] were initialised to them, :1766-1768) */
		}
	}
	r[0] = topf; r[1] = botf; r[2] = topb; r[3] = botb;
	return 1;
}

/* SeedAligner::instantiateSeeds + searchAllSeeds for exact seeds (aligner_seed.cpp:498-720);
 * seed extraction per instantiateSeq (:471-493) / SStringExpandable windowGetDna. */
int bt2o_seed_search(const bt2o_index *ix, const uint8_t *codes, const uint8_t *quals, int len,
                     int seedlen, int interval, int offset, int nofw, int norc,
                     int max_seeds, uint64_t *out_ranges) {
	(void)quals;
	int nseeds = 1;
	if(len - offset > seedlen) nseeds += (len - offset - seedlen) / interval;
	if(nseeds > max_seeds) return -1;
	memset(out_ranges, 0, sizeof(uint64_t) * 2 * (size_t)max_seeds * 4);
	int sl = seedlen < len ? seedlen : len;
	uint8_t seq[1024];
	if(sl > 1024) return -1;
	for(int fwi = 0; fwi < 2; fwi++) {
		if((fwi == 0 && nofw) || (fwi == 1 && norc)) continue;
		for(int i = 0; i < nseeds; i++) {
			int depth = i * interval + offset;
			for(int k = 0; k < sl; k++) {
				if(fwi == 0) seq[k] = codes[depth + k];
				else { int c = codes[depth + sl - 1 - k]; seq[k] = (uint8_t)(c > 3 ? 4 : 3 - c); }
			}
			uint64_t r[4];
			if(exact_seed(ix, seq, sl, r)) {
				memcpy(out_ranges + ((size_t)fwi * max_seeds + i) * 4, r, sizeof(r));
			}
		}
	}
	return nseeds;
}

static int mm_pen(const bt2o_scoring *sc, int q);

/* SeedAligner::oneMmSearch (aligner_seed.cpp:975-1325) with repex = false, rep1mm = true.
 * Hits are appended in the reference's loop order.  out: 6 x int64 per hit (top, bot, pos, chr,
 * qchr as ASCII, score); out_fw per hit. */
static void bi_step_o(const bt2o_index *ix, int mirror, uint64_t top, uint64_t bot, uint64_t topp,
                      uint64_t t[4], uint64_t b[4], uint64_t tp[4], uint64_t bp[4]) {
	bt2o_rank4(ix, mirror, top, t);
	bt2o_rank4(ix, mirror, bot, b);
	uint64_t acc = topp;
	for(int j = 0; j < 4; j++) { tp[j] = acc; acc += b[j] - t[j]; bp[j] = acc; }
}

int bt2o_one_mm(const bt2o_index *ix, const bt2o_scoring *sc, const uint8_t *codes, const uint8_t *quals, int len,
                int64_t minsc, int nofw, int norc, int max_hits, int64_t *out, int *out_fw) {
	static const char dna[] = "ACGTN";
	int nh = 0, ns = 0;
	for(int i = 0; i < len; i++) ns += codes[i] > 3;
	if(ns > 1 || len < 2 || !ix->has_bw) return 0;
	const int nceil = (int)(sc->nceil_const + sc->nceil_linear * (double)len);
	uint8_t *pat[2][2];                          /* [fw?0:1][ebwtfw?0:1] = patFw, patFwRev, patRc, patRcRev */
	uint8_t *qu[2];                              /* qual, qualRev */
	for(int a = 0; a < 2; a++) for(int b = 0; b < 2; b++) pat[a][b] = (uint8_t *)malloc((size_t)len);
	qu[0] = (uint8_t *)malloc((size_t)len); qu[1] = (uint8_t *)malloc((size_t)len);
	for(int i = 0; i < len; i++) {
		int c = codes[i], rc = codes[len - 1 - i]; rc = rc > 3 ? 4 : 3 - rc;
		pat[0][0][i] = (uint8_t)c; pat[0][1][len - 1 - i] = (uint8_t)c;
		pat[1][0][i] = (uint8_t)rc; pat[1][1][len - 1 - i] = (uint8_t)rc;
		qu[0][i] = quals[i]; qu[1][len - 1 - i] = quals[i];
	}
	const int ftab_len = ix->fw.ftab_chars;
	for(int fwi = 0; fwi < 2; fwi++) {
		const int fw = fwi == 0;
		if((fw && nofw) || (!fw && norc)) continue;
		for(int pass = 0; pass < 2; pass++) {
			const int ebwtfw = pass == 0, mir = ebwtfw ? 0 : 1;
			const bt2o_ebwt *e = mir ? &ix->bw : &ix->fw, *ep = mir ? &ix->fw : &ix->bw;
			const uint8_t *seq = pat[fw ? 0 : 1][ebwtfw ? 0 : 1];
			const uint8_t *qual = fw ? (ebwtfw ? qu[0] : qu[1]) : (ebwtfw ? qu[1] : qu[0]);
			const int nea = ebwtfw ? (len >> 1) : ((len >> 1) + (len & 1));
			int skip = 0;
			for(int dep = 0; dep < nea; dep++) if(seq[len - dep - 1] > 3) { skip = 1; break; }
			if(skip) continue;
			uint64_t top, bot, topp, botp, t[4], b[4], tp[4], bp[4];
			int dep;
			if(ftab_len > 1 && ftab_len <= nea) {
				uint64_t fi = ftab_seq_to_int(seq, len - ftab_len, ftab_len, 1), fip = ftab_seq_to_int(seq, len - ftab_len, ftab_len, 0);
				top = ftab_hi(e, fi); bot = ftab_lo(e, fi + 1);
				topp = ftab_hi(ep, fip); botp = ftab_lo(ep, fip + 1);
				if(bot <= top) continue;
				dep = ftab_len;
			} else {
				int c = seq[len - 1];
				top = topp = e->fchr[c]; bot = botp = e->fchr[c + 1];
				if(bot <= top) continue;
				dep = 1;
			}
			int dead = 0;
			for(; dep < nea; dep++) {
				int rdc = seq[len - dep - 1];
				bi_step_o(ix, mir, top, bot, topp, t, b, tp, bp);
				if(b[rdc] <= t[rdc]) { dead = 1; break; }
				topp = tp[rdc]; botp = bp[rdc]; top = t[rdc]; bot = b[rdc];
			}
			if(dead) continue;
			for(; dep < len; dep++) {
				int rdc = seq[len - dep - 1], quc = qual[len - dep - 1];
				if(rdc > 3 && nceil == 0) break;
				if(bot - top == 1 && top == e->z_off) break;
				bi_step_o(ix, mir, top, bot, topp, t, b, tp, bp);
				if(ns == 0 || rdc > 3) {
					for(int j = 0; j < 4; j++) {
						if(j == rdc || b[j] == t[j]) continue;
						uint64_t topm = t[j], botm = b[j], topmp = tp[j], botmp = bp[j];
						int depm = dep + 1;
						for(; depm < len; depm++) {
							int rdcm = seq[len - depm - 1];
							if(rdcm > 3) break;
							uint64_t tm[4], bm[4], tmp[4], bmp[4];
							bi_step_o(ix, mir, topm, botm, topmp, tm, bm, tmp, bmp);
							if(bm[rdcm] <= tm[rdcm]) break;
							topmp = tmp[rdcm]; botmp = bmp[rdcm]; topm = tm[rdcm]; botm = bm[rdcm];
						}
						if(depm != len) continue;
						int off5p = dep;
						if(fw == ebwtfw) off5p = len - off5p - 1;
						int pen = rdc > 3 ? -sc->n_pen : -mm_pen(sc, quc - 33);
						int64_t score = (int64_t)(len - 1) * sc->match_bonus + pen;
						int valid = 1;
						if(sc->local) {
							int64_t lf = 0, lb = 0;
							for(int i = 0; i < len && valid; i++) {
								if(i == dep) { if(lf + pen <= 0) valid = 0; lf += pen; } else lf += sc->match_bonus;
								if(len - i - 1 == dep) { if(lb + pen <= 0) valid = 0; lb += pen; } else lb += sc->match_bonus;
							}
						}
						if(valid && score >= minsc) {
							if(nh < max_hits) {
								int64_t *o = out + 6 * nh;
								o[0] = ebwtfw ? topm : topmp; o[1] = ebwtfw ? botm : botmp; o[2] = off5p;
								o[3] = dna[j]; o[4] = dna[rdc]; o[5] = score;
								out_fw[nh] = fw;
							}
							nh++;
						}
					}
				}
				if(rdc > 3 || b[rdc] <= t[rdc] || dep == len - 1) break;
				topp = tp[rdc]; botp = bp[rdc]; top = t[rdc]; bot = b[rdc];
			}
		}
	}
	for(int a = 0; a < 2; a++) for(int b2 = 0; b2 < 2; b2++) free(pat[a][b2]);
	free(qu[0]); free(qu[1]);
	return nh;
}

/* SwDriver::extend (aligner_sw_driver.cpp:299-484): extend a seed hit to the left with the
 * forward index and to the right with the mirror index while the BW range keeps its size and
 * the (single) extending character equals the read character; at most 255 each way.
 * One direction; `mirror` selects the index, seq/i0/step address the read characters. */
static uint64_t extend_one(const bt2o_index *ix, int mirror, uint64_t top, uint64_t bot,
                           const uint8_t *seq, int64_t i0, int step, uint64_t lim) {
	const bt2o_ebwt *e = mirror ? &ix->bw : &ix->fw;
	uint64_t n = 0;
	for(uint64_t ii = 0; ii < lim; ii++) {
		int rdc = seq[i0 + (int64_t)ii * step];
		if(bot - top > 1) {
			uint64_t t[4], b[4];
			bt2o_rank4(ix, mirror, top, t);
			bt2o_rank4(ix, mirror, bot, b);
			int nonz = -1, abort = 0;
			uint64_t orig = bot - top;
			for(int j = 0; j < 4; j++) {
				if(b[j] > t[j]) {
					if(nonz >= 0) { abort = 1; break; }
					nonz = j; top = t[j]; bot = b[j];
				}
			}
			if(abort || (nonz != rdc && rdc <= 3) || bot - top < orig) break;
		} else {
			/* int Ebwt::mapLF1(row&, l) (bt2_idx.h:2451-2473): -1 and row unchanged at the "$" row */
			int c = -1;
			if(top != e->z_off) { c = bt2o_rowL(ix, mirror, top); top = bt2o_rank1(ix, mirror, top, c); }
			if(c != rdc && rdc <= 3) break;
			bot = top + 1;
		}
		if(++n == 255) break;
	}
	return n;
}

void bt2o_extend(const bt2o_index *ix, const uint8_t *codes, int len, int fw, uint64_t off, uint64_t seedlen,
                 uint64_t topf, uint64_t botf, uint64_t topb, uint64_t botb, uint64_t nlex_nrex[2]) {
	uint8_t *rc = (uint8_t *)malloc((size_t)len + 1);
	for(int i = 0; i < len; i++) { int c = codes[len - 1 - i]; rc[i] = (uint8_t)(c > 3 ? 4 : 3 - c); }
	const uint8_t *seq = fw ? codes : rc;
	uint64_t rdlen = (uint64_t)len;
	uint64_t lim = fw ? off : rdlen - seedlen - off;
	nlex_nrex[0] = nlex_nrex[1] = 0;
	if(lim > 0) {
		int64_t i0 = fw ? (int64_t)off - 1 : (int64_t)(rdlen - off - seedlen) - 1;
		nlex_nrex[0] = extend_one(ix, 0, topf, botf, seq, i0, -1, lim);
	}
	lim = fw ? rdlen - seedlen - off : off;
	if(lim > 0 && ix->has_bw) {
		int64_t i0 = fw ? (int64_t)(seedlen + off) : (int64_t)(rdlen - off);
		nlex_nrex[1] = extend_one(ix, 1, topb, botb, seq, i0, +1, lim);
	}
	free(rc);
}

/* ------------------------------------------------------------------------------------ */
/* seed-extension DP: end-to-end fill, candidate gather, backtrace                       */
/* ------------------------------------------------------------------------------------ */
#define DPNEG (-(1 << 28))

/* Scoring::initPens, COST_MODEL_QUAL (scoring.h:103-132) */
static int mm_pen(const bt2o_scoring *sc, int q) {
	if(q < 0) q = 0;
	int ii = q < 40 ? q : 40;
	float frac = (float)ii / 40.0f;
	return sc->mmp_min + (int)(frac * (float)(sc->mmp_max - sc->mmp_min));
}

void bt2o_scoring_default(bt2o_scoring *sc, int local) {
	sc->match_bonus = local ? 2 : 0;
	sc->mmp_max = 6; sc->mmp_min = 2; sc->n_pen = 1;
	sc->rdgap_const = 5; sc->rdgap_linear = 3; sc->rfgap_const = 5; sc->rfgap_linear = 3;
	sc->gapbar = 4; sc->local = local;
	sc->nceil_const = 0.0; sc->nceil_linear = (double)0.15f;
}

/* Scoring::score(rdc, refm, q) (scoring.h:241-251) with refc a code 0..4 instead of a mask */
static int cell_score(const bt2o_scoring *sc, int rdc, int refc, int q) {
	if(rdc > 3 || refc > 3) return -sc->n_pen;
	if(rdc == refc) return sc->match_bonus;
	return -mm_pen(sc, q);
}

/* SwAligner::ungappedAlign (aligner_sw.cpp:286-487).  out6: score, rowi, rowf, ns, refns, nedits;
 * editmask[i] = 1 where row i (strand orientation) carries an edit.  Returns 0 / -1 / 1. */
int bt2o_ungapped(const bt2o_index *ix, const bt2o_scoring *sc, const uint8_t *codes, const uint8_t *quals, int len, int fw,
                  uint64_t tidx, int64_t off, int64_t tlen, int ohang, int64_t minsc, int64_t *out6, uint8_t *editmask) {
	const int nceil = (int)(sc->nceil_const + sc->nceil_linear * (double)len);
	int ns = 0;
	const int64_t rfi = off, rff = off + len;
	int64_t leftNs = 0, rightNs = 0;
	for(int k = 0; k < 6; k++) out6[k] = 0;
	if(len <= 0) return 0;
	if(rfi < 0) { if(ohang) leftNs = -rfi; else return 0; }
	if(rff > tlen) { if(ohang) rightNs = rff - tlen; else return 0; }
	if(leftNs + rightNs > nceil) return 0;
	uint8_t *rf = (uint8_t *)malloc((size_t)len);
	bt2o_get_stretch(ix, tidx, off, len, rf);                   /* off-end positions come back as 4 (N) */
	for(int i = 0; i < len; i++) if((int64_t)i < leftNs || (int64_t)i >= (int64_t)len - rightNs) rf[i] = 4;
	int64_t score = 0;
	int rowi = 0, rowf = len - 1, rc = 1;
	const int monotone = sc->match_bonus == 0;                  /* Scoring::monotone (scoring.h) */
#define RD(i) (fw ? (int)codes[i] : ((int)codes[len - 1 - (i)] > 3 ? 4 : 3 - (int)codes[len - 1 - (i)]))
#define QU(i) ((int)(fw ? quals[i] : quals[len - 1 - (i)]) - 33)
	if(monotone) {
		for(int i = 0; i < len && rc == 1; i++) {
			int rdc = RD(i);
			if(rdc > 3 || rf[i] > 3) ns++;
			score += cell_score(sc, rdc, rf[i], QU(i));
			if(score < minsc || ns > nceil) rc = 0;
		}
	} else {
		int64_t scoreMax = 0;
		int lastfloor = 0, sols = 0;
		rowi = -1;
		for(int i = 0; i < len; i++) {
			int rdc = RD(i);
			if(rdc > 3 || rf[i] > 3) ns++;
			score += cell_score(sc, rdc, rf[i], QU(i));
			if(score >= minsc && score >= scoreMax) {
				scoreMax = score; rowf = i;
				if(rowi != lastfloor) { rowi = lastfloor; sols++; }
			}
			if(score <= 0) { score = 0; lastfloor = i + 1; }
		}
		if(ns > nceil || scoreMax < minsc) rc = 0;
		else if(sols > 1) rc = -1;
		score = scoreMax;
	}
	if(rc == 1) {
		int refns = 0, ned = 0;
		memset(editmask, 0, (size_t)len);
		for(int i = rowi; i <= rowf; i++) {
			if(rf[i] > 3 || RD(i) != rf[i]) { editmask[i] = 1; ned++; if(rf[i] > 3) refns++; }
		}
		out6[0] = score; out6[1] = rowi; out6[2] = rowf; out6[3] = ns; out6[4] = refns; out6[5] = ned;
	}
#undef RD
#undef QU
	free(rf);
	return rc;
}

typedef struct { int nedsz, celsz, row, col, gaps, score, ns, ct; } bt_frame;

/* Optional log of the backtrace attempts of the next bt2o_dp call, in the order SwAligner::nextAlignment makes them
 * (aligner_sw.cpp:757-1120): one pair per candidate that passed the start filter (and, local mode, the domination
 * filter) = per RNG reseed of the reference: [candidate score, index of the alignment it produced or -1, index of
 * the candidate in the sorted list].  Used by the
 * tests that replay the reference's sequential policy, where the per-read RNG state depends on the attempt count. */
static int64_t *g_attempt_log = NULL;
static int g_attempt_cap = 0, g_attempt_n = 0;
void bt2o_dp_attempt_log(int64_t *buf, int cap) { g_attempt_log = buf; g_attempt_cap = cap; g_attempt_n = 0; }
int bt2o_dp_attempt_count(void) { return g_attempt_n; }

/* One SwAligner session (aligner_sw.cpp:155-271 initRef, :500-729 align, :737-1146
 * nextAlignment).  End-to-end: fill recurrences aligner_swsse_ee_u8.cpp:931-993, gather :1176-1208,
 * backtrace :1283-1877.  Local: fill aligner_swsse_loc_i16.cpp:938-1367 (scores floored at 0 by
 * signed saturation), gather :1420-1535, backtrace :1615-2218 (moves only from cells with score
 * > 0), domination filter aligner_sw.cpp:946-971.  The backtrace is restated literally INCLUDING
 * its reported-through marks, remaining-option masks and branch stack, so that the device
 * kernel's shortcut can be checked against it.
 * Outputs use the same layout as oracle/ref_glue_dp.cpp:ref_dp(). */
int bt2o_dp(const bt2o_index *ix, const bt2o_scoring *sc, const uint8_t *codes, const uint8_t *quals, int len, int fw,
            uint64_t tidx, int64_t refl, int64_t refr, int triml, int corel, int corer, int64_t minsc, int nceil,
            int max_cands, int max_alns, int max_edits,
            int64_t *summary, int64_t *cands, int64_t *alns, int32_t *edits) {
	const int nrow = len, ncol = (int)(refr - refl + 1);
	const int local = sc->local;
	const int rdo = sc->rdgap_const + sc->rdgap_linear, rde = sc->rdgap_linear;
	const int rfo = sc->rfgap_const + sc->rfgap_linear, rfe = sc->rfgap_linear;
	const int NEGV = local ? 0 : DPNEG;          /* "minus infinity" of the mode */
	const int FLOOR = local ? 0 : DPNEG;         /* a source must be > FLOOR to be moved from */
	summary[0] = 0; summary[1] = DPNEG; summary[2] = 0; summary[3] = 0;
	if(ncol <= 0 || nrow <= 0) return -1;
	uint8_t *rd = (uint8_t *)malloc((size_t)nrow), *qu = (uint8_t *)malloc((size_t)nrow);
	uint8_t *rf = (uint8_t *)malloc((size_t)ncol + 2);
	for(int i = 0; i < nrow; i++) {
		if(fw) { rd[i] = codes[i]; qu[i] = quals[i]; }
		else { int c = codes[nrow - 1 - i]; rd[i] = (uint8_t)(c > 3 ? 4 : 3 - c); qu[i] = quals[nrow - 1 - i]; }
	}
	bt2o_get_stretch(ix, tidx, refl, ncol + 1, rf);   /* one extra character (aligner_sw.cpp:170-173) */
	size_t ncell = (size_t)nrow * (size_t)ncol;
	int *H = (int *)malloc(ncell * sizeof(int)), *E = (int *)malloc(ncell * sizeof(int)), *F = (int *)malloc(ncell * sizeof(int));
	uint16_t *mask = (uint16_t *)calloc(ncell, sizeof(uint16_t));
#define AT(M, i, j) M[(size_t)(i) * (size_t)ncol + (size_t)(j)]
#define MAX2(a, b) ((a) > (b) ? (a) : (b))
	int best = DPNEG;
	for(int j = 0; j < ncol; j++) {
		for(int i = 0; i < nrow; i++) {
			int bar = (i < sc->gapbar) || (nrow - 1 - i < sc->gapbar);
			int upH = i > 0 ? AT(H, i - 1, j) : NEGV, upF = i > 0 ? AT(F, i - 1, j) : NEGV;
			int f = bar ? NEGV : MAX2(upF - rfe, upH - rfo);
			int diag = (i == 0) ? 0 : (j == 0 ? NEGV : AT(H, i - 1, j - 1));
			int hd = diag + cell_score(sc, rd[i], rf[j], (int)qu[i] - 33);
			int lH = j > 0 ? AT(H, i, j - 1) : NEGV, lE = j > 0 ? AT(E, i, j - 1) : NEGV;
			int e = MAX2(lE - rde, bar ? NEGV : lH - rdo);
			if(local) { if(f < 0) f = 0; if(e < 0) e = 0; }
			int h = hd; if(e > h) h = e; if(f > h) h = f;
			if(local && h < 0) h = 0;
			AT(H, i, j) = h; AT(E, i, j) = e; AT(F, i, j) = f;
			if(local && h > best) best = h;
		}
	}
	if(!local) for(int j = 0; j < ncol; j++) if(AT(H, nrow - 1, j) > best) best = AT(H, nrow - 1, j);
	summary[1] = best;
	if(best < minsc) goto done;
	{
	/* gather + sort (score desc, row desc, col desc) */
	int ncand = 0;
	int *crow = (int *)malloc(ncell * sizeof(int)), *ccol = (int *)malloc(ncell * sizeof(int));
	if(!local) {
		for(int j = 0; j < ncol; j++) if(AT(H, nrow - 1, j) >= minsc) { crow[ncand] = nrow - 1; ccol[ncand++] = j; }
	} else {
		int bonus = sc->match_bonus;
		int minrow = (int)((minsc + bonus - 1) / bonus) - 1;
		for(int j = 0; j < ncol; j++)
			for(int i = (minrow > 0 ? minrow : 0); i < nrow; i++) {
				if(AT(H, i, j) < minsc) continue;
				int match = rd[i] == rf[j];                    /* mask test: N(16) & (1<<4) also "matches" */
				int match_succ = (i < nrow - 1) && rd[i + 1] == rf[j + 1];
				if(match && !match_succ) { crow[ncand] = i; ccol[ncand++] = j; }
			}
	}
	/* simple O(n^2) insertion sort on (score desc, row desc, col desc) */
	for(int a2 = 1; a2 < ncand; a2++) {
		int cr = crow[a2], cj = ccol[a2], cs = AT(H, cr, cj), b2 = a2 - 1;
		while(b2 >= 0) {
			int bs = AT(H, crow[b2], ccol[b2]);
			int less = bs < cs || (bs == cs && (crow[b2] < cr || (crow[b2] == cr && ccol[b2] < cj)));
			if(!less) break;
			crow[b2 + 1] = crow[b2]; ccol[b2 + 1] = ccol[b2]; b2--;
		}
		crow[b2 + 1] = cr; ccol[b2 + 1] = cj;
	}
	summary[0] = ncand > 0; summary[2] = ncand;
	for(int a2 = 0; a2 < ncand && a2 < max_cands; a2++) {
		cands[3 * a2] = crow[a2]; cands[3 * a2 + 1] = ccol[a2]; cands[3 * a2 + 2] = AT(H, crow[a2], ccol[a2]);
	}
	int naln = 0, ned_total = 0, ndone = 0;
	int32_t *ned = (int32_t *)malloc(4 * sizeof(int32_t) * (size_t)(nrow + ncol + 4));
	int *btcells = (int *)malloc(2 * sizeof(int) * (size_t)(nrow + ncol + 4));
	bt_frame *stack = (bt_frame *)malloc(sizeof(bt_frame) * (size_t)(nrow + ncol + 4));
	int *donerow = (int *)malloc(sizeof(int) * (size_t)(ncand + 1)), *donecol = (int *)malloc(sizeof(int) * (size_t)(ncand + 1));
	int SQ = nrow >> 4; if(SQ == 0) SQ = 1;
	static const char dna[] = "ACGTN";
	for(int ci = 0; ci < ncand; ci++) {
		int row = crow[ci], col = ccol[ci];
		if(AT(mask, row, col) & 1) continue;                       /* BT_CAND_FATE_FILT_START */
		if(local) {
			int dom = 0;
			for(int k = 0; k < ndone && !dom; k++) {
				int dr = donerow[k] - row, dc = donecol[k] - col;
				if(dr < 0) dr = -dr;
				if(dc < 0) dc = -dc;
				if(dr <= SQ && dc <= SQ) dom = 1;
			}
			if(dom) continue;
			donerow[ndone] = row; donecol[ndone] = col; ndone++;
		}
		const int start_row = row;
		const int attempt = g_attempt_n;
		if(g_attempt_log && g_attempt_n < g_attempt_cap) { g_attempt_log[3 * g_attempt_n] = AT(H, row, col); g_attempt_log[3 * g_attempt_n + 1] = -1; g_attempt_log[3 * g_attempt_n + 2] = ci; }
		g_attempt_n++;
		int nned = 0, ncells = 0, nstack = 0, gaps = 0, score = 0, ns = 0, ct = 0 /*0 H,1 E,2 F*/;
		int ok = 1, trim_beg = 0;
		for(;;) {
			int reported = AT(mask, row, col) & 1;
			int can_move = !reported, empty = 0, branch = 0, cur = -1;
			if(!reported && row > 0) {
				int gaps_allowed = !((row < sc->gapbar) || (nrow - 1 - row < sc->gapbar));
				uint16_t *mk = &AT(mask, row, col);
				if(ct == 1) {
					int sc_cur = AT(E, row, col), m = 0;
					if(AT(H, row, col - 1) > FLOOR && AT(H, row, col - 1) - rdo == sc_cur) m |= 1;
					if(AT(E, row, col - 1) > FLOOR && AT(E, row, col - 1) - rde == sc_cur) m |= 2;
					int orig = m;
					if(*mk & (1 << 7)) m = (*mk >> 8) & 3;
					if(m == 3) { cur = 4; *mk = (uint16_t)((*mk & ~(7 << 7)) | (1 << 7) | (2 << 8)); branch = 1; }
					else if(m == 2) { cur = 5; *mk = (uint16_t)((*mk & ~(7 << 7)) | (1 << 7)); }
					else if(m == 1) { cur = 4; *mk = (uint16_t)((*mk & ~(7 << 7)) | (1 << 7)); }
					else { empty = 1; can_move = (orig == 0); }
				} else if(ct == 2) {
					int sc_cur = AT(F, row, col), m = 0;
					if(AT(H, row - 1, col) > FLOOR && AT(H, row - 1, col) - rfo == sc_cur) m |= 1;
					if(AT(F, row - 1, col) > FLOOR && AT(F, row - 1, col) - rfe == sc_cur) m |= 2;
					int orig = m;
					if(*mk & (1 << 10)) m = (*mk >> 11) & 3;
					if(m == 3) { cur = 2; *mk = (uint16_t)((*mk & ~(7 << 10)) | (1 << 10) | (2 << 11)); branch = 1; }
					else if(m == 2) { cur = 3; *mk = (uint16_t)((*mk & ~(7 << 10)) | (1 << 10)); }
					else if(m == 1) { cur = 2; *mk = (uint16_t)((*mk & ~(7 << 10)) | (1 << 10)); }
					else { empty = 1; can_move = (orig == 0); }
				} else {
					int sc_cur = AT(H, row, col), m = 0;
					int sdiag = cell_score(sc, rd[row], rf[col], (int)qu[row] - 33);
					if(gaps_allowed) {
						if(AT(H, row - 1, col) > FLOOR && sc_cur == AT(H, row - 1, col) - rfo) m |= 1;
						if(col > 0 && AT(H, row, col - 1) > FLOOR && sc_cur == AT(H, row, col - 1) - rdo) m |= 2;
						if(AT(F, row - 1, col) > FLOOR && sc_cur == AT(F, row - 1, col) - rfe) m |= 4;
						if(col > 0 && AT(E, row, col - 1) > FLOOR && sc_cur == AT(E, row, col - 1) - rde) m |= 8;
					}
					if(col > 0 && AT(H, row - 1, col - 1) > FLOOR && sc_cur == AT(H, row - 1, col - 1) + sdiag) m |= 16;
					int orig = m;
					if(*mk & (1 << 1)) m = (*mk >> 2) & 31;
					int opts = 0, sel = -1;
					for(int b = 0; b < 5; b++) opts += (m >> b) & 1;
					if(opts >= 1) {
						if(m & 16) sel = 4; else if(m & 1) sel = 0; else if(m & 4) sel = 2; else if(m & 2) sel = 1; else sel = 3;
						int rem = opts > 1 ? (m & ~(1 << sel)) : 0;
						*mk = (uint16_t)((*mk & ~(31 << 1)) | (1 << 1) | (rem << 2));
						if(opts > 1) branch = 1;
						cur = sel == 4 ? 1 : (sel == 0 ? 2 : (sel == 1 ? 4 : (sel == 2 ? 3 : 5)));
					} else { empty = 1; can_move = (orig == 0); }
				}
			}
			AT(mask, row, col) |= 1;                                  /* setReportedThrough */
			if(!can_move) {
				if(nstack > 0) {
					bt_frame *fr = &stack[--nstack];
					ncells = fr->celsz; nned = fr->nedsz; row = fr->row; col = fr->col; gaps = fr->gaps;
					score = fr->score; ns = fr->ns; ct = fr->ct;
					continue;
				}
				ok = 0; break;
			}
			if(empty || row == 0) { btcells[2 * ncells] = row; btcells[2 * ncells + 1] = col; ncells++; trim_beg = row; break; }
			if(branch) {
				bt_frame *fr = &stack[nstack++];
				fr->nedsz = nned; fr->celsz = ncells; fr->row = row; fr->col = col; fr->gaps = gaps; fr->score = score; fr->ns = ns; fr->ct = ct;
			}
			btcells[2 * ncells] = row; btcells[2 * ncells + 1] = col; ncells++;
			int32_t *e4 = ned + 4 * nned;
			if(cur == 1) {
				int rdc = rd[row], rfc = rf[col];
				if(rdc > 3 || rfc > 3 || rdc != rfc) {
					e4[0] = row; e4[1] = dna[rfc]; e4[2] = dna[rdc]; e4[3] = 3; nned++;
					score -= (rdc > 3 || rfc > 3) ? sc->n_pen : mm_pen(sc, (int)qu[row] - 33);
				} else score += sc->match_bonus;
				if(rdc > 3 || rfc > 3) ns++;
				row--; col--; ct = 0;
			} else if(cur == 2 || cur == 3) {
				e4[0] = row; e4[1] = '-'; e4[2] = dna[rd[row]]; e4[3] = 2; nned++;
				score -= cur == 2 ? rfo : rfe; gaps++; row--; ct = cur == 2 ? 0 : 2;
			} else {
				e4[0] = row + 1; e4[1] = dna[rf[col]]; e4[2] = '-'; e4[3] = 1; nned++;
				score -= cur == 4 ? rdo : rde; gaps++; col--; ct = cur == 4 ? 0 : 1;
			}
		}
		if(!ok) continue;
		int core = 0;
		for(int k = 0; k < ncells; k++) {
			int d = btcells[2 * k + 1] - btcells[2 * k] + triml;
			if(d >= corel && d <= corer) { core = 1; break; }
		}
		if(!core) continue;
		{
			int rdc = rd[row], rfc = rf[col];
			int32_t *e4 = ned + 4 * nned;
			if(rdc > 3 || rfc > 3 || rdc != rfc) {
				e4[0] = row; e4[1] = dna[rfc]; e4[2] = dna[rdc]; e4[3] = 3; nned++;
				score -= (rdc > 3 || rfc > 3) ? sc->n_pen : mm_pen(sc, (int)qu[row] - 33);
			} else score += sc->match_bonus;
			if(rdc > 3 || rfc > 3) ns++;
		}
		if(ns > nceil) continue;
		if(g_attempt_log && attempt < g_attempt_cap) g_attempt_log[3 * attempt + 1] = naln;
		if(naln < max_alns) {
			const int trim_end = nrow - 1 - start_row;
			const int ext = nrow - trim_beg - trim_end;
			int64_t *o = alns + 8 * naln;
			o[0] = score; o[1] = ns; o[2] = gaps; o[3] = refl + col;
			o[4] = fw ? trim_beg : trim_end; o[5] = fw ? trim_end : trim_beg; o[6] = nned; o[7] = fw ? 1 : 0;
			/* SwResult::reverse; AlnRes::setShape shifts positions by the upstream trim (aligner_result.cpp:101-109);
			 * AlnRes::invertEdits for reverse-complement alignments (edit.cpp:50-78) */
			for(int k = 0; k < nned; k++) {
				const int32_t *src = fw ? ned + 4 * (nned - 1 - k) : ned + 4 * k;
				if(ned_total < max_edits) {
					int32_t *dst = edits + 4 * ned_total;
					dst[0] = src[0] - trim_beg; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
					if(!fw) dst[0] = ext - dst[0] - (src[3] == 1 ? 0 : 1);
				}
				ned_total++;
			}
		}
		naln++;
	}
	summary[3] = naln;
	free(ned); free(btcells); free(stack); free(donerow); free(donecol);
	free(crow); free(ccol);
	}
done:
	free(rd); free(qu); free(rf); free(H); free(E); free(F); free(mask);
	return 0;
#undef AT
#undef MAX2
}

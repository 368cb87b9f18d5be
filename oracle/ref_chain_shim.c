/*
 * oracle/ref_chain_shim.c — TEST INFRASTRUCTURE.  Calls the REFERENCE's own mem_chain2aln
 * (bwa-0.7.8/bwamem.c:730, compiled unmodified from /root/reference into oracle/_ref/libbwa_ref.so) on
 * flat arrays of reads / chains / seeds.  Contains no algorithm: it builds the reference's structs and loops.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "bwamem.h"     /* the reference header, -I$(REF) */

char *bwa_pg = 0;       /* defined in the reference's main.c:28, which is not part of libbwa */

/* mem_seed_t / mem_chain_t are private to bwamem.c (bwamem.c:168-177); same layout restated here */
typedef struct { int64_t rbeg; int32_t qbeg, len; } seed_t;
typedef struct { int n, m; int64_t pos; seed_t *seeds; } chain_t;
void mem_chain2aln(const mem_opt_t *opt, int64_t l_pac, const uint8_t *pac, int l_query, const uint8_t *query,
                   const chain_t *c, mem_alnreg_v *av);

typedef struct { int a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop; int8_t mat[25]; } flat_opt_t;

int ref_chain2aln_flat(const flat_opt_t *fo, int64_t l_pac, const uint8_t *pac, int n_reads, const int64_t *read_off,
                       const int32_t *read_len, const uint8_t *qpool, int n_chains, const int32_t *chain_read,
                       const int64_t *chain_seed0, const int32_t *chain_nseeds, const seed_t *seeds, int64_t out_cap,
                       mem_alnreg_t *out, int32_t *out_read, int64_t *n_out)
{
	mem_opt_t *opt = mem_opt_init();
	int r, c = 0;
	int64_t n = 0;
	opt->a = fo->a; opt->b = fo->b; opt->o_del = fo->o_del; opt->e_del = fo->e_del; opt->o_ins = fo->o_ins; opt->e_ins = fo->e_ins;
	opt->pen_clip5 = fo->pen_clip5; opt->pen_clip3 = fo->pen_clip3; opt->w = fo->w; opt->zdrop = fo->zdrop;
	memcpy(opt->mat, fo->mat, 25);
	for (r = 0; r < n_reads; ++r) {
		mem_alnreg_v av;
		size_t k;
		av.n = av.m = 0; av.a = 0;
		for (; c < n_chains && chain_read[c] == r; ++c) {
			chain_t ch;
			ch.n = ch.m = chain_nseeds[c]; ch.pos = 0; ch.seeds = (seed_t *)(seeds + chain_seed0[c]);
			mem_chain2aln(opt, l_pac, pac, read_len[r], qpool + read_off[r], &ch, &av);
		}
		for (k = 0; k < av.n; ++k) {
			if (n >= out_cap) { free(av.a); free(opt); return -1; }
			out[n] = av.a[k]; out_read[n] = r; ++n;
		}
		free(av.a);
	}
	*n_out = n;
	free(opt);
	return 0;
}

/* test_capi.c — the host-side C API (include/gsb200_gsc.h) used from plain C, the way
 * r-wrappers.c would: no Python, no C++.  Exit status 0 = every invariant held.
 *   gcc -std=c11 -Wall -Wextra -Werror -pedantic -Iinclude tests/c/test_capi.c \
 *       -Lgenomicsimulation_b200 -lgsb200_gsc -lgsb200 -Wl,-rpath,$PWD/genomicsimulation_b200 -lm */
#include "gsb200_gsc.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

static unsigned long long lcg = 88172645463325252ull;
static double urand(void) { lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17; return ((lcg >> 12) + 0.5) / 4503599627370496.0; }

int main(void) {
    enum { F = 16, M = 2000, C = 4 };
    gsb_SimData* d = gsb_create_simdata(0, M);
    if (!d) { fprintf(stderr, "no CUDA device: %s\n", gsb200_last_error()); return 77; }
    gsb_set_print(d, NULL);

    /* founders, a 4-chromosome map of 100 cM each, one effect set */
    char* chars = malloc((size_t) F * 2 * M);
    for (size_t i = 0; i < (size_t) F * 2 * M; ++i) chars[i] = urand() < 0.5 ? 'A' : 'T';
    gsb_GroupNum founders = gsb_load_genotypes(d, F, chars, NULL);
    CHECK(founders == 1 && gsb_get_group_size(d, founders) == F);
    double lambda[C]; unsigned off[C + 1]; double* dists = malloc(sizeof(double) * M); unsigned* mix = malloc(sizeof(unsigned) * M);
    unsigned char has[C];
    for (int c = 0; c < C; ++c) { lambda[c] = 1.0; off[c] = c * (M / C); has[c] = 1; }
    off[C] = M;
    for (unsigned m = 0; m < M; ++m) { mix[m] = m; dists[m] = (double) (m % (M / C)) / (M / C - 1); }
    gsb_MapID map = gsb_load_map(d, C, lambda, off, dists, mix, has);
    CHECK(map == 1);
    unsigned* cumn = malloc(sizeof(unsigned) * M); char* allele = malloc(2 * M); double* eff = malloc(sizeof(double) * 2 * M);
    for (unsigned m = 0; m < M; ++m) { cumn[m] = 2 * (m + 1); allele[2 * m] = 'A'; allele[2 * m + 1] = 'T'; eff[2 * m] = urand() - 0.5; eff[2 * m + 1] = urand() - 0.5; }
    gsb_EffectID e = gsb_load_effects(d, cumn, allele, eff, NULL);
    CHECK(e == 1);

    /* a breeding scheme in native draw mode */
    gsb_set_draw_mode(d, GSB_DRAWS_NATIVE, 42);
    gsb_GenOptions g = GSB_BASIC_OPT;
    g.will_track_pedigree = 1;
    g.family_size = 2;
    gsb_GroupNum f1 = gsb_make_random_crosses(d, founders, 500, 0, 0 /* GSC_NO_MAP: first map */, g);
    CHECK(f1 == 2 && gsb_get_group_size(d, f1) == 1000 && gsb_last_status(d) == GSB200_OK);
    g.family_size = 1;
    gsb_GroupNum best = gsb_split_by_bv(d, f1, e, 100, 0);
    CHECK(best == 3 && gsb_get_group_size(d, best) == 100 && gsb_get_group_size(d, f1) == 900);
    gsb_GroupNum dh = gsb_make_doubled_haploids(d, best, map, g);
    CHECK(dh == 4 && gsb_get_group_size(d, dh) == 100);

    /* selection picked the top: min(best) >= max(rest) */
    gsb_DecimalMatrix b1 = gsb_calculate_bvs(d, best, e), b2 = gsb_calculate_bvs(d, f1, e);
    CHECK(b1.dim2 == 100 && b2.dim2 == 900);
    double lo = b1.data[0], hi = b2.data[0];
    for (unsigned i = 0; i < b1.dim2; ++i) if (b1.data[i] < lo) lo = b1.data[i];
    for (unsigned i = 0; i < b2.dim2; ++i) if (b2.data[i] > hi) hi = b2.data[i];
    CHECK(lo >= hi);
    gsb_delete_dmatrix(&b1); gsb_delete_dmatrix(&b2);

    /* doubled haploids are homozygous mosaics of their parent */
    char* g_dh = malloc((size_t) 100 * 2 * M); char* g_par = malloc((size_t) 100 * 2 * M);
    unsigned ped1[100], ids[100];
    CHECK(gsb_get_group_data(d, dh, 100, g_dh, NULL, ped1, NULL, NULL, NULL) == 100);
    CHECK(gsb_get_group_data(d, best, 100, g_par, ids, NULL, NULL, NULL, NULL) == 100);
    for (unsigned i = 0; i < 100; ++i) {
        CHECK(ped1[i] == ids[i]);
        for (unsigned m = 0; m < M; ++m) {
            const char a = g_dh[(size_t) i * 2 * M + 2 * m], b = g_dh[(size_t) i * 2 * M + 2 * m + 1];
            CHECK(a == b);
            CHECK(a == g_par[(size_t) i * 2 * M + 2 * m] || a == g_par[(size_t) i * 2 * M + 2 * m + 1]);
        }
    }
    /* local GEBVs by 5 blocks per chromosome sum to the GEBV */
    gsb_MarkerBlocks blocks = gsb_create_evenlength_blocks_each_chr(d, map, 5);
    CHECK(blocks.num_blocks == 20);
    gsb_DecimalMatrix loc = gsb_calculate_local_bvs(d, dh, blocks, e), bv = gsb_calculate_bvs(d, dh, e);
    CHECK(loc.dim1 == 200 && loc.dim2 == 20);
    for (unsigned i = 0; i < 100; ++i) {
        double s = 0;
        for (unsigned b = 0; b < 20; ++b) s += loc.data[(size_t) (2 * i) * 20 + b] + loc.data[(size_t) (2 * i + 1) * 20 + b];
        CHECK(s - bv.data[i] < 1e-9 && bv.data[i] - s < 1e-9);
    }
    gsb_delete_dmatrix(&loc); gsb_delete_dmatrix(&bv); gsb_delete_markerblocks(&blocks);
    gsb_delete_group(d, f1);
    CHECK(gsb_get_n_genotypes(d) == F + 200);
    printf("capi ok: %u genotypes, current id %u\n", gsb_get_n_genotypes(d), gsb_get_current_id(d));
    gsb_delete_simdata(d);
    free(chars); free(dists); free(mix); free(cumn); free(allele); free(eff); free(g_dh); free(g_par);
    return 0;
}

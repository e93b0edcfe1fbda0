/* drive.c — the e2e loop of bench.py (gsb_shard_select_and_cross + gsb_shard_get_local_bvs + gsb_delete_group per generation)
 * against the stub engine: per-call host time of the mirror, nothing else. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "gsb200_gsc.h"

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }

int main(int argc, char** argv) {
    const unsigned n = argc > 1 ? atoi(argv[1]) : 100000, founders = 1000, M = 32, top = n / 100, steps = argc > 2 ? atoi(argv[2]) : 30;
    gsb_SimData* d = gsb_create_simdata(0, M);
    gsb_set_draw_mode(d, GSB_DRAWS_NATIVE, 1);
    char* chars = malloc((size_t) n * 2 * M);
    memset(chars, 'A', (size_t) n * 2 * M);
    gsb_load_genotypes(d, founders, chars, NULL);
    gsb_GroupNum g = gsb_load_genotypes(d, n, chars, NULL);
    double lam = 1.0, dists[32]; unsigned off[2] = {0, M}, idx[32]; unsigned char has[32];
    for (unsigned i = 0; i < M; ++i) { dists[i] = i / (double) M; idx[i] = i; has[i] = 1; }
    gsb_MapID map = gsb_load_map(d, 1, &lam, off, dists, idx, has);
    unsigned cum[33]; char al[32]; double eff[32];
    for (unsigned i = 0; i <= M; ++i) cum[i] = i;
    for (unsigned i = 0; i < M; ++i) { al[i] = 'A'; eff[i] = 0.1; }
    gsb_EffectID e = gsb_load_effects(d, cum, al, eff, NULL);
    char handle[128];
    if (gsb_shard_init(d, 0, 1, n, top) || gsb_shard_handle(d, handle) || gsb_shard_connect(d, handle)) { puts("shard set-up failed"); return 1; }
    gsb_GenOptions opt = GSB_BASIC_OPT;
    double* bv = malloc(8 * (size_t) n);
    double ta = 0, tb = 0, tc = 0;
    for (unsigned s = 0; s < steps + 3; ++s) {
        if (s == 3) ta = tb = tc = 0;
        const double t0 = now();
        gsb_GroupNum ng = gsb_shard_select_and_cross(d, g, n, e, top, 0, n, map, opt);
        const double t1 = now();
        gsb_shard_get_local_bvs(d, n, bv);
        const double t2 = now();
        gsb_delete_group(d, g);
        const double t3 = now();
        if (!ng) { puts("select_and_cross failed"); return 1; }
        g = ng;
        ta += t1 - t0; tb += t2 - t1; tc += t3 - t2;
    }
    printf("n = %u: select_and_cross %.3f ms, local_bvs %.3f ms, delete_group %.3f ms, total %.3f ms per generation (host only)\n",
           n, ta / steps, tb / steps, tc / steps, (ta + tb + tc) / steps);
    /* the single-GPU calls a user of the reference makes: gsb_make_random_crosses from 1000 parents + gsb_get_group_bvs + gsb_delete_group */
    {
        gsb_GenOptions o2 = GSB_BASIC_OPT;
        o2.will_track_pedigree = 1;
        double tx = 0, ty = 0, tz = 0;
        for (unsigned s = 0; s < steps + 3; ++s) {
            if (s == 3) tx = ty = tz = 0;
            const double t0 = now();
            gsb_GroupNum ng = gsb_make_random_crosses(d, 1, n, 0, map, o2);
            const double t1 = now();
            gsb_get_group_bvs(d, ng, e, n, bv);
            const double t2 = now();
            gsb_delete_group(d, ng);
            const double t3 = now();
            if (!ng) { puts("make_random_crosses failed"); return 1; }
            tx += t1 - t0; ty += t2 - t1; tz += t3 - t2;
        }
        printf("n = %u: make_random_crosses %.3f ms, get_group_bvs %.3f ms, delete_group %.3f ms, total %.3f ms per batch (host only)\n",
               n, tx / steps, ty / steps, tz / steps, (tx + ty + tz) / steps);
    }
#ifdef GSB_HOST_PROFILE
    { extern void gsb_host_profile_report(unsigned); gsb_host_profile_report(steps + 3); }
#endif
    gsb_delete_simdata(d);
    return 0;
}

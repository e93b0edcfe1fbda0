/* stub_engine.c — a do-nothing stand-in for libgsb200.so, for PROFILING THE HOST MIRROR on a box without a GPU
 * (profiles/scripts/host_profile/run.sh) and for MODEL-CHECKING ITS BOOKKEEPING on CPU (tests/test_host_bookkeeping.py).  Not a product path and not a test oracle: every engine call succeeds at once,
 * so what the clock sees is the host mirror's own bookkeeping per generation (group scans, row lists, ids, pedigrees).
 * Only the calls the sharded generation makes return data; the rest are declared without prototypes on purpose. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t n_markers; } ctx_t;
typedef struct { uint64_t max_total, n_total; uint32_t max_pool, top_n, n_new; uint32_t *k1, *k2, *ids; double* bv; } shard_t;

int gsb200_create(ctx_t** out, int dev, uint32_t m) { (void) dev; *out = calloc(1, sizeof(ctx_t)); (*out)->n_markers = m; return 0; }
void gsb200_destroy(ctx_t* c) { free(c); }
const char* gsb200_last_error(void) { return "stub"; }
uint32_t gsb200_store_planes(const ctx_t* c) { (void) c; return 1; }
void gsb200_shard_range(uint64_t n_total, uint32_t rank, uint32_t world, uint64_t* lo, uint64_t* hi) {
    const uint64_t q = n_total / world, r = n_total % world;
    *lo = rank * q + (rank < r ? rank : r);
    *hi = *lo + q + (rank < r ? 1 : 0);
}
int gsb200_shard_create(ctx_t* c, uint32_t rank, uint32_t world, uint64_t max_total, uint32_t max_pool, shard_t** out) {
    (void) c; (void) rank; (void) world;
    shard_t* s = calloc(1, sizeof *s);
    s->max_total = max_total; s->max_pool = max_pool;
    s->k1 = malloc(4 * max_total); s->k2 = malloc(4 * max_total); s->ids = malloc(4 * (size_t) max_pool); s->bv = malloc(8 * max_total);
    for (uint64_t i = 0; i < max_total; ++i) { s->k1[i] = (uint32_t) ((i * 2654435761u) % max_pool); s->k2[i] = (uint32_t) ((i * 40503u + 7) % max_pool); s->bv[i] = (double) i; }
    for (uint32_t i = 0; i < max_pool; ++i) s->ids[i] = i + 1;
    *out = s;
    return 0;
}
void gsb200_shard_destroy(shard_t* s) { if (s) { free(s->k1); free(s->k2); free(s->ids); free(s->bv); free(s); } }
int gsb200_shard_handle(shard_t* s, void* h) { (void) s; memset(h, 0, 128); return 0; }
int gsb200_shard_connect(shard_t* s, const void* h) { (void) s; (void) h; return 0; }
int gsb200_shard_select_run(shard_t* s, uint32_t n_local, const uint32_t* rows, uint32_t first, const uint32_t* ids, uint64_t n_total,
                            uint32_t eff, uint32_t top_n, int low) {
    (void) n_local; (void) rows; (void) first; (void) ids; (void) eff; (void) low;
    s->n_total = n_total; s->top_n = top_n;
    return 0;
}
int gsb200_shard_cross_begin_run(shard_t* s, uint32_t n, uint32_t map, const uint32_t* out_rows, uint32_t first, const void* draws, int want) {
    (void) map; (void) out_rows; (void) first; (void) draws; (void) want;
    s->n_new = n;
    return 0;
}
int gsb200_shard_cross_picks(shard_t* s, const uint32_t** k1, const uint32_t** k2, const uint32_t** ids) { *k1 = s->k1; *k2 = s->k2; *ids = s->ids; return 0; }
int gsb200_shard_poll(shard_t* s) { (void) s; return 0; }
int gsb200_shard_finish(shard_t* s) { (void) s; return 0; }
int gsb200_shard_local_bvs(shard_t* s, uint32_t n_local, double* out) { memcpy(out, s->bv, 8 * (size_t) n_local); return 0; }
uint64_t gsb200_shard_generation_size(const shard_t* s) { return s->n_total; }
uint32_t gsb200_shard_pool_size(const shard_t* s) { return s->top_n; }
int gsb200_shard_selected() { return 0; }

int gsb200_cross_random_pairs_begin(ctx_t* c, uint32_t n_crosses, uint32_t family, const uint32_t* rows1, uint32_t n1, const uint32_t* rows2, uint32_t n2,
                                    uint32_t map1, uint32_t map2, const uint32_t* out_rows, uint32_t first_out, const void* draws, uint32_t* k1, uint32_t* k2) {
    (void) c; (void) rows1; (void) rows2; (void) map1; (void) map2; (void) out_rows; (void) first_out; (void) draws;
    const uint64_t n = (uint64_t) n_crosses * family;
    const uint32_t m2 = rows2 ? n2 : n1;
    if (k1 && k2) {          /* cheap on purpose (no division per offspring): what is timed is the mirror, not the stub */
        uint32_t a = 0, b = 7 % m2;
        for (uint64_t i = 0; i < n; ++i) {
            k1[i] = a; k2[i] = b;
            if (family == 1 || (i + 1) % family == 0) { a += 3; if (a >= n1) a -= n1 < 3 ? a : n1; b += 5; if (b >= m2) b -= m2 < 5 ? b : m2; }
        }
    }
    return 0;
}

#define STUB(name) int name() { return 0; }
STUB(gsb200_allele_counts) STUB(gsb200_allele_presence) STUB(gsb200_clone) STUB(gsb200_cross) STUB(gsb200_cross_random_pairs_end) STUB(gsb200_doubled_haploid) STUB(gsb200_effects_delete) STUB(gsb200_effects_set) STUB(gsb200_gebv)
STUB(gsb200_local_gebv) STUB(gsb200_map_delete) STUB(gsb200_map_set) STUB(gsb200_self_n_times) STUB(gsb200_store_dictionary)
STUB(gsb200_store_download_chars) STUB(gsb200_store_download_chars_range) STUB(gsb200_store_reserve) STUB(gsb200_store_upload_chars)
STUB(gsb200_top_n)

/* oracle/gs_oracle.c — TEST INFRASTRUCTURE ONLY. See gs_oracle.h.
 *
 * Sequential restatement of the reference hot path. Genotypes are kept exactly as
 * the reference keeps them: one char[2*M] per genotype, alleles of marker m at
 * [2m] and [2m+1] (core.h:778-783). Storage order is the order of `rows`.
 */
#include "gs_oracle.h"
#include "draws.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    unsigned n_chr;
    double* lambda;
    unsigned* chr_offset;
    double* dists;
    unsigned* marker_index;
    unsigned char* has_dists;
} Map;

typedef struct {
    unsigned* cumn;
    char* allele;
    double* eff;
    double* centre;   /* NULL if none */
} Effects;

typedef struct {
    char* alleles;
    char* name;
    unsigned id, ped1, ped2, group;
} Row;

struct gso_Sim {
    unsigned M;
    Row* rows; unsigned n, cap;
    Map* maps; unsigned n_maps;
    Effects* effs; unsigned n_effs;
    unsigned current_id;
};

/* ------------------------------------------------------------------ store */

gso_Sim* gso_new(unsigned n_markers) {
    gso_Sim* s = calloc(1, sizeof *s);
    s->M = n_markers;
    return s;
}

static void row_free(Row* r) { free(r->alleles); free(r->name); }

void gso_free(gso_Sim* s) {
    if (!s) return;
    for (unsigned i = 0; i < s->n; ++i) row_free(&s->rows[i]);
    free(s->rows);
    for (unsigned i = 0; i < s->n_maps; ++i) {
        free(s->maps[i].lambda); free(s->maps[i].chr_offset); free(s->maps[i].dists);
        free(s->maps[i].marker_index); free(s->maps[i].has_dists);
    }
    free(s->maps);
    for (unsigned i = 0; i < s->n_effs; ++i) {
        free(s->effs[i].cumn); free(s->effs[i].allele); free(s->effs[i].eff); free(s->effs[i].centre);
    }
    free(s->effs);
    free(s);
}

static Row* push_row(gso_Sim* s) {
    if (s->n == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 64;
        s->rows = realloc(s->rows, sizeof(Row) * s->cap);
    }
    Row* r = &s->rows[s->n++];
    memset(r, 0, sizeof *r);
    /* offspring slots start zeroed: '\0' = unknown allele (core.c:75) */
    r->alleles = calloc(2 * (size_t) s->M + 1, 1);
    return r;
}

/* smallest positive integer not used by any current group (core.c:3937-3957) */
static unsigned new_group_num(const gso_Sim* s) {
    unsigned gn = 1;
    for (;;) {
        int used = 0;
        for (unsigned i = 0; i < s->n; ++i) if (s->rows[i].group == gn) { used = 1; break; }
        if (!used) return gn;
        ++gn;
    }
}

unsigned gso_add_genotypes(gso_Sim* s, unsigned n, const char* chars, const char* const* names) {
    if (n == 0) return 0;
    unsigned group = new_group_num(s);
    for (unsigned i = 0; i < n; ++i) {
        Row* r = push_row(s);
        memcpy(r->alleles, chars + 2 * (size_t) s->M * i, 2 * (size_t) s->M);
        if (names && names[i]) r->name = strdup(names[i]);
        r->id = ++s->current_id;
        r->group = group;
    }
    return group;
}

static void* dup_mem(const void* p, size_t bytes) {
    void* q = malloc(bytes ? bytes : 1);
    if (p && bytes) memcpy(q, p, bytes);
    return q;
}

unsigned gso_add_map(gso_Sim* s, unsigned n_chr, const double* lambda, const unsigned* chr_offset,
                     const double* dists, const unsigned* marker_index, const unsigned char* has_dists) {
    s->maps = realloc(s->maps, sizeof(Map) * (s->n_maps + 1));
    Map* m = &s->maps[s->n_maps];
    unsigned total = chr_offset[n_chr];
    m->n_chr = n_chr;
    m->lambda = dup_mem(lambda, sizeof(double) * n_chr);
    m->chr_offset = dup_mem(chr_offset, sizeof(unsigned) * (n_chr + 1));
    m->dists = dup_mem(dists, sizeof(double) * total);
    m->marker_index = dup_mem(marker_index, sizeof(unsigned) * total);
    m->has_dists = dup_mem(has_dists, n_chr);
    return s->n_maps++;
}

unsigned gso_add_effects(gso_Sim* s, const unsigned* cumn, const char* allele, const double* eff,
                         const double* centre) {
    s->effs = realloc(s->effs, sizeof(Effects) * (s->n_effs + 1));
    Effects* e = &s->effs[s->n_effs];
    unsigned nnz = s->M ? cumn[s->M - 1] : 0;
    e->cumn = dup_mem(cumn, sizeof(unsigned) * s->M);
    e->allele = dup_mem(allele, nnz);
    e->eff = dup_mem(eff, sizeof(double) * nnz);
    e->centre = centre ? dup_mem(centre, sizeof(double) * s->M) : NULL;
    return s->n_effs++;
}

unsigned gso_n_genotypes(const gso_Sim* s) { return s->n; }
unsigned gso_current_id(const gso_Sim* s) { return s->current_id; }

/* core.c:4190-4207: group 0 "is not a group so it does not have a size" */
unsigned gso_group_size(const gso_Sim* s, unsigned group) {
    if (group == 0) return 0;
    unsigned c = 0;
    for (unsigned i = 0; i < s->n; ++i) c += (s->rows[i].group == group);
    return c;
}

static int in_group(const Row* r, unsigned group) { return group == 0 || r->group == group; }

unsigned gso_export_group(const gso_Sim* s, unsigned group, unsigned cap, char* genotypes,
                          unsigned* ids, unsigned* ped1, unsigned* ped2, unsigned* groups,
                          unsigned* global_indexes) {
    unsigned k = 0;
    size_t rowlen = 2 * (size_t) s->M;
    for (unsigned i = 0; i < s->n && k < cap; ++i) {
        const Row* r = &s->rows[i];
        if (!in_group(r, group)) continue;
        if (genotypes) memcpy(genotypes + rowlen * k, r->alleles, rowlen);
        if (ids) ids[k] = r->id;
        if (ped1) ped1[k] = r->ped1;
        if (ped2) ped2[k] = r->ped2;
        if (groups) groups[k] = r->group;
        if (global_indexes) global_indexes[k] = i;
        ++k;
    }
    return k;
}

const char* gso_member_name(const gso_Sim* s, unsigned group, unsigned i) {
    for (unsigned r = 0; r < s->n; ++r) {
        if (!in_group(&s->rows[r], group)) continue;
        if (i == 0) return s->rows[r].name;
        --i;
    }
    return NULL;
}

/* indexes (storage positions) of the members of a group, in storage order */
static unsigned* group_members(const gso_Sim* s, unsigned group, unsigned* n_out) {
    unsigned* ix = malloc(sizeof(unsigned) * (s->n ? s->n : 1));
    unsigned k = 0;
    for (unsigned i = 0; i < s->n; ++i) if (in_group(&s->rows[i], group)) ix[k++] = i;
    *n_out = k;
    return ix;
}

/* ---------------------------------------------------------------- meiosis */

static int cmp_double_asc(const void* a, const void* b) {   /* core.c:1324-1332 */
    double x = *(const double*) a, y = *(const double*) b;
    return x < y ? -1 : (x > y);
}

/* One gamete (core.c:7850-7928) or, with `doubled`, a doubled haploid
 * (core.c:7949-8033). Per chromosome, in this order: Rf_rpois(lambda) crossovers
 * (7873), that many unif_rand() positions (7888-7890), qsort (7891-7893), one
 * unif_rand() for the starting strand (7896), then the marker walk that flips
 * strand while dists[i] > crossover (7900-7909 simple, 7911-7921 reorder). */
static void gamete(const gso_Sim* s, const char* parent, char* out, unsigned map_index, int doubled) {
    if (parent == NULL || map_index >= s->n_maps) return;
    const Map* map = &s->maps[map_index];
    unsigned cap = 100;
    double* where = malloc(sizeof(double) * cap);
    for (unsigned chr = 0; chr < map->n_chr; ++chr) {
        int num = (int) gsdraw_rpois(map->lambda[chr]);
        if ((unsigned) num > cap) { cap = (unsigned) num; where = realloc(where, sizeof(double) * cap); }
        for (int i = 0; i < num; ++i) where[i] = gsdraw_unif();
        if (num > 1) qsort(where, (size_t) num, sizeof(double), cmp_double_asc);

        int which = (gsdraw_unif() > 0.5);
        int upto = 0;
        unsigned lo = map->chr_offset[chr], hi = map->chr_offset[chr + 1];
        for (unsigned p = lo; p < hi; ++p) {
            /* unlinked maps have no dists and expected_n_crossovers 0, so num == 0
             * and dists is never read (core.c:6128-6152) */
            while (upto < num && map->has_dists[chr] && map->dists[p] > where[upto]) {
                which = 1 - which;
                ++upto;
            }
            unsigned pos = map->marker_index[p];
            out[2 * pos] = parent[2 * pos + which];
            if (doubled) out[2 * pos + 1] = out[2 * pos];
        }
    }
    free(where);
}

void gso_generate_gamete(const gso_Sim* s, const char* parent, char* out, unsigned map_index) {
    gamete(s, parent, out, map_index, 0);
}
void gso_generate_doubled_haploid(const gso_Sim* s, const char* parent, char* out, unsigned map_index) {
    gamete(s, parent, out, map_index, 1);
}
void gso_generate_clone(const gso_Sim* s, const char* parent, char* out) {   /* core.c:8047-8055 */
    for (unsigned j = 0; j < s->M; ++j) { out[2 * j] = parent[2 * j]; out[2 * j + 1] = parent[2 * j + 1]; }
}

/* --------------------------------------------------------------- scaffold */

enum { GEN_CROSS, GEN_SELF, GEN_DH, GEN_CLONE };

typedef struct {
    int kind;
    unsigned map1, map2;
    unsigned n_gens;        /* selfing */
    int inherit_names;      /* clones */
} Generator;

/* core.c:8900-8926: n rounds of two gametes, ping-ponging between a temporary
 * and the output slot so that the last round lands in the output. */
static void self_n(const gso_Sim* s, const char* parent, char* output, unsigned map, unsigned n) {
    char* tmpchild = calloc(2 * (size_t) s->M + 1, 1);
    const char* tmpparent = parent;
    unsigned odd = n % 2;
    for (unsigned i = 0; i < n; ++i) {
        char* dst = (i % 2 == odd) ? tmpchild : output;
        gamete(s, tmpparent, dst, map, 0);
        gamete(s, tmpparent, dst + 1, map, 0);
        tmpparent = dst;
    }
    free(tmpchild);
}

/* The generic crossing loop (core.c:8258-8353). `pairs` lists the parent pairs
 * (storage indexes) in the order the parentChooser yields them; RNG draws made
 * by the chooser itself happen in `next_pair`, interleaved exactly as at 8301. */
typedef int (*PairFn)(gso_Sim*, void* ctx, unsigned* counter, unsigned pair[2]);

static unsigned scaffold(gso_Sim* s, gso_GenOptions g, PairFn next_pair, void* ctx, Generator gen) {
    if (g.family_size < 1) return 0;
    unsigned out_group = g.retain ? new_group_num(s) : 0;   /* 8285-8291 */
    unsigned first_new = s->n;
    unsigned counter = 0, made = 0;
    unsigned pair[2];
    /* names are issued per 1000-genotype buffer from d->current_id at flush time
     * (8168-8180, 8307-8309, 8334-8335; gsc_set_names core.c:552-591) */
    unsigned block_start = first_new;

    gsdraw_mark(GSDRAW_G);                                   /* 8299 */
    while (next_pair(s, ctx, &counter, pair)) {              /* 8301 */
        ++counter;                                           /* 8302 */
        for (unsigned f = 0; f < g.family_size; ++f) {
            if (s->n - block_start >= 1000) {                /* buffer full: 8307-8320 */
                for (unsigned j = block_start; j < s->n; ++j) {
                    Row* r = &s->rows[j];
                    if (g.name_offspring) {
                        char buf[512];
                        snprintf(buf, sizeof buf, "%s%u", g.name_prefix ? g.name_prefix : "",
                                 s->current_id + 1 + (j - block_start));
                        free(r->name); r->name = strdup(buf);
                    }
                }
                if (g.allocate_ids) for (unsigned j = block_start; j < s->n; ++j) s->rows[j].id = ++s->current_id;
                block_start = s->n;
            }
            Row* child = push_row(s);
            const Row* p0 = &s->rows[pair[0]];
            const Row* p1 = &s->rows[pair[1]];
            switch (gen.kind) {
            case GEN_CROSS:                                  /* 8416-8423 */
                gamete(s, p0->alleles, child->alleles, gen.map1, 0);
                gamete(s, p1->alleles, child->alleles + 1, gen.map2, 0);
                break;
            case GEN_SELF: self_n(s, p0->alleles, child->alleles, gen.map1, gen.n_gens); break;
            case GEN_DH:   gamete(s, p0->alleles, child->alleles, gen.map1, 1); break;   /* 8998-9006 */
            case GEN_CLONE:                                  /* 9097-9108 */
                if (gen.inherit_names && p0->name) child->name = strdup(p0->name);
                gso_generate_clone(s, p0->alleles, child->alleles);
                break;
            }
            child->group = out_group;                        /* 8325 */
            if (g.track_pedigree) { child->ped1 = p0->id; child->ped2 = p1->id; }   /* 8326-8329 */
            ++made;
        }
    }
    gsdraw_mark(GSDRAW_E);                                   /* 8332 */

    for (unsigned j = block_start; j < s->n; ++j) {          /* tail block: 8334-8335 */
        if (g.name_offspring) {
            char buf[512];
            snprintf(buf, sizeof buf, "%s%u", g.name_prefix ? g.name_prefix : "",
                     s->current_id + 1 + (j - block_start));
            free(s->rows[j].name); s->rows[j].name = strdup(buf);
        }
    }
    if (g.allocate_ids) for (unsigned j = block_start; j < s->n; ++j) s->rows[j].id = ++s->current_id;

    if (counter > 0 && g.retain) return out_group;           /* 8344-8348 */
    for (unsigned j = first_new; j < s->n; ++j) row_free(&s->rows[j]);   /* 8349-8351 */
    s->n = first_new;
    (void) made;
    return 0;
}

/* core.c:8556-8581 */
static unsigned randomdraw_replacementrules(unsigned max, unsigned cap, unsigned* uses, unsigned no_collision) {
    if (max < 1 || (max == 1 && no_collision == 0)) return (unsigned) -1;
    unsigned ix = 0;
    if (cap > 0) {
        do { ix = (unsigned) round(gsdraw_unif() * (max - 1)); } while (ix == no_collision || uses[ix] >= cap);
    } else {
        do { ix = (unsigned) round(gsdraw_unif() * (max - 1)); } while (ix == no_collision);
    }
    return ix;
}

typedef struct {
    unsigned n_crosses;
    unsigned *members1, n1, cap1, *uses1;
    unsigned *members2, n2, cap2, *uses2;
    const unsigned *first, *second;   /* targeted */
    unsigned bad;
    unsigned cursor;                  /* selfing: next member */
} Chooser;

static int pair_random(gso_Sim* s, void* ctx, unsigned* pcounter, unsigned pair[2]) {   /* core.c:8364-8402 */
    Chooser* c = ctx; (void) s; unsigned counter = *pcounter;
    if (counter < c->n_crosses && (c->cap1 == 0 || counter < c->cap1 * c->n1)) {
        unsigned a = randomdraw_replacementrules(c->n1, c->cap1, c->uses1, (unsigned) -1);
        unsigned b = randomdraw_replacementrules(c->n1, c->cap1, c->uses1, a);
        if (c->cap1 > 0) { c->uses1[a] += 1; c->uses1[b] += 1; }
        pair[0] = c->members1[a]; pair[1] = c->members1[b];
        return 1;
    }
    return 0;
}

static int pair_random_between(gso_Sim* s, void* ctx, unsigned* pcounter, unsigned pair[2]) {   /* core.c:8594-8631 */
    Chooser* c = ctx; (void) s; unsigned counter = *pcounter;
    if (counter < c->n_crosses && (c->cap1 == 0 || counter < c->cap1 * c->n1) &&
        (c->cap2 == 0 || counter < c->cap2 * c->n2)) {
        unsigned a = randomdraw_replacementrules(c->n1, c->cap1, c->uses1, (unsigned) -1);
        unsigned b = randomdraw_replacementrules(c->n2, c->cap2, c->uses2, (unsigned) -1);
        if (c->cap1 > 0) c->uses1[a] += 1;
        if (c->cap2 > 0) c->uses2[b] += 1;
        pair[0] = c->members1[a]; pair[1] = c->members2[b];
        return 1;
    }
    return 0;
}

static int pair_targeted(gso_Sim* s, void* ctx, unsigned* counter, unsigned pair[2]) {   /* core.c:8751-8776 */
    Chooser* c = ctx;
    while (*counter < c->n_crosses) {
        unsigned a = c->first[*counter], b = c->second[*counter];
        if (a != (unsigned) -1 && b != (unsigned) -1 && a < s->n && b < s->n) {
            pair[0] = a; pair[1] = b;
            return 1;
        }
        ++c->bad;      /* invalid pair: skipped and counted (8771-8773) */
        ++(*counter);
    }
    return 0;
}

static int pair_selfing(gso_Sim* s, void* ctx, unsigned* counter, unsigned pair[2]) {   /* core.c:8876-8887, 9070-9087 */
    Chooser* c = ctx; (void) s; (void) counter;
    if (c->cursor >= c->n1) return 0;
    pair[0] = pair[1] = c->members1[c->cursor++];
    return 1;
}

/* core.c:8430-8456 */
static unsigned random_cross_checks(const gso_Sim* s, unsigned group, unsigned n_crosses, unsigned cap) {
    unsigned g_size = gso_group_size(s, group);
    if (g_size == 0) return 0;
    if (n_crosses < 1) return 0;
    if (cap > 0 && cap * g_size < n_crosses) return 0;
    return g_size;
}

unsigned gso_make_random_crosses(gso_Sim* s, unsigned group, unsigned n_crosses, unsigned cap,
                                 unsigned map_index, gso_GenOptions g) {   /* core.c:8483-8532 */
    unsigned g_size = random_cross_checks(s, group, n_crosses * 2, cap);
    if (g_size < 2 || s->n_maps < 1 || map_index >= s->n_maps) return 0;
    Chooser c = {0};
    c.n_crosses = n_crosses; c.cap1 = cap;
    c.members1 = group_members(s, group, &c.n1);
    c.uses1 = cap > 0 ? calloc(g_size, sizeof(unsigned)) : NULL;
    Generator gen = { GEN_CROSS, map_index, map_index, 0, 0 };
    unsigned out = scaffold(s, g, pair_random, &c, gen);
    free(c.members1); free(c.uses1);
    return out;
}

unsigned gso_make_random_crosses_between(gso_Sim* s, unsigned group1, unsigned group2, unsigned n_crosses,
                                         unsigned cap1, unsigned cap2, unsigned map1, unsigned map2,
                                         gso_GenOptions g) {   /* core.c:8669-8740 */
    unsigned s1 = random_cross_checks(s, group1, n_crosses, cap1);
    unsigned s2 = random_cross_checks(s, group2, n_crosses, cap2);
    if (s1 == 0 || s2 == 0 || s->n_maps < 1 || map1 >= s->n_maps || map2 >= s->n_maps) return 0;
    Chooser c = {0};
    c.n_crosses = n_crosses; c.cap1 = cap1; c.cap2 = cap2;
    c.members1 = group_members(s, group1, &c.n1);
    c.members2 = group_members(s, group2, &c.n2);
    c.uses1 = cap1 > 0 ? calloc(s1, sizeof(unsigned)) : NULL;
    c.uses2 = cap2 > 0 ? calloc(s2, sizeof(unsigned)) : NULL;
    Generator gen = { GEN_CROSS, map1, map2, 0, 0 };
    unsigned out = scaffold(s, g, pair_random_between, &c, gen);
    free(c.members1); free(c.members2); free(c.uses1); free(c.uses2);
    return out;
}

unsigned gso_make_targeted_crosses(gso_Sim* s, unsigned n, const unsigned* first, const unsigned* second,
                                   unsigned map1, unsigned map2, gso_GenOptions g) {   /* core.c:8810-8865 */
    if (n < 1 || s->n_maps < 1 || map1 >= s->n_maps || map2 >= s->n_maps) return 0;
    Chooser c = {0};
    c.n_crosses = n; c.first = first; c.second = second;
    Generator gen = { GEN_CROSS, map1, map2, 0, 0 };
    unsigned out = scaffold(s, g, pair_targeted, &c, gen);
    if (n - c.bad == 0) return 0;
    return out;
}

unsigned gso_make_all_unidirectional_crosses(gso_Sim* s, unsigned group, unsigned map_index, gso_GenOptions g) {
    /* core.c:9173-9212: upper triangle of (member i, member j), i < j, as a targeted cross */
    unsigned n; unsigned* members = group_members(s, group, &n);
    if (group == 0 || n < 2) { free(members); return 0; }
    unsigned n_crosses = n * (n - 1) / 2, k = 0;
    unsigned* a = malloc(sizeof(unsigned) * n_crosses);
    unsigned* b = malloc(sizeof(unsigned) * n_crosses);
    for (unsigned i = 0; i < n; ++i) for (unsigned j = i + 1; j < n; ++j) { a[k] = members[i]; b[k] = members[j]; ++k; }
    unsigned out = gso_make_targeted_crosses(s, n_crosses, a, b, map_index, map_index, g);
    free(a); free(b); free(members);
    return out;
}

static unsigned one_parent_op(gso_Sim* s, unsigned group, Generator gen, gso_GenOptions g) {
    Chooser c = {0};
    c.members1 = group_members(s, group, &c.n1);
    unsigned out = scaffold(s, g, pair_selfing, &c, gen);
    free(c.members1);
    return out;
}

unsigned gso_self_n_times(gso_Sim* s, unsigned n, unsigned group, unsigned map_index, gso_GenOptions g) {
    if (n < 1 || s->n_maps == 0 || map_index >= s->n_maps) return 0;    /* core.c:8951-8988 */
    Generator gen = { GEN_SELF, map_index, map_index, n, 0 };
    return one_parent_op(s, group, gen, g);
}
unsigned gso_make_doubled_haploids(gso_Sim* s, unsigned group, unsigned map_index, gso_GenOptions g) {
    if (s->n_maps == 0 || map_index >= s->n_maps) return 0;              /* core.c:9027-9057 */
    Generator gen = { GEN_DH, map_index, map_index, 0, 0 };
    return one_parent_op(s, group, gen, g);
}
unsigned gso_make_clones(gso_Sim* s, unsigned group, int inherit_names, gso_GenOptions g) {
    Generator gen = { GEN_CLONE, 0, 0, 0, inherit_names };               /* core.c:9135-9154 */
    return one_parent_op(s, group, gen, g);
}

/* ------------------------------------------------------------- evaluation */

/* core.c:9566-9618: per genotype, markers in order, every effect entry of the
 * marker whose allele matches either strand contributes count*eff; then the sum
 * of centres (if any) is subtracted once. */
unsigned gso_calculate_bvs(const gso_Sim* s, unsigned group, unsigned eff_index, double* out) {
    if (eff_index >= s->n_effs) return 0;
    const Effects* e = &s->effs[eff_index];
    unsigned k = 0;
    for (unsigned i = 0; i < s->n; ++i) {
        if (!in_group(&s->rows[i], group)) continue;
        const char* g = s->rows[i].alleles;
        double sum = 0;
        for (unsigned m = 0; m < s->M; ++m) {
            double msum = 0.;
            for (unsigned eix = (m > 0 ? e->cumn[m - 1] : 0); eix < e->cumn[m]; ++eix) {
                double asum = ((e->allele[eix] == g[2 * m]) + (e->allele[eix] == g[2 * m + 1])) * e->eff[eix];
                msum += asum;
            }
            sum += msum;
        }
        out[k++] = sum;
    }
    if (e->centre != NULL) {
        double summed = 0.;
        for (unsigned m = 0; m < s->M; ++m) summed += e->centre[m];
        for (unsigned i = 0; i < k; ++i) out[i] -= summed;
    }
    return k;
}

/* core.c:9636-9667 */
unsigned gso_calculate_allele_counts(const gso_Sim* s, unsigned group, char allele, double* out) {
    unsigned k = 0;
    for (unsigned i = 0; i < s->n; ++i) {
        if (!in_group(&s->rows[i], group)) continue;
        const char* g = s->rows[i].alleles;
        for (unsigned m = 0; m < s->M; ++m)
            out[(size_t) k * s->M + m] = (g[2 * m] == allele) + (g[2 * m + 1] == allele);
        ++k;
    }
    return k;
}

/* core.c:9979-10090: per haplotype and block, the FIRST effect entry matching the
 * allele contributes (gotallele flags); with centres, the whole centre[m] is
 * subtracted from each haplotype (10075-10076). */
unsigned gso_calculate_local_bvs(const gso_Sim* s, unsigned group, unsigned n_blocks,
                                 const unsigned* offsets, const unsigned* markers,
                                 unsigned eff_index, double* out) {
    if (eff_index >= s->n_effs) return 0;
    const Effects* e = &s->effs[eff_index];
    unsigned k = 0;
    for (unsigned i = 0; i < s->n; ++i) {
        if (!in_group(&s->rows[i], group)) continue;
        const char* g = s->rows[i].alleles;
        double* h1 = out + (size_t) (2 * k) * n_blocks;
        double* h2 = out + (size_t) (2 * k + 1) * n_blocks;
        for (unsigned j = 0; j < n_blocks; ++j) {
            h1[j] = 0.; h2[j] = 0.;
            for (unsigned q = offsets[j]; q < offsets[j + 1]; ++q) {
                unsigned m = markers[q];
                int got1 = 0, got2 = 0;
                for (unsigned eix = (m > 0 ? e->cumn[m - 1] : 0); eix < e->cumn[m]; ++eix) {
                    if (!got1 && e->allele[eix] == g[2 * m]) { h1[j] += e->eff[eix]; got1 = 1; }
                    if (!got2 && e->allele[eix] == g[2 * m + 1]) { h2[j] += e->eff[eix]; got2 = 1; }
                }
                if (e->centre != NULL) { h1[j] -= e->centre[m]; h2[j] -= e->centre[m]; }
            }
        }
        ++k;
    }
    return 2 * k;
}

/* core.c:9697-9827: the first marker of a chromosome goes to its first block
 * "for floating point reasons" (9744-9746); marker m moves to the next block
 * while dists[m] > (blocks so far)/(float)n (9751-9752, float division). */
unsigned gso_create_evenlength_blocks_each_chr(const gso_Sim* s, unsigned map_index, unsigned n,
                                               unsigned* offsets, unsigned* markers) {
    if (s->n_maps < 1 || n < 1) return 0;
    if (map_index >= s->n_maps) map_index = 0;
    const Map* map = &s->maps[map_index];
    unsigned nb = n * map->n_chr, t = 0;
    for (unsigned chr = 0; chr < map->n_chr; ++chr) {
        unsigned lo = map->chr_offset[chr], len = map->chr_offset[chr + 1] - lo;
        unsigned first = chr * n, blk = first;
        if (len == 0) { for (unsigned b = 0; b < n; ++b) offsets[first + b] = t; continue; }
        offsets[blk] = t;
        markers[t++] = map->marker_index[lo];
        for (unsigned m = 1; m < len; ++m) {
            while (blk - first < n - 1 && map->dists[lo + m] > (blk - first + 1) / (float) n) {
                ++blk;
                offsets[blk] = t;
            }
            markers[t++] = map->marker_index[lo + m];
        }
        while (blk - first < n - 1) { ++blk; offsets[blk] = t; }
    }
    offsets[nb] = t;
    return nb;
}

unsigned gso_make_group_from(gso_Sim* s, unsigned n, const unsigned* ix) {   /* core.c:2810-2841 */
    if (n < 1) return 0;
    unsigned ng = new_group_num(s), ok = 0;
    for (unsigned i = 0; i < n; ++i) if (ix[i] < s->n) { s->rows[ix[i]].group = ng; ++ok; }
    return ok ? ng : 0;
}

typedef struct { double v; unsigned ix; } Ranked;
static int cmp_desc(const void* a, const void* b) {      /* core.c:1308-1316 */
    double x = ((const Ranked*) a)->v, y = ((const Ranked*) b)->v;
    return x > y ? -1 : (x < y);
}
static int cmp_asc(const void* a, const void* b) {       /* core.c:1341-1349 */
    double x = ((const Ranked*) a)->v, y = ((const Ranked*) b)->v;
    return x < y ? -1 : (x > y);
}

/* core.c:9463-9517. Order among equal breeding values is "undefined"
 * (core.c:1305-1306); callers must compare selections modulo ties at the cut. */
unsigned gso_split_by_bv(gso_Sim* s, unsigned group, unsigned eff_index, unsigned top_n, int low_is_best) {
    if (eff_index >= s->n_effs) return 0;
    unsigned n = gso_group_size(s, group);
    if (n == 0) return 0;
    unsigned cnt; unsigned* members = group_members(s, group, &cnt);
    unsigned out;
    if (n <= top_n) {
        out = gso_make_group_from(s, n, members);
    } else {
        double* bv = malloc(sizeof(double) * n);
        gso_calculate_bvs(s, group, eff_index, bv);
        Ranked* r = malloc(sizeof(Ranked) * n);
        for (unsigned i = 0; i < n; ++i) { r[i].v = bv[i]; r[i].ix = members[i]; }
        qsort(r, n, sizeof(Ranked), low_is_best ? cmp_asc : cmp_desc);
        unsigned* top = malloc(sizeof(unsigned) * (top_n ? top_n : 1));
        for (unsigned i = 0; i < top_n; ++i) top[i] = r[i].ix;
        out = gso_make_group_from(s, top_n, top);
        free(top); free(r); free(bv);
    }
    free(members);
    return out;
}

void gso_delete_group(gso_Sim* s, unsigned group) {      /* core.c:4658-4692 + condense 1546-1601 */
    unsigned k = 0;
    for (unsigned i = 0; i < s->n; ++i) {
        if (s->rows[i].group == group) { row_free(&s->rows[i]); continue; }
        s->rows[k++] = s->rows[i];
    }
    s->n = k;
}

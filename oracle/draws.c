/* oracle/draws.c — TEST INFRASTRUCTURE ONLY. See draws.h. */
#include "draws.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static uint64_t g_s[4] = { 0x9E3779B97F4A7C15ull, 0xBF58476D1CE4E5B9ull,
                           0x94D049BB133111EBull, 0x2545F4914F6CDD1Dull };

static int      g_recording = 0;
static double*  g_rec_val = NULL;
static uint8_t* g_rec_kind = NULL;
static size_t   g_rec_len = 0, g_rec_cap = 0;

static int            g_replaying = 0;
static const double*  g_rp_val = NULL;
static const uint8_t* g_rp_kind = NULL;
static size_t         g_rp_len = 0, g_rp_pos = 0;
static int            g_rp_error = 0;

static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static uint64_t next_u64(void) {            /* xoshiro256++ */
    const uint64_t result = rotl(g_s[0] + g_s[3], 23) + g_s[0];
    const uint64_t t = g_s[1] << 17;
    g_s[2] ^= g_s[0]; g_s[3] ^= g_s[1]; g_s[1] ^= g_s[2]; g_s[0] ^= g_s[3];
    g_s[2] ^= t; g_s[3] = rotl(g_s[3], 45);
    return result;
}

void gsdraw_seed(uint64_t seed) {           /* splitmix64 expansion */
    for (int i = 0; i < 4; ++i) {
        seed += 0x9E3779B97F4A7C15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        g_s[i] = z ^ (z >> 31);
    }
}

static double raw_unif(void) {              /* 53-bit, open interval (0,1) */
    uint64_t x;
    do { x = next_u64() >> 11; } while (x == 0);
    return (double) x * 0x1.0p-53;
}

static void rec_push(int kind, double v) {
    if (!g_recording) return;
    if (g_rec_len == g_rec_cap) {
        size_t ncap = g_rec_cap ? 2 * g_rec_cap : 4096;
        g_rec_val = (double*) realloc(g_rec_val, ncap * sizeof(double));
        g_rec_kind = (uint8_t*) realloc(g_rec_kind, ncap);
        g_rec_cap = ncap;
    }
    g_rec_val[g_rec_len] = v;
    g_rec_kind[g_rec_len] = (uint8_t) kind;
    ++g_rec_len;
}

/* pop the next U/P entry from the replay tape, skipping G/E markers */
static double replay_pop(int kind) {
    while (g_rp_pos < g_rp_len &&
           (g_rp_kind[g_rp_pos] == GSDRAW_G || g_rp_kind[g_rp_pos] == GSDRAW_E)) {
        ++g_rp_pos;
    }
    if (g_rp_pos >= g_rp_len || g_rp_kind[g_rp_pos] != kind) {
        g_rp_error = 1;
        return kind == GSDRAW_U ? 0.25 : 0.0;
    }
    return g_rp_val[g_rp_pos++];
}

double gsdraw_unif(void) {
    double u = g_replaying ? replay_pop(GSDRAW_U) : raw_unif();
    rec_push(GSDRAW_U, u);
    return u;
}

/* Exact Poisson by sequential-search inversion; mu is split into pieces of at
 * most 16 so exp(-mu) never underflows (a sum of independent Poissons is Poisson). */
static double pois_inv(double mu) {
    double u = raw_unif();
    double p = exp(-mu), F = p;
    unsigned k = 0;
    while (u > F && k < 100000u) { ++k; p *= mu / k; F += p; }
    return (double) k;
}

double gsdraw_rpois(double mu) {
    double k = 0.0;
    if (g_replaying) {
        k = replay_pop(GSDRAW_P);
    } else if (mu > 0.0) {
        double left = mu;
        while (left > 0.0) {
            double m = left > 16.0 ? 16.0 : left;
            left -= m;
            k += pois_inv(m);
        }
    }
    rec_push(GSDRAW_P, k);
    return k;
}

void gsdraw_mark(int kind) { rec_push(kind, 0.0); }

void gsdraw_record(int on) {
    g_recording = on ? 1 : 0;
    if (on) g_rec_len = 0;
}
size_t gsdraw_tape_len(void) { return g_rec_len; }
void gsdraw_tape_copy(double* values, uint8_t* kinds) {
    if (g_rec_len == 0) return;
    memcpy(values, g_rec_val, g_rec_len * sizeof(double));
    memcpy(kinds, g_rec_kind, g_rec_len);
}

void gsdraw_replay_start(const double* values, const uint8_t* kinds, size_t n) {
    g_rp_val = values; g_rp_kind = kinds; g_rp_len = n; g_rp_pos = 0;
    g_rp_error = 0; g_replaying = 1;
}
size_t gsdraw_replay_stop(void) {
    /* swallow trailing markers so "consumed == len" means the tape was used up */
    while (g_replaying && g_rp_pos < g_rp_len &&
           (g_rp_kind[g_rp_pos] == GSDRAW_G || g_rp_kind[g_rp_pos] == GSDRAW_E)) {
        ++g_rp_pos;
    }
    g_replaying = 0;
    return g_rp_pos;
}
int gsdraw_replay_error(void) { return g_rp_error; }

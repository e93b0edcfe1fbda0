/* oracle/draws.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The "draw source" shared by both oracles:
 *   - oracle/_ref  (the unmodified reference core behind the R shim): unif_rand()
 *     and Rf_rpois() forward here (oracle/shim/shim_impl.c);
 *   - oracle/gs_oracle.c (the plain-C restatement).
 *
 * It replaces libR's RNG (Mersenne-Twister + nmath rpois — module "R", version
 * not pinned by the reference's DESCRIPTION; not present in /root/reference).
 * Three modes:
 *   free-running : xoshiro256++ uniforms in the open interval (0,1) and an
 *                  exact inversion Poisson sampler;
 *   recording    : additionally logs every draw the caller sees, in call order,
 *                  as (kind, value) pairs — the "tape";
 *   replaying    : serves draws from a caller-supplied tape instead of the RNG.
 *
 * Tape grammar for the crossing path (verified against
 * sim-operations.c:8299-8332, 8556-8581, 7867-7896):
 *   G  { pick draws (U...)  { per family member { per gamete { per chr:
 *        P k, U x k, U strand } } } }*  E
 */
#ifndef GSB200_ORACLE_DRAWS_H
#define GSB200_ORACLE_DRAWS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    GSDRAW_U = 1,  /* a unif_rand() result; value = u                      */
    GSDRAW_P = 2,  /* an Rf_rpois() result; value = k (as double)          */
    GSDRAW_G = 3,  /* GetRNGstate() marker; value = 0                      */
    GSDRAW_E = 4   /* PutRNGstate() marker; value = 0                      */
};

void   gsdraw_seed(uint64_t seed);
double gsdraw_unif(void);
double gsdraw_rpois(double mu);
void   gsdraw_mark(int kind);            /* G / E markers */

/* recording */
void   gsdraw_record(int on);            /* on=1 starts (and clears) a recording */
size_t gsdraw_tape_len(void);
void   gsdraw_tape_copy(double* values, uint8_t* kinds);

/* replaying: the arrays must stay alive until gsdraw_replay_stop(). */
void   gsdraw_replay_start(const double* values, const uint8_t* kinds, size_t n);
size_t gsdraw_replay_stop(void);         /* returns number of entries consumed */
int    gsdraw_replay_error(void);        /* nonzero if the grammar went out of step */

#ifdef __cplusplus
}
#endif
#endif

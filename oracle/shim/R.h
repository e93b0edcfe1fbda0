/* oracle/shim/R.h — TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for the seven libR symbols that the reference core
 * (/root/reference/src/sim-operations.c, pulled in through sim-operations.h:75-78)
 * needs: Rprintf, error (-> Rf_error), unif_rand, Rf_rpois, GetRNGstate,
 * PutRNGstate, R_CheckUserInterrupt.  R itself is not installed in this image.
 * Rinternals.h, Rmath.h and R_ext/Utils.h in this directory all forward here.
 */
#ifndef GSB200_ORACLE_SHIM_R_H
#define GSB200_ORACLE_SHIM_R_H

#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <math.h>

void   Rprintf(const char* fmt, ...);
void   Rf_error(const char* fmt, ...) __attribute__((noreturn));
#define error Rf_error

double unif_rand(void);        /* open interval (0,1) */
double Rf_rpois(double mu);    /* exact sampler; 0 for mu <= 0 */
void   GetRNGstate(void);
void   PutRNGstate(void);
void   R_CheckUserInterrupt(void);

#endif

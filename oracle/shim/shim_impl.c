/* oracle/shim/shim_impl.c — TEST INFRASTRUCTURE ONLY.
 * Implements the libR symbols declared in shim/R.h on top of oracle/draws.c.
 */
#include <R.h>
#include <string.h>
#include "../draws.h"

/* 0 = discard, 1 = stdout, 2 = capture into a buffer readable by tests */
static int    g_print_mode = 0;
static char*  g_cap = NULL;
static size_t g_cap_len = 0, g_cap_cap = 0;

void shim_set_print_mode(int mode) { g_print_mode = mode; g_cap_len = 0; }
size_t shim_captured_len(void) { return g_cap_len; }
void shim_captured_copy(char* out) { if (g_cap_len) memcpy(out, g_cap, g_cap_len); }

void Rprintf(const char* fmt, ...) {
    va_list ap;
    if (g_print_mode == 0) return;
    va_start(ap, fmt);
    if (g_print_mode == 1) {
        vprintf(fmt, ap);
    } else {
        char tmp[4096];
        int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
        if (n > 0) {
            if ((size_t) n >= sizeof tmp) n = sizeof tmp - 1;
            if (g_cap_len + (size_t) n > g_cap_cap) {
                g_cap_cap = 2 * (g_cap_len + (size_t) n) + 1024;
                g_cap = (char*) realloc(g_cap, g_cap_cap);
            }
            memcpy(g_cap + g_cap_len, tmp, (size_t) n);
            g_cap_len += (size_t) n;
        }
    }
    va_end(ap);
}

void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[oracle shim] R error(): ");
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    abort();
}

double unif_rand(void) { return gsdraw_unif(); }
double Rf_rpois(double mu) { return gsdraw_rpois(mu); }
void GetRNGstate(void) { gsdraw_mark(GSDRAW_G); }
void PutRNGstate(void) { gsdraw_mark(GSDRAW_E); }
void R_CheckUserInterrupt(void) { }

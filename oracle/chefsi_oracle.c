/*
 * chefsi_oracle.c -- TEST INFRASTRUCTURE ONLY (see chefsi_oracle.h).
 *
 * Plain-C restatement of the reference's CheFSI filter path.  Used as the checker;
 * never linked into the product.  Parity pin: tests/test_oracle_vs_reference.py and
 * tests/golden/ (vectors produced by the compiled reference).
 */
#include "chefsi_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* ---- real instantiation ---- */
#define T double
#define ORACLE_COMPLEX 0
#define FN(name) CAT(name, _d)
#include "chefsi_oracle_impl.inc"
#undef T
#undef ORACLE_COMPLEX
#undef FN

/* ---- complex instantiation ---- */
#define T double _Complex
#define ORACLE_COMPLEX 1
#define FN(name) CAT(name, _z)
#include "chefsi_oracle_impl.inc"
#undef T
#undef ORACLE_COMPLEX
#undef FN

void oracle_lap_plus_diag(const chefsi_grid_t *g, int ncol, double a, double b, double c,
                          const double *v, const double *x, size_t ldi, double *y, size_t ldo)
{
    lap_plus_diag_d(g, NULL, ncol, a, b, c, v, x, ldi, y, ldo);
}
void oracle_lap_plus_diag_kpt(const chefsi_grid_t *g, const double kvec[3], int ncol, double a,
                              double b, double c, const double *v, const double _Complex *x,
                              size_t ldi, double _Complex *y, size_t ldo)
{
    lap_plus_diag_z(g, kvec, ncol, a, b, c, v, x, ldi, y, ldo);
}
void oracle_vnl_mult(const chefsi_grid_t *g, const chefsi_nloc_t *nl, int ncol, const double *x,
                     size_t ldi, double *Hx, size_t ldo)
{
    vnl_mult_d(g, nl, NULL, ncol, x, ldi, Hx, ldo);
}
void oracle_vnl_mult_kpt(const chefsi_grid_t *g, const chefsi_nloc_t *nl, const double kvec[3],
                         int ncol, const double _Complex *x, size_t ldi, double _Complex *Hx,
                         size_t ldo)
{
    vnl_mult_z(g, nl, kvec, ncol, x, ldi, Hx, ldo);
}
void oracle_hamiltonian_mult(const chefsi_grid_t *g, const chefsi_nloc_t *nl, const double *veff,
                             int ncol, double c, const double *x, size_t ldi, double *Hx,
                             size_t ldo)
{
    hamiltonian_mult_d(g, nl, veff, NULL, ncol, c, x, ldi, Hx, ldo);
}
void oracle_hamiltonian_mult_kpt(const chefsi_grid_t *g, const chefsi_nloc_t *nl,
                                 const double *veff, const double kvec[3], int ncol, double c,
                                 const double _Complex *x, size_t ldi, double _Complex *Hx,
                                 size_t ldo)
{
    hamiltonian_mult_z(g, nl, veff, kvec, ncol, c, x, ldi, Hx, ldo);
}
void oracle_chebyshev_filter(const chefsi_grid_t *g, const chefsi_nloc_t *nl, const double *veff,
                             double *X, size_t ldi, double *Y, size_t ldo, int ncol, int m,
                             double a, double b, double a0)
{
    chebyshev_filter_d(g, nl, veff, NULL, X, ldi, Y, ldo, ncol, m, a, b, a0);
}
void oracle_chebyshev_filter_kpt(const chefsi_grid_t *g, const chefsi_nloc_t *nl,
                                 const double *veff, const double kvec[3], double _Complex *X,
                                 size_t ldi, double _Complex *Y, size_t ldo, int ncol, int m,
                                 double a, double b, double a0)
{
    chebyshev_filter_z(g, nl, veff, kvec, X, ldi, Y, ldo, ncol, m, a, b, a0);
}

/* splitmix64 finaliser as the counter-based generator */
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
double oracle_random_value(unsigned long long seed, long long col, long long idx)
{
    uint64_t h = mix64(mix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(col + 1)) + (uint64_t)idx);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}
void oracle_fill_random(double *buf, size_t n_per_col, size_t ld, int ncol, long long first_col,
                        unsigned long long seed)
{
#pragma omp parallel for
    for (int n = 0; n < ncol; n++)
        for (size_t i = 0; i < n_per_col; i++)
            buf[(size_t)n * ld + i] = oracle_random_value(seed, first_col + n, (long long)i);
}

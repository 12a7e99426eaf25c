/*
 * chefsi_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the reference's Chebyshev-filter path, used solely as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The product (sparc_b200/, libchefsi_b200.so) never includes, links or calls this.
 *
 * Parity pin: every function here is checked in tests/test_oracle_vs_reference.py
 * against the UNMODIFIED reference routines compiled into oracle/_ref/ (all cell types,
 * both boundary conditions, real and complex), and against the vectors committed under
 * tests/golden/ that were produced by those reference routines
 * (tests/golden/make_golden.py).  The compiled reference itself reproduces the
 * reference's own Si8 / BaTiO3 .refout energies at np=1 (DESIGN.md "Oracle").
 *
 * The problem description structs are the product's public ones (include/chefsi_b200.h).
 */
#ifndef CHEFSI_ORACLE_H
#define CHEFSI_ORACLE_H

#include <complex.h>
#include "../include/chefsi_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* y = (a*Lap + b*diag(v) + c) x for ncol columns; v may be NULL (then b := 0).
 * Follows Lap_plus_diag_vec_mult_orth (lapVecRoutines.c:306-611) for cell_typ 0 and
 * Lap_plus_diag_vec_mult_nonorth (lapVecRoutines.c:940-1331) for cell_typ 11..17. */
void oracle_lap_plus_diag(const chefsi_grid_t *g, int ncol, double a, double b, double c,
                          const double *v, const double *x, size_t ldi, double *y, size_t ldo);
/* complex variant with Bloch-phase halos: lapVecRoutinesKpt.c:179-513, :567-991 */
void oracle_lap_plus_diag_kpt(const chefsi_grid_t *g, const double kvec[3], int ncol, double a,
                              double b, double c, const double *v, const double _Complex *x,
                              size_t ldi, double _Complex *y, size_t ldo);

/* Hx += Vnl x : nlocVecRoutines.c:798-883 (real), :889-999 (complex) */
void oracle_vnl_mult(const chefsi_grid_t *g, const chefsi_nloc_t *nl, int ncol, const double *x,
                     size_t ldi, double *Hx, size_t ldo);
void oracle_vnl_mult_kpt(const chefsi_grid_t *g, const chefsi_nloc_t *nl, const double kvec[3],
                         int ncol, const double _Complex *x, size_t ldi, double _Complex *Hx,
                         size_t ldo);

/* Hx = (-1/2 Lap + Veff + c) x + Vnl x : hamiltonianVecRoutines.c:45-121, :132-242 */
void oracle_hamiltonian_mult(const chefsi_grid_t *g, const chefsi_nloc_t *nl, const double *veff,
                             int ncol, double c, const double *x, size_t ldi, double *Hx,
                             size_t ldo);
void oracle_hamiltonian_mult_kpt(const chefsi_grid_t *g, const chefsi_nloc_t *nl,
                                 const double *veff, const double kvec[3], int ncol, double c,
                                 const double _Complex *x, size_t ldi, double _Complex *Hx,
                                 size_t ldo);

/* eigenSolver.c:722-798 / eigenSolverKpt.c:458-535.  X in/out, Y out. */
void oracle_chebyshev_filter(const chefsi_grid_t *g, const chefsi_nloc_t *nl, const double *veff,
                             double *X, size_t ldi, double *Y, size_t ldo, int ncol, int m,
                             double a, double b, double a0);
void oracle_chebyshev_filter_kpt(const chefsi_grid_t *g, const chefsi_nloc_t *nl,
                                 const double *veff, const double kvec[3], double _Complex *X,
                                 size_t ldi, double _Complex *Y, size_t ldo, int ncol, int m,
                                 double a, double b, double a0);

/* Counter-based U(-0.5,0.5) start vectors (Init_orbital,
 * orbitalElecDensInit.c:388-392 draws from the same interval).  n_per_col = Nd for real
 * data, 2*Nd for complex (re,im interleaved). */
double oracle_random_value(unsigned long long seed, long long col, long long idx);
void oracle_fill_random(double *buf, size_t n_per_col, size_t ld, int ncol, long long first_col,
                        unsigned long long seed);

#ifdef __cplusplus
}
#endif
#endif

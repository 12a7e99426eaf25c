/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference routines (compiled into oracle/_ref/libsparc_ref.so
 * from /root/reference/src) on a flat problem description, so that the C restatement
 * (chefsi_oracle.c) and the CUDA path can be compared with the reference itself on
 * arbitrary synthetic inputs.  Compiled against the reference's own headers; contains no
 * reference code.  Built only where /root/reference exists; the resulting
 * oracle/_ref/libref_harness.so travels to the GPU box.
 *
 * A SPARC_OBJ is filled with exactly the fields the path reads:
 *   isddft.h:391 cell_typ, :396-398 Nx.., :401-403 range_*, :411 dV, :464 order,
 *   :467-481 stencil tables, :667-669 BC*, :540 Veff_loc_dmcomm, :452-460 projector
 *   tables, :661-663 k*_loc, :375 Nspinor_eig, :367 spin_start_indx, :355 bandcomm_index.
 * Each real atom becomes its own "type" with a single l = 0 channel of nproj radial
 * functions, which reproduces any per-projector Gamma list through the reference's
 * scaling loop (nlocVecRoutines.c:841-863).
 */
#include <complex.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mpi.h>

#include "isddft.h"
#include "eigenSolver.h"
#include "eigenSolverKpt.h"
#include "hamiltonianVecRoutines.h"
#include "lapVecRoutines.h"
#include "lapVecRoutinesKpt.h"
#include "nlocVecRoutines.h"
#include "linearSolver.h"

#include "chefsi_oracle.h"

typedef struct {
    SPARC_OBJ S;
    int DMVertices[6];
    int Nd;
    double k1, k2, k3;
    double *veff;
} ref_problem_t;

static double *dup_coefs(const double *src, int n)
{
    double *p = (double *)malloc(sizeof(double) * n);
    memcpy(p, src, sizeof(double) * n);
    return p;
}

void *ref_problem_create(const chefsi_grid_t *g, const chefsi_nloc_t *nl, const double *veff,
                         const double *kvec)
{
    ref_problem_t *P = (ref_problem_t *)calloc(1, sizeof(ref_problem_t));
    SPARC_OBJ *S = &P->S;
    const int n = g->FDn + 1;
    S->order = 2 * g->FDn;
    S->cell_typ = g->cell_typ;
    S->BCx = g->BCx; S->BCy = g->BCy; S->BCz = g->BCz;
    S->Nx = g->Nx; S->Ny = g->Ny; S->Nz = g->Nz;
    S->Nd = g->Nx * g->Ny * g->Nz;
    S->Nd_d_dmcomm = S->Nd;
    S->range_x = g->range_x; S->range_y = g->range_y; S->range_z = g->range_z;
    S->dV = g->dV;
    S->D2_stencil_coeffs_x = dup_coefs(g->D2_x, n);
    S->D2_stencil_coeffs_y = dup_coefs(g->D2_y, n);
    S->D2_stencil_coeffs_z = dup_coefs(g->D2_z, n);
    S->D2_stencil_coeffs_xy = dup_coefs(g->D2_xy, n);
    S->D2_stencil_coeffs_xz = dup_coefs(g->D2_xz, n);
    S->D2_stencil_coeffs_yz = dup_coefs(g->D2_yz, n);
    S->D1_stencil_coeffs_x = dup_coefs(g->D1_x, n);
    S->D1_stencil_coeffs_y = dup_coefs(g->D1_y, n);
    S->D1_stencil_coeffs_z = dup_coefs(g->D1_z, n);
    S->D1_stencil_coeffs_xy = dup_coefs(g->D1_xy, n);
    S->D1_stencil_coeffs_yx = dup_coefs(g->D1_yx, n);
    S->D1_stencil_coeffs_xz = dup_coefs(g->D1_xz, n);
    S->D1_stencil_coeffs_zx = dup_coefs(g->D1_zx, n);
    S->D1_stencil_coeffs_yz = dup_coefs(g->D1_yz, n);
    S->D1_stencil_coeffs_zy = dup_coefs(g->D1_zy, n);
    S->CyclixFlag = 0;
    S->usefock = 0;
    S->ixc[2] = 0;
    S->is_hubbard = 0;
    S->SOC_Flag = 0;
    S->spin_typ = 0;
    S->Nspinor_eig = 1;
    S->Nspinor_spincomm = 1;
    S->spin_start_indx = 0;
    S->bandcomm_index = 0;
    S->kptcomm_topo = MPI_COMM_NULL;       /* != the communicator passed below */
    S->kptcomm_inter = MPI_COMM_NULL;      /* Lanczos: no ranks outside the Cartesian topology (eigenSolver.c:2105) */
    S->comm_dist_graph_phi = MPI_COMM_SELF;
    S->comm_dist_graph_psi = MPI_COMM_SELF;
    S->kptcomm_topo_dist_graph = MPI_COMM_SELF;

    P->Nd = S->Nd;
    P->DMVertices[0] = 0; P->DMVertices[1] = g->Nx - 1;
    P->DMVertices[2] = 0; P->DMVertices[3] = g->Ny - 1;
    P->DMVertices[4] = 0; P->DMVertices[5] = g->Nz - 1;
    for (int i = 0; i < 6; i++) S->DMVertices[i] = P->DMVertices[i]; /* phi domain = the whole grid (poisson_residual) */
    P->veff = (double *)calloc(S->Nd, sizeof(double));
    if (veff) memcpy(P->veff, veff, sizeof(double) * S->Nd);
    S->Veff_loc_dmcomm = P->veff;

    if (kvec) { P->k1 = kvec[0]; P->k2 = kvec[1]; P->k3 = kvec[2]; }
    S->k1_loc = &P->k1; S->k2_loc = &P->k2; S->k3_loc = &P->k3;

    /* projector tables: one "type" per real atom */
    const int natom = nl ? nl->n_atom : 0;
    S->n_atom = natom;
    S->Ntypes = natom;
    S->nAtomv = (int *)calloc(natom + 1, sizeof(int));
    S->localPsd = (int *)calloc(natom + 1, sizeof(int));
    S->psd = (PSD_OBJ *)calloc(natom + 1, sizeof(PSD_OBJ));
    S->IP_displ = (int *)calloc(natom + 1, sizeof(int));
    S->Atom_Influence_nloc = (ATOM_NLOC_INFLUENCE_OBJ *)calloc(natom + 1, sizeof(ATOM_NLOC_INFLUENCE_OBJ));
    S->nlocProj = (NLOC_PROJ_OBJ *)calloc(natom + 1, sizeof(NLOC_PROJ_OBJ));
    for (int t = 0; t < natom; t++) {
        const int nproj = nl->IP_displ[t + 1] - nl->IP_displ[t];
        S->IP_displ[t] = nl->IP_displ[t];
        S->IP_displ[t + 1] = nl->IP_displ[t + 1];
        S->nAtomv[t] = 1;
        S->localPsd[t] = 4;
        S->psd[t].lmax = 0;
        S->psd[t].ppl = (int *)calloc(1, sizeof(int));
        S->psd[t].ppl[0] = nproj;
        S->psd[t].Gamma = dup_coefs(nl->gamma + nl->IP_displ[t], nproj > 0 ? nproj : 1);
        S->nlocProj[t].nproj = nproj;
        int cnt = 0;
        for (int J = 0; J < nl->n_img; J++) cnt += (nl->img_atom[J] == t);
        ATOM_NLOC_INFLUENCE_OBJ *A = &S->Atom_Influence_nloc[t];
        A->n_atom = cnt;
        A->coords = (double *)calloc(3 * cnt + 1, sizeof(double));
        A->atom_index = (int *)calloc(cnt + 1, sizeof(int));
        A->ndc = (int *)calloc(cnt + 1, sizeof(int));
        A->grid_pos = (int **)calloc(cnt + 1, sizeof(int *));
        S->nlocProj[t].Chi = (double **)calloc(cnt + 1, sizeof(double *));
        S->nlocProj[t].Chi_c = (double _Complex **)calloc(cnt + 1, sizeof(double _Complex *));
        int q = 0;
        for (int J = 0; J < nl->n_img; J++) {
            if (nl->img_atom[J] != t) continue;
            const int ndc = nl->img_ndc[J];
            A->atom_index[q] = t;
            A->ndc[q] = ndc;
            memcpy(A->coords + 3 * q, nl->img_coords + 3 * J, 3 * sizeof(double));
            A->grid_pos[q] = (int *)malloc(sizeof(int) * (ndc > 0 ? ndc : 1));
            memcpy(A->grid_pos[q], nl->grid_pos + nl->pos_off[J], sizeof(int) * ndc);
            const size_t nchi = (size_t)ndc * nproj;
            S->nlocProj[t].Chi[q] = (double *)malloc(sizeof(double) * (nchi ? nchi : 1));
            S->nlocProj[t].Chi_c[q] = (double _Complex *)malloc(sizeof(double _Complex) * (nchi ? nchi : 1));
            for (size_t i = 0; i < nchi; i++) {
                S->nlocProj[t].Chi[q][i] = nl->chi[nl->chi_off[J] + i];
                S->nlocProj[t].Chi_c[q][i] = nl->chi[nl->chi_off[J] + i];
            }
            q++;
        }
    }
    return P;
}

void ref_problem_destroy(void *h)
{
    /* test harness: leak the small tables, free the big ones */
    ref_problem_t *P = (ref_problem_t *)h;
    if (!P) return;
    SPARC_OBJ *S = &P->S;
    for (int t = 0; t < S->Ntypes; t++) {
        for (int q = 0; q < S->Atom_Influence_nloc[t].n_atom; q++) {
            free(S->Atom_Influence_nloc[t].grid_pos[q]);
            free(S->nlocProj[t].Chi[q]);
            free(S->nlocProj[t].Chi_c[q]);
        }
    }
    free(P->veff);
    free(P);
}

static int dims1[3] = {1, 1, 1};

void ref_lap_plus_diag(void *h, int ncol, double a, double b, double c, int use_v, const double *x,
                       int ldi, double *y, int ldo)
{
    ref_problem_t *P = (ref_problem_t *)h;
    const double *v = use_v ? P->veff : NULL;
    for (int n = 0; n < ncol; n++) {
        if (P->S.cell_typ == 0)
            Lap_plus_diag_vec_mult_orth(&P->S, P->Nd, P->DMVertices, 1, a, b, c, v, x + (size_t)n * ldi,
                                        ldi, y + (size_t)n * ldo, ldo, MPI_COMM_SELF, dims1);
        else
            Lap_plus_diag_vec_mult_nonorth(&P->S, P->Nd, P->DMVertices, 1, a, b, c, v,
                                           x + (size_t)n * ldi, ldi, y + (size_t)n * ldo, ldo,
                                           MPI_COMM_SELF, MPI_COMM_SELF, dims1);
    }
}

void ref_lap_plus_diag_kpt(void *h, int ncol, double a, double b, double c, int use_v,
                           const double _Complex *x, int ldi, double _Complex *y, int ldo)
{
    ref_problem_t *P = (ref_problem_t *)h;
    const double *v = use_v ? P->veff : NULL;
    for (int n = 0; n < ncol; n++) {
        if (P->S.cell_typ == 0)
            Lap_plus_diag_vec_mult_orth_kpt(&P->S, P->Nd, P->DMVertices, 1, a, b, c, v,
                                            x + (size_t)n * ldi, ldi, y + (size_t)n * ldo, ldo,
                                            MPI_COMM_SELF, dims1, 0);
        else
            Lap_plus_diag_vec_mult_nonorth_kpt(&P->S, P->Nd, P->DMVertices, 1, a, b, c, v,
                                               x + (size_t)n * ldi, ldi, y + (size_t)n * ldo, ldo,
                                               MPI_COMM_SELF, MPI_COMM_SELF, dims1, 0);
    }
}

void ref_vnl_mult(void *h, int ncol, double *x, int ldi, double *Hx, int ldo)
{
    ref_problem_t *P = (ref_problem_t *)h;
    Vnl_vec_mult(&P->S, P->Nd, P->S.Atom_Influence_nloc, P->S.nlocProj, ncol, x, ldi, Hx, ldo,
                 MPI_COMM_SELF);
}

void ref_vnl_mult_kpt(void *h, int ncol, double _Complex *x, int ldi, double _Complex *Hx, int ldo)
{
    ref_problem_t *P = (ref_problem_t *)h;
    Vnl_vec_mult_kpt(&P->S, P->Nd, P->S.Atom_Influence_nloc, P->S.nlocProj, ncol, x, ldi, Hx, ldo, 0,
                     MPI_COMM_SELF);
}

void ref_hamiltonian_mult(void *h, int ncol, double c, double *x, int ldi, double *Hx, int ldo)
{
    ref_problem_t *P = (ref_problem_t *)h;
    Hamiltonian_vectors_mult(&P->S, P->Nd, P->DMVertices, P->veff, P->S.Atom_Influence_nloc,
                             P->S.nlocProj, ncol, c, x, ldi, Hx, ldo, 0, MPI_COMM_SELF);
}

void ref_hamiltonian_mult_kpt(void *h, int ncol, double c, double _Complex *x, int ldi,
                              double _Complex *Hx, int ldo)
{
    ref_problem_t *P = (ref_problem_t *)h;
    Hamiltonian_vectors_mult_kpt(&P->S, P->Nd, P->DMVertices, P->veff, P->S.Atom_Influence_nloc,
                                 P->S.nlocProj, ncol, c, x, ldi, Hx, ldo, 0, 0, MPI_COMM_SELF);
}

double ref_chebyshev_filter(void *h, double *X, int ldi, double *Y, int ldo, int ncol, int m,
                            double a, double b, double a0)
{
    ref_problem_t *P = (ref_problem_t *)h;
    double t = 0.0;
    ChebyshevFiltering(&P->S, P->DMVertices, X, ldi, Y, ldo, ncol, m, a, b, a0, 0, 0, MPI_COMM_SELF, &t);
    return t;
}

double ref_chebyshev_filter_kpt(void *h, double _Complex *X, int ldi, double _Complex *Y, int ldo,
                                int ncol, int m, double a, double b, double a0)
{
    ref_problem_t *P = (ref_problem_t *)h;
    double t = 0.0;
    ChebyshevFiltering_kpt(&P->S, P->DMVertices, X, ldi, Y, ldo, ncol, m, a, b, a0, 0, 0,
                           MPI_COMM_SELF, &t);
    return t;
}

/* ---- the rows around the filter: the reference's own Lanczos and AAR on the same flat problem ---- */
void Jacobi_preconditioner(SPARC_OBJ *pSPARC, int N, double c, double *r, double *f, MPI_Comm comm); /* electrostatics.c:1682 */

/* Lanczos (src/eigenSolver.c:1920-2129) from the start vector x0 */
void ref_lanczos(void *h, double *x0, double tol_min, double tol_max, int maxit, double *eigmin, double *eigmax)
{
    ref_problem_t *P = (ref_problem_t *)h;
    MPI_Request req = MPI_REQUEST_NULL;
    Lanczos(&P->S, P->DMVertices, P->veff, P->S.Atom_Influence_nloc, P->S.nlocProj, eigmin, eigmax, x0, tol_min, tol_max, maxit,
            0, 0, MPI_COMM_SELF, &req);
}

/* Lanczos_kpt (src/eigenSolverKpt.c:1361-1566) from the complex start vector x0, at the problem's k-point */
void ref_lanczos_kpt(void *h, double _Complex *x0, double tol_min, double tol_max, int maxit, double *eigmin, double *eigmax)
{
    ref_problem_t *P = (ref_problem_t *)h;
    MPI_Request req = MPI_REQUEST_NULL;
    Lanczos_kpt(&P->S, P->DMVertices, P->veff, P->S.Atom_Influence_nloc, P->S.nlocProj, eigmin, eigmax, x0, tol_min, tol_max,
                maxit, 0, 0, MPI_COMM_SELF, &req);
}

/* DP_Solve_Generalized_EigenProblem (src/eigenSolver.c:1262-1375, LAPACK branch: LAPACKE_dsygvd itype 1 'V' 'U') on
 * caller-provided n x n column-major Hp, Mp (both overwritten, as in the reference): lambda[n], Q = eig_vecs */
void DP_Solve_Generalized_EigenProblem(SPARC_OBJ *pSPARC, int spn_i);
void DP_Solve_Generalized_EigenProblem_kpt(SPARC_OBJ *pSPARC, int kpt, int spn_i);
void ref_subspace_eig(void *h, int n, double *Hp, double *Mp, double *lambda, double *Q)
{
    ref_problem_t *P = (ref_problem_t *)h;
    SPARC_OBJ *S = &P->S;
    struct DP_CheFSI_s dp;
    memset(&dp, 0, sizeof(dp));
    dp.Ns_dp = n; dp.rank_kpt = 0; dp.Hp_local = Hp; dp.Mp_local = Mp; dp.eig_vecs = Q; dp.kpt_comm = MPI_COMM_SELF;
    void *save_dp = S->DP_CheFSI;
    double *save_lambda = S->lambda;
    const int save_lapack = S->useLAPACK, save_std = S->StandardEigenFlag;
    S->DP_CheFSI = &dp; S->lambda = lambda; S->useLAPACK = 1; S->StandardEigenFlag = 0;
    DP_Solve_Generalized_EigenProblem(S, 0);
    S->DP_CheFSI = save_dp; S->lambda = save_lambda; S->useLAPACK = save_lapack; S->StandardEigenFlag = save_std;
}

/* DP_Solve_Generalized_EigenProblem_kpt (src/eigenSolverKpt.c:836-930, LAPACKE_zhegvd) */
void ref_subspace_eig_kpt(void *h, int n, double _Complex *Hp, double _Complex *Mp, double *lambda, double _Complex *Q)
{
    ref_problem_t *P = (ref_problem_t *)h;
    SPARC_OBJ *S = &P->S;
    struct DP_CheFSI_kpt_s dp;
    memset(&dp, 0, sizeof(dp));
    dp.Ns_dp = n; dp.rank_kpt = 0; dp.Hp_local = Hp; dp.Mp_local = Mp; dp.eig_vecs = Q; dp.kpt_comm = MPI_COMM_SELF;
    void *save_dp = S->DP_CheFSI_kpt;
    double *save_lambda = S->lambda;
    const int save_lapack = S->useLAPACK, save_ns = S->Nstates, save_nk = S->Nkpts_kptcomm;
    S->DP_CheFSI_kpt = &dp; S->lambda = lambda; S->useLAPACK = 1; S->Nstates = n; S->Nkpts_kptcomm = 1;
    DP_Solve_Generalized_EigenProblem_kpt(S, 0, 0);
    S->DP_CheFSI_kpt = save_dp; S->lambda = save_lambda; S->useLAPACK = save_lapack; S->Nstates = save_ns; S->Nkpts_kptcomm = save_nk;
}

/* CalculateDensity_psi (src/electronDensity.c:104-200) for one k-point, no spin: rho[Nd] (zeroed by the caller, as
 * Calculate_elecDens does with calloc, :33) += occfac * kptwt * occ[n] |X_n|^2, then the 1/dV scaling (:190-196) */
void CalculateDensity_psi(SPARC_OBJ *pSPARC, double *rho);
void ref_density(void *h, int ncol, void *X, int is_complex, double *occ, double occfac, double kptwt, double *rho)
{
    ref_problem_t *P = (ref_problem_t *)h;
    SPARC_OBJ *S = &P->S;
    SPARC_OBJ keep = *S;
    S->spincomm_index = 0; S->kptcomm_index = 0; S->bandcomm_index = 0; S->dmcomm = MPI_COMM_SELF;
    S->Nstates = ncol; S->band_start_indx = 0; S->band_end_indx = ncol - 1;
    S->Nspinor = 1; S->Nspinor_spincomm = 1; S->spinor_start_indx = 0; S->spin_typ = 0;
    S->Nkpts_kptcomm = 1; S->Nkpts = 1; S->kptWts_loc = &kptwt; S->occfac = occfac; S->occ = occ;
    S->isGammaPoint = !is_complex;
    S->Xorb = (double *)X; S->Xorb_kpt = (double _Complex *)X;
    S->npspin = 1; S->npkpt = 1; S->npband = 1; S->blacscomm = MPI_COMM_SELF;
    CalculateDensity_psi(S, rho);
    *S = keep;
}

/* AAR (src/linearSolver.c:38-146) with the operator pair SPARC uses: poisson_residual (lapVecRoutines.c:61) and
 * Jacobi_preconditioner; x is the start vector on entry and the solution on return */
void ref_aar(void *h, double c, double *x, double *b, double omega, double beta, int m, int p, double tol, int max_iter)
{
    ref_problem_t *P = (ref_problem_t *)h;
    AAR(&P->S, poisson_residual, Jacobi_preconditioner, c, P->Nd, x, b, omega, beta, m, p, tol, max_iter, MPI_COMM_SELF);
}

/* Gradient_vectors_dir (src/gradVecRoutines.c:32-51) and Gradient_vectors_dir_kpt (src/gradVecRoutinesKpt.c:35-55):
 * (D_dir + c) x on ncol columns; the k-point routine takes a pointer to the k component along dir */
void Gradient_vectors_dir(const SPARC_OBJ *pSPARC, const int DMnd, const int *DMVertices, const int ncol, const double c,
                          const double *x, const int ldi, double *Dx, const int ldo, const int dir, MPI_Comm comm);
void Gradient_vectors_dir_kpt(const SPARC_OBJ *pSPARC, const int DMnd, const int *DMVertices, const int ncol, const double c,
                              const double _Complex *x, const int ldi, double _Complex *Dx, const int ldo, const int dir,
                              const double *kpt_vec, MPI_Comm comm);
void ref_gradient_dir(void *h, int ncol, double c, const double *x, int ldi, double *Dx, int ldo, int dir)
{
    ref_problem_t *P = (ref_problem_t *)h;
    Gradient_vectors_dir(&P->S, (int)P->Nd, P->DMVertices, ncol, c, x, ldi, Dx, ldo, dir, MPI_COMM_SELF);
}
void ref_gradient_dir_kpt(void *h, int ncol, double c, const double _Complex *x, int ldi, double _Complex *Dx, int ldo, int dir,
                          double kdir)
{
    ref_problem_t *P = (ref_problem_t *)h;
    Gradient_vectors_dir_kpt(&P->S, (int)P->Nd, P->DMVertices, ncol, c, x, ldi, Dx, ldo, dir, &kdir, MPI_COMM_SELF);
}

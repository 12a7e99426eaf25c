/*
 * chefsi_b200.h -- C ABI of libchefsi_b200.so
 *
 * B200-native (sm_100a CUDA, FP64) implementation of ONE hot path of SPARC: the
 * Chebyshev-filtered subspace iteration (CheFSI) filter, i.e. the repeated
 * (H - cI) * X product fused with the three-term Chebyshev recurrence.
 *
 * Everything here is plain C: pointers, sizes, POD structs.  No CUDA, torch or
 * MPI types cross this boundary.  The reference-side binding (the functions
 * SPARC itself calls) lives in sparc_b200/csrc/sparc_shim.c, which flattens
 * SPARC_OBJ into the structs below; see INTEGRATION.md.
 *
 * Reference interfaces replaced (paths relative to the SPARC tree):
 *   chefsi_chebyshev_filter        <- ChebyshevFiltering            src/eigenSolver.c:722-798
 *                                     (decl. src/include/eigenSolver.h:93-97)
 *   chefsi_chebyshev_filter_kpt    <- ChebyshevFiltering_kpt        src/eigenSolverKpt.c:458-535
 *                                     (decl. src/include/eigenSolverKpt.h:41-45)
 *   chefsi_hamiltonian_mult        <- Hamiltonian_vectors_mult      src/hamiltonianVecRoutines.c:45-121
 *   chefsi_hamiltonian_mult_kpt    <- Hamiltonian_vectors_mult_kpt  src/hamiltonianVecRoutines.c:132-242
 *                                     (decl. src/include/hamiltonianVecRoutines.h:30-47)
 *   chefsi_laplacian_mult[_kpt]    <- Lap_vec_mult[_kpt]            src/lapVecRoutines.c:37-58, src/lapVecRoutinesKpt.c
 *                                     (the Laplacian of the Poisson residual, src/lapVecRoutines.c:61-79)
 *   chefsi_set_grid                <- the SPARC_OBJ fields read by Lap_plus_diag_vec_mult_{orth,nonorth}[_kpt]
 *                                     (src/lapVecRoutines.c:306,940; src/lapVecRoutinesKpt.c:179,567)
 *   chefsi_set_projectors          <- ATOM_NLOC_INFLUENCE_OBJ / NLOC_PROJ_OBJ / PSD_OBJ.Gamma / IP_displ
 *                                     as read by Vnl_vec_mult[_kpt] (src/nlocVecRoutines.c:798,889)
 *   chefsi_set_veff                <- Veff_loc argument / Transfer_Veff_loc (src/electronicGroundState.c:1313)
 *
 * Conventions
 *   - Grid functions are stored x-fastest: index = k*Nx*Ny + j*Nx + i (lapVecRoutines.c:313).
 *   - A block of orbitals is column-major: column n starts at ptr + n*ld (ld >= Nd).
 *   - Complex data is interleaved (re,im) doubles, identical to C99 `double _Complex`.
 *   - All functions return 0 on success, non-zero on failure; chefsi_last_error() then
 *     returns a message.  There is NO CPU fallback: if no CUDA device is usable the
 *     call fails.
 *   - A context is bound to one CUDA device and is not thread-safe (the reference calls
 *     the path from its single MPI-rank thread).
 */
#ifndef CHEFSI_B200_H
#define CHEFSI_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHEFSI_MAX_FDN 12 /* FD_ORDER <= 24 */

/* Discretisation of one (unsplit) domain.  Field meaning and provenance:
 * src/include/isddft.h:391 (cell_typ), :396-398 (Nx..), :401-403 (range_*), :411 (dV),
 * :464 (order), :467-481 (stencil coefficient tables), :667-669 (BC*).
 * Coefficient tables are the values SPARC computed in src/initialization.c:2112-2178;
 * this library never recomputes them. */
typedef struct chefsi_grid {
    int Nx, Ny, Nz;
    int BCx, BCy, BCz; /* 0 periodic, 1 Dirichlet (zero halo) */
    int FDn;           /* order / 2 */
    int cell_typ;      /* 0 orthogonal; 11..17 non-orthogonal flavours */
    double dV;
    double range_x, range_y, range_z; /* cell lengths, used for Bloch phases */
    double D2_x[CHEFSI_MAX_FDN + 1], D2_y[CHEFSI_MAX_FDN + 1], D2_z[CHEFSI_MAX_FDN + 1];
    double D2_xy[CHEFSI_MAX_FDN + 1], D2_xz[CHEFSI_MAX_FDN + 1], D2_yz[CHEFSI_MAX_FDN + 1];
    double D1_x[CHEFSI_MAX_FDN + 1], D1_y[CHEFSI_MAX_FDN + 1], D1_z[CHEFSI_MAX_FDN + 1];
    double D1_xy[CHEFSI_MAX_FDN + 1], D1_yx[CHEFSI_MAX_FDN + 1];
    double D1_xz[CHEFSI_MAX_FDN + 1], D1_zx[CHEFSI_MAX_FDN + 1];
    double D1_yz[CHEFSI_MAX_FDN + 1], D1_zy[CHEFSI_MAX_FDN + 1];
} chefsi_grid_t;

/* Kleinman-Bylander nonlocal projectors, flattened over atom types.
 * "image" = one entry of Atom_Influence_nloc[ityp] (an atom or one of its periodic
 * images whose rc-sphere touches the domain), in the order ityp-major, iat-minor
 * (the loop order of nlocVecRoutines.c:807-831). */
typedef struct chefsi_nloc {
    int n_atom;             /* real atoms; alpha has IP_displ[n_atom] rows per column   */
    const int *IP_displ;    /* [n_atom+1]   isddft.h:460, nlocVecRoutines.c:769-791     */
    const double *gamma;    /* [IP_displ[n_atom]]  PSD_OBJ.Gamma expanded to one value   */
                            /* per (atom, projector) in alpha order (nlocVecRoutines.c:841-863) */
    int n_img;
    const int *img_atom;    /* [n_img]   atom_index  (isddft.h:206)                      */
    const int *img_ndc;     /* [n_img]   grid points in the rc-sphere (isddft.h:213)     */
    const double *img_coords; /* [3*n_img] image coordinates (isddft.h:205); only used   */
                            /* for the k-point Bloch factor (nlocVecRoutines.c:911-921)  */
    const long long *pos_off; /* [n_img+1] offsets into grid_pos                         */
    const long long *chi_off; /* [n_img+1] offsets into chi                              */
    const int *grid_pos;    /* linear grid indices of each sphere (isddft.h:214)         */
    const double *chi;      /* per image: ndc x nproj, column-major, real values         */
                            /* (NLOC_PROJ_OBJ.Chi, or creal(Chi_c): nlocVecRoutines.c:731) */
} chefsi_nloc_t;

typedef struct chefsi_ctx chefsi_ctx_t;

/* flags for the filter entry points */
#define CHEFSI_FLAG_NO_X_COPYBACK 1 /* host entry points: do not copy the clobbered X back   */
                                    /* (the caller reuses X as scratch, eigenSolver.c:364)   */
#define CHEFSI_FLAG_KEEP_Y 2        /* keep Y = p_m(H) X0 on the device for chefsi_subspace_project / _rotate            */
                                    /* (single-device context; needs a successful chefsi_subspace_reserve[_kpt])             */
#define CHEFSI_FLAG_NO_Y_COPYBACK 4 /* with KEEP_Y: do not copy Y to the host at all (the caller's next steps are           */
                                    /* chefsi_subspace_project and chefsi_subspace_rotate, which read the device copy)      */

/* ---- lifetime ---------------------------------------------------------------------- */
int chefsi_device_count(void); /* usable CUDA devices (0 when there is none or the driver is absent) */
int chefsi_create(chefsi_ctx_t **ctx, int device);
/* One context that owns `ndev` GPUs of a box (device ordinals in `devices`): SPARC's band-parallel axis inside one
 * process (NP_BAND_PARAL = ndev, src/parallelization.c:403-428).  The host-buffer entry points split their columns
 * NB = ceil(ncol / ndev) per device and run the devices concurrently; chefsi_set_veff / chefsi_set_projectors upload to
 * the first device and replicate with ncclBroadcast over NVLink (Transfer_Veff_loc's MPI_Bcast,
 * src/electronicGroundState.c:1313-1385); NCCL is loaded at run time, cudaMemcpyPeer is used when it is absent.
 * chefsi_subspace_project / _rotate on such a context split Hp / Mp / Y Q by column blocks and read the other devices'
 * blocks through NVLink peer memory inside the GEMM kernels; chefsi_lanczos / chefsi_poisson_aar (single vectors) run on
 * the first device.  The device-resident entry points take single-device contexts only. */
int chefsi_create_multi(chefsi_ctx_t **ctx, const int *devices, int ndev);
/* devices of the context (1 for a single-device context), whether the replication runs over NCCL, and the
 * number / bytes of broadcasts so far */
int chefsi_multi_info(const chefsi_ctx_t *ctx, int *ndev, int *uses_nccl, unsigned long long *bcast_calls,
                      unsigned long long *bcast_bytes);
void chefsi_destroy(chefsi_ctx_t *ctx);
const char *chefsi_last_error(const chefsi_ctx_t *ctx);
const char *chefsi_version(void);

/* ---- problem description (host pointers; copied to the device) ---------------------- */
int chefsi_set_grid(chefsi_ctx_t *ctx, const chefsi_grid_t *grid);
int chefsi_set_projectors(chefsi_ctx_t *ctx, const chefsi_nloc_t *nloc); /* NULL: no Vnl */
int chefsi_set_veff(chefsi_ctx_t *ctx, const double *veff_host);        /* Nd doubles   */
int chefsi_set_kpoint(chefsi_ctx_t *ctx, double k1, double k2, double k3);

/* ---- host-buffer entry points: what the reference's functions bind to ---------------
 * X is in/out (ends as p_{m-1}(H) X0, as in the reference), Y is out (p_m(H) X0).
 * a = lambda_cutoff, b = eigmax, a0 = eigmin (eigenSolver.c:729).                       */
int chefsi_chebyshev_filter(chefsi_ctx_t *ctx, double *X, size_t ldi, double *Y, size_t ldo,
                            int ncol, int m, double a, double b, double a0, int flags);
int chefsi_chebyshev_filter_kpt(chefsi_ctx_t *ctx, void *X, size_t ldi, void *Y, size_t ldo,
                                int ncol, int m, double a, double b, double a0, int flags);
/* Hx = (-1/2 Lap + Veff + c) x + Vnl x */
int chefsi_hamiltonian_mult(chefsi_ctx_t *ctx, int ncol, double c, const double *x, size_t ldi,
                            double *Hx, size_t ldo);
int chefsi_hamiltonian_mult_kpt(chefsi_ctx_t *ctx, int ncol, double c, const void *x,
                                size_t ldi, void *Hx, size_t ldo);

/* y = (a Lap + c) x: no potential, no projectors (Lap_vec_mult: Lap_plus_diag_vec_mult_* with b = 0, v = NULL,
 * lapVecRoutines.c:37-58,321-322).  a must be non-zero.  Same kernels as the Hamiltonian apply. */
int chefsi_laplacian_mult(chefsi_ctx_t *ctx, int ncol, double a, double c, const double *x, size_t ldi,
                          double *y, size_t ldo);
int chefsi_laplacian_mult_kpt(chefsi_ctx_t *ctx, int ncol, double a, double c, const void *x, size_t ldi,
                              void *y, size_t ldo);

/* Dx = (D_dir + c) x: the first-derivative stencil along lattice direction dir (0: x, 1: y, 2: z) with the weights
 * D1_{x,y,z} of chefsi_grid_t, periodic wrap or zero halo along dir.  Replaces Gradient_vectors_dir
 * (src/gradVecRoutines.c:32-51 -> Gradient_vec_dir :59-311 -> Calc_DX :318-442) and, for complex columns,
 * Gradient_vectors_dir_kpt (src/gradVecRoutinesKpt.c:35-55 -> :63-340): kdir is the k-point component along dir (what
 * the reference passes as *kpt_vec); halo values from beyond the low / high face are multiplied by exp(-/+ i kdir L_dir)
 * (:179-191).  SURVEY.md 8f-4 ("gradient ops sharing the stencil": GGA density gradients, force / stress terms). */
int chefsi_gradient_mult(chefsi_ctx_t *ctx, int ncol, double c, const double *x, size_t ldi, double *Dx, size_t ldo,
                         int dir);
int chefsi_gradient_mult_kpt(chefsi_ctx_t *ctx, int ncol, double c, const void *x, size_t ldi, void *Dx, size_t ldo,
                             int dir, double kdir);

/* ---- Rayleigh-Ritz projection and subspace rotation with the block resident on the device -------
 * Replaces, for real (Gamma-point) data on a single-device context:
 *   chefsi_subspace_project <- DP_Project_Hamiltonian  src/eigenSolver.c:939-1086 (Project_Hamiltonian :1477-1669):
 *                              HY = H Y (c = 0), Mp = Y^T Y, Hp = Y^T HY; Hp, Mp: ncol x ncol, column-major, ld = ldp (host)
 *   chefsi_subspace_rotate  <- DP_Subspace_Rotation    src/eigenSolver.c:1386-1443 (Subspace_Rotation :1854-1918):
 *                              X = Y Q, Q: ncol x ncol (host, ld = ldq), X: host block, ld = ldx
 * The three GEMMs run on the FP64 tensor cores (DMMA); only Hp, Mp, Q and the rotated block cross PCIe.
 * chefsi_subspace_reserve allocates the two resident blocks for `ncol` columns (fails when they do not fit).
 * chefsi_subspace_project takes Y from the device when the last filter call kept it (CHEFSI_FLAG_KEEP_Y, same host
 * address and column count), otherwise it uploads the host block. chefsi_subspace_rotate needs a preceding project;
 * Q == NULL: the eigenvectors chefsi_subspace_eig left on the device. */
int chefsi_subspace_reserve(chefsi_ctx_t *ctx, int ncol);
int chefsi_subspace_project(chefsi_ctx_t *ctx, const double *Y, size_t ldy, int ncol, double *Hp, double *Mp, size_t ldp);
int chefsi_subspace_rotate(chefsi_ctx_t *ctx, const double *Q, size_t ldq, int ncol, double *X, size_t ldx);
/* k-point (complex, interleaved re/im) variants: Hp = Y^H H Y, Mp = Y^H Y (DP_Project_Hamiltonian_kpt,
 * src/eigenSolverKpt.c:676-790: zgemm ConjTrans), X = Y Q (DP_Subspace_Rotation_kpt, :947-1010); set the k-point first */
int chefsi_subspace_reserve_kpt(chefsi_ctx_t *ctx, int ncol);
int chefsi_subspace_project_kpt(chefsi_ctx_t *ctx, const void *Y, size_t ldy, int ncol, void *Hp, void *Mp, size_t ldp);
int chefsi_subspace_rotate_kpt(chefsi_ctx_t *ctx, const void *Q, size_t ldq, int ncol, void *X, size_t ldx);

/* ---- the subspace eigenproblem and the density -----------------------------------------------------
 * chefsi_subspace_eig[_kpt] <- DP_Solve_Generalized_EigenProblem      src/eigenSolver.c:1262-1375 (LAPACKE_dsygvd, itype 1, 'V'),
 *                              DP_Solve_Generalized_EigenProblem_kpt  src/eigenSolverKpt.c:836-930 (LAPACKE_zhegvd):
 *   Hp q = lambda Mp q.  Hp == Mp == NULL: the matrices chefsi_subspace_project[_kpt] left on the device (single-device
 *   context, same ncol); otherwise they are uploaded (host, column-major, ld = ldp) -- the form a multi-device context
 *   needs.  lambda: ncol ascending eigenvalues (host).  Q (host, column-major, ld = ldq; may be NULL): eigenvectors,
 *   Q^H Mp Q = I.  They also stay on the device: chefsi_subspace_rotate[_kpt] with Q == NULL uses them.
 *   The reference calls a LAPACK library here; this calls the same routine of cuSOLVER (cusolverDnDsygvd / Zhegvd),
 *   loaded with dlopen on first use (CHEFSI_B200_CUSOLVER_LIB overrides the name); fails if it is absent.
 * chefsi_density_accumulate[_kpt] <- the loop body of CalculateDensity_psi  src/electronDensity.c:104-200:
 *   rho[i] += sum_n g[n] |X[i + n ldx]|^2 for one k-point / spin block (X, rho: host; g[n] = occfac * w_k * occ[n]).
 *   With the band store on, chefsi_subspace_rotate keeps a device copy of each rotated block keyed by its host address
 *   and the density reads that copy (consuming it); a block that is not resident is uploaded.  Filtering a host block
 *   drops its copy.  chefsi_band_store(ctx, n): keep up to n blocks (k-points x spins of the rank); 0 frees them. */
int chefsi_subspace_eig(chefsi_ctx_t *ctx, int ncol, const double *Hp, const double *Mp, size_t ldp, double *lambda, double *Q,
                        size_t ldq);
int chefsi_subspace_eig_kpt(chefsi_ctx_t *ctx, int ncol, const void *Hp, const void *Mp, size_t ldp, double *lambda, void *Q,
                            size_t ldq);
int chefsi_band_store(chefsi_ctx_t *ctx, int max_blocks);
int chefsi_density_accumulate(chefsi_ctx_t *ctx, const double *X, size_t ldx, int ncol, const double *g, double *rho);
int chefsi_density_accumulate_kpt(chefsi_ctx_t *ctx, const void *X, size_t ldx, int ncol, const double *g, double *rho);

/* Extreme eigenvalues of H = -1/2 Lap + Veff + Vnl by the Lanczos iteration with every vector resident on the device
 * (SURVEY.md 8f-2): the body of Lanczos (src/eigenSolver.c:1920-2129) at one rank.  x0: start vector (host, Nd doubles);
 * stops when |eigmin - previous| <= tol_min and |eigmax - previous| <= tol_max, or after maxit steps.
 * Fails (the caller falls back to the reference routine) if x0 is an eigenvector of H.
 * chefsi_lanczos_kpt: the body of Lanczos_kpt (src/eigenSolverKpt.c:1361-1566) for the k-point set with
 * chefsi_set_kpoint; x0: Nd complex numbers (interleaved re/im).  The reference's complex dot product keeps the real
 * part only (VectorDotProduct_complex, src/tools.c:815-826), and so does this. */
int chefsi_lanczos(chefsi_ctx_t *ctx, const double *x0, double tol_min, double tol_max, int maxit, double *eigmin,
                   double *eigmax, int *iterations);
int chefsi_lanczos_kpt(chefsi_ctx_t *ctx, const void *x0, double tol_min, double tol_max, int maxit, double *eigmin,
                       double *eigmax, int *iterations);

/* Alternating Anderson-Richardson solve of -(Lap + c) x = b with the Jacobi preconditioner, every vector resident on the
 * device: AAR (src/linearSolver.c:38-146) with res_fun = poisson_residual (src/lapVecRoutines.c:61) and
 * precond_fun = Jacobi_preconditioner (src/electrostatics.c:1682) -- the Poisson solve of every SCF iteration
 * (electrostatics.c:1658) and the Kerker preconditioner (mixing.c:501).  x: start vector in, solution out (host, Nd);
 * b: right-hand side (host); omega, beta, m (<= 16), p, tol (relative to ||b||), max_iter as in the reference.  Real data. */
int chefsi_poisson_aar(chefsi_ctx_t *ctx, double c, double *x, const double *b, double omega, double beta, int m, int p,
                       double tol, int max_iter, int *iterations, double *res_norm);

/* ---- band-parallel ranks, ONE PROCESS PER GPU: the subspace products over CUDA IPC peer memory (ranks.cu) -------------
 * The reference's band communicator (NB = ceil(Ns / P) columns per rank, src/parallelization.c:403-428) needs an exchange
 * for Mp = Y^T Y, Hp = Y^T H Y and X = Y Q: BP2DP (MPI_Alltoallv, src/parallelization.c:2535; eigenSolver.c:977-990) or
 * pdgemr2d + pdgemm (Project_Hamiltonian :1504-1582, Subspace_Rotation :1854-1918).  Here rank I computes the column
 * block I of Hp / Mp / Y Q and its GEMM kernels read the other ranks' resident blocks Y_J in place, through addresses
 * obtained with CUDA IPC (NVLink peer memory between GPUs): the all-gather is fused into the product.  The caller owns
 * the handle exchange and the barriers between the steps (sparc_b200/band_parallel.py, torch.distributed):
 *   rank_load / KEEP_Y filter -> barrier -> rank_project -> (all-gather of the blocks, eigensolve)
 *   -> rank_rotate_prepare -> barrier -> rank_rotate -> barrier.
 * ncols[nranks]: columns of every rank (rank order = column order); peerY / peerT [nranks]: device addresses of the
 * other ranks' Y (and, complex data, T = i Y) blocks valid in THIS process (entry `rank` ignored; peerT may be NULL for
 * real data).  Blocks of Hp / Mp / Q are host arrays of Ns rows x ncols[rank] columns, column-major. */
#define CHEFSI_IPC_HANDLE_BYTES 64
int chefsi_rank_load(chefsi_ctx_t *ctx, const void *Y, size_t ldy, int ncol, int is_complex);
void *chefsi_resident_ptr(chefsi_ctx_t *ctx, int which); /* 0: Y, 1: W, 2: T -- for ranks that share one process */
int chefsi_ipc_export(chefsi_ctx_t *ctx, int which, void *handle64);
int chefsi_ipc_open(chefsi_ctx_t *ctx, const void *handle64, void **dptr);
int chefsi_ipc_close(chefsi_ctx_t *ctx, void *dptr);
int chefsi_rank_project(chefsi_ctx_t *ctx, int is_complex, int nranks, int rank, const int *ncols, void *const *peerY,
                        void *Hp_blk, void *Mp_blk, size_t ldp);
/* Hp and Mp are Hermitian: with chefsi_rank_project_shared only one element of every mirrored pair is formed (rank I forms
 * block (J, I) for the ranks J that follow it cyclically at a distance below P / 2; for even P the two ranks of an antipodal
 * pair share that block: the lower one forms the first half of its columns, the upper one the rows that mirror the other
 * half; the diagonal block uses upper-triangle tiles), the other elements come back as zeros and the caller mirrors them
 * after the all-gather.  chefsi_rank_block_part tells which part of block (rows of rank J, columns of `rank`) `rank` forms:
 * local rows [r0, r1) x columns [c0, c1); chefsi_rank_forms_block: whether that part is non-empty for equal column counts. */
int chefsi_rank_project_shared(chefsi_ctx_t *ctx, int is_complex, int nranks, int rank, const int *ncols, void *const *peerY,
                               void *Hp_blk, void *Mp_blk, size_t ldp);
int chefsi_rank_forms_block(int J, int rank, int nranks);
void chefsi_rank_block_part(int J, int rank, int nranks, int ncJ, int ncI, int *r0, int *r1, int *c0, int *c1);
int chefsi_rank_rotate_prepare(chefsi_ctx_t *ctx, int is_complex);
int chefsi_rank_rotate(chefsi_ctx_t *ctx, int is_complex, int nranks, int rank, const int *ncols, void *const *peerY,
                       void *const *peerT, const void *Q_blk, size_t ldq, void *X_blk, size_t ldx);

/* ---- device-resident entry points ---------------------------------------------------
 * Buffers are device pointers (256-byte aligned) holding ncol columns in the library's INTERNAL
 * layout: chefsi_device_ld(ctx) elements (doubles, or complex pairs for the _kpt variants) per
 * column; for large orthogonal grids every xy-plane carries a halo pad that the TMA tiles of the
 * streaming kernel read (DESIGN.md "Data layout in HBM").  chefsi_pack_device /
 * chefsi_unpack_device convert from / to the reference's dense layout (column n at n*ld_dense,
 * x fastest) on the device; chefsi_fill_random_device writes the internal layout directly.
 * The three buffers rotate through the recurrence; on return *y_slot / *x_slot say
 * which of {0:bufA, 1:bufB, 2:bufC} hold Y = p_m(H)X0 and X = p_{m-1}(H)X0.
 * bufA holds X0 on entry.  All work is enqueued on the context's stream and the call
 * returns after enqueueing (use chefsi_synchronize).                                    */
size_t chefsi_device_ld(const chefsi_ctx_t *ctx);
int chefsi_chebyshev_filter_device(chefsi_ctx_t *ctx, double *bufA, double *bufB, double *bufC,
                                   int ncol, int m, double a, double b, double a0,
                                   int *y_slot, int *x_slot);
int chefsi_chebyshev_filter_kpt_device(chefsi_ctx_t *ctx, void *bufA, void *bufB, void *bufC,
                                       int ncol, int m, double a, double b, double a0,
                                       int *y_slot, int *x_slot);
int chefsi_hamiltonian_mult_device(chefsi_ctx_t *ctx, int ncol, double c, const double *x,
                                   double *Hx);
int chefsi_hamiltonian_mult_kpt_device(chefsi_ctx_t *ctx, int ncol, double c, const void *x,
                                       void *Hx);
/* chefsi_gradient_mult[_kpt] on a device-resident block in the internal layout (enqueued, no synchronisation) */
int chefsi_gradient_mult_device(chefsi_ctx_t *ctx, int ncol, double c, const void *x, void *Dx, int dir, double kdir,
                                int is_complex);
/* Building blocks of the optional DOMAIN SPLIT (z slabs over ranks, sparc_b200/domain_split.py), real data:
 * the reference exchanges FDn halo planes per face before every stencil application
 * (Lap_plus_diag_vec_mult_orth, src/lapVecRoutines.c:387-442,494-534) and all-reduces the projector
 * inner products alpha over the domain communicator (Vnl_vec_mult, src/nlocVecRoutines.c:834-838).
 * The caller owns the exchange (NCCL send/recv, all-reduce); these calls are the local pieces in between,
 * on device-resident blocks in the internal layout of the rank's slab (grid set with its halo planes and a
 * Dirichlet z face):
 *   stencil_step   out = s1 * ((-1/2 Lap + Veff + c) x) - s2 * xprev        (no projector part; xprev may be NULL)
 *   nloc_project   alpha[IP_displ[atom] * ncol + col * nproj(atom) + p] = dV * sum over the LOCAL sphere points
 *                  of Chi x, written to the caller's device buffer alpha_out (n_proj_total * ncol doubles),
 *                  which the caller all-reduces over the ranks of the split
 *   nloc_expand    out += scale * Chi Gamma alpha_in    (alpha_in: the reduced buffer, device)             */
int chefsi_stencil_step_device(chefsi_ctx_t *ctx, const double *x, const double *xprev, double *out, int ncol,
                               double c, double s1, double s2);
int chefsi_nloc_project_device(chefsi_ctx_t *ctx, const double *x, int ncol, double *alpha_out);
int chefsi_nloc_expand_device(chefsi_ctx_t *ctx, double *out, int ncol, double scale, const double *alpha_in);
int chefsi_synchronize(chefsi_ctx_t *ctx);
/* Page-lock / unlock a caller-owned host range so the host entry points' chunk pipeline copies at
 * full PCIe rate and overlaps with the kernels (cudaHostRegister; the caller's malloc'd orbital
 * arrays, src/orbitalElecDensInit.c:364).  Optional: pageable memory works, only slower.  */
int chefsi_host_register(chefsi_ctx_t *ctx, void *ptr, size_t bytes);
int chefsi_host_unregister(chefsi_ctx_t *ctx, void *ptr);
int chefsi_pack_device(chefsi_ctx_t *ctx, const void *dense, size_t ld_dense, void *packed, int ncol,
                       int is_complex);
int chefsi_unpack_device(chefsi_ctx_t *ctx, const void *packed, void *dense, size_t ld_dense, int ncol,
                         int is_complex);

/* Fill ncol device columns with the synthetic start vectors of SURVEY.md section 8(d):
 * U(-0.5,0.5) from a counter-based generator keyed on (seed, global column, grid index),
 * so any column block is reproducible on CPU and GPU (mirrors Init_orbital,
 * src/orbitalElecDensInit.c:388-392).  is_complex != 0 fills (re,im) pairs.            */
int chefsi_fill_random_device(chefsi_ctx_t *ctx, void *buf, int ncol, long long first_col,
                              unsigned long long seed, int is_complex);

/* ---- introspection (timing, counters) ------------------------------------------------ */
typedef struct chefsi_stats {
    unsigned long long kernel_launches; /* kernels of this library launched so far         */
    double last_filter_ms;              /* device time of the last filter call (events)    */
    double last_stencil_ms;             /* summed device time of its stencil-step kernels  */
    double last_nloc_ms;                /* summed device time of its projector kernels     */
    int last_stencil_launches;
    int last_path;                      /* 0 = 3-D brick kernel, 1 = TMA streaming kernel (orthogonal), 2 = z-march kernel, 3 = TMA streaming kernel (non-orthogonal) */
    int last_nloc_atomic;               /* 1: the last projector expand took the overlapping-sphere branch      */
                                        /* (FP64 atomics, nlocVecRoutines.c:866-881 scatter-add)                */
    int last_alpha_reduced;             /* 1: per-atom alpha sums were formed by alpha_reduce_kernel            */
    unsigned int round_barrier_timeouts; /* times a streaming kernel's producer gave up on the round barrier (results are */
                                        /* unaffected; a non-zero count means the xy-halo L2 sharing was lost: a perf cliff) */
    int reserved_;
    unsigned int density_resident_blocks; /* chefsi_density_accumulate calls served from the band store (no upload)  */
    unsigned int density_uploaded_blocks; /* ... that had to upload their block from the host                        */
    unsigned int band_store_misses;       /* rotated blocks the band store could not keep (full, or out of memory)   */
    unsigned int reserved2_;
} chefsi_stats_t;
int chefsi_get_stats(const chefsi_ctx_t *ctx, chefsi_stats_t *out);
/* when on, every kernel of a filter call is bracketed by CUDA events (adds host
 * overhead; used by bench.py's roofline leg, off by default) */
int chefsi_set_profiling(chefsi_ctx_t *ctx, int on);
void *chefsi_stream(chefsi_ctx_t *ctx); /* cudaStream_t the context launches on */

#ifdef __cplusplus
}
#endif
#endif /* CHEFSI_B200_H */

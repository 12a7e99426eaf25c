/*
 * multi.cu -- one context that owns several GPUs of a box: SPARC's band-parallel axis inside one process.
 *
 * SPARC splits the columns of the orbital block over `npband` band communicators (NB = ceil(Ns / npband); rank r
 * owns columns [r NB, min((r+1) NB, Ns)), src/parallelization.c:403-428) and the filter has no communication
 * along that axis (eigenSolver.c:325 filters Nband_bandcomm local columns).  chefsi_create_multi makes a leader
 * context with one single-device child per GPU; the host-buffer entry points (ChebyshevFiltering,
 * Hamiltonian_vectors_mult, Lap_vec_mult) split their columns by the same rule and run every child's chunk
 * pipeline concurrently (one host thread per device, each with its own three streams).
 *
 * Replicated data -- Veff every SCF iteration, the projector tables once per ionic step -- is uploaded to the first
 * device only and BROADCAST over NVLink: ncclBroadcast (single process, ncclCommInitAll; the library is loaded with
 * dlopen so that the one-GPU path has no NCCL dependency), the counterpart of Transfer_Veff_loc's MPI_Bcast over
 * blacscomm (src/electronicGroundState.c:1313-1385).  When NCCL cannot be used (not installed, or the device list
 * names one device twice, which the single-GPU tests do) the same copies go through cudaMemcpyPeerAsync.
 */
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "chefsi_internal.h"

/* the handful of NCCL declarations used (nccl.h, NCCL 2.x ABI), resolved at run time */
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum { kNcclChar = 0 };
typedef ncclResult_t (*PFN_ncclCommInitAll)(ncclComm_t *, int, const int *);
typedef ncclResult_t (*PFN_ncclCommDestroy)(ncclComm_t);
typedef ncclResult_t (*PFN_ncclGroup)(void);
typedef ncclResult_t (*PFN_ncclBroadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char *(*PFN_ncclGetErrorString)(ncclResult_t);

struct MultiState {
    std::vector<chefsi_ctx *> kids;
    /* subspace steps: per device the column block of Hp / Mp / Q it owns (ncol x nc_r, column-major, ld = ncol) */
    std::vector<void *> d_hp, d_mp, d_q, d_qr, d_qi;
    size_t blk_bytes = 0;
    int sub_ncol = 0, sub_complex = 0;      /* the blocks currently resident on the kids */
    void *nccl_lib = nullptr;
    std::vector<ncclComm_t> comms;
    PFN_ncclCommDestroy commDestroy = nullptr;
    PFN_ncclGroup groupStart = nullptr, groupEnd = nullptr;
    PFN_ncclBroadcast broadcast = nullptr;
    PFN_ncclGetErrorString errStr = nullptr;
    bool use_nccl = false;
    unsigned long long bcast_calls = 0, bcast_bytes = 0;
};

static int multi_fail_from(chefsi_ctx *lead, chefsi_ctx *kid, int rc)
{
    if (rc && kid && kid != lead) {
        strncpy(lead->err, kid->err, sizeof(lead->err) - 1);
        lead->err[sizeof(lead->err) - 1] = 0;
    }
    return rc;
}

static void multi_try_nccl(chefsi_ctx *lead, const int *devices, int n)
{
    MultiState *ms = lead->multi;
    if (getenv("CHEFSI_B200_NO_NCCL")) return;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return; /* NCCL wants distinct devices */
    const char *names[] = {getenv("CHEFSI_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm) continue;
        ms->nccl_lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (ms->nccl_lib) break;
    }
    if (!ms->nccl_lib) return;
    PFN_ncclCommInitAll initAll = (PFN_ncclCommInitAll)dlsym(ms->nccl_lib, "ncclCommInitAll");
    ms->commDestroy = (PFN_ncclCommDestroy)dlsym(ms->nccl_lib, "ncclCommDestroy");
    ms->groupStart = (PFN_ncclGroup)dlsym(ms->nccl_lib, "ncclGroupStart");
    ms->groupEnd = (PFN_ncclGroup)dlsym(ms->nccl_lib, "ncclGroupEnd");
    ms->broadcast = (PFN_ncclBroadcast)dlsym(ms->nccl_lib, "ncclBroadcast");
    ms->errStr = (PFN_ncclGetErrorString)dlsym(ms->nccl_lib, "ncclGetErrorString");
    if (!initAll || !ms->commDestroy || !ms->groupStart || !ms->groupEnd || !ms->broadcast) return;
    ms->comms.assign(n, nullptr);
    if (initAll(ms->comms.data(), n, devices) != 0) { ms->comms.clear(); return; }
    ms->use_nccl = true;
}

/* bufs[0] on kid 0 holds the data; replicate `bytes` bytes into bufs[r] on every other kid, then wait */
static int multi_bcast(chefsi_ctx *lead, const std::vector<void *> &bufs, size_t bytes)
{
    MultiState *ms = lead->multi;
    const int n = (int)ms->kids.size();
    if (n < 2 || bytes == 0) return 0;
    ms->bcast_calls++;
    ms->bcast_bytes += bytes;
    if (ms->use_nccl) {
        ncclResult_t r = ms->groupStart();
        for (int i = 0; i < n && r == 0; i++) {
            cudaSetDevice(ms->kids[i]->device);
            r = ms->broadcast(bufs[0], bufs[i], bytes, kNcclChar, 0, ms->comms[i], ms->kids[i]->stream);
        }
        const ncclResult_t r2 = ms->groupEnd();
        if (r == 0) r = r2;
        if (r != 0) return chefsi_fail(lead, "ncclBroadcast: %s", ms->errStr ? ms->errStr(r) : "error");
    } else {
        for (int i = 1; i < n; i++) {
            CHEFSI_CUDA(lead, cudaSetDevice(ms->kids[i]->device));
            if (ms->kids[i]->device == ms->kids[0]->device)
                CHEFSI_CUDA(lead, cudaMemcpyAsync(bufs[i], bufs[0], bytes, cudaMemcpyDeviceToDevice, ms->kids[i]->stream));
            else
                CHEFSI_CUDA(lead, cudaMemcpyPeerAsync(bufs[i], ms->kids[i]->device, bufs[0], ms->kids[0]->device, bytes, ms->kids[i]->stream));
        }
    }
    for (int i = 0; i < n; i++) {
        CHEFSI_CUDA(lead, cudaSetDevice(ms->kids[i]->device));
        CHEFSI_CUDA(lead, cudaStreamSynchronize(ms->kids[i]->stream));
    }
    return 0;
}

extern "C" int chefsi_create_multi(chefsi_ctx_t **out, const int *devices, int ndev)
{
    if (!out || !devices || ndev < 1) return 1;
    *out = nullptr;
    chefsi_ctx *lead = new (std::nothrow) chefsi_ctx();
    if (!lead) return 1;
    lead->multi = new (std::nothrow) MultiState();
    if (!lead->multi) { delete lead; return 1; }
    lead->device = devices[0];
    for (int i = 0; i < ndev; i++) {
        chefsi_ctx *kid = nullptr;
        if (chefsi_create(&kid, devices[i]) != 0) {
            for (chefsi_ctx *k : lead->multi->kids) chefsi_destroy(k);
            delete lead->multi;
            delete lead;
            return 1; /* chefsi_last_error(NULL) has the reason */
        }
        lead->multi->kids.push_back(kid);
    }
    /* peer access for the non-NCCL replication path and for pinned buffers shared by all devices */
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < ndev; j++)
            if (devices[i] != devices[j]) {
                int can = 0;
                cudaSetDevice(devices[i]);
                if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                    if (e != cudaSuccess) cudaGetLastError(); /* already enabled */
                }
            }
    multi_try_nccl(lead, devices, ndev);
    lead->num_sms = lead->multi->kids[0]->num_sms;
    *out = lead;
    return 0;
}

/* single-column solvers (Lanczos, AAR) do not split: they run on the first device, which holds every replicated table */
chefsi_ctx *multi_first(chefsi_ctx *lead) { return lead->multi ? lead->multi->kids[0] : lead; }

int multi_size(const chefsi_ctx *lead) { return lead->multi ? (int)lead->multi->kids.size() : 1; }
int multi_uses_nccl(const chefsi_ctx *lead) { return lead->multi && lead->multi->use_nccl; }

void multi_destroy(chefsi_ctx *lead)
{
    MultiState *ms = lead->multi;
    for (size_t i = 0; i < ms->comms.size(); i++)
        if (ms->comms[i]) ms->commDestroy(ms->comms[i]);
    for (size_t r = 0; r < ms->kids.size(); r++) {
        cudaSetDevice(ms->kids[r]->device);
        if (r < ms->d_hp.size()) { cudaFree(ms->d_hp[r]); cudaFree(ms->d_mp[r]); cudaFree(ms->d_q[r]); cudaFree(ms->d_qr[r]); cudaFree(ms->d_qi[r]); }
    }
    for (chefsi_ctx *k : ms->kids) chefsi_destroy(k);
    if (ms->nccl_lib) dlclose(ms->nccl_lib);
    delete ms;
    lead->multi = nullptr;
    delete lead;
}

int multi_set_grid(chefsi_ctx *lead, const chefsi_grid_t *g)
{
    for (chefsi_ctx *k : lead->multi->kids) {
        const int rc = chefsi_set_grid(k, g);
        if (rc) return multi_fail_from(lead, k, rc);
    }
    chefsi_ctx *k0 = lead->multi->kids[0];
    lead->grid = k0->grid; lead->Nd = k0->Nd; lead->ld = k0->ld; lead->lay = k0->lay; lead->have_grid = true;
    return 0;
}

int multi_set_kpoint(chefsi_ctx *lead, double k1, double k2, double k3)
{
    for (chefsi_ctx *k : lead->multi->kids) {
        const int rc = chefsi_set_kpoint(k, k1, k2, k3);
        if (rc) return multi_fail_from(lead, k, rc);
    }
    return 0;
}

/* Veff: upload to the first device, broadcast (Transfer_Veff_loc's MPI_Bcast, electronicGroundState.c:1373) */
int multi_set_veff(chefsi_ctx *lead, const double *veff_host)
{
    MultiState *ms = lead->multi;
    int rc = chefsi_set_veff(ms->kids[0], veff_host);
    if (rc) return multi_fail_from(lead, ms->kids[0], rc);
    std::vector<void *> bufs;
    for (chefsi_ctx *k : ms->kids) { bufs.push_back(k->d_veff); k->have_veff = ms->kids[0]->have_veff; }
    if (!veff_host) return 0;
    return multi_bcast(lead, bufs, lead->Nd * sizeof(double));
}

void chefsi_free_nloc(NlocDev &d); /* chefsi_api.cu */

/* projector tables: built and uploaded on the first device, replicated device-to-device */
int multi_set_projectors(chefsi_ctx *lead, const chefsi_nloc_t *nl)
{
    MultiState *ms = lead->multi;
    chefsi_ctx *k0 = ms->kids[0];
    int rc = chefsi_set_projectors(k0, nl);
    if (rc) return multi_fail_from(lead, k0, rc);
    const NlocDev &s = k0->nl;
    for (size_t r = 1; r < ms->kids.size(); r++) {
        chefsi_ctx *k = ms->kids[r];
        CHEFSI_CUDA(lead, cudaSetDevice(k->device));
        chefsi_free_nloc(k->nl);
        if (s.n_img == 0) continue;
        NlocDev &d = k->nl;
        d = s; /* scalars; the pointers are replaced below */
        d.h_img_coords = (double *)malloc(sizeof(double) * 3 * (size_t)s.n_img);
        memcpy(d.h_img_coords, s.h_img_coords, sizeof(double) * 3 * (size_t)s.n_img);
        d.IP_displ = nullptr; d.gamma = nullptr; d.img_atom = nullptr; d.img_ndc = nullptr; d.pos_off = nullptr;
        d.chiT_off = nullptr; d.grid_pos = nullptr; d.chiT = nullptr; d.img_aoff = nullptr; d.img_phase = nullptr;
        d.atom_img_off = nullptr; d.atom_img = nullptr;
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.IP_displ, sizeof(int) * ((size_t)s.n_atom + 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.gamma, sizeof(double) * (size_t)(s.ntot ? s.ntot : 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.img_atom, sizeof(int) * (size_t)s.n_img));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.img_ndc, sizeof(int) * (size_t)s.n_img));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.pos_off, sizeof(long long) * ((size_t)s.n_img + 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.chiT_off, sizeof(long long) * ((size_t)s.n_img + 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.grid_pos, sizeof(int) * (size_t)(s.total_pts ? s.total_pts : 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.chiT, sizeof(double) * (size_t)(s.n_chiT ? s.n_chiT : 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.img_aoff, sizeof(int) * ((size_t)s.n_img + 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.img_phase, sizeof(double2) * (size_t)s.n_img));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.atom_img_off, sizeof(int) * ((size_t)s.n_atom + 1)));
        CHEFSI_CUDA(lead, cudaMalloc((void **)&d.atom_img, sizeof(int) * (size_t)s.n_img));
    }
    if (s.n_img == 0) return 0;
#define CHEFSI_BCAST_FIELD(field, bytes)                                                       \
    do {                                                                                       \
        std::vector<void *> bufs;                                                              \
        for (chefsi_ctx *k : ms->kids) bufs.push_back((void *)k->nl.field);                    \
        if (multi_bcast(lead, bufs, (bytes))) return 1;                                        \
    } while (0)
    CHEFSI_BCAST_FIELD(IP_displ, sizeof(int) * ((size_t)s.n_atom + 1));
    CHEFSI_BCAST_FIELD(gamma, sizeof(double) * (size_t)s.ntot);
    CHEFSI_BCAST_FIELD(img_atom, sizeof(int) * (size_t)s.n_img);
    CHEFSI_BCAST_FIELD(img_ndc, sizeof(int) * (size_t)s.n_img);
    CHEFSI_BCAST_FIELD(pos_off, sizeof(long long) * ((size_t)s.n_img + 1));
    CHEFSI_BCAST_FIELD(chiT_off, sizeof(long long) * ((size_t)s.n_img + 1));
    CHEFSI_BCAST_FIELD(grid_pos, sizeof(int) * (size_t)s.total_pts);
    CHEFSI_BCAST_FIELD(chiT, sizeof(double) * (size_t)s.n_chiT);
    CHEFSI_BCAST_FIELD(img_aoff, sizeof(int) * ((size_t)s.n_img + 1));
    CHEFSI_BCAST_FIELD(img_phase, sizeof(double2) * (size_t)s.n_img);
    CHEFSI_BCAST_FIELD(atom_img_off, sizeof(int) * ((size_t)s.n_atom + 1));
    CHEFSI_BCAST_FIELD(atom_img, sizeof(int) * (size_t)s.n_img);
#undef CHEFSI_BCAST_FIELD
    return 0;
}

/* run fn(kid, first column, ncol) on every kid's column slice concurrently: NB = ceil(ncol / n) (parallelization.c:403-428) */
template <class F>
static int multi_split(chefsi_ctx *lead, int ncol, F fn)
{
    MultiState *ms = lead->multi;
    const int n = (int)ms->kids.size();
    const int nb = (ncol + n - 1) / n;
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    for (int r = 1; r < n; r++) {
        const int c0 = r * nb < ncol ? r * nb : ncol, c1 = (r + 1) * nb < ncol ? (r + 1) * nb : ncol;
        if (c1 > c0) th.emplace_back([&, r, c0, c1] { rcs[r] = fn(ms->kids[r], c0, c1 - c0); });
    }
    rcs[0] = fn(ms->kids[0], 0, nb < ncol ? nb : ncol);
    for (std::thread &t : th) t.join();
    double ms_max = 0;
    unsigned long long launches = 0;
    for (int r = 0; r < n; r++) {
        if (ms->kids[r]->stats.last_filter_ms > ms_max) ms_max = ms->kids[r]->stats.last_filter_ms;
        launches += ms->kids[r]->stats.kernel_launches;
    }
    lead->stats = ms->kids[0]->stats;
    lead->stats.last_filter_ms = ms_max;
    lead->stats.kernel_launches = launches;
    for (int r = 0; r < n; r++)
        if (rcs[r]) return multi_fail_from(lead, ms->kids[r], rcs[r]);
    return 0;
}

int multi_filter_host(chefsi_ctx *lead, void *X, size_t ldi, void *Y, size_t ldo, int ncol, int m, double a, double b, double a0,
                      int flags, bool is_complex)
{
    const size_t esz = is_complex ? 16 : 8;
    return multi_split(lead, ncol, [=](chefsi_ctx *k, int c0, int nc) {
        char *x = (char *)X + (size_t)c0 * ldi * esz, *y = (char *)Y + (size_t)c0 * ldo * esz;
        return is_complex ? chefsi_chebyshev_filter_kpt(k, x, ldi, y, ldo, nc, m, a, b, a0, flags)
                          : chefsi_chebyshev_filter(k, (double *)x, ldi, (double *)y, ldo, nc, m, a, b, a0, flags);
    });
}

int multi_hmult_host(chefsi_ctx *lead, int ncol, double c, const void *x, size_t ldi, void *Hx, size_t ldo, bool is_complex)
{
    const size_t esz = is_complex ? 16 : 8;
    return multi_split(lead, ncol, [=](chefsi_ctx *k, int c0, int nc) {
        const char *xi = (const char *)x + (size_t)c0 * ldi * esz;
        char *yo = (char *)Hx + (size_t)c0 * ldo * esz;
        return is_complex ? chefsi_hamiltonian_mult_kpt(k, nc, c, xi, ldi, yo, ldo)
                          : chefsi_hamiltonian_mult(k, nc, c, (const double *)xi, ldi, (double *)yo, ldo);
    });
}

int multi_lapmult_host(chefsi_ctx *lead, int ncol, double a, double c, const void *x, size_t ldi, void *y, size_t ldo, bool is_complex)
{
    const size_t esz = is_complex ? 16 : 8;
    return multi_split(lead, ncol, [=](chefsi_ctx *k, int c0, int nc) {
        const char *xi = (const char *)x + (size_t)c0 * ldi * esz;
        char *yo = (char *)y + (size_t)c0 * ldo * esz;
        return is_complex ? chefsi_laplacian_mult_kpt(k, nc, a, c, xi, ldi, yo, ldo)
                          : chefsi_laplacian_mult(k, nc, a, c, (const double *)xi, ldi, (double *)yo, ldo);
    });
}

int multi_gradmult_host(chefsi_ctx *lead, int ncol, double c, const void *x, size_t ldi, void *Dx, size_t ldo, int dir, double kdir,
                        bool is_complex)
{
    const size_t esz = is_complex ? 16 : 8;
    return multi_split(lead, ncol, [=](chefsi_ctx *k, int c0, int nc) {
        const char *xi = (const char *)x + (size_t)c0 * ldi * esz;
        char *yo = (char *)Dx + (size_t)c0 * ldo * esz;
        return is_complex ? chefsi_gradient_mult_kpt(k, nc, c, xi, ldi, yo, ldo, dir, kdir)
                          : chefsi_gradient_mult(k, nc, c, (const double *)xi, ldi, (double *)yo, ldo, dir);
    });
}

int multi_synchronize(chefsi_ctx *lead)
{
    for (chefsi_ctx *k : lead->multi->kids) {
        const int rc = chefsi_synchronize(k);
        if (rc) return multi_fail_from(lead, k, rc);
    }
    return 0;
}

void multi_set_profiling(chefsi_ctx *lead, int on)
{
    for (chefsi_ctx *k : lead->multi->kids) k->profiling = on;
}

void multi_bcast_stats(const chefsi_ctx *lead, unsigned long long *calls, unsigned long long *bytes)
{
    *calls = lead->multi ? lead->multi->bcast_calls : 0;
    *bytes = lead->multi ? lead->multi->bcast_bytes : 0;
}

/* ---- Rayleigh-Ritz steps over several devices (SURVEY.md 8e "exchange step after", 8f-1) --------------------------------
 * The filtered block is split by columns: device r holds Y_r (and H Y_r).  Mp = Y^H Y and Hp = Y^H H Y need every pair of
 * column blocks, X = Y Q every block for every output block.  The reference re-distributes the block with MPI_Alltoallv
 * (BP2DP, src/parallelization.c:2535) before its dgemm; here nothing is re-distributed: device I computes the column
 * block I of Hp / Mp and of X, and its GEMM kernels READ the other devices' Y_J straight through NVLink peer memory
 * (the A operand of gemm_tn / gemm_nn is a peer pointer; cp.async pulls the tiles while the DMMAs of the previous chunk
 * run), i.e. the all-gather is fused into the product.  Only the Ns x Ns matrices and the rotated block cross PCIe. */
static void kid_range(int ncol, int n, int r, int *c0, int *nc)
{
    const int nb = (ncol + n - 1) / n;
    const int a = r * nb < ncol ? r * nb : ncol, b = (r + 1) * nb < ncol ? (r + 1) * nb : ncol;
    *c0 = a;
    *nc = b - a;
}

template <class F>
static int multi_parallel(chefsi_ctx *lead, F fn)
{
    MultiState *ms = lead->multi;
    const int n = (int)ms->kids.size();
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    for (int r = 1; r < n; r++) th.emplace_back([&, r] { rcs[r] = fn(r); });
    rcs[0] = fn(0);
    for (std::thread &t : th) t.join();
    for (int r = 0; r < n; r++)
        if (rcs[r]) return multi_fail_from(lead, ms->kids[r], rcs[r]);
    return 0;
}

int multi_subspace_reserve(chefsi_ctx *lead, int ncol, bool is_complex)
{
    MultiState *ms = lead->multi;
    const int n = (int)ms->kids.size();
    const int nb = (ncol + n - 1) / n;
    const size_t need = (size_t)ncol * nb * sizeof(double) * (is_complex ? 2 : 1);
    for (int r = 0; r < n; r++) {
        int c0, nc;
        kid_range(ncol, n, r, &c0, &nc);
        if (nc <= 0) continue;
        const int rc = is_complex ? chefsi_subspace_reserve_kpt(ms->kids[r], nc) : chefsi_subspace_reserve(ms->kids[r], nc);
        if (rc) return multi_fail_from(lead, ms->kids[r], rc);
    }
    if (need > ms->blk_bytes) {
        ms->d_hp.resize(n, nullptr); ms->d_mp.resize(n, nullptr); ms->d_q.resize(n, nullptr); ms->d_qr.resize(n, nullptr); ms->d_qi.resize(n, nullptr);
        for (int r = 0; r < n; r++) {
            CHEFSI_CUDA(lead, cudaSetDevice(ms->kids[r]->device));
            cudaFree(ms->d_hp[r]); cudaFree(ms->d_mp[r]); cudaFree(ms->d_q[r]); cudaFree(ms->d_qr[r]); cudaFree(ms->d_qi[r]);
            ms->d_hp[r] = ms->d_mp[r] = ms->d_q[r] = ms->d_qr[r] = ms->d_qi[r] = nullptr;
            CHEFSI_CUDA(lead, cudaMalloc(&ms->d_hp[r], need));
            CHEFSI_CUDA(lead, cudaMalloc(&ms->d_mp[r], need));
            CHEFSI_CUDA(lead, cudaMalloc(&ms->d_q[r], need));
            CHEFSI_CUDA(lead, cudaMalloc(&ms->d_qr[r], need));
            CHEFSI_CUDA(lead, cudaMalloc(&ms->d_qi[r], need));
        }
        ms->blk_bytes = need;
    }
    return 0;
}

int multi_subspace_project(chefsi_ctx *lead, const void *Y, size_t ldy, int ncol, void *Hp, void *Mp, size_t ldp, bool is_complex)
{
    MultiState *ms = lead->multi;
    const int n = (int)ms->kids.size();
    if (ncol <= 0 || ldp < (size_t)ncol || ldy < lead->Nd) return chefsi_fail(lead, "subspace_project: bad dimensions");
    if (multi_subspace_reserve(lead, ncol, is_complex)) return 1;
    const int words = is_complex ? 2 : 1;
    const size_t esz = sizeof(double) * words, row = lead->Nd * esz;
    /* phase 1 (per device): make Y_r resident if the filter did not leave it there, W_r = H Y_r, T_r = -i W_r .. done later */
    int rc = multi_parallel(lead, [&](int r) {
        chefsi_ctx *k = ms->kids[r];
        int c0, nc;
        kid_range(ncol, n, r, &c0, &nc);
        if (nc <= 0) return 0;
        const char *yslice = (const char *)Y + (size_t)c0 * ldy * esz;
        if (cudaSetDevice(k->device) != cudaSuccess) return chefsi_fail(k, "cudaSetDevice failed");
        if (!(k->res_ncol == nc && k->res_host == (const void *)yslice && k->res_complex == (int)is_complex)) {
            if (k->res_unwritten_host == (const void *)yslice)
                return chefsi_fail(k, "subspace_project: the host copy of this Y block was never written (NO_Y_COPYBACK) and the device copy is gone");
            if (cudaMemcpy2DAsync(k->d_res_Y, k->ld * esz, yslice, ldy * esz, row, nc, cudaMemcpyHostToDevice, k->stream) != cudaSuccess)
                return chefsi_fail(k, "subspace_project: upload of the Y block failed");
            k->res_ncol = nc; k->res_host = yslice; k->res_complex = is_complex;
        }
        const int rc1 = is_complex ? chefsi_hamiltonian_mult_kpt_device(k, nc, 0.0, k->d_res_Y, k->d_res_W)
                                   : chefsi_hamiltonian_mult_device(k, nc, 0.0, (const double *)k->d_res_Y, (double *)k->d_res_W);
        if (rc1) return rc1;
        return cudaStreamSynchronize(k->stream) == cudaSuccess ? 0 : chefsi_fail(k, "subspace_project: H Y failed");
    });
    if (rc) return rc;
    /* phase 2 (per device I): column block I of Mp and Hp; the A operand Y_J is read through peer memory.  Both matrices
       are Hermitian: of every off-diagonal block pair only one is formed (rank_forms_block), the diagonal blocks use
       upper-triangle tiles, and the host mirrors the rest below */
    rc = multi_parallel(lead, [&](int I) {
        chefsi_ctx *k = ms->kids[I];
        int c0I, ncI;
        kid_range(ncol, n, I, &c0I, &ncI);
        if (ncI <= 0) return 0;
        if (cudaSetDevice(k->device) != cudaSuccess) return chefsi_fail(k, "cudaSetDevice failed");
        const size_t K = lead->Nd * words, ldv = k->ld * words;
        double *dMp = (double *)ms->d_mp[I], *dHp = (double *)ms->d_hp[I];
        const double *Yi = (const double *)k->d_res_Y, *Wi = (const double *)k->d_res_W;
        for (int pass = 0; pass < (is_complex ? 2 : 1); pass++) {
            const double *By = Yi, *Bw = Wi;
            if (pass == 1) { /* imaginary parts: B -> -i B, formed locally */
                if (launch_rot90(k, k->d_res_Y, k->d_res_T, lead->Nd, k->ld, ncI, -1.0) < 0) return 1;
                By = (const double *)k->d_res_T;
            }
            for (int which = 0; which < 2; which++) { /* 0: Mp (B = Y_I), 1: Hp (B = H Y_I) */
                if (pass == 1 && which == 1) {
                    if (launch_rot90(k, k->d_res_W, k->d_res_T, lead->Nd, k->ld, ncI, -1.0) < 0) return 1;
                    Bw = (const double *)k->d_res_T;
                }
                const double *B = which ? Bw : By;
                double *Cblk = (which ? dHp : dMp) + pass; /* interleaved complex: imaginary parts at +1 */
                for (int J = 0; J < n; J++) {
                    int c0J, ncJ;
                    kid_range(ncol, n, J, &c0J, &ncJ);
                    if (ncJ <= 0) continue;
                    int r0, r1, q0, q1; /* the part of block (J, I) device I forms */
                    rank_block_part(J, I, n, ncJ, ncI, &r0, &r1, &q0, &q1);
                    if (r1 <= r0 || q1 <= q0) continue;
                    const size_t lda = ms->kids[J]->ld * words;
                    const double *A = (const double *)ms->kids[J]->d_res_Y; /* peer pointer when J != I */
                    const int sym = J == I ? (pass == 0 ? +1 : -1) : 0; /* real part symmetric, imaginary part antisymmetric */
                    const int nl = launch_gemm_tn(k, A + (size_t)r0 * lda, lda, B + (size_t)q0 * ldv, ldv, r1 - r0, q1 - q0, K, 1.0,
                                                  Cblk + ((size_t)q0 * ncol + c0J + r0) * words, ncol, words, sym);
                    if (nl < 0) return 1;
                    k->stats.kernel_launches += nl;
                }
                if (pass == 1 && which == 0 && cudaStreamSynchronize(k->stream) != cudaSuccess) return chefsi_fail(k, "subspace_project: GEMM failed"); /* d_res_T is reused */
            }
        }
        const size_t w = (size_t)ncol * esz;
        if (cudaMemcpy2DAsync((char *)Mp + (size_t)c0I * ldp * esz, ldp * esz, dMp, w, w, ncI, cudaMemcpyDeviceToHost, k->stream) != cudaSuccess ||
            cudaMemcpy2DAsync((char *)Hp + (size_t)c0I * ldp * esz, ldp * esz, dHp, w, w, ncI, cudaMemcpyDeviceToHost, k->stream) != cudaSuccess ||
            cudaStreamSynchronize(k->stream) != cudaSuccess)
            return chefsi_fail(k, "subspace_project: copy of Hp / Mp failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    });
    if (rc) return rc;
    /* every element (rows of J, columns of I) that device I left out = conjugate of its mirror image, which device J formed */
    for (int I = 0; I < n; I++)
        for (int J = 0; J < n; J++) {
            if (J == I) continue;
            int c0I, ncI, c0J, ncJ, r0, r1, q0, q1;
            kid_range(ncol, n, I, &c0I, &ncI);
            kid_range(ncol, n, J, &c0J, &ncJ);
            if (ncI <= 0 || ncJ <= 0) continue;
            rank_block_part(J, I, n, ncJ, ncI, &r0, &r1, &q0, &q1);
            for (int which = 0; which < 2; which++) {
                double *M = (double *)(which ? Hp : Mp);
                for (int c = 0; c < ncI; c++)
                    for (int r = 0; r < ncJ; r++) {
                        if (r >= r0 && r < r1 && c >= q0 && c < q1) continue; /* formed by device I */
                        double *dst = M + ((size_t)(c0I + c) * ldp + c0J + r) * words;
                        const double *src = M + ((size_t)(c0J + r) * ldp + c0I + c) * words;
                        dst[0] = src[0];
                        if (is_complex) dst[1] = -src[1];
                    }
            }
        }
    ms->sub_ncol = ncol;
    ms->sub_complex = is_complex;
    return 0;
}

int multi_subspace_rotate(chefsi_ctx *lead, const void *Q, size_t ldq, int ncol, void *X, size_t ldx, bool is_complex)
{
    MultiState *ms = lead->multi;
    const int n = (int)ms->kids.size();
    if (ms->sub_ncol != ncol || ms->sub_complex != (int)is_complex)
        return chefsi_fail(lead, "subspace_rotate: no resident block of %d columns (call chefsi_subspace_project first)", ncol);
    if (!Q) return chefsi_fail(lead, "subspace_rotate: a multi-device context needs the host copy of Q");
    if (ldq < (size_t)ncol || ldx < lead->Nd) return chefsi_fail(lead, "subspace_rotate: bad dimensions");
    const int words = is_complex ? 2 : 1;
    const size_t esz = sizeof(double) * words;
    /* phase 1: complex data needs i Y_J next to Y_J on every device before anybody multiplies */
    if (is_complex) {
        const int rc = multi_parallel(lead, [&](int r) {
            chefsi_ctx *k = ms->kids[r];
            int c0, nc;
            kid_range(ncol, n, r, &c0, &nc);
            if (nc <= 0) return 0;
            if (cudaSetDevice(k->device) != cudaSuccess) return chefsi_fail(k, "cudaSetDevice failed");
            if (launch_rot90(k, k->d_res_Y, k->d_res_T, lead->Nd, k->ld, nc, 1.0) < 0) return 1;
            return cudaStreamSynchronize(k->stream) == cudaSuccess ? 0 : chefsi_fail(k, "subspace_rotate: rotation pass failed");
        });
        if (rc) return rc;
    }
    /* phase 2 (per device I): X_I = sum_J Y_J Q[J, I], Y_J read through peer memory, accumulated in W_I */
    const int rc = multi_parallel(lead, [&](int I) {
        chefsi_ctx *k = ms->kids[I];
        int c0I, ncI;
        kid_range(ncol, n, I, &c0I, &ncI);
        if (ncI <= 0) return 0;
        if (cudaSetDevice(k->device) != cudaSuccess) return chefsi_fail(k, "cudaSetDevice failed");
        const size_t w = (size_t)ncol * esz;
        double *dQ = (double *)ms->d_q[I];
        if (cudaMemcpy2DAsync(dQ, w, (const char *)Q + (size_t)c0I * ldq * esz, ldq * esz, w, ncI, cudaMemcpyHostToDevice, k->stream) != cudaSuccess)
            return chefsi_fail(k, "subspace_rotate: upload of Q failed");
        const double *Qr = dQ, *Qi = nullptr;
        if (is_complex) {
            if (launch_split_complex(k, dQ, ncol, ncol, ncI, (double *)ms->d_qr[I], (double *)ms->d_qi[I]) < 0) return 1;
            Qr = (const double *)ms->d_qr[I];
            Qi = (const double *)ms->d_qi[I];
        }
        const size_t K = lead->Nd * words;
        int first = 1;
        for (int J = 0; J < n; J++) {
            int c0J, ncJ;
            kid_range(ncol, n, J, &c0J, &ncJ);
            if (ncJ <= 0) continue;
            chefsi_ctx *kj = ms->kids[J];
            int nl = launch_gemm_nn(k, (const double *)kj->d_res_Y, kj->ld * words, Qr + c0J, ncol, K, ncJ, ncI, (double *)k->d_res_W, k->ld * words, first ? 0 : 1);
            if (nl < 0) return 1;
            if (is_complex) {
                nl = launch_gemm_nn(k, (const double *)kj->d_res_T, kj->ld * words, Qi + c0J, ncol, K, ncJ, ncI, (double *)k->d_res_W, k->ld * words, 1);
                if (nl < 0) return 1;
            }
            first = 0;
            k->stats.kernel_launches += is_complex ? 2 : 1;
        }
        if (cudaMemcpy2DAsync((char *)X + (size_t)c0I * ldx * esz, ldx * esz, k->d_res_W, k->ld * esz, lead->Nd * esz, ncI, cudaMemcpyDeviceToHost, k->stream) != cudaSuccess ||
            cudaStreamSynchronize(k->stream) != cudaSuccess)
            return chefsi_fail(k, "subspace_rotate: GEMM / copy of the rotated block failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    });
    if (rc) return rc;
    for (chefsi_ctx *k : ms->kids) k->res_ncol = 0; /* consumed */
    ms->sub_ncol = 0;
    return 0;
}

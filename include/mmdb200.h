/*
 * mmdb200.h — C ABI of libmmdb200.so: the B200 (sm_100a) two-electron engine that replaces the
 * hot path of jjgoings/McMurchie-Davidson.
 *
 * Boundary: this is what the reference's Python layer binds (via ctypes)
 * instead of its Cython modules.  Plain ints / doubles / pointers only; every entry point returns
 * 0 on success and a non-zero code on failure, with mmdb_last_error() giving a thread-local
 * message.  There is no CPU fallback: without a CUDA device every compute entry fails.
 *
 * Reference interfaces replaced (paths into the reference tree):
 *   cython/basis.pxi:6-120   cdef class Basis          -> mmdb_basis_create (shell table)
 *   cython/twoe.pyx:36-50    ERI(a,b,c,d)              -> mmdb_eri_shell_quartets
 *   cython/twoe.pyx:12-31    doERIs(N,TwoE,bfs)        -> mmdb_eri_dense
 *   mmd/molecule.py:95-99    Schwarz dict (pq|pq)      -> mmdb_schwarz
 *   cython/fock.pyx:13-87    formPT(P,P_old,...)       -> mmdb_fock_direct
 *   mmd/scf.py:97-98         einsum J/K                -> mmdb_jk_incore
 *   cython/onee.pyx:14-192   S,T,V,Mu,RxDel            -> mmdb_onee   (enabler)
 *
 * Conventions
 *   - "device function index": basis functions are numbered shell by shell, Cartesian components
 *     in the reference's order (mmd/molecule.py:108-114): p = x,y,z ; d = xx,xy,xz,yy,yz,zz.
 *     bf0[s] is the first function of shell s; all matrices (P, G, Q, TwoE) use these indices.
 *   - shells carry per-primitive coefficients c_k = N_k(L) * d_k * N_contr WITHOUT the
 *     per-component factor 1/sqrt((2l-1)!!(2m-1)!!(2n-1)!!) of cython/basis.pxi:102-105; the
 *     library applies that factor per Cartesian component on output.
 *   - pointers named *_dev are CUDA device pointers on the handle's device; `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream).  Calls with device pointers are
 *     asynchronous on `stream`; calls with host pointers synchronise before returning.
 */
#ifndef MMDB200_H
#define MMDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmdb_basis mmdb_basis; /* opaque: shell table + shell-pair tables + work buffers on one GPU */

#define MMDB_OK 0
#define MMDB_ERR_INVALID 1
#define MMDB_ERR_CUDA 2
#define MMDB_ERR_UNSUPPORTED 3
#define MMDB_ERR_NOMEM 4

#define MMDB_MAX_AM 2       /* s, p, d  (north star: (ss|ss) ... (dd|dd)) */
#define MMDB_NCLASS_PAIR 6  /* ss ps pp ds dp dd, index = la*(la+1)/2 + lb, la >= lb */

int mmdb_version(void);
const char *mmdb_last_error(void);
int mmdb_device_count(int *count);

/* ---- basis ------------------------------------------------------------------------------- */
/* nshell shells; shell s: angular momentum am[s] (0..2), centre[3s..3s+2] (bohr), primitives
 * prim_off[s] .. prim_off[s]+nprim[s]-1 into exps[]/coefs[] (coefs as described above),
 * first device function index bf0[s].  prim_cut: primitive pairs whose magnitude estimate
 * |c_a c_b| exp(-mu |AB|^2) (pi/p)^1.5-scaled falls below prim_cut are dropped (<= 0 keeps all). */
int mmdb_basis_create(int device, int nshell, const int *am, const int *nprim, const int *prim_off,
                      const double *centre, const double *exps, const double *coefs, const int *bf0,
                      double prim_cut, mmdb_basis **out);
int mmdb_basis_destroy(mmdb_basis *b);
int mmdb_basis_nbf(const mmdb_basis *b, int *nbf);
/* number of shell pairs / primitive pairs kept per pair class (arrays of MMDB_NCLASS_PAIR) */
int mmdb_basis_pair_counts(const mmdb_basis *b, int64_t *npairs, int64_t *nprimpairs);
/* shell pair p of class pc -> (shell A, shell B), am[A] >= am[B] */
int mmdb_basis_pair_shells(const mmdb_basis *b, int pc, int *shA, int *shB);

/* ---- ERIs -------------------------------------------------------------------------------- */
/* Contracted shell quartets (bra pair | ket pair) of one class: bra pairs index pair class pc_bra,
 * ket pairs index class pc_ket.  out_dev[q * nfn + ((a*nb+b)*nc+c)*nd+d] receives (ab|cd) for
 * quartet q with the per-component normalisation applied; nfn = na*nb*nc*nd of the class.
 * impl: 0 = default (register-resident class-specialised kernel where one exists),
 *       1 = force the generic runtime-L kernel (cross-check). */
int mmdb_eri_shell_quartets(mmdb_basis *b, int pc_bra, int pc_ket, int64_t n, const int32_t *bra_idx_dev,
                            const int32_t *ket_idx_dev, double *out_dev, int impl, void *stream);

/* Schwarz table (mmd/molecule.py:95-99): Q_dev[p*N+q] = (pq|pq), symmetric, N = nbf.  Also caches
 * sqrt(Q) and the per-shell-pair maxima inside the handle for mmdb_fock_direct. */
int mmdb_schwarz(mmdb_basis *b, double *Q_dev, void *stream);

/* Install a caller-supplied Schwarz table instead (the reference lets the caller pass any `screen`
 * dict to formPT): Q_tri[p(p+1)/2+q] for p >= q, host memory. */
int mmdb_set_schwarz_host(mmdb_basis *b, const double *Q_tri);

/* Dense (N,N,N,N) row-major tensor, all 8 permutational images written (cython/twoe.pyx:12-31).
 * TwoE_dev must hold N^4 doubles; every element is written (no pre-zeroing needed). */
int mmdb_eri_dense(mmdb_basis *b, double *TwoE_dev, void *stream);

/* ---- Fock builds ------------------------------------------------------------------------- */
/* In-core J/K (mmd/scf.py:97-98): J_pq = sum_rs (pq|rs) P_sr, K_pq = sum_rs (ps|qr) P_sr.
 * P/J/K are (N,N) row-major real planes; the imaginary plane pointers may be NULL (real density).
 * One pass over TwoE serves both planes. */
int mmdb_jk_incore(int device, const double *TwoE_dev, int N, const double *P_re_dev, const double *P_im_dev,
                   double *J_re_dev, double *J_im_dev, double *K_re_dev, double *K_im_dev, void *stream);

typedef struct {
    int64_t candidates;       /* shell quartets examined by the screen on this shard            */
    int64_t quartets;         /* contracted shell quartets evaluated (survived shell-level bound) */
    int64_t prim_quartets;    /* primitive shell quartets evaluated                              */
    int64_t fn_quartets;      /* contracted basis-function quartets produced                     */
    int64_t slow_quartets;    /* of `quartets`: digested per function (diagonal-type / complex)  */
    int64_t launches;         /* kernels launched by this call (2 density screens + 4 per class-pair chunk) */
    double model_flops;       /* sum over classes of prim_quartets(class) * F(class), SURVEY §8d */
    int64_t class_quartets[MMDB_NCLASS_PAIR * MMDB_NCLASS_PAIR];
    int64_t class_prim_quartets[MMDB_NCLASS_PAIR * MMDB_NCLASS_PAIR];
    float class_ms[MMDB_NCLASS_PAIR * MMDB_NCLASS_PAIR];        /* ERI+digestion kernel time per class (flags bit0) */
    float class_screen_ms[MMDB_NCLASS_PAIR * MMDB_NCLASS_PAIR]; /* screening kernel time per class (flags bit0)     */
    int64_t far_entries;      /* list entries (virtual bra pair, ket pair) on the far-field lists: every primitive quartet asymptotic */
    int64_t near_entries;     /* block-digestible list entries with at least one primitive quartet in the tabulated Boys range */
    int64_t exec_prim_quartets; /* primitive quartets actually evaluated: generally contracted s shells share theirs, so this is
                                   below prim_quartets (which, like quartets, is counted in the reference's own shells) unless flags bit2 */
} mmdb_fock_stats;

/* Direct Fock build (cython/fock.pyx:13-87): G += contributions of every canonical basis-function
 * quartet i>=j, k>=l, ij>=kl with sqrt(Q_ij) sqrt(Q_kl) max|{4dP_ij,4dP_kl,dP_ik,dP_il,dP_jk,dP_jl}| >= tol,
 * eri scaled by its degeneracy, six updates of fock.pyx:79-85 — G is the reference's
 * UN-symmetrised matrix.  dP = P - P_old is formed by the caller.  mmdb_schwarz must have been
 * called on the handle.  G planes must be zeroed by the caller (the call accumulates), which lets
 * shards of one build add into one buffer.  Work is restricted to shard `shard` of `nshards`
 * (static cost-balanced split: ket-pair row j of every class pair belongs to shard j % nshards); nshards = 1 does the
 * whole build.
 * flags: bit0 = time each class launch with CUDA events (fills stats->class_ms; synchronises);
 *        bit1 = DETERMINISTIC accumulation: contributions are rounded to multiples of 2^-50 and added with 64-bit
 *               integer atomics, so G is bitwise reproducible for any schedule / shard count.  G then holds scaled
 *               integers: sum the shards as int64 and finish with mmdb_fixed_to_double.
 *        bit2 = plain shell classes only.  By default generally contracted s shells (two consecutive s shells on one
 *               centre over the same primitives, e.g. cc-pVDZ) are evaluated as ONE two-component pseudo-shell so their
 *               primitive integrals are computed once; G is the same, the shell-level screen is a superset.  With bit2
 *               every reference shell is its own shell: the statistics are then exactly the reference's shell quartets.
 * Streams: the call is asynchronous with respect to the host and ordered on `stream` — everything enqueued on
 * `stream` after it sees the complete G.  Internally it forks onto two handle-owned streams (small class pairs;
 * the screening pipeline that runs one class pair ahead of the ERI kernels) and joins them back with events,
 * so ONE build per handle may be in flight at a time.  With `stats` (or flags bit0) the call synchronises. */
int mmdb_fock_direct(mmdb_basis *b, const double *dP_re_dev, const double *dP_im_dev, double tol,
                     double *G_re_dev, double *G_im_dev, int shard, int nshards, int flags,
                     mmdb_fock_stats *stats, void *stream);

int mmdb_fixed_to_double(int device, double *G_dev, int64_t n, void *stream);

/* ---- multi-GPU: the one exchange step of the sharded direct build ------------------------------------
 * One process (or thread) per GPU builds its shard with mmdb_fock_direct(..., shard, nshards, ...) into its own G and
 * then sums the partial matrices: an FP64 sum all-reduce over NCCL (NVLink 5 / NVSwitch).  Rank 0 obtains a unique id
 * (128 bytes) and hands it to the other ranks through the host program's own rendezvous; every rank then calls
 * mmdb_comm_init.  fixed_point != 0 reduces the 2^50-scaled 64-bit integers of a deterministic build (flags bit1)
 * exactly; finish with mmdb_fixed_to_double.  The shard rule (lib.cu screen kernels): KET pair row j of every class
 * pair belongs to shard j % nshards; rows are ordered by contraction depth, so the deal is cost-balanced. */
typedef struct mmdb_comm mmdb_comm;
int mmdb_comm_unique_id(unsigned char *id128);
int mmdb_comm_init(int device, int nranks, int rank, const unsigned char *id128, mmdb_comm **out);
int mmdb_allreduce_G(mmdb_comm *comm, double *G_dev, int64_t n, int fixed_point, void *stream);
int mmdb_comm_destroy(mmdb_comm *comm);

/* Host-buffer convenience forms (the reference-facing calls: host numpy in, host numpy out;
 * H2D/D2H inside).  P, P_old, G are complex128 interleaved (N,N) like the reference's arrays;
 * Q is the reference's triangular table of N(N+1)/2 values keyed p(p+1)/2+q. */
int mmdb_formPT_host(mmdb_basis *b, const double *P_c128, const double *P_old_c128, double tol,
                     double *G_c128, mmdb_fock_stats *stats);
/* the two host passes mmdb_formPT_host makes, exported for bindings that stage the planes themselves (multi-GPU):
 * re/im[x] = P[x] - P_old[x] (cython/fock.pyx:24), *has_im = any non-zero imaginary difference; and the inverse
 * interleave of the G planes (im may be NULL). */
int mmdb_c128_diff_split_host(const double *P_c128, const double *P_old_c128, int64_t n, double *re, double *im, int *has_im);
int mmdb_c128_join_host(const double *re, const double *im, int64_t n, double *out_c128);
int mmdb_schwarz_host(mmdb_basis *b, double *Q_tri);
int mmdb_eri_dense_host(mmdb_basis *b, double *TwoE_host);

/* ---- one-electron integrals (enabler; not on the graded path) ---------------------------- */
/* S, T, V (N,N); M (3,N,N) dipole about `origin`; L (3,N,N) RxDel about `origin`
 * (cython/onee.pyx; mmd/molecule.py:253-276).  natom nuclei with charges Z and positions xyz. */
int mmdb_onee_host(mmdb_basis *b, int natom, const double *Z, const double *xyz, const double *origin,
                   double *S, double *T, double *V, double *M, double *L);

/* ---- nuclear gradient (cython/grad.pyx + mmd/forces.py of the reference) ------------------------
 * dE/dX of the closed-shell RHF energy, [natom][3] each (host): one-electron part (kinetic, nuclear attraction incl. the
 * Hellmann-Feynman operator term, overlap x energy-weighted density), two-electron part (derivative ERIs contracted with
 * the two-particle density 16 P_ij P_kl - 4 P_ik P_jl - 4 P_il P_jk as they are produced — no derivative tensor), and
 * nuclear repulsion.  P = C_occ C_occ^T (no factor 2) and W = P F P: real (N,N) host matrices in device function order;
 * shell_atom[s] = atom index of shell s.  Forces are minus the sum of the three parts.  d shells use f-type shifted
 * primitives inside the kernels.  mmdb_schwarz must have been called. */
int mmdb_gradient_host(mmdb_basis *b, int natom, const double *Z, const double *xyz, const int *shell_atom,
                       const double *P, const double *W, double *grad_1e, double *grad_2e, double *grad_nuc);

/* ---- post-SCF consumers of the dense tensor (next row) ------------------- */
/* AO->MO transformation (mmd/postscf.py:21-41) by four cuBLAS DGEMM quarter transformations and the
 * closed-shell MP2 correlation energy (mmd/postscf.py:59-70).  Row-major (N,N,N,N) tensors; C_dev (N,N)
 * real MO coefficients (column = orbital); work_dev = N^4 doubles of scratch; e2_host may be NULL. */
int mmdb_ao2mo_mp2(int device, const double *TwoE_dev, int N, int nocc, const double *C_dev, const double *eps_dev,
                   double *MO_dev, double *work_dev, double *e2_host, void *stream);

/* ---- utilities --------------------------------------------------------------------------- */
/* F_0..F_mmax(T[i]) evaluated on the device with the kernels' Boys routine: out[i*(mmax+1)+m]. */
int mmdb_boys_host(int device, int mmax, int64_t n, const double *T, double *out);
/* The same through the class kernels' TEMPLATED routine for total angular momentum L (0..8): Taylor table below the
 * class's own switch-over T_max(L) = 37,41,44,47,50,53,55,58,60, alpha-free asymptotic series at and above it
 * (csrc/core.cuh prim_Fs<L>).  out[i*(L+1)+m] = F_m(T[i]). */
int mmdb_boys_class_host(int device, int L, int64_t n, const double *T, double *out);
/* Peak FP64 FMA issue rate probe: runs a register-resident DFMA loop, returns TFLOP/s. */
int mmdb_fp64_peak(int device, double *tflops, float *ms);
/* FLOP model of SURVEY §8(d): flops per primitive shell quartet of class (la lb|lc ld). */
double mmdb_class_flops(int la, int lb, int lc, int ld);

#ifdef __cplusplus
}
#endif
#endif /* MMDB200_H */

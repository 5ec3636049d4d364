/*
 * md_oracle.c — CPU restatement of the reference's two-electron hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path; it is never on
 * the product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it (see oracle/README.md).  Parity status: PINNED — tests/test_oracle.py checks this
 * restatement against (a) the 50 Boys known-answer values of the reference's tests/test012.py,
 * (b) golden vectors produced by the reference itself built in-container (oracle/_ref via
 * oracle/build_ref.py; fixtures + generating script under tests/golden/).
 *
 * Each function cites the reference file:line it restates (paths relative to the reference tree).
 * The arithmetic follows the reference formula-for-formula (same recursions, same loop nests, same
 * accumulation order); the only deliberate departures are
 *   - boys(): the reference calls SciPy's hyp1f1 (third-party, scipy>=0.19 per requirements.txt:2;
 *     SciPy 1.18.1 in this image).  Restated here from the published series
 *     F_m(T) = exp(-T) * sum_k (2T)^k / ((2m+1)(2m+3)...(2m+2k+1))        (Kummer; all terms > 0)
 *     evaluated in long double, with erf + upward recursion for integer m at large T and the
 *     asymptotic Gamma(m+1/2)/(2 T^(m+1/2)) beyond;
 *   - R(): leaves read a per-call table (-2p)^n F_n(T) instead of re-evaluating boys at every leaf
 *     (identical values, exponentially fewer hyp1f1 evaluations).
 *
 * Build: make -C oracle   ->  oracle/libmdoracle.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>


/* ---- tiny pthread parallel-for (no libgomp in this image); used only for the embarrassingly
   parallel fills below, never for accumulations whose order matters ------------------------- */
#include <pthread.h>
#include <unistd.h>
typedef void (*mdo_body)(long i, void *ctx);
typedef struct { long n; volatile long next; mdo_body fn; void *ctx; } mdo_pf;
static void *pf_worker(void *arg)
{
    mdo_pf *pf = (mdo_pf *)arg;
    for (;;) {
        long i = __sync_fetch_and_add(&pf->next, 1);
        if (i >= pf->n) break;
        pf->fn(i, pf->ctx);
    }
    return NULL;
}
static int g_threads = 0;
void mdo_set_threads(int n) { g_threads = n; }
int mdo_get_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)(n > 64 ? 64 : n) : 1;
}
static void parallel_for(long n, mdo_body fn, void *ctx)
{
    int nt = mdo_get_threads();
    if (nt > n) nt = (int)n;
    mdo_pf pf = {n, 0, fn, ctx};
    if (nt <= 1) { pf_worker(&pf); return; }
    pthread_t th[64];
    for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, pf_worker, &pf);
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

#define MDO_PI 3.141592653589793238462643383279 /* cython/util.pxi:7 */
#define MDO_MAXN 32

/* ---- cython/util.pxi:54-55  boys(m,T) = hyp1f1(m+1/2, m+3/2, -T)/(2m+1) ------------------- */
static long double boys_series(long double m, long double T)
{
    long double term = 1.0L / (2.0L * m + 1.0L);
    long double sum = term;
    for (int k = 1; k < 20000; ++k) {
        term *= 2.0L * T / (2.0L * m + 2.0L * k + 1.0L);
        sum += term;
        if (term < 1e-22L * sum) break;
    }
    return expl(-T) * sum;
}

double mdo_boys(double m, double T)
{
    if (T < 0) T = 0;
    if (T <= 45.0) return (double)boys_series((long double)m, (long double)T);
    long double Tl = (long double)T;
    double mi = floor(m);
    if (mi == m && m >= 0 && m < MDO_MAXN) {
        /* F_0 = sqrt(pi/T)/2 erf(sqrt T); F_{k+1} = ((2k+1) F_k - exp(-T)) / (2T)  (stable: T >> k) */
        if (T < 2.0 * m + 40.0) return (double)boys_series((long double)m, Tl);
        long double F = 0.5L * sqrtl((long double)MDO_PI / Tl) * erfl(sqrtl(Tl));
        long double eT = (T < 11000.0) ? expl(-Tl) : 0.0L;
        for (int k = 0; k < (int)m; ++k) F = ((2.0L * k + 1.0L) * F - eT) / (2.0L * Tl);
        return (double)F;
    }
    if (T <= 700.0) return (double)boys_series((long double)m, Tl);
    return (double)(tgammal((long double)m + 0.5L) / (2.0L * powl(Tl, (long double)m + 0.5L)));
}

/* ---- cython/util.pxi:13-28  E(i,j,t,Qx,a,b), n = 0 branch --------------------------------- */
double mdo_E(int i, int j, int t, double Qx, double a, double b)
{
    double p = a + b;
    double u = a * b / p;
    if (t < 0 || t > i + j) return 0.0;
    if (i == 0 && j == 0 && t == 0) return exp(-u * Qx * Qx);
    if (j == 0)
        return (1 / (2 * p)) * mdo_E(i - 1, j, t - 1, Qx, a, b) - (u * Qx / a) * mdo_E(i - 1, j, t, Qx, a, b) +
               (t + 1) * mdo_E(i - 1, j, t + 1, Qx, a, b);
    return (1 / (2 * p)) * mdo_E(i, j - 1, t - 1, Qx, a, b) + (u * Qx / b) * mdo_E(i, j - 1, t, Qx, a, b) +
           (t + 1) * mdo_E(i, j - 1, t + 1, Qx, a, b);
}

/* ---- cython/util.pxi:33-50  R(t,u,v,n,p,PCx,PCy,PCz,RPC) ---------------------------------- */
/* leaf[n] = pow(-2p,n)*boys(n,T) precomputed by the caller (see header). */
static double R_rec(int t, int u, int v, int n, const double *leaf, double X, double Y, double Z)
{
    double val = 0.0;
    if (t == 0 && u == 0 && v == 0) {
        val += leaf[n];
    } else if (t == 0 && u == 0) {
        if (v > 1) val += (v - 1) * R_rec(t, u, v - 2, n + 1, leaf, X, Y, Z);
        val += Z * R_rec(t, u, v - 1, n + 1, leaf, X, Y, Z);
    } else if (t == 0) {
        if (u > 1) val += (u - 1) * R_rec(t, u - 2, v, n + 1, leaf, X, Y, Z);
        val += Y * R_rec(t, u - 1, v, n + 1, leaf, X, Y, Z);
    } else {
        if (t > 1) val += (t - 1) * R_rec(t - 2, u, v, n + 1, leaf, X, Y, Z);
        val += X * R_rec(t - 1, u, v, n + 1, leaf, X, Y, Z);
    }
    return val;
}

static void R_leaves(int nmax, double p, double RPC, double *leaf)
{
    double T = p * RPC * RPC; /* util.pxi:34 */
    for (int n = 0; n <= nmax; ++n) leaf[n] = pow(-2 * p, n) * mdo_boys((double)n, T); /* util.pxi:37 */
}

double mdo_R(int t, int u, int v, int n, double p, double PCx, double PCy, double PCz, double RPC)
{
    double leaf[MDO_MAXN];
    int nmax = n + t + u + v;
    if (nmax >= MDO_MAXN) return NAN;
    R_leaves(nmax, p, RPC, leaf);
    return R_rec(t, u, v, n, leaf, PCx, PCy, PCz);
}

/* ---- cython/twoe.pyx:56-94  electron_repulsion -------------------------------------------- */
double mdo_electron_repulsion(double a, const long *lmn1, const double *A, double b, const long *lmn2,
                              const double *B, double c, const long *lmn3, const double *C, double d,
                              const long *lmn4, const double *D)
{
    long l1 = lmn1[0], m1 = lmn1[1], n1 = lmn1[2];
    long l2 = lmn2[0], m2 = lmn2[1], n2 = lmn2[2];
    long l3 = lmn3[0], m3 = lmn3[1], n3 = lmn3[2];
    long l4 = lmn4[0], m4 = lmn4[1], n4 = lmn4[2];
    double p = a + b;
    double q = c + d;
    double alpha = p * q / (p + q);
    double Px = (a * A[0] + b * B[0]) / p, Py = (a * A[1] + b * B[1]) / p, Pz = (a * A[2] + b * B[2]) / p;
    double Qx = (c * C[0] + d * D[0]) / q, Qy = (c * C[1] + d * D[1]) / q, Qz = (c * C[2] + d * D[2]) / q;
    double RPQ = sqrt(pow(Px - Qx, 2) + pow(Py - Qy, 2) + pow(Pz - Qz, 2));
    double leaf[MDO_MAXN];
    int L = (int)(l1 + m1 + n1 + l2 + m2 + n2 + l3 + m3 + n3 + l4 + m4 + n4);
    R_leaves(L, alpha, RPQ, leaf);

    /* E values depend only on the loop index of their own dimension: tabulate once (the reference
       re-evaluates the same six calls inside the innermost loop, twoe.pyx:83-88). */
    double E1[16], E2[16], E3[16], E4[16], E5[16], E6[16];
    for (int t = 0; t <= l1 + l2; ++t) E1[t] = mdo_E(l1, l2, t, A[0] - B[0], a, b);
    for (int t = 0; t <= m1 + m2; ++t) E2[t] = mdo_E(m1, m2, t, A[1] - B[1], a, b);
    for (int t = 0; t <= n1 + n2; ++t) E3[t] = mdo_E(n1, n2, t, A[2] - B[2], a, b);
    for (int t = 0; t <= l3 + l4; ++t) E4[t] = mdo_E(l3, l4, t, C[0] - D[0], c, d);
    for (int t = 0; t <= m3 + m4; ++t) E5[t] = mdo_E(m3, m4, t, C[1] - D[1], c, d);
    for (int t = 0; t <= n3 + n4; ++t) E6[t] = mdo_E(n3, n4, t, C[2] - D[2], c, d);

    double val = 0.0;
    for (int t = 0; t <= l1 + l2; ++t)
        for (int u = 0; u <= m1 + m2; ++u)
            for (int v = 0; v <= n1 + n2; ++v)
                for (int tau = 0; tau <= l3 + l4; ++tau)
                    for (int nu = 0; nu <= m3 + m4; ++nu)
                        for (int phi = 0; phi <= n3 + n4; ++phi)
                            val += E1[t] * E2[u] * E3[v] * E4[tau] * E5[nu] * E6[phi] *
                                   (((tau + nu + phi) & 1) ? -1.0 : 1.0) *
                                   R_rec(t + tau, u + nu, v + phi, 0, leaf, Px - Qx, Py - Qy, Pz - Qz);
    val *= 2 * pow(MDO_PI, 2.5) / (p * q * sqrt(p + q)); /* twoe.pyx:93 */
    return val;
}

/* ---- basis carrier: flat arrays mirroring cdef class Basis (cython/basis.pxi:6-14) -------- */
typedef struct {
    long nbf;
    const double *origin; /* [nbf*3] */
    const long *shell;    /* [nbf*3] (l,m,n) */
    const long *nprim;    /* [nbf]   num_exps */
    const long *off;      /* [nbf]   offset into exps/coefs/norm */
    const double *exps, *coefs, *norm;
} mdo_basis;

static double fact2l(long n) /* (−1)!! = 1: the old-SciPy semantics basis.pxi relies on */
{
    double r = 1.0;
    for (long k = n; k > 1; k -= 2) r *= (double)k;
    return r;
}

/* ---- cython/basis.pxi:87-120  Basis.normalize --------------------------------------------- */
void mdo_normalize(const long *lmn, long K, const double *exps, double *coefs /* in: raw, out: scaled */,
                   double *norm)
{
    long l = lmn[0], m = lmn[1], n = lmn[2], L = l + m + n;
    for (long ia = 0; ia < K; ++ia)
        norm[ia] = sqrt(pow(2, 2 * (l + m + n) + 1.5) * pow(exps[ia], l + m + n + 1.5) / fact2l(2 * l - 1) /
                        fact2l(2 * m - 1) / fact2l(2 * n - 1) / pow(M_PI, 1.5));
    double prefactor = pow(M_PI, 1.5) * fact2l(2 * l - 1) * fact2l(2 * m - 1) * fact2l(2 * n - 1) / pow(2.0, L);
    double N = 0.0;
    for (long ia = 0; ia < K; ++ia)
        for (long ib = 0; ib < K; ++ib)
            N += norm[ia] * norm[ib] * coefs[ia] * coefs[ib] / pow(exps[ia] + exps[ib], L + 1.5);
    N *= prefactor;
    N = pow(N, -0.5);
    for (long ia = 0; ia < K; ++ia) coefs[ia] *= N;
}

/* ---- cython/twoe.pyx:36-50  ERI(a,b,c,d) --------------------------------------------------- */
static double eri_bf(const mdo_basis *bs, long a, long b, long c, long d)
{
    double eri = 0.0;
    const double *ea = bs->exps + bs->off[a], *eb = bs->exps + bs->off[b], *ec = bs->exps + bs->off[c],
                 *ed = bs->exps + bs->off[d];
    const double *ca = bs->coefs + bs->off[a], *cb = bs->coefs + bs->off[b], *cc = bs->coefs + bs->off[c],
                 *cd = bs->coefs + bs->off[d];
    const double *na = bs->norm + bs->off[a], *nb = bs->norm + bs->off[b], *nc = bs->norm + bs->off[c],
                 *nd = bs->norm + bs->off[d];
    for (long ja = 0; ja < bs->nprim[a]; ++ja)
        for (long jb = 0; jb < bs->nprim[b]; ++jb)
            for (long jc = 0; jc < bs->nprim[c]; ++jc)
                for (long jd = 0; jd < bs->nprim[d]; ++jd)
                    eri += na[ja] * nb[jb] * nc[jc] * nd[jd] * ca[ja] * cb[jb] * cc[jc] * cd[jd] *
                           mdo_electron_repulsion(ea[ja], bs->shell + 3 * a, bs->origin + 3 * a, eb[jb],
                                                  bs->shell + 3 * b, bs->origin + 3 * b, ec[jc], bs->shell + 3 * c,
                                                  bs->origin + 3 * c, ed[jd], bs->shell + 3 * d, bs->origin + 3 * d);
    return eri;
}

static mdo_basis mk(long nbf, const double *origin, const long *shell, const long *nprim, const long *off,
                    const double *exps, const double *coefs, const double *norm)
{
    mdo_basis b = {nbf, origin, shell, nprim, off, exps, coefs, norm};
    return b;
}

double mdo_ERI(long nbf, const double *origin, const long *shell, const long *nprim, const long *off,
               const double *exps, const double *coefs, const double *norm, long a, long b, long c, long d)
{
    mdo_basis bs = mk(nbf, origin, shell, nprim, off, exps, coefs, norm);
    return eri_bf(&bs, a, b, c, d);
}

/* batched: out[n] = ERI(idx[4n..4n+3]) */
typedef struct { const mdo_basis *bs; const long *idx; double *out; } batch_ctx;
static void batch_body(long q, void *v)
{
    batch_ctx *c = (batch_ctx *)v;
    c->out[q] = eri_bf(c->bs, c->idx[4 * q], c->idx[4 * q + 1], c->idx[4 * q + 2], c->idx[4 * q + 3]);
}
void mdo_ERI_batch(long nbf, const double *origin, const long *shell, const long *nprim, const long *off,
                   const double *exps, const double *coefs, const double *norm, long n, const long *idx,
                   double *out)
{
    mdo_basis bs = mk(nbf, origin, shell, nprim, off, exps, coefs, norm);
    batch_ctx c = {&bs, idx, out};
    parallel_for(n, batch_body, &c);
}

/* ---- cython/twoe.pyx:12-31  doERIs(N, TwoE, bfs) ------------------------------------------- */
typedef struct { const mdo_basis *bs; long N; double *out; } fill_ctx;
static void doeris_row(long i, void *v);
void mdo_doERIs(long N, double *TwoE, const double *origin, const long *shell, const long *nprim,
                const long *off, const double *exps, const double *coefs, const double *norm)
{
    mdo_basis bs = mk(N, origin, shell, nprim, off, exps, coefs, norm);
    fill_ctx c = {&bs, N, TwoE};
    parallel_for(N, doeris_row, &c); /* rows i are independent: every element is written with one value */
}
static void doeris_row(long i, void *v)
{
    fill_ctx *c = (fill_ctx *)v;
    const mdo_basis bs = *c->bs;
    long N = c->N;
    double *TwoE = c->out;
#define T4(i, j, k, l) TwoE[(((i)*N + (j)) * N + (k)) * N + (l)]
    {
        for (long j = 0; j <= i; ++j) {
            long ij = i * (i + 1) / 2 + j;
            for (long k = 0; k < N; ++k)
                for (long l = 0; l <= k; ++l) {
                    long kl = k * (k + 1) / 2 + l;
                    if (ij >= kl) {
                        double val = eri_bf(&bs, i, j, k, l);
                        T4(i, j, k, l) = val; T4(k, l, i, j) = val; T4(j, i, l, k) = val; T4(l, k, j, i) = val;
                        T4(j, i, k, l) = val; T4(l, k, i, j) = val; T4(i, j, l, k) = val; T4(k, l, j, i) = val;
                    }
                }
        }
    }
#undef T4
}

static void schwarz_row(long p, void *v)
{
    fill_ctx *c = (fill_ctx *)v;
    for (long q = 0; q <= p; ++q) c->out[p * (p + 1) / 2 + q] = eri_bf(c->bs, p, q, p, q);
}
/* ---- mmd/molecule.py:95-99  Schwarz table screen[p(p+1)/2+q] = (pq|pq) --------------------- */
void mdo_schwarz(long N, double *screen, const double *origin, const long *shell, const long *nprim,
                 const long *off, const double *exps, const double *coefs, const double *norm)
{
    mdo_basis bs = mk(N, origin, shell, nprim, off, exps, coefs, norm);
    fill_ctx c = {&bs, N, screen};
    parallel_for(N, schwarz_row, &c);
}

/* ---- cython/fock.pyx:13-87  formPT(P, P_old, bfs, nbasis, screen, tol) --------------------- */
/* P, P_old, G are complex128 stored interleaved (re,im), row-major (N,N).  Returns the
   UN-symmetrised G exactly like the reference; the number of quartets that passed the screen is
   written to *ncomputed (may be NULL).  The reference's single-threaded accumulation order is
   kept (no OpenMP here: G is a shared accumulator and order matters for bit-reproducibility). */
void mdo_formPT(long N, const double *P, const double *P_old, const double *screen, double tol, double *G,
                long *ncomputed, const double *origin, const long *shell, const long *nprim, const long *off,
                const double *exps, const double *coefs, const double *norm)
{
    mdo_basis bs = mk(N, origin, shell, nprim, off, exps, coefs, norm);
    double *dP = (double *)malloc(sizeof(double) * 2 * N * N);
    for (long x = 0; x < 2 * N * N; ++x) { dP[x] = P[x] - P_old[x]; G[x] = 0.0; } /* fock.pyx:22,24 */
    long ncomp = 0;
#define DPRE(i, j) dP[2 * ((i)*N + (j))]
#define DPIM(i, j) dP[2 * ((i)*N + (j)) + 1]
#define CABS(i, j, s) hypot((s)*DPRE(i, j), (s)*DPIM(i, j))
#define GADD(i, j, k, l, f)                                                                             \
    do { G[2 * ((i)*N + (j))] += (f)*DPRE(k, l) * eri; G[2 * ((i)*N + (j)) + 1] += (f)*DPIM(k, l) * eri; } while (0)
    for (long i = 0; i < N; ++i)
        for (long j = 0; j <= i; ++j) {
            long ij = i * (i + 1) / 2 + j;
            for (long k = 0; k < N; ++k)
                for (long l = 0; l <= k; ++l) {
                    long kl = k * (k + 1) / 2 + l;
                    if (ij < kl) continue;
                    double bound = sqrt(screen[ij]) * sqrt(screen[kl]); /* fock.pyx:46-47 */
                    double dmax = CABS(i, j, 4.0);                     /* fock.pyx:49-54 */
                    double x;
                    x = CABS(k, l, 4.0); if (x > dmax) dmax = x;
                    x = CABS(i, k, 1.0); if (x > dmax) dmax = x;
                    x = CABS(i, l, 1.0); if (x > dmax) dmax = x;
                    x = CABS(j, k, 1.0); if (x > dmax) dmax = x;
                    x = CABS(j, l, 1.0); if (x > dmax) dmax = x;
                    bound *= dmax;
                    if (bound < tol) continue; /* NaN bound is NOT skipped, like the reference */
                    double s12 = (i == j) ? 1.0 : 2.0, s34 = (k == l) ? 1.0 : 2.0; /* fock.pyx:60-70 */
                    double s1234 = (i == k) ? ((j == l) ? 1.0 : 2.0) : 2.0;
                    double eri = s12 * s34 * s1234 * eri_bf(&bs, i, j, k, l); /* fock.pyx:74-75 */
                    ++ncomp;
                    GADD(i, j, k, l, 1.0);   /* fock.pyx:79 */
                    GADD(k, l, i, j, 1.0);   /* fock.pyx:80 */
                    GADD(i, k, j, l, -0.25); /* fock.pyx:82 */
                    GADD(j, l, i, k, -0.25); /* fock.pyx:83 */
                    GADD(i, l, j, k, -0.25); /* fock.pyx:84 */
                    GADD(k, j, i, l, -0.25); /* fock.pyx:85 */
                }
        }
    if (ncomputed) *ncomputed = ncomp;
    free(dP);
}

/* ---- mmd/scf.py:97-98  J = einsum('pqrs,sr->pq'), K = einsum('psqr,sr->pq') ---------------- */
/* TwoE real (N,N,N,N); P, J, K complex interleaved (N,N). */
void mdo_jk_incore(long N, const double *TwoE, const double *P, double *J, double *K)
{
    for (long p = 0; p < N; ++p)
        for (long q = 0; q < N; ++q) {
            double jr = 0, ji = 0, kr = 0, ki = 0;
            for (long r = 0; r < N; ++r)
                for (long s = 0; s < N; ++s) {
                    double pr = P[2 * (s * N + r)], pi = P[2 * (s * N + r) + 1];
                    double ej = TwoE[((p * N + q) * N + r) * N + s];
                    double ek = TwoE[((p * N + s) * N + q) * N + r];
                    jr += ej * pr; ji += ej * pi;
                    kr += ek * pr; ki += ek * pi;
                }
            J[2 * (p * N + q)] = jr; J[2 * (p * N + q) + 1] = ji;
            K[2 * (p * N + q)] = kr; K[2 * (p * N + q) + 1] = ki;
        }
}

/* =============================================================================================
 * One-electron integrals (cython/onee.pyx) — checker for the enabler kernel csrc/onee.cu
 * ============================================================================================= */
static double E_neg(int i, int j, int t, double Q, double a, double b)
{
    /* E with possibly negative j as the reference's kinetic() may call it: every such term is
       multiplied by a zero coefficient there; return 0 like the reference's recursion bottom. */
    if (i < 0 || j < 0) return 0.0;
    return mdo_E(i, j, t, Q, a, b);
}

/* onee.pyx:71-78 */
static double p_overlap(double a, const long *l1, const double *A, double b, const long *l2, const double *B)
{
    return mdo_E(l1[0], l2[0], 0, A[0] - B[0], a, b) * mdo_E(l1[1], l2[1], 0, A[1] - B[1], a, b) *
           mdo_E(l1[2], l2[2], 0, A[2] - B[2], a, b) * pow(MDO_PI / (a + b), 1.5);
}

/* onee.pyx:108-137 */
static double p_kinetic(double a, const long *l1, const double *A, double b, const long *l2, const double *B)
{
    double T[3], S[3];
    for (int d = 0; d < 3; ++d) S[d] = mdo_E(l1[d], l2[d], 0, A[d] - B[d], a, b);
    for (int d = 0; d < 3; ++d) {
        double Ad = (2 * l2[d] + 1) * b, Bd = -2 * pow(b, 2), Cd = -0.5 * l2[d] * (l2[d] - 1);
        T[d] = Ad * mdo_E(l1[d], l2[d], 0, A[d] - B[d], a, b) + Bd * mdo_E(l1[d], l2[d] + 2, 0, A[d] - B[d], a, b) +
               Cd * E_neg(l1[d], l2[d] - 2, 0, A[d] - B[d], a, b);
    }
    double Tx = T[0] * S[1] * S[2], Ty = T[1] * S[0] * S[2], Tz = T[2] * S[0] * S[1];
    return (Tx + Ty + Tz) * pow(MDO_PI / (a + b), 1.5);
}

/* onee.pyx:81-106 */
static double p_dipole(double a, const long *l1, const double *A, double b, const long *l2, const double *B,
                       const double *C, int dir)
{
    double p = a + b, S[3], D;
    for (int d = 0; d < 3; ++d) S[d] = mdo_E(l1[d], l2[d], 0, A[d] - B[d], a, b);
    double Pd = (a * A[dir] + b * B[dir]) / p;
    D = mdo_E(l1[dir], l2[dir], 1, A[dir] - B[dir], a, b) + (Pd - C[dir]) * S[dir];
    S[dir] = D;
    return S[0] * S[1] * S[2] * pow(MDO_PI / p, 1.5);
}

/* onee.pyx:140-176 */
static double p_angular(double a, const long *l1, const double *A, double b, const long *l2, const double *B,
                        const double *C, int dir)
{
    double S0[3], S1[3], D1[3];
    for (int d = 0; d < 3; ++d) {
        double Q = A[d] - B[d];
        S0[d] = mdo_E(l1[d], l2[d], 0, Q, a, b);
        S1[d] = mdo_E(l1[d] + 1, l2[d], 0, Q, a, b) + (A[d] - C[d]) * mdo_E(l1[d], l2[d], 0, Q, a, b); /* util.pxi:27-28 */
        D1[d] = l2[d] * E_neg(l1[d], l2[d] - 1, 0, Q, a, b) - 2 * b * mdo_E(l1[d], l2[d] + 1, 0, Q, a, b);
    }
    double pf = pow(MDO_PI / (a + b), 1.5);
    if (dir == 0) return -S0[0] * (S1[1] * D1[2] - S1[2] * D1[1]) * pf;
    if (dir == 1) return -S0[1] * (S1[2] * D1[0] - S1[0] * D1[2]) * pf;
    return -S0[2] * (S1[0] * D1[1] - S1[1] * D1[0]) * pf;
}

/* onee.pyx:178-192 */
static double p_nuclear(double a, const long *l1, const double *A, double b, const long *l2, const double *B,
                        const double *C)
{
    double p = a + b;
    double P[3];
    for (int d = 0; d < 3; ++d) P[d] = (a * A[d] + b * B[d]) / p;
    double RPC = sqrt((P[0] - C[0]) * (P[0] - C[0]) + (P[1] - C[1]) * (P[1] - C[1]) + (P[2] - C[2]) * (P[2] - C[2]));
    double leaf[MDO_MAXN];
    int L = (int)(l1[0] + l1[1] + l1[2] + l2[0] + l2[1] + l2[2]);
    R_leaves(L, p, RPC, leaf);
    double val = 0.0;
    for (int t = 0; t <= l1[0] + l2[0]; ++t)
        for (int u = 0; u <= l1[1] + l2[1]; ++u)
            for (int v = 0; v <= l1[2] + l2[2]; ++v)
                val += mdo_E(l1[0], l2[0], t, A[0] - B[0], a, b) * mdo_E(l1[1], l2[1], u, A[1] - B[1], a, b) *
                       mdo_E(l1[2], l2[2], v, A[2] - B[2], a, b) *
                       R_rec(t, u, v, 0, leaf, P[0] - C[0], P[1] - C[1], P[2] - C[2]);
    return val * 2 * MDO_PI / p;
}

typedef struct {
    const mdo_basis *bs; long N; long natom; const double *Z, *xyz, *origin;
    double *S, *T, *V, *M, *L;
} onee_ctx;

/* mmd/molecule.py:253-276: rows i, columns j <= i, mirrored (L antisymmetric, diagonal ends -L_ii) */
static void onee_row(long i, void *v)
{
    onee_ctx *c = (onee_ctx *)v;
    const mdo_basis *bs = c->bs;
    long N = c->N;
    for (long j = 0; j <= i; ++j) {
        double s = 0, t = 0, mu[3] = {0, 0, 0}, ll[3] = {0, 0, 0};
        const long *la = bs->shell + 3 * i, *lb = bs->shell + 3 * j;
        const double *A = bs->origin + 3 * i, *B = bs->origin + 3 * j;
        for (long ia = 0; ia < bs->nprim[i]; ++ia)
            for (long ib = 0; ib < bs->nprim[j]; ++ib) {
                double ea = bs->exps[bs->off[i] + ia], eb = bs->exps[bs->off[j] + ib];
                double w = bs->norm[bs->off[i] + ia] * bs->norm[bs->off[j] + ib] * bs->coefs[bs->off[i] + ia] *
                           bs->coefs[bs->off[j] + ib];
                s += w * p_overlap(ea, la, A, eb, lb, B);
                t += w * p_kinetic(ea, la, A, eb, lb, B);
                for (int d = 0; d < 3; ++d) {
                    mu[d] += w * p_dipole(ea, la, A, eb, lb, B, c->origin, d);
                    ll[d] += w * p_angular(ea, la, A, eb, lb, B, c->origin, d);
                }
            }
        double vv = 0.0;
        for (long at = 0; at < c->natom; ++at) {
            double va = 0.0;
            for (long ia = 0; ia < bs->nprim[i]; ++ia)
                for (long ib = 0; ib < bs->nprim[j]; ++ib) {
                    double ea = bs->exps[bs->off[i] + ia], eb = bs->exps[bs->off[j] + ib];
                    double w = bs->norm[bs->off[i] + ia] * bs->norm[bs->off[j] + ib] * bs->coefs[bs->off[i] + ia] *
                               bs->coefs[bs->off[j] + ib];
                    va += w * p_nuclear(ea, la, A, eb, lb, B, c->xyz + 3 * at);
                }
            vv += -c->Z[at] * va;
        }
        c->S[i * N + j] = c->S[j * N + i] = s;
        c->T[i * N + j] = c->T[j * N + i] = t;
        c->V[i * N + j] = c->V[j * N + i] = vv;
        for (int d = 0; d < 3; ++d) {
            c->M[d * N * N + i * N + j] = c->M[d * N * N + j * N + i] = mu[d];
            c->L[d * N * N + i * N + j] = ll[d];
            c->L[d * N * N + j * N + i] = -ll[d];
        }
    }
}

void mdo_onee(long N, long natom, const double *Z, const double *xyz, const double *origin3, double *S, double *T,
              double *V, double *M, double *L, const double *origin, const long *shell, const long *nprim,
              const long *off, const double *exps, const double *coefs, const double *norm)
{
    mdo_basis bs = mk(N, origin, shell, nprim, off, exps, coefs, norm);
    onee_ctx c = {&bs, N, natom, Z, xyz, origin3, S, T, V, M, L};
    parallel_for(N, onee_row, &c);
}

/* ==============================================================================================
 * Nuclear-derivative integrals (cython/grad.pyx).  Restated for the NEXT row of the
 * scope table: the reference differentiates the Hermite coefficient of the differentiated centre,
 *     Ex(i,j,t) = 2a E(i+1,j,t) - i E(i-1,j,t)                                  (grad.pyx:104-113)
 * and sums t one order higher; by linearity of the Hermite expansion that is the same as
 *     d/dA_x [a b|..] = 2a [a+1_x b|..] - l_x [a-1_x b|..]
 * applied to every primitive with the UNSHIFTED function's norm and coefficient (grad.pyx:74-101),
 * which is what is coded here (and checked against the reference's own values in tests/).
 * ============================================================================================== */
static void shifted(const long *lmn, int x, int delta, long *out)
{
    out[0] = lmn[0]; out[1] = lmn[1]; out[2] = lmn[2];
    out[x] += delta;
}

/* cython/grad.pyx:74-101 ERIx(a,b,c,d, x, center): centre = 0..3 for 'a'..'d' */
static double erix_bf(const mdo_basis *bs, const long f[4], int x, int center)
{
    double val = 0.0;
    const double *e[4], *c[4], *n[4];
    for (int k = 0; k < 4; ++k) {
        e[k] = bs->exps + bs->off[f[k]];
        c[k] = bs->coefs + bs->off[f[k]];
        n[k] = bs->norm + bs->off[f[k]];
    }
    long up[3], dn[3];
    const long *base = bs->shell + 3 * f[center];
    shifted(base, x, +1, up);
    shifted(base, x, -1, dn);
    const long lx = base[x];
    for (long ja = 0; ja < bs->nprim[f[0]]; ++ja)
        for (long jb = 0; jb < bs->nprim[f[1]]; ++jb)
            for (long jc = 0; jc < bs->nprim[f[2]]; ++jc)
                for (long jd = 0; jd < bs->nprim[f[3]]; ++jd) {
                    const long j[4] = {ja, jb, jc, jd};
                    const long *l[4] = {bs->shell + 3 * f[0], bs->shell + 3 * f[1], bs->shell + 3 * f[2], bs->shell + 3 * f[3]};
                    const double w = n[0][ja] * n[1][jb] * n[2][jc] * n[3][jd] * c[0][ja] * c[1][jb] * c[2][jc] * c[3][jd];
                    const double alpha = e[center][j[center]];
                    l[center] = up;
                    double t = 2.0 * alpha *
                               mdo_electron_repulsion(e[0][ja], l[0], bs->origin + 3 * f[0], e[1][jb], l[1], bs->origin + 3 * f[1],
                                                      e[2][jc], l[2], bs->origin + 3 * f[2], e[3][jd], l[3], bs->origin + 3 * f[3]);
                    if (lx > 0) {
                        l[center] = dn;
                        t -= (double)lx *
                             mdo_electron_repulsion(e[0][ja], l[0], bs->origin + 3 * f[0], e[1][jb], l[1], bs->origin + 3 * f[1],
                                                    e[2][jc], l[2], bs->origin + 3 * f[2], e[3][jd], l[3], bs->origin + 3 * f[3]);
                    }
                    val += w * t;
                }
    return val;
}

/* batched: out[n] = ERIx(idx[4n..4n+3], x = xc[2n], center = xc[2n+1]) */
typedef struct { const mdo_basis *bs; const long *idx; const long *xc; double *out; } erix_ctx;
static void erix_body(long q, void *v)
{
    erix_ctx *c = (erix_ctx *)v;
    c->out[q] = erix_bf(c->bs, c->idx + 4 * q, (int)c->xc[2 * q], (int)c->xc[2 * q + 1]);
}
void mdo_ERIx_batch(long nbf, const double *origin, const long *shell, const long *nprim, const long *off,
                    const double *exps, const double *coefs, const double *norm, long n, const long *idx,
                    const long *xc, double *out)
{
    mdo_basis bs = mk(nbf, origin, shell, nprim, off, exps, coefs, norm);
    erix_ctx c = {&bs, idx, xc, out};
    parallel_for(n, erix_body, &c);
}

/* cython/grad.pyx:11-24 Sx and :27-40 Tx (overlapX :352-397, kineticX :400-506): derivative with respect to
   the centre of a (center = 0) or b (center = 1); kind 0 = overlap, 1 = kinetic energy */
double mdo_onee_x(long nbf, const double *origin, const long *shell, const long *nprim, const long *off,
                  const double *exps, const double *coefs, const double *norm, long a, long b, long x, long center,
                  long kind)
{
    mdo_basis bs = mk(nbf, origin, shell, nprim, off, exps, coefs, norm);
    const long f[2] = {a, b};
    long up[3], dn[3];
    const long *base = bs.shell + 3 * f[center];
    shifted(base, (int)x, +1, up);
    shifted(base, (int)x, -1, dn);
    const long lx = base[x];
    double val = 0.0;
    for (long ja = 0; ja < bs.nprim[a]; ++ja)
        for (long jb = 0; jb < bs.nprim[b]; ++jb) {
            const double ea = bs.exps[bs.off[a] + ja], eb = bs.exps[bs.off[b] + jb];
            const double w = bs.norm[bs.off[a] + ja] * bs.norm[bs.off[b] + jb] * bs.coefs[bs.off[a] + ja] * bs.coefs[bs.off[b] + jb];
            const long *l[2] = {bs.shell + 3 * a, bs.shell + 3 * b};
            const double alpha = center == 0 ? ea : eb;
            l[center] = up;
            double t = 2.0 * alpha * (kind == 0 ? p_overlap(ea, l[0], bs.origin + 3 * a, eb, l[1], bs.origin + 3 * b)
                                                : p_kinetic(ea, l[0], bs.origin + 3 * a, eb, l[1], bs.origin + 3 * b));
            if (lx > 0) {
                l[center] = dn;
                t -= (double)lx * (kind == 0 ? p_overlap(ea, l[0], bs.origin + 3 * a, eb, l[1], bs.origin + 3 * b)
                                             : p_kinetic(ea, l[0], bs.origin + 3 * a, eb, l[1], bs.origin + 3 * b));
            }
            val += w * t;
        }
    return val;
}

/* cython/grad.pyx:509-550 nuclear_attractionXa: derivative of the OPERATOR 1/|r - C| with respect to C_x
   (Hellmann-Feynman term): -sum E E E R_{t+1_x,u,v} * 2 pi / p */
static double p_nuclear_dC(double a, const long *l1, const double *A, double b, const long *l2, const double *B,
                           const double *C, int x)
{
    double p = a + b;
    double P[3];
    for (int d = 0; d < 3; ++d) P[d] = (a * A[d] + b * B[d]) / p;
    double RPC = sqrt((P[0] - C[0]) * (P[0] - C[0]) + (P[1] - C[1]) * (P[1] - C[1]) + (P[2] - C[2]) * (P[2] - C[2]));
    double leaf[MDO_MAXN];
    int L = (int)(l1[0] + l1[1] + l1[2] + l2[0] + l2[1] + l2[2]) + 1;
    R_leaves(L, p, RPC, leaf);
    double val = 0.0;
    for (int t = 0; t <= l1[0] + l2[0]; ++t)
        for (int u = 0; u <= l1[1] + l2[1]; ++u)
            for (int v = 0; v <= l1[2] + l2[2]; ++v)
                val -= mdo_E(l1[0], l2[0], t, A[0] - B[0], a, b) * mdo_E(l1[1], l2[1], u, A[1] - B[1], a, b) *
                       mdo_E(l1[2], l2[2], v, A[2] - B[2], a, b) *
                       R_rec(t + (x == 0), u + (x == 1), v + (x == 2), 0, leaf, P[0] - C[0], P[1] - C[1], P[2] - C[2]);
    return val * 2 * MDO_PI / p;
}

/* cython/grad.pyx:43-71 VxA (mode 2: operator derivative) and VxB (mode 0 / 1: derivative with respect to the
   centre of a / b, grad.pyx:553-629) for the nucleus at C */
double mdo_V_x(long nbf, const double *origin, const long *shell, const long *nprim, const long *off,
               const double *exps, const double *coefs, const double *norm, long a, long b, const double *C, long x,
               long mode)
{
    mdo_basis bs = mk(nbf, origin, shell, nprim, off, exps, coefs, norm);
    const long f[2] = {a, b};
    double val = 0.0;
    long up[3], dn[3];
    long lx = 0;
    if (mode < 2) {
        const long *base = bs.shell + 3 * f[mode];
        shifted(base, (int)x, +1, up);
        shifted(base, (int)x, -1, dn);
        lx = base[x];
    }
    for (long ja = 0; ja < bs.nprim[a]; ++ja)
        for (long jb = 0; jb < bs.nprim[b]; ++jb) {
            const double ea = bs.exps[bs.off[a] + ja], eb = bs.exps[bs.off[b] + jb];
            const double w = bs.norm[bs.off[a] + ja] * bs.norm[bs.off[b] + jb] * bs.coefs[bs.off[a] + ja] * bs.coefs[bs.off[b] + jb];
            const long *l[2] = {bs.shell + 3 * a, bs.shell + 3 * b};
            double t;
            if (mode == 2) {
                t = p_nuclear_dC(ea, l[0], bs.origin + 3 * a, eb, l[1], bs.origin + 3 * b, C, (int)x);
            } else {
                const double alpha = mode == 0 ? ea : eb;
                l[mode] = up;
                t = 2.0 * alpha * p_nuclear(ea, l[0], bs.origin + 3 * a, eb, l[1], bs.origin + 3 * b, C);
                if (lx > 0) {
                    l[mode] = dn;
                    t -= (double)lx * p_nuclear(ea, l[0], bs.origin + 3 * a, eb, l[1], bs.origin + 3 * b, C);
                }
            }
            val += w * t;
        }
    return val;
}

// grad.cu — nuclear gradient of the RHF energy on the device.
//
// Replaces the reference's cython/grad.pyx (Sx, Tx, VxA, VxB, ERIx; 629 lines of un-memoised recursion) and the
// N^4-tensor-per-atom-and-direction contraction of mmd/forces.py:61-92 by two kernels:
//
//   grad1e_kernel   one thread per ordered basis-function pair (i,j): derivative overlap, kinetic and nuclear-attraction
//                   integrals with respect to the centres of i and j, and the operator (Hellmann-Feynman) derivative of
//                   every nuclear attraction term, contracted with P and W = P F P on the fly (mmd/forces.py:20-59, 94-99);
//   grad2e_kernel   one thread per (bra shell pair, ket shell pair, ket component pair): the derivative of (ab|cd) with
//                   respect to the two BRA centres, obtained from Hermite coefficients one order higher on the
//                   differentiated function (cython/grad.pyx:104-113: 2a E^{i+1,j} - i E^{i-1,j}), contracted with the
//                   two-particle density as it is produced:
//                       dE2/dX = 1/2 sum_{ijkl} Gamma_ijkl d(ij|kl)/dX_i ,  Gamma = 16 P_ij P_kl - 4 P_ik P_jl - 4 P_il P_jk
//                   (all ordered quartets; equals einsum(P, 2Jx - Kx) of mmd/forces.py:86-92 for the real symmetric RHF
//                   density P = C_occ C_occ^T).  No derivative tensor is ever stored.
// d functions need f-type shifted primitives: the Hermite tables go one order past the (dd|dd) energy kernels (L = 9).
// Runtime-L code with thread-local tables: forces are a consumer of the path, not its hot loop.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "handle.h"
#include "kernels_b.cuh"

using namespace mmdb;

void mmdb_make_boys_table(int L, std::vector<double> &tab);     // lib.cu

namespace {

constexpr int GR_AM = 2;                     // highest angular momentum of a basis function
constexpr int GR_LB = 2 * GR_AM + 1;         // Hermite order of a differentiated pair (5)
constexpr int GR_L = 4 * GR_AM + 1;          // order of the Hermite Coulomb table (9)

// normalised Hermite coefficients E_t^{ij}/E_0^{00} for i <= imax, j <= jmax (cython/util.pxi:13-26)
typedef double EGR[3][GR_AM + 2][GR_AM + 2][GR_LB + 1];

__device__ void build_E_gr(EGR &E, int imax, int jmax, const double *PA, const double *PB, double oo2p)
{
    for (int dim = 0; dim < 3; ++dim) {
        for (int i = 0; i <= imax; ++i)
            for (int j = 0; j <= jmax; ++j)
                for (int t = 0; t <= GR_LB; ++t) E[dim][i][j][t] = 0.0;
        E[dim][0][0][0] = 1.0;
        for (int i = 1; i <= imax; ++i)
            for (int t = 0; t <= i; ++t) {
                double x = PA[dim] * E[dim][i - 1][0][t];
                if (t > 0) x += oo2p * E[dim][i - 1][0][t - 1];
                if (t + 1 <= i - 1) x += (double)(t + 1) * E[dim][i - 1][0][t + 1];
                E[dim][i][0][t] = x;
            }
        for (int j = 1; j <= jmax; ++j)
            for (int i = 0; i <= imax; ++i) {
                if (i + j > imax + jmax - 1) continue;       // the (imax, jmax) corner is never used: only ONE function is raised
                for (int t = 0; t <= i + j; ++t) {
                    double x = PB[dim] * E[dim][i][j - 1][t];
                    if (t > 0) x += oo2p * E[dim][i][j - 1][t - 1];
                    if (t + 1 <= i + j - 1) x += (double)(t + 1) * E[dim][i][j - 1][t + 1];
                    E[dim][i][j][t] = x;
                }
            }
    }
}

// derivative Hermite coefficient of dimension `dim` with respect to the first (which = 0) or second centre:
//   2 alpha E^{i+1,j}_t - i E^{i-1,j}_t     (cython/grad.pyx:104-113)
__device__ __forceinline__ double dE(const EGR &E, int dim, int i, int j, int t, int which, double alpha)
{
    if (which == 0) {
        double x = 2.0 * alpha * E[dim][i + 1][j][t];
        if (i > 0) x -= (double)i * E[dim][i - 1][j][t];
        return x;
    }
    double x = 2.0 * alpha * E[dim][i][j + 1][t];
    if (j > 0) x -= (double)j * E[dim][i][j - 1][t];
    return x;
}

struct Grad2eArgs {
    const PairHdr *braH, *ketH;
    const PrimPair *braP, *ketP;
    const double2 *braAB;                 // individual exponents (a, b) of every bra primitive pair
    const double *Qs_bra, *Qs_ket;
    int nbra, nket, la, lb, lc, ld;
    int N;
    const double *P;                      // (N,N) real symmetric density, P = C_occ C_occ^T
    const int *shell_atom;
    const double *boys_tab;               // order la+lb+lc+ld+1
    double cut;                           // shell quartets with Qs_bra Qs_ket max|P|^2 * 16 below this are skipped
    double pmax2;
    double *grad;                         // [natom][3]
};

__global__ void __launch_bounds__(64) grad2e_kernel(const Grad2eArgs g)
{
    extern __shared__ double s_boys[];
    for (int x = threadIdx.x; x < BOYS_ROWS * BOYS_STRIDE; x += blockDim.x) s_boys[x] = g.boys_tab[x];
    __syncthreads();
    const int la = g.la, lb = g.lb, lc = g.lc, ld = g.ld;
    const int NA = ncart(la), NB = ncart(lb), ND = ncart(ld), NCD = ncart(lc) * ND;
    const int LBRA = la + lb + 1, L = la + lb + lc + ld + 1;
    const long long total = (long long)g.nbra * g.nket * NCD;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        const long long e = w / NCD;
        const int cd = (int)(w % NCD), c = cd / ND, d = cd % ND;
        const int i = (int)(e % g.nbra), j = (int)(e / g.nbra);
        if (g.Qs_bra[i] * g.Qs_ket[j] * g.pmax2 < g.cut) continue;
        const PairHdr bh = g.braH[i], kh = g.ketH[j];
        const int cx = cart_pow_rt(lc, c, 0), cy = cart_pow_rt(lc, c, 1), cz = cart_pow_rt(lc, c, 2);
        const int dx = cart_pow_rt(ld, d, 0), dy = cart_pow_rt(ld, d, 1), dz = cart_pow_rt(ld, d, 2);
        double gA[3] = {0.0, 0.0, 0.0}, gB[3] = {0.0, 0.0, 0.0};
        const bool sameAB = bh.shA == bh.shB;
        const int N = g.N;
        const int k = kh.bfA + c, l = kh.bfB + d;
        const double Pkl = g.P[(size_t)k * N + l];
        const double scd = comp_scale_rt(lc, c) * comp_scale_rt(ld, d) * ((kh.shA == kh.shB) ? 0.5 : 1.0);    // 1/2 f_ket
        for (int ib = 0; ib < bh.pnum; ++ib) {
            const PrimPair b = g.braP[bh.poff + ib];
            const double2 ab = g.braAB[bh.poff + ib];
            EGR Eb;
            {
                const double PA[3] = {b.PAx, b.PAy, b.PAz};
                const double PB[3] = {b.PAx + bh.ABx, b.PAy + bh.ABy, b.PAz + bh.ABz};
                build_E_gr(Eb, la + 1, lb + 1, PA, PB, 0.5 / b.p);
            }
            double G[nherm(GR_LB)];
            const int nhb = nherm(LBRA);
            for (int x = 0; x < nhb; ++x) G[x] = 0.0;
            for (int ik = 0; ik < kh.pnum; ++ik) {
                const PrimPair kq = g.ketP[kh.poff + ik];
                const double rs = fast_rsqrt(b.p + kq.p);
                const double alpha = b.p * kq.p * (rs * rs);
                const double X = b.Px - kq.Px, Y = b.Py - kq.Py, Z = b.Pz - kq.Pz;
                double Fs[GR_L + 1];
                boys_eval_rt(L, alpha * (X * X + Y * Y + Z * Z), s_boys, Fs);
                {
                    double s = b.cc * kq.cc * sqrt(b.p * kq.p) * rs * (1.0 / SQRTPI_2);
                    const double m2a = -2.0 * alpha;
                    for (int n = 0; n <= L; ++n) { Fs[n] *= s; s *= m2a; }
                }
                double R[nherm(GR_L)];
                build_R_rt(R, L, Fs, X, Y, Z);
                ETabRT Ek;
                {
                    const double QC[3] = {kq.PAx, kq.PAy, kq.PAz};
                    const double QD[3] = {kq.PAx + kh.ABx, kq.PAy + kh.ABy, kq.PAz + kh.ABz};
                    build_E_rt(Ek, lc, ld, QC, QD, 0.5 / kq.p);
                }
                for (int tau = 0; tau <= cx + dx; ++tau)
                    for (int nu = 0; nu <= cy + dy; ++nu)
                        for (int phi = 0; phi <= cz + dz; ++phi) {
                            double coef = Ek[0][cx][dx][tau] * Ek[1][cy][dy][nu] * Ek[2][cz][dz][phi];
                            if ((tau + nu + phi) & 1) coef = -coef;
                            for (int t = 0; t <= LBRA; ++t)
                                for (int u = 0; u <= LBRA - t; ++u)
                                    for (int v = 0; v <= LBRA - t - u; ++v)
                                        G[hidx(t, u, v)] = fma(coef, R[hidx(t + tau, u + nu, v + phi)], G[hidx(t, u, v)]);
                        }
            }
            for (int a = 0; a < NA; ++a)
                for (int bb = 0; bb < NB; ++bb) {
                    const int ii = bh.bfA + a, jj = bh.bfB + bb;
                    const double gam = 16.0 * g.P[(size_t)ii * N + jj] * Pkl - 4.0 * g.P[(size_t)ii * N + k] * g.P[(size_t)jj * N + l] -
                                       4.0 * g.P[(size_t)ii * N + l] * g.P[(size_t)jj * N + k];
                    const double wgt = gam * scd * comp_scale_rt(la, a) * comp_scale_rt(lb, bb);
                    if (wgt == 0.0) continue;
                    const int pa[3] = {cart_pow_rt(la, a, 0), cart_pow_rt(la, a, 1), cart_pow_rt(la, a, 2)};
                    const int pb[3] = {cart_pow_rt(lb, bb, 0), cart_pow_rt(lb, bb, 1), cart_pow_rt(lb, bb, 2)};
                    for (int which = 0; which < (sameAB ? 1 : 2); ++which) {
                        const double alpha = which == 0 ? ab.x : ab.y;
                        for (int x = 0; x < 3; ++x) {
                            // dimension x carries the derivative coefficient (one Hermite order more), the other two the plain ones
                            const int tmax[3] = {pa[0] + pb[0] + (x == 0), pa[1] + pb[1] + (x == 1), pa[2] + pb[2] + (x == 2)};
                            double acc = 0.0;
                            for (int t = 0; t <= tmax[0]; ++t) {
                                const double e0 = (x == 0) ? dE(Eb, 0, pa[0], pb[0], t, which, alpha) : Eb[0][pa[0]][pb[0]][t];
                                for (int u = 0; u <= tmax[1]; ++u) {
                                    const double e1 = (x == 1) ? dE(Eb, 1, pa[1], pb[1], u, which, alpha) : Eb[1][pa[1]][pb[1]][u];
                                    for (int v = 0; v <= tmax[2]; ++v) {
                                        const double e2 = (x == 2) ? dE(Eb, 2, pa[2], pb[2], v, which, alpha) : Eb[2][pa[2]][pb[2]][v];
                                        acc = fma(e0 * e1 * e2, G[hidx(t, u, v)], acc);
                                    }
                                }
                            }
                            if (which == 0) gA[x] = fma(wgt, acc, gA[x]);
                            else gB[x] = fma(wgt, acc, gB[x]);
                        }
                    }
                }
        }
        const int atA = g.shell_atom[bh.shA], atB = g.shell_atom[bh.shB];
        for (int x = 0; x < 3; ++x) {
            if (gA[x] != 0.0) atomicAdd(&g.grad[3 * atA + x], gA[x]);
            if (!sameAB && gB[x] != 0.0) atomicAdd(&g.grad[3 * atB + x], gB[x]);
        }
    }
}

// ---- one-electron part ---------------------------------------------------------------------------------------------
constexpr int G1_MAXI = GR_AM + 2;           // i up to la+1
constexpr int G1_MAXJ = GR_AM + 4;           // j up to lb+3 (kinetic energy of a once-raised function)
constexpr int G1_MAXT = 2 * GR_AM + 5;
struct E1G { double v[G1_MAXI][G1_MAXJ][G1_MAXT]; };

// E_t^{ij} including exp(-mu Q^2), as csrc/onee.cu build_E1
__device__ void build_E1g(E1G &E, int imax, int jmax, double Q, double a, double b)
{
    const double p = a + b, u = a * b / p, oo2p = 1.0 / (2 * p);
    const double PA = -(u * Q / a), PB = (u * Q / b);
    for (int i = 0; i <= imax; ++i)
        for (int j = 0; j <= jmax; ++j)
            for (int t = 0; t < G1_MAXT; ++t) E.v[i][j][t] = 0.0;
    E.v[0][0][0] = exp(-u * Q * Q);
    for (int i = 1; i <= imax; ++i)
        for (int t = 0; t <= i; ++t) {
            double x = PA * E.v[i - 1][0][t];
            if (t > 0) x += oo2p * E.v[i - 1][0][t - 1];
            if (t + 1 < G1_MAXT) x += (t + 1) * E.v[i - 1][0][t + 1];
            E.v[i][0][t] = x;
        }
    for (int j = 1; j <= jmax; ++j)
        for (int i = 0; i <= imax; ++i)
            for (int t = 0; t <= i + j; ++t) {
                double x = PB * E.v[i][j - 1][t];
                if (t > 0) x += oo2p * E.v[i][j - 1][t - 1];
                if (t + 1 < G1_MAXT) x += (t + 1) * E.v[i][j - 1][t + 1];
                E.v[i][j][t] = x;
            }
}
// 1-D kinetic factor of (i | -1/2 d^2/dx^2 | j), cython/onee.pyx:108-137
__device__ __forceinline__ double kin1(const E1G &E, int i, int j, double b)
{
    double x = (2 * j + 1) * b * E.v[i][j][0] - 2.0 * b * b * E.v[i][j + 2][0];
    if (j >= 2) x += -0.5 * j * (j - 1) * E.v[i][j - 2][0];
    return x;
}
// derivative with respect to the first / second centre of a 1-D factor f(i,j):  2a f(i+1,j) - i f(i-1,j)
template <class F>
__device__ __forceinline__ double d1(F f, int i, int j, int which, double a, double b)
{
    if (which == 0) return 2.0 * a * f(i + 1, j) - (i > 0 ? i * f(i - 1, j) : 0.0);
    return 2.0 * b * f(i, j + 1) - (j > 0 ? j * f(i, j - 1) : 0.0);
}

struct FnInfoG { int shell, comp; };
struct Grad1eArgs {
    int N, natom;
    const FnInfoG *fn;
    const int *am, *nprim, *poff, *shell_atom;
    const double *centre, *exps, *coefs;
    const double *Z, *xyz;
    const double *const *boys;           // boys[L], L <= 2 GR_AM + 1
    const double *P, *W;                 // (N,N) real symmetric: density and energy-weighted density P F P
    double *grad;                        // [natom][3]
};

__global__ void __launch_bounds__(64) grad1e_kernel(const Grad1eArgs g)
{
    const long long npair = (long long)g.N * g.N;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < npair; w += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(w / g.N), j = (int)(w % g.N);
        const FnInfoG fa = g.fn[i], fb = g.fn[j];
        const int la = g.am[fa.shell], lb = g.am[fb.shell];
        const int pw1[3] = {cart_pow_rt(la, fa.comp, 0), cart_pow_rt(la, fa.comp, 1), cart_pow_rt(la, fa.comp, 2)};
        const int pw2[3] = {cart_pow_rt(lb, fb.comp, 0), cart_pow_rt(lb, fb.comp, 1), cart_pow_rt(lb, fb.comp, 2)};
        const double *A = g.centre + 3 * fa.shell, *B = g.centre + 3 * fb.shell;
        const double scale = comp_scale_rt(la, fa.comp) * comp_scale_rt(lb, fb.comp);
        const double Pij = g.P[(size_t)i * g.N + j], Wij = g.W[(size_t)i * g.N + j];
        if (Pij == 0.0 && Wij == 0.0) continue;
        double acc[2][3] = {{0, 0, 0}, {0, 0, 0}};         // centre of i / centre of j
        const int Lt = la + lb + 1;
        for (int ia = 0; ia < g.nprim[fa.shell]; ++ia)
            for (int ib = 0; ib < g.nprim[fb.shell]; ++ib) {
                const double a = g.exps[g.poff[fa.shell] + ia], b = g.exps[g.poff[fb.shell] + ib];
                const double cc = g.coefs[g.poff[fa.shell] + ia] * g.coefs[g.poff[fb.shell] + ib] * scale;
                const double p = a + b;
                const double pref = pow(M_PI / p, 1.5);
                E1G E[3];
                for (int dm = 0; dm < 3; ++dm) build_E1g(E[dm], pw1[dm] + 1, pw2[dm] + 3, A[dm] - B[dm], a, b);
                double S1[3], T1[3];
                for (int dm = 0; dm < 3; ++dm) { S1[dm] = E[dm].v[pw1[dm]][pw2[dm]][0]; T1[dm] = kin1(E[dm], pw1[dm], pw2[dm], b); }
                for (int which = 0; which < 2; ++which)
                    for (int x = 0; x < 3; ++x) {
                        const int y = (x + 1) % 3, z = (x + 2) % 3;
                        const double dS = d1([&](int ii, int jj) { return E[x].v[ii][jj][0]; }, pw1[x], pw2[x], which, a, b);
                        const double dT = d1([&](int ii, int jj) { return kin1(E[x], ii, jj, b); }, pw1[x], pw2[x], which, a, b);
                        const double ds3 = dS * S1[y] * S1[z] * pref;                                            // cython/grad.pyx:352-397
                        const double dt3 = (dT * S1[y] * S1[z] + dS * (T1[y] * S1[z] + S1[y] * T1[z])) * pref;  // grad.pyx:400-506
                        acc[which][x] += cc * (2.0 * Pij * dt3 - 2.0 * Wij * ds3);
                    }
                // nuclear attraction: centre derivatives (grad.pyx:553-629) and operator derivative (grad.pyx:509-550)
                const double Px = (a * A[0] + b * B[0]) / p, Py = (a * A[1] + b * B[1]) / p, Pz = (a * A[2] + b * B[2]) / p;
                for (int at = 0; at < g.natom; ++at) {
                    const double X = Px - g.xyz[3 * at], Y = Py - g.xyz[3 * at + 1], Zc = Pz - g.xyz[3 * at + 2];
                    double Fs[KB_MAXL + 2];
                    boys_eval_rt(Lt, p * (X * X + Y * Y + Zc * Zc), g.boys[Lt], Fs);
                    double sc = 1.0;
                    for (int n = 0; n <= Lt; ++n) { Fs[n] *= sc; sc *= -2.0 * p; }
                    double R[nherm(2 * GR_AM + 1)];
                    build_R_rt(R, Lt, Fs, X, Y, Zc);
                    const double vpre = cc * (2.0 * M_PI / p) * (-g.Z[at]) * 2.0 * Pij;
                    double op[3] = {0, 0, 0};
                    for (int t = 0; t <= pw1[0] + pw2[0]; ++t)
                        for (int u = 0; u <= pw1[1] + pw2[1]; ++u)
                            for (int v = 0; v <= pw1[2] + pw2[2]; ++v) {
                                const double eee = E[0].v[pw1[0]][pw2[0]][t] * E[1].v[pw1[1]][pw2[1]][u] * E[2].v[pw1[2]][pw2[2]][v];
                                op[0] -= eee * R[hidx(t + 1, u, v)];
                                op[1] -= eee * R[hidx(t, u + 1, v)];
                                op[2] -= eee * R[hidx(t, u, v + 1)];
                            }
                    for (int x = 0; x < 3; ++x) atomicAdd(&g.grad[3 * at + x], vpre * op[x]);
                    for (int which = 0; which < 2; ++which)
                        for (int x = 0; x < 3; ++x) {
                            const int tm[3] = {pw1[0] + pw2[0] + (x == 0), pw1[1] + pw2[1] + (x == 1), pw1[2] + pw2[2] + (x == 2)};
                            double val = 0.0;
                            for (int t = 0; t <= tm[0]; ++t) {
                                const double e0 = (x == 0) ? d1([&](int ii, int jj) { return E[0].v[ii][jj][t]; }, pw1[0], pw2[0], which, a, b)
                                                           : E[0].v[pw1[0]][pw2[0]][t];
                                for (int u = 0; u <= tm[1]; ++u) {
                                    const double e1 = (x == 1) ? d1([&](int ii, int jj) { return E[1].v[ii][jj][u]; }, pw1[1], pw2[1], which, a, b)
                                                               : E[1].v[pw1[1]][pw2[1]][u];
                                    for (int v = 0; v <= tm[2]; ++v) {
                                        const double e2 = (x == 2) ? d1([&](int ii, int jj) { return E[2].v[ii][jj][v]; }, pw1[2], pw2[2], which, a, b)
                                                                   : E[2].v[pw1[2]][pw2[2]][v];
                                        val += e0 * e1 * e2 * R[hidx(t, u, v)];
                                    }
                                }
                            }
                            acc[which][x] += vpre * val;
                        }
                }
            }
        const int atA = g.shell_atom[fa.shell], atB = g.shell_atom[fb.shell];
        for (int x = 0; x < 3; ++x) {
            atomicAdd(&g.grad[3 * atA + x], acc[0][x]);
            atomicAdd(&g.grad[3 * atB + x], acc[1][x]);
        }
    }
}

}  // namespace

// dE/dX of the RHF energy, split into its one-electron (kinetic + nuclear attraction + overlap/energy-weighted-density),
// two-electron and nuclear-repulsion parts, each [natom][3] (host).  P = C_occ C_occ^T and W = P F P are real (N,N) host
// matrices in device function order; shell_atom[s] is the atom a shell sits on.  mmd/forces.py:8-99 of the reference.
extern "C" int mmdb_gradient_host(mmdb_basis *b, int natom, const double *Z, const double *xyz, const int *shell_atom,
                                  const double *P, const double *W, double *grad_1e, double *grad_2e, double *grad_nuc)
{
    if (!b || natom <= 0 || !Z || !xyz || !shell_atom || !P || !W || !grad_1e || !grad_2e || !grad_nuc)
        return fail(MMDB_ERR_INVALID, "mmdb_gradient_host: bad arguments");
    if (!b->have_schwarz) return fail(MMDB_ERR_INVALID, "mmdb_gradient_host: call mmdb_schwarz first");
    CU(cudaSetDevice(b->device));
    const int N = b->nbf, ns = b->nshell;
    const size_t N2 = (size_t)N * N;
    // nuclear repulsion (mmd/forces.py:45-51)
    for (int a = 0; a < natom; ++a)
        for (int x = 0; x < 3; ++x) {
            double s = 0.0;
            for (int c = 0; c < natom; ++c) {
                if (c == a) continue;
                const double dx = xyz[3 * a] - xyz[3 * c], dy = xyz[3 * a + 1] - xyz[3 * c + 1], dz = xyz[3 * a + 2] - xyz[3 * c + 2];
                const double r = std::sqrt(dx * dx + dy * dy + dz * dz);
                if (r > 1e-12) s += -(xyz[3 * a + x] - xyz[3 * c + x]) * Z[a] * Z[c] / (r * r * r);
            }
            grad_nuc[3 * a + x] = s;
        }
    std::vector<FnInfoG> fn(N, FnInfoG{-1, 0});
    std::vector<int> am(ns), np(ns), po(ns);
    std::vector<double> cen(3 * ns);
    for (int s = 0; s < ns; ++s) {
        am[s] = b->sh[s].am; np[s] = b->sh[s].nprim; po[s] = b->sh[s].poff;
        cen[3 * s] = b->sh[s].x; cen[3 * s + 1] = b->sh[s].y; cen[3 * s + 2] = b->sh[s].z;
        for (int c = 0; c < ncart(am[s]); ++c) fn[b->sh[s].bf0 + c] = FnInfoG{s, c};
    }
    double pmax = 0.0;
    for (size_t x = 0; x < N2; ++x) pmax = std::max(pmax, std::fabs(P[x]));
    // Boys tables up to order 9 (one past the energy kernels)
    std::vector<double *> tabs(GR_L + 1, nullptr);
    for (int L = 0; L <= GR_L; ++L) {
        std::vector<double> t;
        mmdb_make_boys_table(L, t);
        CU(cudaMalloc(&tabs[L], t.size() * sizeof(double)));
        CU(cudaMemcpy(tabs[L], t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    char *buf = nullptr;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t bytes_e = sizeof(double) * b->exps.size();
    const size_t total = up(sizeof(FnInfoG) * N) + 4 * up(sizeof(int) * ns) + up(sizeof(double) * 3 * ns) + 2 * up(bytes_e) +
                         up(sizeof(double) * natom) + up(sizeof(double) * 3 * natom) + up(sizeof(double *) * (GR_L + 1)) +
                         2 * up(sizeof(double) * N2) + 2 * up(sizeof(double) * 3 * natom);
    CU(cudaMalloc(&buf, total));
    char *cur = buf;
    auto put = [&](const void *src, size_t n) -> void * {
        void *dst = cur;
        if (src) cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
        else cudaMemset(dst, 0, n);
        cur += up(n);
        return dst;
    };
    Grad1eArgs g1;
    g1.N = N; g1.natom = natom;
    g1.fn = (const FnInfoG *)put(fn.data(), sizeof(FnInfoG) * N);
    g1.am = (const int *)put(am.data(), sizeof(int) * ns);
    g1.nprim = (const int *)put(np.data(), sizeof(int) * ns);
    g1.poff = (const int *)put(po.data(), sizeof(int) * ns);
    g1.shell_atom = (const int *)put(shell_atom, sizeof(int) * ns);
    g1.centre = (const double *)put(cen.data(), sizeof(double) * 3 * ns);
    g1.exps = (const double *)put(b->exps.data(), bytes_e);
    g1.coefs = (const double *)put(b->coefs.data(), bytes_e);
    g1.Z = (const double *)put(Z, sizeof(double) * natom);
    g1.xyz = (const double *)put(xyz, sizeof(double) * 3 * natom);
    g1.boys = (const double *const *)put(tabs.data(), sizeof(double *) * (GR_L + 1));
    g1.P = (const double *)put(P, sizeof(double) * N2);
    g1.W = (const double *)put(W, sizeof(double) * N2);
    double *d_g1 = (double *)put(nullptr, sizeof(double) * 3 * natom);
    double *d_g2 = (double *)put(nullptr, sizeof(double) * 3 * natom);
    g1.grad = d_g1;
    {
        const long long npair = (long long)N * N;
        const int grid = (int)std::min<long long>((npair + 63) / 64, (long long)b->nsm * 32);
        grad1e_kernel<<<grid, 64>>>(g1);
    }
    int rc = MMDB_OK;
    for (int cb = 0; cb < MMDB_NCLASS_PAIR && rc == MMDB_OK; ++cb)
        for (int ck = 0; ck < MMDB_NCLASS_PAIR; ++ck) {          // every ORDERED class pair: the derivative acts on the bra
            PairClass &B = b->pc[cb], &K = b->pc[ck];
            if (B.npairs == 0 || K.npairs == 0) continue;
            Grad2eArgs g2;
            g2.braH = B.hdr_dev; g2.ketH = K.hdr_dev; g2.braP = B.prim_dev; g2.ketP = K.prim_dev; g2.braAB = B.prim_ab_dev;
            g2.Qs_bra = B.Qs_dev; g2.Qs_ket = K.Qs_dev;
            g2.nbra = B.npairs; g2.nket = K.npairs; g2.la = B.la; g2.lb = B.lb; g2.lc = K.la; g2.ld = K.lb;
            g2.N = N; g2.P = g1.P; g2.shell_atom = g1.shell_atom;
            g2.boys_tab = tabs[B.la + B.lb + K.la + K.lb + 1];
            g2.cut = 1e-15; g2.pmax2 = 16.0 * pmax * pmax;
            g2.grad = d_g2;
            const long long totalw = (long long)B.npairs * K.npairs * ncart(K.la) * ncart(K.lb);
            const int grid = (int)std::min<long long>((totalw + 63) / 64, (long long)b->nsm * 32);
            grad2e_kernel<<<grid, 64, BOYS_ROWS * BOYS_STRIDE * sizeof(double)>>>(g2);
        }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail(MMDB_ERR_CUDA, std::string("gradient kernels: ") + cudaGetErrorString(e));
    if (rc == MMDB_OK) {
        cudaMemcpy(grad_1e, d_g1, sizeof(double) * 3 * natom, cudaMemcpyDeviceToHost);
        cudaMemcpy(grad_2e, d_g2, sizeof(double) * 3 * natom, cudaMemcpyDeviceToHost);
    }
    cudaFree(buf);
    for (double *t : tabs) cudaFree(t);
    return rc;
}

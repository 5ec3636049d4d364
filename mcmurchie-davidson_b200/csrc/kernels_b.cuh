// kernels_b.cuh — generic (runtime angular momentum) shell-quartet kernel: the independent cross-check of the
// class-specialised kernels (impl = 1 of mmdb_eri_shell_quartets; it stores integrals, it does not digest).  One thread per (shell quartet, ket component pair); R, E and the Hermite
// intermediate are thread-local arrays with runtime indexing.
#pragma once
#include "core.cuh"
#include "kernels_a.cuh"

namespace mmdb {

constexpr int KB_THREADS = 128;
constexpr int KB_MAXAM = 2;
constexpr int KB_MAXL = 4 * KB_MAXAM;

__device__ __forceinline__ int cart_pow_rt(int l, int c, int dim)
{
    // l <= 2, reference component order
    const int tab[3][6][3] = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}},
                              {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}},
                              {{2, 0, 0}, {1, 1, 0}, {1, 0, 1}, {0, 2, 0}, {0, 1, 1}, {0, 0, 2}}};
    return tab[l][c][dim];
}

__device__ __forceinline__ double comp_scale_rt(int l, int c)
{
    return (l == 2 && (c == 0 || c == 3 || c == 5)) ? 0.57735026918962576451 : 1.0;
}

typedef double ETabRT[3][KB_MAXAM + 1][KB_MAXAM + 1][2 * KB_MAXAM + 1];

__device__ __forceinline__ void build_E_rt(ETabRT &E, int la, int lb, const double *PA, const double *PB, double oo2p)
{
    for (int dim = 0; dim < 3; ++dim) {
        E[dim][0][0][0] = 1.0;
        for (int i = 1; i <= la; ++i)
            for (int t = 0; t <= i; ++t) {
                double x = 0.0;
                if (t > 0) x = oo2p * E[dim][i - 1][0][t - 1];
                if (t <= i - 1) x = fma(PA[dim], E[dim][i - 1][0][t], x);
                if (t + 1 <= i - 1) x = fma((double)(t + 1), E[dim][i - 1][0][t + 1], x);
                E[dim][i][0][t] = x;
            }
        for (int j = 1; j <= lb; ++j)
            for (int i = 0; i <= la; ++i)
                for (int t = 0; t <= i + j; ++t) {
                    double x = 0.0;
                    if (t > 0) x = oo2p * E[dim][i][j - 1][t - 1];
                    if (t <= i + j - 1) x = fma(PB[dim], E[dim][i][j - 1][t], x);
                    if (t + 1 <= i + j - 1) x = fma((double)(t + 1), E[dim][i][j - 1][t + 1], x);
                    E[dim][i][j][t] = x;
                }
    }
}

__device__ __forceinline__ void build_R_rt(double *R, int L, const double *Fs, double X, double Y, double Z)
{
    R[0] = Fs[L];
    for (int n = L - 1; n >= 0; --n) {
        for (int d = L - n; d >= 1; --d)
            for (int t = d; t >= 0; --t)
                for (int u = d - t; u >= 0; --u) {
                    const int v = d - t - u;
                    double x;
                    if (t > 0) {
                        x = X * R[hidx(t - 1, u, v)];
                        if (t > 1) x = fma((double)(t - 1), R[hidx(t - 2, u, v)], x);
                    } else if (u > 0) {
                        x = Y * R[hidx(t, u - 1, v)];
                        if (u > 1) x = fma((double)(u - 1), R[hidx(t, u - 2, v)], x);
                    } else {
                        x = Z * R[hidx(t, u, v - 1)];
                        if (v > 1) x = fma((double)(v - 1), R[hidx(t, u, v - 2)], x);
                    }
                    R[hidx(t, u, v)] = x;
                }
        R[0] = Fs[n];
    }
}

static __global__ void __launch_bounds__(KB_THREADS) eri_generic_kernel(const EriArgs a, int la, int lb, int lc, int ld)
{
    extern __shared__ double s_boys[];
    const int NA = ncart(la), NB = ncart(lb), NC = ncart(lc), ND = ncart(ld);
    const int NAB = NA * NB, NCD = NC * ND;
    const int LBRA = la + lb, L = la + lb + lc + ld;
    (void)NC;
    const unsigned long long n = a.count_dev ? *a.count_dev : a.n;
    const unsigned long long total = n * (unsigned long long)NCD;
    if ((unsigned long long)blockIdx.x * blockDim.x >= total) return;      // no work: skip the table staging
    for (int x = threadIdx.x; x < BOYS_ROWS * BOYS_STRIDE; x += blockDim.x) s_boys[x] = __ldg(a.boys_tab + x);
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const unsigned long long e = w / NCD;
        const int cd = (int)(w % NCD);
        const int c = cd / ND, d = cd % ND;
        const int cx = cart_pow_rt(lc, c, 0), cy = cart_pow_rt(lc, c, 1), cz = cart_pow_rt(lc, c, 2);
        const int dx = cart_pow_rt(ld, d, 0), dy = cart_pow_rt(ld, d, 1), dz = cart_pow_rt(ld, d, 2);
        const uint2 ij = __ldg(a.list + (long long)e * a.list_step);
        const PairHdr bh = ld_hdr(a.braH + ij.x);
        const PairHdr kh = ld_hdr(a.ketH + ij.y);
        double out[36];
        for (int x = 0; x < NAB; ++x) out[x] = 0.0;
        const int ib0 = 0, ib1 = bh.pnum;
        for (int ib = ib0; ib < ib1; ++ib) {
            const PrimPair b = ld_prim(a.braP + bh.poff + ib);
            ETabRT Eb;
            {
                const double PA[3] = {b.PAx, b.PAy, b.PAz};
                const double PB[3] = {b.PAx + bh.ABx, b.PAy + bh.ABy, b.PAz + bh.ABz};
                build_E_rt(Eb, la, lb, PA, PB, 0.5 / b.p);
            }
            double G[nherm(2 * KB_MAXAM)];
            const int nhb = nherm(LBRA);
            for (int x = 0; x < nhb; ++x) G[x] = 0.0;
            for (int ik = 0; ik < kh.pnum; ++ik) {
                const PrimPair k = ld_prim(a.ketP + kh.poff + ik);
                const double rs = fast_rsqrt(b.p + k.p);
                const double alpha = b.p * k.p * (rs * rs);
                const double X = b.Px - k.Px, Y = b.Py - k.Py, Z = b.Pz - k.Pz;
                const double T = alpha * (X * X + Y * Y + Z * Z);
                double Fs[KB_MAXL + 1];
                boys_eval_rt(L, T, s_boys, Fs);
                {
                    double s = b.cc * k.cc * sqrt(b.p * k.p) * rs * (1.0 / SQRTPI_2);      // PrimPair::cc carries 1/sqrt(p) and sqrt(sqrt(pi)/2); boys_eval_rt returns the true F_m
                    const double m2a = -2.0 * alpha;
                    for (int nn = 0; nn <= L; ++nn) { Fs[nn] *= s; s *= m2a; }
                }
                double R[nherm(KB_MAXL)];
                build_R_rt(R, L, Fs, X, Y, Z);
                ETabRT Ek;
                {
                    const double QC[3] = {k.PAx, k.PAy, k.PAz};
                    const double QD[3] = {k.PAx + kh.ABx, k.PAy + kh.ABy, k.PAz + kh.ABz};
                    build_E_rt(Ek, lc, ld, QC, QD, 0.5 / k.p);
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
            for (int ab = 0; ab < NAB; ++ab) {
                const int aa = ab / NB, bb = ab % NB;
                const int ax = cart_pow_rt(la, aa, 0), ay = cart_pow_rt(la, aa, 1), az = cart_pow_rt(la, aa, 2);
                const int bx = cart_pow_rt(lb, bb, 0), by = cart_pow_rt(lb, bb, 1), bz = cart_pow_rt(lb, bb, 2);
                double acc = out[ab];
                for (int t = 0; t <= ax + bx; ++t)
                    for (int u = 0; u <= ay + by; ++u)
                        for (int v = 0; v <= az + bz; ++v)
                            acc = fma(Eb[0][ax][bx][t] * Eb[1][ay][by][u] * Eb[2][az][bz][v], G[hidx(t, u, v)], acc);
                out[ab] = acc;
            }
        }
        const double scd = comp_scale_rt(lc, c) * comp_scale_rt(ld, d);
        double *o = a.out + e * (unsigned long long)(NAB * NCD);
        for (int ab = 0; ab < NAB; ++ab)
            o[ab * NCD + cd] = out[ab] * (scd * comp_scale_rt(la, ab / NB) * comp_scale_rt(lb, ab % NB));
    }
}

}  // namespace mmdb

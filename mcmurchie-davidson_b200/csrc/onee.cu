// onee.cu — one-electron integrals S, T, V, dipole M, angular momentum L on the device.
//
// Enabler outside the graded two-electron path: the reference's
// Python-object loops (cython/onee.pyx:14-192, mmd/molecule.py:253-276) are O(N^2 N_atoms) at
// interpreter speed and make the large configurations unrunnable end-to-end.  Same formulas,
// evaluated with one thread per basis-function pair i >= j, mirrored like the reference does.
#include <algorithm>
#include <string>
#include <vector>

#include "handle.h"
#include "kernels_b.cuh"

using namespace mmdb;

namespace {

constexpr int OE_MAXI = 3 + 1;   // i up to la+1 (angular momentum operator)
constexpr int OE_MAXJ = 2 + 2 + 1;   // j up to lb+2 (kinetic energy)
constexpr int OE_MAXT = 8;

struct FnInfo { int shell, comp; };

struct OneeArgs {
    int N, natom, nshell;
    const FnInfo *fn;
    const int *am, *nprim, *poff;
    const double *centre, *exps, *coefs;
    const double *Z, *xyz;
    double ox, oy, oz;
    const double *const *boys;   // boys[L] tables (global)
    double *S, *T, *V, *M, *L;
};

// E_t^{ij} including the exp(-mu Q^2) factor, cython/util.pxi:13-26, tabulated for i<=imax, j<=jmax
struct E1D { double v[OE_MAXI][OE_MAXJ][OE_MAXT]; };

__device__ void build_E1(E1D &E, int imax, int jmax, double Q, double a, double b)
{
    const double p = a + b, u = a * b / p, oo2p = 1.0 / (2 * p);
    const double PA = -(u * Q / a), PB = (u * Q / b);
    for (int i = 0; i <= imax; ++i)
        for (int j = 0; j <= jmax; ++j)
            for (int t = 0; t < OE_MAXT; ++t) E.v[i][j][t] = 0.0;
    E.v[0][0][0] = exp(-u * Q * Q);
    for (int i = 1; i <= imax; ++i)
        for (int t = 0; t <= i; ++t) {
            double x = PA * E.v[i - 1][0][t];
            if (t > 0) x += oo2p * E.v[i - 1][0][t - 1];
            if (t + 1 < OE_MAXT) x += (t + 1) * E.v[i - 1][0][t + 1];
            E.v[i][0][t] = x;
        }
    for (int j = 1; j <= jmax; ++j)
        for (int i = 0; i <= imax; ++i)
            for (int t = 0; t <= i + j; ++t) {
                double x = PB * E.v[i][j - 1][t];
                if (t > 0) x += oo2p * E.v[i][j - 1][t - 1];
                if (t + 1 < OE_MAXT) x += (t + 1) * E.v[i][j - 1][t + 1];
                E.v[i][j][t] = x;
            }
}

__global__ void __launch_bounds__(64) onee_kernel(const OneeArgs g)
{
    const long long npair = (long long)g.N * (g.N + 1) / 2;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < npair; w += (long long)gridDim.x * blockDim.x) {
        // w -> (i >= j)
        int i = (int)((sqrt(8.0 * (double)w + 1.0) - 1.0) * 0.5);
        while ((long long)i * (i + 1) / 2 > w) --i;
        while ((long long)(i + 1) * (i + 2) / 2 <= w) ++i;
        const int j = (int)(w - (long long)i * (i + 1) / 2);
        const FnInfo fa = g.fn[i], fb = g.fn[j];
        const int la = g.am[fa.shell], lb = g.am[fb.shell];
        const int l1 = cart_pow_rt(la, fa.comp, 0), m1 = cart_pow_rt(la, fa.comp, 1), n1 = cart_pow_rt(la, fa.comp, 2);
        const int l2 = cart_pow_rt(lb, fb.comp, 0), m2 = cart_pow_rt(lb, fb.comp, 1), n2 = cart_pow_rt(lb, fb.comp, 2);
        const double *A = g.centre + 3 * fa.shell, *B = g.centre + 3 * fb.shell;
        const double scale = comp_scale_rt(la, fa.comp) * comp_scale_rt(lb, fb.comp);
        double s = 0, t = 0, v = 0, mu[3] = {0, 0, 0}, ll[3] = {0, 0, 0};
        for (int ia = 0; ia < g.nprim[fa.shell]; ++ia)
            for (int ib = 0; ib < g.nprim[fb.shell]; ++ib) {
                const double a = g.exps[g.poff[fa.shell] + ia], b = g.exps[g.poff[fb.shell] + ib];
                const double cc = g.coefs[g.poff[fa.shell] + ia] * g.coefs[g.poff[fb.shell] + ib];
                const double p = a + b;
                const double pref = pow(M_PI / p, 1.5);
                E1D Ex, Ey, Ez;
                build_E1(Ex, l1 + 1, l2 + 2, A[0] - B[0], a, b);
                build_E1(Ey, m1 + 1, m2 + 2, A[1] - B[1], a, b);
                build_E1(Ez, n1 + 1, n2 + 2, A[2] - B[2], a, b);
                const double Sx = Ex.v[l1][l2][0], Sy = Ey.v[m1][m2][0], Sz = Ez.v[n1][n2][0];
                // overlap, onee.pyx:71-78
                s += cc * Sx * Sy * Sz * pref;
                // kinetic, onee.pyx:108-137
                {
                    const double Bq = -2.0 * b * b;
                    double Tx = (2 * l2 + 1) * b * Sx + Bq * Ex.v[l1][l2 + 2][0];
                    if (l2 >= 2) Tx += -0.5 * l2 * (l2 - 1) * Ex.v[l1][l2 - 2][0];
                    double Ty = (2 * m2 + 1) * b * Sy + Bq * Ey.v[m1][m2 + 2][0];
                    if (m2 >= 2) Ty += -0.5 * m2 * (m2 - 1) * Ey.v[m1][m2 - 2][0];
                    double Tz = (2 * n2 + 1) * b * Sz + Bq * Ez.v[n1][n2 + 2][0];
                    if (n2 >= 2) Tz += -0.5 * n2 * (n2 - 1) * Ez.v[n1][n2 - 2][0];
                    t += cc * (Tx * Sy * Sz + Ty * Sx * Sz + Tz * Sx * Sy) * pref;
                }
                const double Px = (a * A[0] + b * B[0]) / p, Py = (a * A[1] + b * B[1]) / p, Pz = (a * A[2] + b * B[2]) / p;
                // dipole, onee.pyx:81-106
                {
                    const double Dx = Ex.v[l1][l2][1] + (Px - g.ox) * Sx;
                    const double Dy = Ey.v[m1][m2][1] + (Py - g.oy) * Sy;
                    const double Dz = Ez.v[n1][n2][1] + (Pz - g.oz) * Sz;
                    mu[0] += cc * Dx * Sy * Sz * pref;
                    mu[1] += cc * Sx * Dy * Sz * pref;
                    mu[2] += cc * Sx * Sy * Dz * pref;
                }
                // angular momentum r x del, onee.pyx:140-176
                {
                    const double S1x = Ex.v[l1 + 1][l2][0] + (A[0] - g.ox) * Sx;
                    const double S1y = Ey.v[m1 + 1][m2][0] + (A[1] - g.oy) * Sy;
                    const double S1z = Ez.v[n1 + 1][n2][0] + (A[2] - g.oz) * Sz;
                    const double D1x = (l2 > 0 ? l2 * Ex.v[l1][l2 - 1][0] : 0.0) - 2 * b * Ex.v[l1][l2 + 1][0];
                    const double D1y = (m2 > 0 ? m2 * Ey.v[m1][m2 - 1][0] : 0.0) - 2 * b * Ey.v[m1][m2 + 1][0];
                    const double D1z = (n2 > 0 ? n2 * Ez.v[n1][n2 - 1][0] : 0.0) - 2 * b * Ez.v[n1][n2 + 1][0];
                    ll[0] += cc * (-Sx * (S1y * D1z - S1z * D1y) * pref);
                    ll[1] += cc * (-Sy * (S1z * D1x - S1x * D1z) * pref);
                    ll[2] += cc * (-Sz * (S1x * D1y - S1y * D1x) * pref);
                }
                // nuclear attraction, onee.pyx:178-192, summed over nuclei like mmd/molecule.py:266-268
                {
                    const int Lt = la + lb;
                    double vsum = 0.0;
                    for (int at = 0; at < g.natom; ++at) {
                        const double X = Px - g.xyz[3 * at], Y = Py - g.xyz[3 * at + 1], Zc = Pz - g.xyz[3 * at + 2];
                        const double Tt = p * (X * X + Y * Y + Zc * Zc);
                        double Fs[KB_MAXL + 1];
                        boys_eval_rt(Lt, Tt, g.boys[Lt], Fs);
                        double sc = 1.0;
                        for (int n = 0; n <= Lt; ++n) { Fs[n] *= sc; sc *= -2.0 * p; }
                        double R[nherm(4)];
                        build_R_rt(R, Lt, Fs, X, Y, Zc);
                        double val = 0.0;
                        for (int tt = 0; tt <= l1 + l2; ++tt)
                            for (int uu = 0; uu <= m1 + m2; ++uu)
                                for (int vv = 0; vv <= n1 + n2; ++vv)
                                    val += Ex.v[l1][l2][tt] * Ey.v[m1][m2][uu] * Ez.v[n1][n2][vv] * R[hidx(tt, uu, vv)];
                        vsum += -g.Z[at] * val;
                    }
                    v += cc * vsum * (2.0 * M_PI / p);
                }
            }
        const size_t N = g.N;
        const size_t ij = (size_t)i * N + j, ji = (size_t)j * N + i;
        g.S[ij] = g.S[ji] = s * scale;
        g.T[ij] = g.T[ji] = t * scale;
        g.V[ij] = g.V[ji] = v * scale;
        for (int d = 0; d < 3; ++d) {
            g.M[d * N * N + ij] = g.M[d * N * N + ji] = mu[d] * scale;
            g.L[d * N * N + ij] = ll[d] * scale;
            g.L[d * N * N + ji] = -(ll[d] * scale);   // mmd/molecule.py:276 (diagonal ends up -L_ii like the reference)
        }
    }
}

}  // namespace

extern "C" int mmdb_onee_host(mmdb_basis *b, int natom, const double *Z, const double *xyz, const double *origin,
                              double *S, double *T, double *V, double *M, double *L)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    CU(cudaSetDevice(b->device));
    const int N = b->nbf, ns = b->nshell;
    const size_t N2 = (size_t)N * N;
    std::vector<FnInfo> fn(N, FnInfo{-1, 0});
    std::vector<int> am(ns), np(ns), po(ns);
    std::vector<double> cen(3 * ns);
    for (int s = 0; s < ns; ++s) {
        am[s] = b->sh[s].am; np[s] = b->sh[s].nprim; po[s] = b->sh[s].poff;
        cen[3 * s] = b->sh[s].x; cen[3 * s + 1] = b->sh[s].y; cen[3 * s + 2] = b->sh[s].z;
        for (int c = 0; c < ncart(am[s]); ++c) fn[b->sh[s].bf0 + c] = FnInfo{s, c};
    }
    for (int i = 0; i < N; ++i)
        if (fn[i].shell < 0) return fail(MMDB_ERR_INVALID, "mmdb_onee_host: basis functions are not covered by shells");
    char *buf = nullptr;
    const size_t bytes_fn = sizeof(FnInfo) * N, bytes_i = sizeof(int) * ns, bytes_c = sizeof(double) * 3 * ns;
    const size_t bytes_e = sizeof(double) * b->exps.size(), bytes_at = sizeof(double) * natom;
    const size_t bytes_tab = sizeof(double *) * (BOYS_MAXL + 1);
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t total = up(bytes_fn) + 3 * up(bytes_i) + up(bytes_c) + 2 * up(bytes_e) + up(bytes_at) + up(3 * bytes_at) +
                         up(bytes_tab) + up(sizeof(double) * N2 * 9);
    CU(cudaMalloc(&buf, total));
    char *cur = buf;
    auto put = [&](const void *src, size_t n) -> void * {
        void *dst = cur;
        if (src) cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
        cur += up(n);
        return dst;
    };
    OneeArgs g;
    g.N = N; g.natom = natom; g.nshell = ns;
    g.fn = (const FnInfo *)put(fn.data(), bytes_fn);
    g.am = (const int *)put(am.data(), bytes_i);
    g.nprim = (const int *)put(np.data(), bytes_i);
    g.poff = (const int *)put(po.data(), bytes_i);
    g.centre = (const double *)put(cen.data(), bytes_c);
    g.exps = (const double *)put(b->exps.data(), bytes_e);
    g.coefs = (const double *)put(b->coefs.data(), bytes_e);
    g.Z = (const double *)put(Z, bytes_at);
    g.xyz = (const double *)put(xyz, 3 * bytes_at);
    g.boys = (const double *const *)put(b->boys_dev, bytes_tab);
    double *outd = (double *)put(nullptr, sizeof(double) * N2 * 9);
    g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
    g.S = outd; g.T = outd + N2; g.V = outd + 2 * N2; g.M = outd + 3 * N2; g.L = outd + 6 * N2;
    const long long npair = (long long)N * (N + 1) / 2;
    const int grid = (int)std::min<long long>((npair + 63) / 64, (long long)b->nsm * 32);
    onee_kernel<<<grid, 64>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(buf);
        return fail(MMDB_ERR_CUDA, std::string("onee kernel: ") + cudaGetErrorString(e));
    }
    cudaMemcpy(S, g.S, sizeof(double) * N2, cudaMemcpyDeviceToHost);
    cudaMemcpy(T, g.T, sizeof(double) * N2, cudaMemcpyDeviceToHost);
    cudaMemcpy(V, g.V, sizeof(double) * N2, cudaMemcpyDeviceToHost);
    cudaMemcpy(M, g.M, sizeof(double) * 3 * N2, cudaMemcpyDeviceToHost);
    cudaMemcpy(L, g.L, sizeof(double) * 3 * N2, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    return MMDB_OK;
}

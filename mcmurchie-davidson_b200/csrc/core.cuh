// core.cuh — device building blocks of the B200 McMurchie-Davidson two-electron engine.
//
// What the reference does per basis-function quartet with naive recursion
// (cython/twoe.pyx:56-94 electron_repulsion, cython/util.pxi:13-50 E and R, util.pxi:54-55 boys)
// is done here per SHELL quartet:
//   - Boys F_0..F_L(T): shared-memory-staged Taylor table (step 1/8, 9 terms) for F_L, downward
//     recursion for the rest, asymptotic + upward recursion for T >= 40;
//   - Hermite Coulomb integrals R_tuv: one in-place recursion over the auxiliary index n, entirely
//     in registers for the class-specialised kernels;
//   - Hermite expansion coefficients E_t^{ij}: per primitive PAIR, rebuilt in registers from
//     (P-A, P-B, 1/2p), normalised by E_0^{00} (which lives in the pair prefactor);
//   - ket Hermite->Cartesian transform per primitive quartet, bra transform once per bra
//     primitive pair after the ket primitives have been summed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace mmdb {

// ------------------------------------------------------------------------------------------
// compile-time geometry of Cartesian shells and Hermite index sets
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int nherm(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }

// Cartesian component c of a shell with angular momentum l, reference order
// (mmd/molecule.py:108-114): p = x,y,z ; d = xx,xy,xz,yy,yz,zz.
__host__ __device__ constexpr int cart_pow(int l, int c, int dim)
{
    int idx = 0;
    for (int i = l; i >= 0; --i)
        for (int j = l - i; j >= 0; --j) {
            if (idx == c) return dim == 0 ? i : (dim == 1 ? j : l - i - j);
            ++idx;
        }
    return 0;
}

// Packed index of Hermite triple (t,u,v): degree-major, independent of the maximum degree.
__host__ __device__ constexpr int hidx(int t, int u, int v)
{
    const int n = t + u + v;
    return n * (n + 1) * (n + 2) / 6 + (n - t) * (n - t + 1) / 2 + v;
}

// 1/sqrt((2l-1)!!(2m-1)!!(2n-1)!!): the per-component part of the primitive norm
// (cython/basis.pxi:102-105).  Only d has non-trivial values.
__host__ __device__ constexpr double comp_scale(int l, int c)
{
    double s = 1.0;
    for (int dim = 0; dim < 3; ++dim) {
        int k = cart_pow(l, c, dim);
        if (k == 2) s *= 0.57735026918962576451;          // 1/sqrt(3)
        if (k == 3) s *= 0.25819888974716112568;          // 1/sqrt(15)
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// Shell TYPE codes of the class kernels.  0, 1, 2 = s, p, d shells.  3 = "S2": a generally contracted s pseudo-shell —
// TWO s-type contractions over ONE primitive set on one centre (cc-pVDZ's first two s functions of every heavy atom),
// i.e. two consecutive basis functions that share every primitive integral and differ only in their contraction
// coefficients.  A pair that contains an S2 shell keeps the geometric part of its pair coefficient in PrimPair::cc
// and one WEIGHT c_a^(i) c_b^(j) per component pair beside it; the kernels evaluate each primitive quartet once and
// apply the weights where the Hermite -> Cartesian coefficients are applied.  (The reference evaluates every
// contracted function quartet from scratch, cython/twoe.pyx:36-50.)
// ------------------------------------------------------------------------------------------
constexpr int SH_S2 = 3;
__host__ __device__ constexpr int am_of(int t) { return t == SH_S2 ? 0 : t; }
__host__ __device__ constexpr int ncomp(int t) { return t == SH_S2 ? 2 : ncart(t); }
__host__ __device__ constexpr int comp_pow(int t, int c, int dim) { return t == SH_S2 ? 0 : cart_pow(t, c, dim); }
__host__ __device__ constexpr double cscale(int t, int c) { return t == SH_S2 ? 1.0 : comp_scale(t, c); }
// weights of a pair of shell types: one per component pair of its S2 members
__host__ __device__ constexpr int nwgt(int ta, int tb) { return (ta == SH_S2 ? 2 : 1) * (tb == SH_S2 ? 2 : 1); }
__host__ __device__ constexpr int widx(int ta, int tb, int a, int b)
{
    return (ta == SH_S2 ? a : 0) * (tb == SH_S2 ? 2 : 1) + (tb == SH_S2 ? b : 0);
}
constexpr int MAX_WGT = 4;

template <int I, int N, class F>
__device__ __forceinline__ void sfor(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        sfor<I + 1, N>(f);
    }
}

// ------------------------------------------------------------------------------------------
// device-resident tables
// ------------------------------------------------------------------------------------------
struct __align__(16) PrimPair {   // one primitive pair of a shell pair, 64 B
    double Px, Py, Pz, p;          // Gaussian product centre, total exponent
    double cc;                     // c_a c_b exp(-mu|AB|^2) * sqrt(2) pi^(5/4) / p^(3/2)  (the extra 1/sqrt(p): see prim_R)
    double PAx, PAy, PAz;          // P - A   (P - B = PA + (A - B))
};

struct __align__(16) PairHdr {    // one shell pair, 64 B
    int bfA, bfB;                  // first device function index of shells A, B
    int poff, pnum;                // primitive-pair range in the class's PrimPair array
    int shA, shB;                  // shell ids (am[A] >= am[B]; equal am: A >= B)
    int pad0, pad1;
    double ABx, ABy, ABz;          // A - B
    double Qs;                     // max over function pairs sqrt|(pq|pq)|  (filled by mmdb_schwarz)
};


// ------------------------------------------------------------------------------------------
// explicit global-space accessors: the kernels receive their pointers inside a by-value struct, so
// the compiler cannot prove the address space and would emit generic LD / ATOM with run-time
// space checks (QSPC + shared-memory CAS fallback).  These force LDG(.nc) and RED.global.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_f64(double *p, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
// Accumulation into G on the per-function digestion path.  fixed == 0: FP64 atomics.
// fixed != 0: DETERMINISTIC mode — the contribution is rounded to a multiple of 2^-50 and added with a
// 64-bit INTEGER atomic; integer addition is associative, so G is bitwise reproducible for any thread
// schedule, any shard count and any all-reduce order (|G| < 2^13 is required; a Fock element is O(10)).
// In deterministic mode the screening kernel sends EVERY quartet to the per-function list, so the block
// digestion kernels (FP64 atomics, warp-level pre-reduction whose grouping depends on the list order) are
// not used at all: reproducibility costs speed, it is meant for parity runs.
constexpr double FIXED_SCALE = 1125899906842624.0;          // 2^50
__device__ __forceinline__ void red_add_g(int fixed, double *p, double v)
{
    if (fixed) {
        const long long q = __double2ll_rn(v * FIXED_SCALE);
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(q) : "memory");
    } else {
        red_add_f64(p, v);
    }
}
__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(p)));
}
__device__ __forceinline__ PrimPair ld_prim(const PrimPair *p)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    PrimPair r;
    r.Px = a.x; r.Py = a.y; r.Pz = b.x; r.p = b.y; r.cc = c.x; r.PAx = c.y; r.PAy = d.x; r.PAz = d.y;
    return r;
}
// Bra-side primitive pairs are read from a structure-of-arrays copy: field f of primitive k of pair i lives at
// S[f * N + row[k] + i], row[k] = number of (pair, primitive) slots of the primitives before k (pairs are sorted
// by primitive count, so primitive k exists for a PREFIX of the pairs and the rows need no padding).  The lanes
// of a warp walk consecutive bra pairs, so each of the eight loads touches a couple of 128-byte lines instead of
// thirty-two (the 64-byte records of different pairs are pnum * 64 bytes apart): ncu showed the L1 data pipe
// of the light classes at 70 % with a third of its wavefronts coming from these loads.
struct BraSrc {
    const double *S;
    const long long *row;
    long long N;
    unsigned pair;
    const double *W;       // weights of the pairs with an S2 member: [MAX_WGT fields][N], same rows as S (or nullptr)
};
template <int NW>
__device__ __forceinline__ void ld_wgt_soa(const BraSrc &src, int k, double (&w)[NW])
{
    if constexpr (NW > 1) {
        const double *q = src.W + (__ldg(src.row + k) + (long long)src.pair);
#pragma unroll
        for (int x = 0; x < NW; ++x) w[x] = __ldg(q + x * src.N);
    } else {
        w[0] = 1.0;
    }
}
// ket side: weights of primitive pair i at W[i * MAX_WGT ..] (warp-uniform address: broadcast loads)
template <int NW>
__device__ __forceinline__ void ld_wgt(const double *__restrict__ W, long long i, double (&w)[NW])
{
    if constexpr (NW == 4) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(W + i * MAX_WGT)), b = __ldg(reinterpret_cast<const double2 *>(W + i * MAX_WGT) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
    } else if constexpr (NW == 2) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(W + i * MAX_WGT));
        w[0] = a.x; w[1] = a.y;
    } else {
        w[0] = 1.0;
    }
}
__device__ __forceinline__ PrimPair ld_prim_soa(const BraSrc &src, int k)
{
    const double *q = src.S + (__ldg(src.row + k) + (long long)src.pair);
    PrimPair r;
    r.Px = __ldg(q); r.Py = __ldg(q + src.N); r.Pz = __ldg(q + 2 * src.N); r.p = __ldg(q + 3 * src.N);
    r.cc = __ldg(q + 4 * src.N); r.PAx = __ldg(q + 5 * src.N); r.PAy = __ldg(q + 6 * src.N); r.PAz = __ldg(q + 7 * src.N);
    return r;
}
__device__ __forceinline__ PairHdr ld_hdr(const PairHdr *p)
{
    const int4 *qi = reinterpret_cast<const int4 *>(p);
    const int4 a = __ldg(qi), b = __ldg(qi + 1);
    const double2 *qd = reinterpret_cast<const double2 *>(p) + 2;
    const double2 c = __ldg(qd), d = __ldg(qd + 1);
    PairHdr r;
    r.bfA = a.x; r.bfB = a.y; r.poff = a.z; r.pnum = a.w; r.shA = b.x; r.shB = b.y; r.pad0 = b.z; r.pad1 = b.w;
    r.ABx = c.x; r.ABy = c.y; r.ABz = d.x; r.Qs = d.y;
    return r;
}

// Boys table: rows T0 = i/8, i = 0..480; columns k = 0..8: F_{L+k}(T0)/k!, column 9: exp(-T0).
// Beyond T = 60 exp(-T)/(2T) is below 2e-17 of F_m(T) for every m <= 8, so the large-T branch is the pure
// asymptotic series (no exp(), no table) — and it is the COMMON branch in extended systems, where most
// surviving quartets are long-range (T = alpha R^2 >> 60).
constexpr int BOYS_ROWS = 481;
constexpr int BOYS_STRIDE = 10;
constexpr double BOYS_TMAX = 60.0;
// Class-specialised kernels switch to the asymptotic series as early as their highest order allows: the
// smallest integer T beyond which sqrt(pi/T)/2 * (2m-1)!!/(2T)^m is within 2e-17 (relative) of F_m(T) for
// every m <= L (checked against 60-digit hyp1f1).  Fewer lanes take the table branch (in extended systems
// it runs with a third of the warp active) and only the first boys_rows(L) table rows are staged.
__host__ __device__ constexpr int boys_tmax_i(int L)
{
    constexpr int t[9] = {37, 41, 44, 47, 50, 53, 55, 58, 60};
    return t[L < 0 ? 0 : (L > 8 ? 8 : L)];
}
__host__ __device__ constexpr int boys_rows(int L) { return boys_tmax_i(L) * 8 + 1; }

// 1/sqrt(x) to full double precision without the slow-path branches of the library routine: MUFU.RSQ64H seed
// (relative error <= 2^-20) + ONE third-order step  y (1 + e/2 + 3e^2/8),  e = 1 - x y^2  (error 5e^3/16 < 2^-58):
// five FP64 instructions with a dependent depth of four, instead of seven / six for two Newton steps.
// x is a sum of exponents, a squared distance or a Boys argument here (normal range, never 0 on the paths that use it).
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = y * e;
    return fma(q, p, y);
}
// sqrt(pi)/2 is folded into the pair coefficients (each PrimPair::cc carries its square root, lib.cu) and its
// inverse into the Boys tables, so neither Boys branch of the class kernels multiplies by it.
constexpr double SQRTPI_2 = 0.88622692545275801365;
constexpr int BOYS_MAXL = 8;

// ------------------------------------------------------------------------------------------
// Boys function  (replaces scipy hyp1f1 of cython/util.pxi:54-55)
// ------------------------------------------------------------------------------------------
// table branch: valid for T < boys_tmax_i(L) (+ rounding)
template <int L>
__device__ __forceinline__ void boys_table(double T, const double *__restrict__ tab, double (&F)[L + 1])
{
    const int row = __double2int_rn(T * 8.0);
    const double *r = tab + row * BOYS_STRIDE;
    const double d = (double)row * 0.125 - T;   // T0 - T, |d| <= 1/16
    // degree-8 polynomials by Estrin's scheme (dependent depth 4 instead of 8): the table branch runs with a
    // minority of the lanes and its latency, not its instruction count, is what the rest of the warp waits for
    const double d2 = d * d, d4 = d2 * d2;
    {
        const double a01 = fma(r[1], d, r[0]), a23 = fma(r[3], d, r[2]), a45 = fma(r[5], d, r[4]), a67 = fma(r[7], d, r[6]);
        const double b0 = fma(a23, d2, a01), b1 = fma(a67, d2, a45);
        F[L] = fma(fma(r[8], d4, b1), d4, b0);
    }
    if constexpr (L > 0) {
        // exp(-T) = exp(-T0) * exp(d)
        const double a01 = 1.0 + d, a23 = fma(1.66666666666666666667e-01, d, 0.5);
        const double a45 = fma(8.33333333333333333333e-03, d, 4.16666666666666666667e-02);
        const double a67 = fma(1.98412698412698412698e-04, d, 1.38888888888888888889e-03);
        const double b0 = fma(a23, d2, a01), b1 = fma(a67, d2, a45);
        const double e = fma(fma(2.48015873015873015873e-05, d4, b1), d4, b0) * r[9];
        const double t2 = T + T;
        sfor<0, L>([&](auto I) {
            constexpr int m = L - decltype(I)::value;     // m = L .. 1
            F[m - 1] = fma(t2, F[m], e) * (1.0 / (2 * m - 1));
        });
    }
}

// runtime-L version for the generic kernel
__device__ __forceinline__ void boys_eval_rt(int L, double T, const double *__restrict__ tab, double *F)
{
    if (T < BOYS_TMAX) {
        const int row = __double2int_rn(T * 8.0);
        const double *r = tab + row * BOYS_STRIDE;
        const double d = (double)row * 0.125 - T;
        double f = r[8];
#pragma unroll
        for (int k = 7; k >= 0; --k) f = fma(f, d, r[k]);
        F[L] = f;
        if (L > 0) {
            double e = 2.48015873015873015873e-05;
            e = fma(e, d, 1.98412698412698412698e-04);
            e = fma(e, d, 1.38888888888888888889e-03);
            e = fma(e, d, 8.33333333333333333333e-03);
            e = fma(e, d, 4.16666666666666666667e-02);
            e = fma(e, d, 1.66666666666666666667e-01);
            e = fma(e, d, 0.5);
            e = fma(e, d, 1.0);
            e = fma(e, d, 1.0);
            e *= r[9];
            const double t2 = T + T;
            for (int m = L; m > 0; --m) F[m - 1] = fma(t2, F[m], e) / (double)(2 * m - 1);
        }
        for (int m = 0; m <= L; ++m) F[m] *= SQRTPI_2;      // the tables carry 2/sqrt(pi)
    } else {
        const double rt = fast_rsqrt(T);
        F[0] = 0.88622692545275801365 * rt;
        const double hit = 0.5 * (rt * rt);
        for (int m = 0; m < L; ++m) F[m + 1] = ((2 * m + 1) * hit) * F[m];
    }
}

// ------------------------------------------------------------------------------------------
// Hermite expansion coefficients, normalised by E_0^{00}  (cython/util.pxi:13-26, n = 0 branch)
//   E^{i,0}_t = 1/(2p) E^{i-1,0}_{t-1} + (P-A) E^{i-1,0}_t + (t+1) E^{i-1,0}_{t+1}
//   E^{i,j}_t = 1/(2p) E^{i,j-1}_{t-1} + (P-B) E^{i,j-1}_t + (t+1) E^{i,j-1}_{t+1}
// ------------------------------------------------------------------------------------------
template <int LA, int LB>
struct ETab {
    double v[3][LA + 1][LB + 1][LA + LB + 1];
};

template <int LA, int LB>
__device__ __forceinline__ void build_E(ETab<LA, LB> &E, const double (&PA)[3], const double (&PB)[3], double oo2p)
{
    sfor<0, 3>([&](auto DIM) {
        constexpr int dim = decltype(DIM)::value;
        E.v[dim][0][0][0] = 1.0;
        sfor<1, LA + 1>([&](auto I) {
            constexpr int i = decltype(I)::value;
            sfor<0, i + 1>([&](auto TT) {
                constexpr int t = decltype(TT)::value;
                double x = 0.0;
                if constexpr (t > 0) x = oo2p * E.v[dim][i - 1][0][t - 1];
                if constexpr (t <= i - 1) x = fma(PA[dim], E.v[dim][i - 1][0][t], x);
                if constexpr (t + 1 <= i - 1) x = fma((double)(t + 1), E.v[dim][i - 1][0][t + 1], x);
                E.v[dim][i][0][t] = x;
            });
        });
        sfor<1, LB + 1>([&](auto J) {
            constexpr int j = decltype(J)::value;
            sfor<0, LA + 1>([&](auto I) {
                constexpr int i = decltype(I)::value;
                sfor<0, i + j + 1>([&](auto TT) {
                    constexpr int t = decltype(TT)::value;
                    double x = 0.0;
                    if constexpr (t > 0) x = oo2p * E.v[dim][i][j - 1][t - 1];
                    if constexpr (t <= i + j - 1) x = fma(PB[dim], E.v[dim][i][j - 1][t], x);
                    if constexpr (t + 1 <= i + j - 1) x = fma((double)(t + 1), E.v[dim][i][j - 1][t + 1], x);
                    E.v[dim][i][j][t] = x;
                });
            });
        });
    });
}

// ------------------------------------------------------------------------------------------
// Hermite Coulomb integrals R^0_{tuv}, t+u+v <= L  (cython/util.pxi:33-50), in place over n.
// Fs[n] = prefactor * (-2 alpha)^n F_n(T) on entry.  Same branch order as the reference:
// lower t if t > 0, else u if u > 0, else v.
// ------------------------------------------------------------------------------------------
template <int L, bool SMEM>
struct RStore;
template <int L>
struct RStore<L, false> {          // registers
    double v[nherm(L)];
    __device__ __forceinline__ double &operator[](int i) { return v[i]; }
    __device__ __forceinline__ const double &operator[](int i) const { return v[i]; }
};
template <int L>
struct RStore<L, true> {           // shared memory, element i of thread tid at base[i * stride] (conflict-free)
    double *base;
    int stride;
    __device__ __forceinline__ double &operator[](int i) { return base[i * stride]; }
    __device__ __forceinline__ const double &operator[](int i) const { return base[i * stride]; }
};

template <int L, class RS>
__device__ __forceinline__ void build_R_impl(RS &R, const double (&Fs)[L + 1], double X, double Y, double Z)
{
    R[0] = Fs[L];
    sfor<0, L>([&](auto NN) {
        constexpr int n = L - 1 - decltype(NN)::value;   // n = L-1 .. 0
        sfor<0, L - n>([&](auto DD) {
            constexpr int d = (L - n) - decltype(DD)::value;   // degree d = L-n .. 1
            sfor<0, d + 1>([&](auto TT) {
                constexpr int t = d - decltype(TT)::value;
                sfor<0, d - t + 1>([&](auto UU) {
                    constexpr int u = (d - t) - decltype(UU)::value;
                    constexpr int v = d - t - u;
                    double x;
                    if constexpr (t > 0) {
                        x = X * R[hidx(t - 1, u, v)];
                        if constexpr (t > 1) x = fma((double)(t - 1), R[hidx(t - 2, u, v)], x);
                    } else if constexpr (u > 0) {
                        x = Y * R[hidx(t, u - 1, v)];
                        if constexpr (u > 1) x = fma((double)(u - 1), R[hidx(t, u - 2, v)], x);
                    } else {
                        x = Z * R[hidx(t, u, v - 1)];
                        if constexpr (v > 1) x = fma((double)(v - 1), R[hidx(t, u, v - 2)], x);
                    }
                    R[hidx(t, u, v)] = x;
                });
            });
        });
        R[0] = Fs[n];
    });
}

// Boys + prefactor scaling + R recursion for one primitive quartet.
// ccb, cck are the pair coefficients DIVIDED by sqrt(p) (PrimPair::cc).  With that normalisation the
// large-T branch — the common one in extended systems — needs neither alpha nor T:
//     s (-2 alpha)^m F_m(T)  ->  ccb cck sqrt(pi)/2 * (-1)^m (2m-1)!! / |PQ|^(2m+1)        (T >= T_max(L))
// i.e. ONE reciprocal square root (of |PQ|^2) and a product chain; the branch is taken on
// p q |PQ|^2 >= T_max (p+q), which needs no division either.  The table branch pays one extra rsqrt
// (sqrt(alpha)) to undo the normalisation.
// Fs[n] = s (-2 alpha)^n F_n(alpha |PQ|^2), the seeds of the R recursion, for one primitive quartet (both Boys
// branches; this is the routine mmdb_boys_class_host probes, per L, across its own T_max(L) switch)
template <int L, bool FAR = false>
__device__ __forceinline__ void prim_Fs(double (&Fs)[L + 1], double pb, double pk, double ccb, double cck, double R2,
                                        const double *__restrict__ boys_tab)
{
    const double c2 = ccb * cck;
    auto asym = [&]() {
        const double ri = fast_rsqrt(R2);
        Fs[0] = c2 * ri;
        if constexpr (L > 0) {
            const double nr2 = -(ri * ri);
            sfor<0, L>([&](auto I) {
                constexpr int m = decltype(I)::value;
                Fs[m + 1] = ((2 * m + 1) * nr2) * Fs[m];
            });
        }
    };
    if constexpr (FAR) {
        // the screening kernel has proved alpha |PQ|^2 >= T_max(L) for EVERY primitive quartet of this list entry
        asym();
    } else {
        const double pp = pb * pk, ps = pb + pk;
        if (pp * R2 >= (double)boys_tmax_i(L) * ps) {
            asym();
        } else {
            const double rs = fast_rsqrt(ps);        // one reciprocal square root serves 1/(p+q) and 1/sqrt(p+q)
            const double alpha = pp * (rs * rs);
            boys_table<L>(alpha * R2, boys_tab, Fs);     // (2/sqrt(pi)) F_m: undoes the sqrt(pi)/2 inside c2
            double s = c2 * (alpha * fast_rsqrt(alpha));     // ccb cck sqrt(p q) / sqrt(p+q)
            const double m2a = -2.0 * alpha;
#pragma unroll
            for (int n = 0; n <= L; ++n) { Fs[n] *= s; s *= m2a; }
        }
    }
}

template <int L, bool FAR = false, class RS>
__device__ __forceinline__ void prim_R(RS &R, double pb, double pk, double ccb, double cck, double X, double Y, double Z,
                                       const double *__restrict__ boys_tab)
{
    const double R2 = X * X + Y * Y + Z * Z;
    double Fs[L + 1];
    prim_Fs<L, FAR>(Fs, pb, pk, ccb, cck, R2, boys_tab);
    build_R_impl<L>(R, Fs, X, Y, Z);
}

// asymptotic branch only (caller has checked p q |PQ|^2 >= T_max (p+q)); c2 = ccb * cck
template <int L, class RS>
__device__ __forceinline__ void prim_R_asym(RS &R, double c2, double R2, double X, double Y, double Z)
{
    double Fs[L + 1];
    const double ri = fast_rsqrt(R2);
    Fs[0] = c2 * ri;
    if constexpr (L > 0) {
        const double nr2 = -(ri * ri);
        sfor<0, L>([&](auto I) {
            constexpr int m = decltype(I)::value;
            Fs[m + 1] = ((2 * m + 1) * nr2) * Fs[m];
        });
    }
    build_R_impl<L>(R, Fs, X, Y, Z);
}

// out-of-line copy for the shared-memory variant: the (long) recursion is emitted once per kernel
// instead of once per ket-component chunk
template <int L, bool FAR = false>
__device__ __noinline__ void prim_R_smem(double *base, int stride, double pb, double pk, double ccb, double cck, double X,
                                         double Y, double Z, const double *__restrict__ boys_tab)
{
    RStore<L, true> R{base, stride};
    prim_R<L, FAR>(R, pb, pk, ccb, cck, X, Y, Z, boys_tab);
}

// one ket primitive pair with the weights of its component pairs (pairs with an S2 member; unused otherwise)
template <int NW>
struct KetPrimT {
    PrimPair p;
    double w[NW];
};

// ---- compile-time layout of the signed ket coefficient products of one chunk (R-major ket transform) -------------------
template <int LC, int LD>
__host__ __device__ constexpr int ket_box(int cd, int dim)
{
    return comp_pow(LC, cd / ncomp(LD), dim) + comp_pow(LD, cd % ncomp(LD), dim);
}
template <int LC, int LD, int CD0>
__host__ __device__ constexpr int ket_coef_off(int cdi, int tau, int nu, int phi)
{
    int off = 0;
    for (int x = 0; x < cdi; ++x)
        off += (ket_box<LC, LD>(CD0 + x, 0) + 1) * (ket_box<LC, LD>(CD0 + x, 1) + 1) * (ket_box<LC, LD>(CD0 + x, 2) + 1);
    const int ny = ket_box<LC, LD>(CD0 + cdi, 1) + 1, nz = ket_box<LC, LD>(CD0 + cdi, 2) + 1;
    return off + (tau * ny + nu) * nz + phi;
}
template <int LC, int LD, int CD0, int NCDC>
__host__ __device__ constexpr int ket_ncoef()
{
    int off = 0;
    for (int x = 0; x < NCDC; ++x)
        off += (ket_box<LC, LD>(CD0 + x, 0) + 1) * (ket_box<LC, LD>(CD0 + x, 1) + 1) * (ket_box<LC, LD>(CD0 + x, 2) + 1);
    return off;
}
// does R[KT,KU,KV] feed any G element of this chunk?
template <int LC, int LD, int CD0, int NCDC, int LBRA>
__host__ __device__ constexpr bool ket_uses_R(int KT, int KU, int KV)
{
    for (int x = 0; x < NCDC; ++x) {
        const int ex = ket_box<LC, LD>(CD0 + x, 0), ey = ket_box<LC, LD>(CD0 + x, 1), ez = ket_box<LC, LD>(CD0 + x, 2);
        for (int tau = 0; tau <= (KT < ex ? KT : ex); ++tau)
            for (int nu = 0; nu <= (KU < ey ? KU : ey); ++nu)
                for (int phi = 0; phi <= (KV < ez ? KV : ez); ++phi)
                    if ((KT - tau) + (KU - nu) + (KV - phi) <= LBRA) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------
// One contracted shell quartet, ket component pairs [CD0, CD0+NCDC), class-specialised.
// out[ab*NCDC + cdi] accumulates (ab|cd) WITHOUT the per-component normalisation.
// ------------------------------------------------------------------------------------------
// SCR_OUT: the contracted block is accumulated in a global scratch column of the thread (element x at outg[x * ostride],
// L2-resident, touched once per bra primitive pair) instead of registers — classes whose Hermite intermediate G already
// fills the register file (dp bra pairs: 20 x 3 doubles) keep three ket component pairs per thread that way instead of
// one, so Boys + R + the bra E table are built a third as often and the digestion shares its bra-block loads.
template <int LA, int LB, int LC, int LD, int CD0, int NCDC, bool RSMEM, bool SERIAL_CHUNKS, bool FAR = false, bool SCR_OUT = false>
__device__ __forceinline__ void eval_quartet_chunk(const PairHdr &bh, const BraSrc &bsrc,
                                                   const PairHdr &kh, const PrimPair *__restrict__ kp, const double *__restrict__ kw,
                                                   const double *__restrict__ boys_tab, double *r_smem, int r_stride,
                                                   int ib0, int ib1, double (&out)[ncomp(LA) * ncomp(LB) * NCDC],
                                                   double *__restrict__ outg = nullptr, long long ostride = 0)
{
    // LA..LD are shell TYPE codes (am_of: 3 = S2 is an s shell with two components)
    constexpr int LBRA = am_of(LA) + am_of(LB), LKET = am_of(LC) + am_of(LD), L = LBRA + LKET;
    constexpr int NA = ncomp(LA), NB = ncomp(LB), ND = ncomp(LD);
    constexpr int NWB = nwgt(LA, LB), NWK = nwgt(LC, LD);
    constexpr int NAB = NA * NB;
    constexpr int NHB = nherm(LBRA);

    if constexpr (!SCR_OUT) {
#pragma unroll
        for (int x = 0; x < NAB * NCDC; ++x) out[x] = 0.0;
    }

    using KetPrim = KetPrimT<NWK>;
    auto ld_kp = [&](int i) -> KetPrim {
        KetPrim r;
        r.p = ld_prim(kp + i);
        ld_wgt<NWK>(kw, (long long)i, r.w);
        return r;
    };
    for (int ib = ib0; ib < ib1; ++ib) {       // [ib0,ib1): this entry's slice of the bra primitive pairs
        const PrimPair b = ld_prim_soa(bsrc, ib);
        double wb[NWB];
        ld_wgt_soa<NWB>(bsrc, ib, wb);
        double G[NHB * NCDC];
#pragma unroll
        for (int x = 0; x < NHB * NCDC; ++x) G[x] = 0.0;

        auto ket_transform = [&](const auto &R, const PrimPair &k, const double *wk) {      // (an array-reference parameter here crashes cudafe++ 12.9)
            ETab<am_of(LC), am_of(LD)> Ek;
            {
                const double QC[3] = {k.PAx, k.PAy, k.PAz};
                const double QD[3] = {k.PAx + kh.ABx, k.PAy + kh.ABy, k.PAz + kh.ABz};
                build_E<am_of(LC), am_of(LD)>(Ek, QC, QD, 0.5 / k.p);
            }
            // ket Hermite -> Cartesian:  G[tuv][cd] += (-1)^(tau+nu+phi) E^cd_tau E^cd_nu E^cd_phi R[t+tau,u+nu,v+phi]
            if constexpr (RSMEM) {
                // R lives in shared memory: R-major ("scatter") order.  The signed coefficient products of the chunk are
                // formed first (registers), then every R element is loaded ONCE and added into all the G elements it
                // feeds — the cd-major order made the compiler keep dozens of R values live across the chunk (or load
                // them again), which is where the 255-register kernels and their spills came from.
                constexpr int NCO = ket_ncoef<LC, LD, CD0, NCDC>();
                double co[NCO];
                sfor<0, NCDC>([&](auto CDI) {
                    constexpr int cdi = decltype(CDI)::value;
                    constexpr int cd = CD0 + cdi;
                    constexpr int c = cd / ND, d = cd % ND;
                    constexpr int ex = comp_pow(LC, c, 0) + comp_pow(LD, d, 0), ey = comp_pow(LC, c, 1) + comp_pow(LD, d, 1),
                                  ez = comp_pow(LC, c, 2) + comp_pow(LD, d, 2);
                    sfor<0, ex + 1>([&](auto TAU) {
                        constexpr int tau = decltype(TAU)::value;
                        sfor<0, ey + 1>([&](auto NU) {
                            constexpr int nu = decltype(NU)::value;
                            const double exy = Ek.v[0][comp_pow(LC, c, 0)][comp_pow(LD, d, 0)][tau] * Ek.v[1][comp_pow(LC, c, 1)][comp_pow(LD, d, 1)][nu];
                            sfor<0, ez + 1>([&](auto PHI) {
                                constexpr int phi = decltype(PHI)::value;
                                double coef = exy * Ek.v[2][comp_pow(LC, c, 2)][comp_pow(LD, d, 2)][phi];
                                if constexpr (NWK > 1) coef *= wk[widx(LC, LD, c, d)];
                                if constexpr ((tau + nu + phi) & 1) coef = -coef;
                                co[ket_coef_off<LC, LD, CD0>(cdi, tau, nu, phi)] = coef;
                            });
                        });
                    });
                });
                sfor<0, L + 1>([&](auto KT_) {
                    constexpr int KT = decltype(KT_)::value;
                    sfor<0, L - KT + 1>([&](auto KU_) {
                        constexpr int KU = decltype(KU_)::value;
                        sfor<0, L - KT - KU + 1>([&](auto KV_) {
                            constexpr int KV = decltype(KV_)::value;
                            if constexpr (ket_uses_R<LC, LD, CD0, NCDC, LBRA>(KT, KU, KV)) {
                                const double r = R[hidx(KT, KU, KV)];
                                sfor<0, NCDC>([&](auto CDI) {
                                    constexpr int cdi = decltype(CDI)::value;
                                    constexpr int cd = CD0 + cdi;
                                    constexpr int c = cd / ND, d = cd % ND;
                                    constexpr int ex = comp_pow(LC, c, 0) + comp_pow(LD, d, 0), ey = comp_pow(LC, c, 1) + comp_pow(LD, d, 1),
                                                  ez = comp_pow(LC, c, 2) + comp_pow(LD, d, 2);
                                    sfor<0, (KT < ex ? KT : ex) + 1>([&](auto TAU) {
                                        constexpr int tau = decltype(TAU)::value;
                                        sfor<0, (KU < ey ? KU : ey) + 1>([&](auto NU) {
                                            constexpr int nu = decltype(NU)::value;
                                            sfor<0, (KV < ez ? KV : ez) + 1>([&](auto PHI) {
                                                constexpr int phi = decltype(PHI)::value;
                                                constexpr int t = KT - tau, u = KU - nu, v = KV - phi;
                                                if constexpr (t + u + v <= LBRA)
                                                    G[hidx(t, u, v) * NCDC + cdi] =
                                                        fma(co[ket_coef_off<LC, LD, CD0>(cdi, tau, nu, phi)], r, G[hidx(t, u, v) * NCDC + cdi]);
                                            });
                                        });
                                    });
                                });
                            }
                        });
                    });
                });
            } else {
            sfor<0, NCDC>([&](auto CDI) {
                constexpr int cdi = decltype(CDI)::value;
                constexpr int cd = CD0 + cdi;
                constexpr int c = cd / ND, d = cd % ND;
                constexpr int cx = comp_pow(LC, c, 0), cy = comp_pow(LC, c, 1), cz = comp_pow(LC, c, 2);
                constexpr int dx = comp_pow(LD, d, 0), dy = comp_pow(LD, d, 1), dz = comp_pow(LD, d, 2);
                sfor<0, cx + dx + 1>([&](auto TAU) {
                    constexpr int tau = decltype(TAU)::value;
                    sfor<0, cy + dy + 1>([&](auto NU) {
                        constexpr int nu = decltype(NU)::value;
                        const double exy = Ek.v[0][cx][dx][tau] * Ek.v[1][cy][dy][nu];
                        sfor<0, cz + dz + 1>([&](auto PHI) {
                            constexpr int phi = decltype(PHI)::value;
                            double coef = exy * Ek.v[2][cz][dz][phi];
                            if constexpr (NWK > 1) coef *= wk[widx(LC, LD, c, d)];
                            if constexpr ((tau + nu + phi) & 1) coef = -coef;
                            sfor<0, LBRA + 1>([&](auto TT) {
                                constexpr int t = decltype(TT)::value;
                                sfor<0, LBRA - t + 1>([&](auto UU) {
                                    constexpr int u = decltype(UU)::value;
                                    sfor<0, LBRA - t - u + 1>([&](auto VV) {
                                        constexpr int v = decltype(VV)::value;
                                        G[hidx(t, u, v) * NCDC + cdi] =
                                            fma(coef, R[hidx(t + tau, u + nu, v + phi)], G[hidx(t, u, v) * NCDC + cdi]);
                                    });
                                });
                            });
                        });
                    });
                });
            });
            }
        };
        auto ket_body = [&](const KetPrim &kq) {
            const PrimPair &k = kq.p;
            const double X = b.Px - k.Px, Y = b.Py - k.Py, Z = b.Pz - k.Pz;
            RStore<L, RSMEM> R;
            if constexpr (RSMEM) {
                R.base = r_smem;
                R.stride = r_stride;
                // a quartet with ONE primitive quartet keeps its R table in shared memory across the
                // ket-component chunks: only the first chunk builds it
                if (!SERIAL_CHUNKS || CD0 == 0 || (ib1 - ib0) * kh.pnum != 1)
                    prim_R_smem<L, FAR>(r_smem, r_stride, b.p, k.p, b.cc, k.cc, X, Y, Z, boys_tab);
            } else {
                prim_R<L, FAR>(R, b.p, k.p, b.cc, k.cc, X, Y, Z, boys_tab);
            }
            ket_transform(R, k, kq.w);
        };
        // Two ket primitives at once (L <= 1 only).  These kernels are latency-bound: ~5 warps per scheduler, and
        // one primitive quartet is essentially ONE dependent chain (|PQ|^2 -> rsqrt -> F_m -> R -> G), so the
        // FP64 pipe idles between dependent instructions.  When both primitive quartets are on the branch-free
        // asymptotic path for every converged lane, the two chains are emitted in one basic block and the
        // compiler interleaves them; otherwise the two bodies run one after the other as before.
        auto ket_body2 = [&](const KetPrim &kqa, const KetPrim &kqc) {
            const PrimPair &ka = kqa.p, &kc = kqc.p;
            const double Xa = b.Px - ka.Px, Ya = b.Py - ka.Py, Za = b.Pz - ka.Pz;
            const double Xc = b.Px - kc.Px, Yc = b.Py - kc.Py, Zc = b.Pz - kc.Pz;
            const double R2a = Xa * Xa + Ya * Ya + Za * Za, R2c = Xc * Xc + Yc * Yc + Zc * Zc;
            bool both;
            if constexpr (FAR) {
                both = true;         // every primitive quartet of a far-list entry is on the asymptotic branch
            } else {
                const double tm = (double)boys_tmax_i(L);
                const bool asym = ((b.p * ka.p) * R2a >= tm * (b.p + ka.p)) && ((b.p * kc.p) * R2c >= tm * (b.p + kc.p));
                both = __all_sync(__activemask(), asym);
            }
            if (both) {
                RStore<L, false> Ra, Rc;
                prim_R_asym<L>(Ra, b.cc * ka.cc, R2a, Xa, Ya, Za);
                prim_R_asym<L>(Rc, b.cc * kc.cc, R2c, Xc, Yc, Zc);
                ket_transform(Ra, ka, kqa.w);
                ket_transform(Rc, kc, kqc.w);
            } else {
                ket_body(kqa);
                ket_body(kqc);
            }
        };
        // The next ket primitive pair is loaded while the current one is consumed.  The light classes
        // alternate between two register sets (loop unrolled by two) so the hand-over costs no moves —
        // the 16 register copies were a fifth of their inner loop; the heavy classes keep one copy of
        // the (large) loop body.
        if constexpr (L <= 2) {
            KetPrim k0 = ld_kp(kh.poff), k1 = k0;
            for (int ik = 0; ik < kh.pnum; ik += 2) {
                const bool two = ik + 1 < kh.pnum;
                if (two) k1 = ld_kp(kh.poff + ik + 1);
                if constexpr (L <= 1 && !RSMEM) {
                    if (two) {
                        const KetPrim ka = k0;
                        if (ik + 2 < kh.pnum) k0 = ld_kp(kh.poff + ik + 2);
                        ket_body2(ka, k1);
                    } else {
                        ket_body(k0);
                    }
                } else {
                    ket_body(k0);
                    if (two) {
                        if (ik + 2 < kh.pnum) k0 = ld_kp(kh.poff + ik + 2);
                        ket_body(k1);
                    }
                }
            }
        } else {
            KetPrim k_next = ld_kp(kh.poff);
            for (int ik = 0; ik < kh.pnum; ++ik) {
                const KetPrim k = k_next;
                if (ik + 1 < kh.pnum) k_next = ld_kp(kh.poff + ik + 1);
                ket_body(k);
            }
        }
        // bra Hermite -> Cartesian:  out[ab][cd] += E^ab_t E^ab_u E^ab_v G[tuv][cd]
        // (the bra E table is built here, after the ket primitives, so it is not live across the ket loop)
        ETab<am_of(LA), am_of(LB)> Eb;
        {
            const double PA[3] = {b.PAx, b.PAy, b.PAz};
            const double PB[3] = {b.PAx + bh.ABx, b.PAy + bh.ABy, b.PAz + bh.ABz};
            build_E<am_of(LA), am_of(LB)>(Eb, PA, PB, 0.5 / b.p);
        }
        sfor<0, NAB>([&](auto ABI) {
            constexpr int ab = decltype(ABI)::value;
            constexpr int a = ab / NB, bb = ab % NB;
            constexpr int ax = comp_pow(LA, a, 0), ay = comp_pow(LA, a, 1), az = comp_pow(LA, a, 2);
            constexpr int bx = comp_pow(LB, bb, 0), by = comp_pow(LB, bb, 1), bz = comp_pow(LB, bb, 2);
            double acc[NCDC];
            if constexpr (SCR_OUT) {
#pragma unroll
                for (int cdi = 0; cdi < NCDC; ++cdi) acc[cdi] = 0.0;
            }
            sfor<0, ax + bx + 1>([&](auto TT) {
                constexpr int t = decltype(TT)::value;
                sfor<0, ay + by + 1>([&](auto UU) {
                    constexpr int u = decltype(UU)::value;
                    const double exy = Eb.v[0][ax][bx][t] * Eb.v[1][ay][by][u];
                    sfor<0, az + bz + 1>([&](auto VV) {
                        constexpr int v = decltype(VV)::value;
                        double coef = exy * Eb.v[2][az][bz][v];
                        if constexpr (NWB > 1) coef *= wb[widx(LA, LB, a, bb)];
#pragma unroll
                        for (int cdi = 0; cdi < NCDC; ++cdi) {
                            if constexpr (SCR_OUT) acc[cdi] = fma(coef, G[hidx(t, u, v) * NCDC + cdi], acc[cdi]);
                            else out[ab * NCDC + cdi] = fma(coef, G[hidx(t, u, v) * NCDC + cdi], out[ab * NCDC + cdi]);
                        }
                    });
                });
            });
            if constexpr (SCR_OUT) {
#pragma unroll
                for (int cdi = 0; cdi < NCDC; ++cdi) {
                    // first bra primitive: plain store; later ones: fire-and-forget reductions (a read-modify-write would
                    // chain NAB * NCDC dependent L2 round trips per bra primitive).  The reader uses ld.global.cg.
                    double *o = outg + (long long)(ab * NCDC + cdi) * ostride;
                    if (ib == ib0) __stcg(o, acc[cdi]);
                    else red_add_f64(o, acc[cdi]);
                }
            }
        });
    }
}

// ------------------------------------------------------------------------------------------
// Digestion of one contracted basis-function integral (ij|kl) into G  (cython/fock.pyx:38-85).
// ------------------------------------------------------------------------------------------
struct DigestArgs {
    int N;
    double tol;
    const double *SQ;      // sqrt((pq|pq)) (N,N)
    const double *Dabs;    // |dP| (N,N)
    const double *dPre;    // Re dP (N,N)
    const double *dPim;    // Im dP or nullptr
    double *Gre;
    double *Gim;           // or nullptr
    int fixed;             // deterministic fixed-point accumulation (see red_add_g)
};

template <bool FIXED>
__device__ __forceinline__ void digest_fn_quartet(const DigestArgs &g, int i, int j, int k, int l, bool sameAB,
                                                  bool sameCD, bool samePair, double val)
{
    // each canonical function quartet i>=j, k>=l, ij>=kl exactly once (cython/fock.pyx:38-44); written
    // branch-free up to the final "live" test so that all fourteen loads are in flight together
    bool live = !(sameAB && i < j) && !(sameCD && k < l);
    const int ii = max(i, j), jj = min(i, j), kk = max(k, l), ll = min(k, l);
    const long long ij = (long long)ii * (ii + 1) / 2 + jj, kl = (long long)kk * (kk + 1) / 2 + ll;
    const bool sw = ij < kl;
    live = live && !(sw && samePair);
    i = sw ? kk : ii; j = sw ? ll : jj; k = sw ? ii : kk; l = sw ? jj : ll;
    const int N = g.N;
    const double *D = g.Dabs;
    const double *P = g.dPre;
    const double sq1 = __ldg(&g.SQ[i * N + j]), sq2 = __ldg(&g.SQ[k * N + l]);
    const double d_ij = __ldg(&D[i * N + j]), d_kl = __ldg(&D[k * N + l]), d_ik = __ldg(&D[i * N + k]);
    const double d_il = __ldg(&D[i * N + l]), d_jk = __ldg(&D[j * N + k]), d_jl = __ldg(&D[j * N + l]);
    const double p_kl = __ldg(&P[k * N + l]), p_ij = __ldg(&P[i * N + j]), p_jl = __ldg(&P[j * N + l]);
    const double p_ik = __ldg(&P[i * N + k]), p_jk = __ldg(&P[j * N + k]), p_il = __ldg(&P[i * N + l]);
    double bound = sq1 * sq2;                                               // fock.pyx:46-47
    const double dmax = fmax(fmax(4.0 * d_ij, 4.0 * d_kl), fmax(fmax(d_ik, d_il), fmax(d_jk, d_jl)));   // fock.pyx:49-54
    bound *= dmax;
    if (!live || bound < g.tol) return;                                     // fock.pyx:56-57
    double deg = (i == j) ? 1.0 : 2.0;                                      // fock.pyx:60-70
    if (k != l) deg *= 2.0;
    if (!(i == k && j == l)) deg *= 2.0;
    const double e = deg * val;                                            // fock.pyx:74-75
    const double eq = -0.25 * e;
    red_add_g(FIXED ? 1 : 0, &g.Gre[i * N + j], p_kl * e);                  // fock.pyx:79
    red_add_g(FIXED ? 1 : 0, &g.Gre[k * N + l], p_ij * e);                  // fock.pyx:80
    red_add_g(FIXED ? 1 : 0, &g.Gre[i * N + k], p_jl * eq);                 // fock.pyx:82
    red_add_g(FIXED ? 1 : 0, &g.Gre[j * N + l], p_ik * eq);                 // fock.pyx:83
    red_add_g(FIXED ? 1 : 0, &g.Gre[i * N + l], p_jk * eq);                 // fock.pyx:84
    red_add_g(FIXED ? 1 : 0, &g.Gre[k * N + j], p_il * eq);                 // fock.pyx:85
    if (g.dPim != nullptr) {
        const double *Q = g.dPim;
        red_add_g(FIXED ? 1 : 0, &g.Gim[i * N + j], __ldg(&Q[k * N + l]) * e);
        red_add_g(FIXED ? 1 : 0, &g.Gim[k * N + l], __ldg(&Q[i * N + j]) * e);
        red_add_g(FIXED ? 1 : 0, &g.Gim[i * N + k], __ldg(&Q[j * N + l]) * eq);
        red_add_g(FIXED ? 1 : 0, &g.Gim[j * N + l], __ldg(&Q[i * N + k]) * eq);
        red_add_g(FIXED ? 1 : 0, &g.Gim[i * N + l], __ldg(&Q[j * N + k]) * eq);
        red_add_g(FIXED ? 1 : 0, &g.Gim[k * N + j], __ldg(&Q[i * N + l]) * eq);
    }
}

// Arguments of the ERI kernels (both families).
struct EriArgs {
    const PairHdr *braH;
    const PrimPair *braP;         // array-of-structures copy (generic kernel)
    const double *braS;           // structure-of-arrays copy of the bra primitive pairs (class kernels), see BraSrc
    const long long *braRow;
    long long braN;
    const double *braPQ;          // direct builds: per bra pair {dP block, sqrt(Q) block} in component order, packed per build (or nullptr)
    const double *braW;           // weights of bra pairs with an S2 member: [MAX_WGT][braN], rows as braS (or nullptr)
    const PairHdr *ketH;
    const PrimPair *ketP;
    const double *ketW;           // weights of ket pairs with an S2 member: [nprimpairs][MAX_WGT] (or nullptr)
    const uint2 *list;            // (bra pair, ket pair) per entry; entry e lives at list[e * list_step]
    long long list_step;          // +1: front-to-back (fast-path list), -1: back-to-front (slow-path list)
    const unsigned long long *count_dev;   // number of entries (device) or nullptr -> use n
    unsigned long long n;
    const double *boys_tab;       // [BOYS_ROWS][BOYS_STRIDE] for this class's L (global)
    double *out;                  // EPI_STORE: [entry][nfn]
    double *scratch;              // per-thread columns of the contracted block (classes with scratch_out): [x][grid threads]
    int same_class;               // bra class == ket class (entries with .x == .y are diagonal quartets)
    DigestArgs dg;                // EPI_DIGEST
};

// EPI_DIGEST: every entry satisfies the block-digestion preconditions (the screening kernel sorts the
// others into a second list); EPI_DIGEST_SLOW: per-function digestion for that second list
enum { EPI_STORE = 0, EPI_DIGEST = 1, EPI_DIGEST_SLOW = 2 };

// Direct-build list entries carry a slice id in the top byte of the bra pair index: slice s covers the bra
// primitive pairs [s*BRA_SLICE, (s+1)*BRA_SLICE).  Deeply contracted quartets ((s8 s8|s8 s8) = 4096 primitive
// quartets) are thereby spread over up to 8 threads; J/K digestion is linear in the integrals, so every slice
// digests its own partial block.  Bounds the longest serial thread (the tail of every launch).
constexpr int BRA_SLICE = 8;
constexpr unsigned SLICE_SHIFT = 24;
constexpr unsigned PAIR_MASK = (1u << SLICE_SHIFT) - 1u;

}  // namespace mmdb

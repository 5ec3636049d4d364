// kernels_a.cuh — class-specialised (compile-time angular momentum) shell-quartet kernels.
// One thread per contracted shell quartet; R, E, the Hermite intermediate and the contracted
// integrals live in registers.  When the (ab|cd) block is too large for the register file the ket
// component pairs are processed in NCHUNK static chunks (R is rebuilt per chunk).
#pragma once
#include <algorithm>

#include "core.cuh"

namespace mmdb {

constexpr int KA_THREADS = 128;
// compile-time A/B hooks (profiles/README.md): CTA sizes of the (ss|ss) / (ps|ss) kernels; -DMMDB_SERIAL_SCRATCH runs the
// ket chunks of the scratch-column classes serially in the thread
// scratch-column classes that digest their three ket component pairs in ONE call (digest_cdn_rt): measured
// (dp|ds) 2.93 -> 2.59 ms, (dp|dp) 2.56 -> 1.80 ms, but (dp|pp) 4.69 -> 5.25 ms (larger code: that kernel is the one with
// instruction-fetch stalls) — so the d-ket classes only.  -DMMDB_DIGEST_MULTI=0 / =2 : none / all of them.
#ifndef MMDB_DIGEST_MULTI
#define MMDB_DIGEST_MULTI 1
#endif
#ifndef MMDB_T_L0
#define MMDB_T_L0 768
#endif
#ifndef MMDB_T_L1
#define MMDB_T_L1 640
#endif

// LA..LD are shell TYPE codes (core.cuh: 3 = S2, a two-component s pseudo-shell); ltot = total angular momentum
__host__ __device__ constexpr int ltot(int la, int lb, int lc, int ld) { return am_of(la) + am_of(lb) + am_of(lc) + am_of(ld); }
__host__ __device__ constexpr bool has_s2(int la, int lb, int lc, int ld) { return la == SH_S2 || lb == SH_S2 || lc == SH_S2 || ld == SH_S2; }

__host__ __device__ constexpr int etab_size(int la, int lb)
{
    int n = 0;
    for (int i = 0; i <= la; ++i)
        for (int j = 0; j <= lb; ++j) n += i + j + 1;
    return n;
}

// Classes that keep the contracted block in a global scratch column (see eval_quartet_chunk, SCR_OUT): dp bra pairs
// with a ket of at least three component pairs.  G (20 Hermite rows x 3 ket component pairs) fills the registers.
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr bool scratch_out()
{
    // (dp|ps) (three ket component pairs in all) stays with one pair per thread and three CTAs per quartet: measured
    // 6.5 ms against 6.9 ms with the scratch column; (dp|pp) went from 6.3 to 5.0 ms, (dp|ds) from 3.3 to 3.1 ms
    return LA == 2 && LB == 1 && ncomp(LC) * ncomp(LD) >= 6;
}

// R_tuv lives in shared memory (one column per thread) for the high-L classes, in registers otherwise
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr bool r_in_smem()
{
    return ltot(LA, LB, LC, LD) >= 5 || scratch_out<LA, LB, LC, LD>();
}

// ket component pairs per chunk
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr int chunk_ncd()
{
    constexpr int NAB = ncomp(LA) * ncomp(LB), NCD = ncomp(LC) * ncomp(LD), NHB = nherm(am_of(LA) + am_of(LB));
    int best = 1;
    if (scratch_out<LA, LB, LC, LD>()) return 3;
    if (r_in_smem<LA, LB, LC, LD>()) {
        // live doubles: ket transform G + ket E table; bra transform G + out + bra E table
        constexpr int nEb = 3 * etab_size(am_of(LA), am_of(LB)), nEk = 3 * etab_size(am_of(LC), am_of(LD));
        for (int c = 1; c <= NCD; ++c)
            if (NCD % c == 0 && NHB * c + nEk <= 112 && (NAB + NHB) * c + nEb <= 124) best = c;
    } else {
        // (classes with an S2 pseudo-shell and L >= 3: smaller chunks, their digestion blocks are wider)
        constexpr int out_max = (has_s2(LA, LB, LC, LD) && ltot(LA, LB, LC, LD) >= 3) ? 24 : 36;
        for (int c = 1; c <= NCD; ++c)
            if (NCD % c == 0 && NAB * c <= out_max && NHB * c <= 40) best = c;
    }
    return best;
}

// chunk = CTA property (true) or serial loop inside the thread (false)
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr bool block_chunks()
{
#ifdef MMDB_SERIAL_SCRATCH      // experiment: the scratch-column classes run their chunks serially in the thread (R kept for one-primitive quartets)
    if (scratch_out<LA, LB, LC, LD>()) return false;
#endif
    return ncomp(LC) * ncomp(LD) / chunk_ncd<LA, LB, LC, LD>() >= 2;
}

// CTA shape.  The L <= 2 classes run ONE large CTA per SM (as many warps as their register use allows:
// 24 / 20 / 12 / 12 / 16) instead of several 128-thread CTAs: every CTA stages its own 38.5 KB copy of the
// Boys table, and five or six copies per SM left the L1 cache with a few tens of KB (ncu: 37% L1 hit rate
// on the primitive-pair, header and density loads, long-scoreboard stalls on the first use of each load).
// The L >= 3 classes keep 128-thread CTAs (they synchronise per quartet, see LOCKSTEP below).
template <int LA, int LB, int LC, int LD, bool FAR = false>
__host__ __device__ constexpr int ka_threads()
{
    constexpr int L = ltot(LA, LB, LC, LD);
    // far-list kernels stage no Boys table, so nothing is gained by one huge CTA per SM: 256-thread CTAs, as many
    // as the register budget of min_blocks allows
    if (FAR && L <= 2) return 256;
    if (has_s2(LA, LB, LC, LD) && L <= 2) {
        // classes with an S2 pseudo-shell: the blocks are larger than those of the plain class of the same L
        // (thread count = register cap; chosen from the ptxas logs so that none of these spills)
        constexpr int NFN = ncomp(LA) * ncomp(LB) * ncomp(LC) * ncomp(LD);
        if (L == 0) return NFN <= 2 ? 640 : (NFN == 4 ? (LC == SH_S2 ? 512 : 640) : (NFN == 8 ? 384 : 256));
        if (L == 1) return NFN <= 6 ? 512 : (NFN <= 12 ? 384 : 256);
        return 256;
    }
    if (L == 0) return MMDB_T_L0;                // 80 registers
    if (L == 1) return MMDB_T_L1;                // 96 registers (768 threads at 80 registers: slower)
    if (L == 2 && LA == 2) return 512;           // (ds|ss): 128 registers
    if (L == 2) return 384;                      // (ps|ps), (pp|ss): <= 170 registers
    // lock-step classes, measured per class: one 8-warp CTA per SM (ONE instruction stream per SM, fetched once
    // for eight warps) wins for (pp|pp) (ds|pp) (dp|dp) (dd|ps) (dd|pp) (dd|ds); (ds|ds) (dp|ps) (dp|pp) (dp|ds)
    // prefer two 4-warp CTAs (two chunks in flight hide each other's barrier and load stalls)
    if (L >= 4 && L <= 6 && !(LA == 2 && LB == 1 && !(LC == 2 && LD == 1)) && !(LA == 2 && LB == 0 && LC == 2)) return 256;
    return KA_THREADS;
}

// minimum co-resident CTAs per SM the compiler must allow for (register cap = 65536 / (threads * MINB))
template <int LA, int LB, int LC, int LD, bool FAR = false>
__host__ __device__ constexpr int min_blocks()
{
    constexpr int L = ltot(LA, LB, LC, LD);
    if (FAR && L == 0) return 4;                 // <= 64 registers
    if (FAR && L == 1) return 3;                 // <= 80 registers
    if (FAR && L == 2) return (LA == 2) ? 2 : 1; // (ds|ss) <= 128 registers; (ps|ps), (pp|ss) as the register use falls
#ifdef MMDB_MINB
    if (L >= 3) return MMDB_MINB;       // experiment switch
#endif
    return (L <= 2 || ka_threads<LA, LB, LC, LD>() == 256) ? 1 : 2;     // measured: tighter caps on the L >= 3 classes only add spills
}

template <int LA, int LB, int LC, int LD, bool FAR = false>
__host__ __device__ constexpr size_t class_smem_bytes()
{
    size_t b = FAR ? 0 : (size_t)boys_rows(ltot(LA, LB, LC, LD)) * BOYS_STRIDE * sizeof(double);
    if (r_in_smem<LA, LB, LC, LD>()) b += (size_t)nherm(ltot(LA, LB, LC, LD)) * ka_threads<LA, LB, LC, LD, FAR>() * sizeof(double);
    return b;
}

// ------------------------------------------------------------------------------------------
// Shell-level digestion (fast path).  Preconditions, checked by the caller: real density and the
// higher-indexed shells of bra and ket differ — then the bra/ket order ij>=kl of cython/fock.pyx:38-44
// is the same for every function quartet of the block.  With A != B and C != D the degeneracy is 8
// throughout; a pair on ONE shell (A == B, only in the ss/pp/dd pair classes) keeps the components
// a >= b (fock.pyx:39: j <= i) with half the weight on a == b (fock.pyx:60-62) — static 0 / 0.5 / 1
// weights selected by a block-uniform flag.  All density / Schwarz elements the block needs are loaded up front
// (independent loads, one latency exposure), the six updates of fock.pyx:79-85 are accumulated in
// registers per destination element, and each destination element receives ONE atomic.
// Orientation: J blocks are stored [hi,lo]; exchange blocks [canonical-bra fn, canonical-ket fn]
// except (lo of canonical bra, hi of canonical ket) which the reference stores transposed (G[k,j]).
// ------------------------------------------------------------------------------------------
struct BlockAddr {
    long long base;   // element offset of (0,0)  (64-bit on purpose: 32-bit offsets compiled to slower address code)
    int s0, s1;       // strides of the first / second block index
};

struct DigestGeom {
    bool sameAB, sameCD;              // bra / ket pair on one shell
    BlockAddr ab, cd;                 // J blocks, P and G share the address
    BlockAddr pac, pad, pbc, pbd;     // P reads of the exchange blocks
    BlockAddr gac, gad, gbc, gbd;     // G writes of the exchange blocks
};

__device__ __forceinline__ DigestGeom make_geom(int N, int bfA, int bfB, int bfC, int bfD)
{
    DigestGeom g;
    const bool aHi = bfA >= bfB, cHi = bfC >= bfD;     // one shell twice: the first index is the row (a >= b kept)
    g.sameAB = bfA == bfB; g.sameCD = bfC == bfD;
    g.ab.base = aHi ? (long long)bfA * N + bfB : (long long)bfB * N + bfA;
    g.ab.s0 = aHi ? N : 1; g.ab.s1 = aHi ? 1 : N;
    g.cd.base = cHi ? (long long)bfC * N + bfD : (long long)bfD * N + bfC;
    g.cd.s0 = cHi ? N : 1; g.cd.s1 = cHi ? 1 : N;
    const int hiB = aHi ? bfA : bfB, hiK = cHi ? bfC : bfD;
    const bool swapped = hiB < hiK;     // the template's bra pair is the canonical KET pair
    auto cross = [&](int bfs, bool sHi, int bft, bool tHi, BlockAddr &p, BlockAddr &gg) {
        // s from the template bra pair, t from the template ket pair
        bool rowIsS;
        if (!swapped) {
            p.base = (long long)bfs * N + bft; p.s0 = N; p.s1 = 1;           // P[s,t]
            rowIsS = !(!sHi && tHi);                                         // (lo bra, hi ket) stored as [t,s]
        } else {
            p.base = (long long)bft * N + bfs; p.s0 = 1; p.s1 = N;           // P[t,s]
            rowIsS = (!tHi && sHi);                                          // normal [t,s]; exception (lo of canonical bra = t, hi of canonical ket = s)
        }
        if (rowIsS) { gg.base = (long long)bfs * N + bft; gg.s0 = N; gg.s1 = 1; }
        else        { gg.base = (long long)bft * N + bfs; gg.s0 = 1; gg.s1 = N; }
    };
    cross(bfA, aHi, bfC, cHi, g.pac, g.gac);
    cross(bfA, aHi, bfD, !cHi, g.pad, g.gad);
    cross(bfB, !aHi, bfC, cHi, g.pbc, g.gbc);
    cross(bfB, !aHi, bfD, !cHi, g.pbd, g.gbd);
    return g;
}

template <int LA, int LB, int LC, int LD, int CD0, int NCDC>
__device__ __forceinline__ void digest_block(const DigestArgs &dg, const DigestGeom &g, bool active, bool ket_uniform,
                                             const double (&out)[ncomp(LA) * ncomp(LB) * NCDC], const double *__restrict__ pq)
{
    constexpr int NA = ncomp(LA), NB = ncomp(LB), NC = ncomp(LC), ND = ncomp(LD);
    // which c / d components this chunk touches
    auto need_c = [](int c) constexpr { for (int x = CD0; x < CD0 + NCDC; ++x) if (x / ND == c) return true; return false; };
    auto need_d = [](int d) constexpr { for (int x = CD0; x < CD0 + NCDC; ++x) if (x % ND == d) return true; return false; };
    const double *__restrict__ P = dg.dPre;
    const double *__restrict__ SQ = dg.SQ;
    double Jcd[NCDC];
#pragma unroll
    for (int x = 0; x < NCDC; ++x) Jcd[x] = 0.0;
    if (active) {
        double *__restrict__ G = dg.Gre;
        const double tol = dg.tol;
        // blocks that persist over the (a,b) loops: everything indexed by the ket shells only, or by b
        double Pcd[NCDC], Qcd[NCDC], Pbc[NB * NC], Pbd[NB * ND], Kbc[NB * NC], Kbd[NB * ND];
        sfor<0, NCDC>([&](auto I) {
            constexpr int cdi = decltype(I)::value;
            constexpr int cd = CD0 + cdi;
            const long long o = g.cd.base + (cd / ND) * g.cd.s0 + (cd % ND) * g.cd.s1;
            Pcd[cdi] = __ldg(&P[o]); Qcd[cdi] = __ldg(&SQ[o]);
        });
        sfor<0, NB>([&](auto B_) {
            constexpr int b = decltype(B_)::value;
            sfor<0, NC>([&](auto C_) {
                constexpr int c = decltype(C_)::value;
                if constexpr (need_c(c)) { Pbc[b * NC + c] = __ldg(&P[g.pbc.base + b * g.pbc.s0 + c * g.pbc.s1]); Kbc[b * NC + c] = 0.0; }
            });
            sfor<0, ND>([&](auto D_) {
                constexpr int d = decltype(D_)::value;
                if constexpr (need_d(d)) { Pbd[b * ND + d] = __ldg(&P[g.pbd.base + b * g.pbd.s0 + d * g.pbd.s1]); Kbd[b * ND + d] = 0.0; }
            });
        });
        sfor<0, NA>([&](auto A_) {
            constexpr int a = decltype(A_)::value;
            double Pac[NC], Pad[ND], Kac[NC], Kad[ND];
            sfor<0, NC>([&](auto C_) {
                constexpr int c = decltype(C_)::value;
                if constexpr (need_c(c)) { Pac[c] = __ldg(&P[g.pac.base + a * g.pac.s0 + c * g.pac.s1]); Kac[c] = 0.0; }
            });
            sfor<0, ND>([&](auto D_) {
                constexpr int d = decltype(D_)::value;
                if constexpr (need_d(d)) { Pad[d] = __ldg(&P[g.pad.base + a * g.pad.s0 + d * g.pad.s1]); Kad[d] = 0.0; }
            });
            sfor<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                constexpr int ab = a * NB + b;
                constexpr double sab = cscale(LA, a) * cscale(LB, b);
                const long long oab = g.ab.base + a * g.ab.s0 + b * g.ab.s1;
                // the bra pair's own density and Schwarz blocks come packed per pair (one or two sectors instead of one
                // sector per element: these kernels are L2-bound on exactly such gathers, profiles/r02_ncu_final_summary.txt)
                const double pab = pq ? __ldg(pq + ab) : __ldg(&P[oab]), qab = pq ? __ldg(pq + NA * NB + ab) : __ldg(&SQ[oab]);
                const double pab4 = 4.0 * fabs(pab);
                double wab = 1.0;
                if constexpr (LA == LB) wab = g.sameAB ? (a > b ? 1.0 : (a == b ? 0.5 : 0.0)) : 1.0;
                double jab = 0.0;
                sfor<0, NCDC>([&](auto J) {
                    constexpr int cdi = decltype(J)::value;
                    constexpr int cd = CD0 + cdi;
                    constexpr int c = cd / ND, d = cd % ND;
                    constexpr double s8 = 8.0 * sab * cscale(LC, c) * cscale(LD, d);
                    double w = wab;
                    if constexpr (LC == LD) w *= g.sameCD ? (c > d ? 1.0 : (c == d ? 0.5 : 0.0)) : 1.0;
                    double dmax = fmax(pab4, 4.0 * fabs(Pcd[cdi]));
                    dmax = fmax(dmax, fmax(fmax(fabs(Pac[c]), fabs(Pad[d])), fmax(fabs(Pbc[b * NC + c]), fabs(Pbd[b * ND + d]))));
                    const double bound = (qab * Qcd[cdi]) * dmax;
                    const double e = (bound < tol) ? 0.0 : (s8 * w) * out[ab * NCDC + cdi];
                    const double eq = -0.25 * e;
                    jab = fma(Pcd[cdi], e, jab);
                    Jcd[cdi] = fma(pab, e, Jcd[cdi]);
                    Kac[c] = fma(Pbd[b * ND + d], eq, Kac[c]);
                    Kbd[b * ND + d] = fma(Pac[c], eq, Kbd[b * ND + d]);
                    Kad[d] = fma(Pbc[b * NC + c], eq, Kad[d]);
                    Kbc[b * NC + c] = fma(Pad[d], eq, Kbc[b * NC + c]);
                });
                red_add_f64(&G[oab], jab);
            });
            sfor<0, NC>([&](auto C_) {
                constexpr int c = decltype(C_)::value;
                if constexpr (need_c(c)) red_add_f64(&G[g.gac.base + a * g.gac.s0 + c * g.gac.s1], Kac[c]);
            });
            sfor<0, ND>([&](auto D_) {
                constexpr int d = decltype(D_)::value;
                if constexpr (need_d(d)) red_add_f64(&G[g.gad.base + a * g.gad.s0 + d * g.gad.s1], Kad[d]);
            });
        });
        sfor<0, NB>([&](auto B_) {
            constexpr int b = decltype(B_)::value;
            sfor<0, NC>([&](auto C_) {
                constexpr int c = decltype(C_)::value;
                if constexpr (need_c(c)) red_add_f64(&G[g.gbc.base + b * g.gbc.s0 + c * g.gbc.s1], Kbc[b * NC + c]);
            });
            sfor<0, ND>([&](auto D_) {
                constexpr int d = decltype(D_)::value;
                if constexpr (need_d(d)) red_add_f64(&G[g.gbd.base + b * g.gbd.s0 + d * g.gbd.s1], Kbd[b * ND + d]);
            });
        });
    }
    // J_cd: every lane of the warp works on the same ket pair in the common case -> one atomic per warp
    if (ket_uniform) {
        sfor<0, NCDC>([&](auto J) {
            constexpr int cdi = decltype(J)::value;
            double v = Jcd[cdi];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            Jcd[cdi] = v;
        });
        if ((threadIdx.x & 31) == 0) {
            sfor<0, NCDC>([&](auto J) {
                constexpr int cdi = decltype(J)::value;
                constexpr int cd = CD0 + cdi;
                red_add_f64(&dg.Gre[g.cd.base + (cd / ND) * g.cd.s0 + (cd % ND) * g.cd.s1], Jcd[cdi]);
            });
        }
    } else if (active) {
        sfor<0, NCDC>([&](auto J) {
            constexpr int cdi = decltype(J)::value;
            constexpr int cd = CD0 + cdi;
            red_add_f64(&dg.Gre[g.cd.base + (cd / ND) * g.cd.s0 + (cd % ND) * g.cd.s1], Jcd[cdi]);
        });
    }
}

// Out-of-line variant of digest_block for the classes whose chunks hold ONE ket component pair (every class
// with a d shell in the ket or L >= 4 with a dp/dd bra): the (c,d) components are run-time arguments, so the
// digestion code is emitted once per kernel instead of once per chunk.  With up to 18 chunks per class the
// unrolled digestion was the bulk of the instruction footprint (ncu: stall_no_instruction dominant).
template <int LA, int LB, int LC, int LD>
__device__ __noinline__ void digest_cd_rt(const DigestArgs &dg, int bfA, int bfB, int bfC, int bfD, int c, int d, double scd,
                                          bool active, bool ket_uniform, const double *__restrict__ out, long long ostride,
                                          const double *__restrict__ pq)
{
    constexpr int NA = ncomp(LA), NB = ncomp(LB);
    const DigestGeom g = make_geom(dg.N, bfA, bfB, bfC, bfD);
    const double *__restrict__ P = dg.dPre;
    const double *__restrict__ SQ = dg.SQ;
    double *__restrict__ G = dg.Gre;
    const long long ocd = g.cd.base + c * g.cd.s0 + d * g.cd.s1;
    double jcd = 0.0;
    if (LC == LD && g.sameCD) {      // one shell twice: components c >= d only, half weight on c == d
        if (c < d) active = false;
        if (c == d) scd *= 0.5;
    }
    if (active) {
        const double tol = dg.tol;
        // the block's integrals first: the reductions below are compiler barriers (asm volatile, memory clobber), a load
        // issued between them would wait for its L2 round trip alone
        double ov[NA * NB];
#pragma unroll
        for (int x = 0; x < NA * NB; ++x) ov[x] = (ostride == 1) ? out[x] : __ldcg(out + (long long)x * ostride);
        const double pcd = __ldg(&P[ocd]), qcd = __ldg(&SQ[ocd]);
        const double pcd4 = 4.0 * fabs(pcd);
        double Pbc[NB], Pbd[NB], Kbc[NB], Kbd[NB], Mb[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            Pbc[b] = __ldg(&P[g.pbc.base + b * g.pbc.s0 + c * g.pbc.s1]);
            Pbd[b] = __ldg(&P[g.pbd.base + b * g.pbd.s0 + d * g.pbd.s1]);
            Kbc[b] = 0.0;
            Kbd[b] = 0.0;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) Mb[b] = fmax(fabs(Pbc[b]), fabs(Pbd[b]));      // shared by every a
        sfor<0, NA>([&](auto A_) {
            constexpr int a = decltype(A_)::value;
            const double pac = __ldg(&P[g.pac.base + a * g.pac.s0 + c * g.pac.s1]);
            const double pad = __ldg(&P[g.pad.base + a * g.pad.s0 + d * g.pad.s1]);
            double kac = 0.0, kad = 0.0;
            const double ma = fmax(pcd4, fmax(fabs(pac), fabs(pad)));                // shared by every b
            sfor<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                constexpr double s8 = 8.0 * cscale(LA, a) * cscale(LB, b);
                const long long oab = g.ab.base + a * g.ab.s0 + b * g.ab.s1;
                const double pab = pq ? __ldg(pq + a * NB + b) : __ldg(&P[oab]), qab = pq ? __ldg(pq + NA * NB + a * NB + b) : __ldg(&SQ[oab]);
                double wab = 1.0;
                if constexpr (LA == LB) wab = g.sameAB ? (a > b ? 1.0 : (a == b ? 0.5 : 0.0)) : 1.0;
                const double dmax = fmax(4.0 * fabs(pab), fmax(ma, Mb[b]));
                const double bound = (qab * qcd) * dmax;
                const double e = (bound < tol) ? 0.0 : (s8 * wab * scd) * ov[a * NB + b];
                const double eq = -0.25 * e;
                red_add_f64(&G[oab], pcd * e);
                jcd = fma(pab, e, jcd);
                kac = fma(Pbd[b], eq, kac);
                Kbd[b] = fma(pac, eq, Kbd[b]);
                kad = fma(Pbc[b], eq, kad);
                Kbc[b] = fma(pad, eq, Kbc[b]);
            });
            red_add_f64(&G[g.gac.base + a * g.gac.s0 + c * g.gac.s1], kac);
            red_add_f64(&G[g.gad.base + a * g.gad.s0 + d * g.gad.s1], kad);
        });
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            red_add_f64(&G[g.gbc.base + b * g.gbc.s0 + c * g.gbc.s1], Kbc[b]);
            red_add_f64(&G[g.gbd.base + b * g.gbd.s0 + d * g.gbd.s1], Kbd[b]);
        }
    }
    if (ket_uniform) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) jcd += __shfl_xor_sync(0xffffffffu, jcd, o);
        if ((threadIdx.x & 31) == 0) red_add_f64(&G[ocd], jcd);
    } else if (active) {
        red_add_f64(&G[ocd], jcd);
    }
}

// Several consecutive ket component pairs of ONE thread at once (the scratch-column classes hold three): the bra pair's
// blocks are loaded once, J_ab is summed over the ket components before its reduction (a third of the J_ab reductions
// of three digest_cd_rt calls), the per-a integrals are loaded together.  Element (ab, cdi) at out[(ab * NCX + cdi) * ostride].
template <int LA, int LB, int LC, int LD, int NCX>
__device__ __noinline__ void digest_cdn_rt(const DigestArgs &dg, int bfA, int bfB, int bfC, int bfD, int cd0, bool active,
                                           bool ket_uniform, const double *__restrict__ out, long long ostride,
                                           const double *__restrict__ pq)
{
    constexpr int NA = ncomp(LA), NB = ncomp(LB), ND = ncomp(LD);
    const DigestGeom g = make_geom(dg.N, bfA, bfB, bfC, bfD);
    const double *__restrict__ P = dg.dPre;
    const double *__restrict__ SQ = dg.SQ;
    double *__restrict__ G = dg.Gre;
    int cc[NCX], dd[NCX];
    long long ocd[NCX];
    double scd[NCX], jcd[NCX];
#pragma unroll
    for (int x = 0; x < NCX; ++x) {
        const int cd = cd0 + x;
        cc[x] = cd / ND; dd[x] = cd % ND;
        ocd[x] = g.cd.base + cc[x] * g.cd.s0 + dd[x] * g.cd.s1;
        // per-component normalisation of d shells (xx, yy, zz: 1/sqrt(3)), cython/basis.pxi:102-105
        scd[x] = ((LC == 2 && (cc[x] == 0 || cc[x] == 3 || cc[x] == 5)) ? 0.57735026918962576451 : 1.0) *
                 ((LD == 2 && (dd[x] == 0 || dd[x] == 3 || dd[x] == 5)) ? 0.57735026918962576451 : 1.0);
        if (LC == LD && g.sameCD) {      // one shell twice: components c >= d only, half weight on c == d
            if (cc[x] < dd[x]) scd[x] = 0.0;
            if (cc[x] == dd[x]) scd[x] *= 0.5;
        }
        jcd[x] = 0.0;
    }
    const bool same_c = cc[0] == cc[NCX - 1], same_d = ND == 1;       // (consecutive cd: c is non-decreasing)
    if (active) {
        const double tol = dg.tol;
        double pcd[NCX], qcd[NCX], pcd4[NCX], Pbc[NB][NCX], Pbd[NB][NCX], Kbc[NB][NCX], Kbd[NB][NCX], Mb[NB][NCX];
#pragma unroll
        for (int x = 0; x < NCX; ++x) {
            pcd[x] = __ldg(&P[ocd[x]]); qcd[x] = __ldg(&SQ[ocd[x]]);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                Pbc[b][x] = __ldg(&P[g.pbc.base + b * g.pbc.s0 + cc[x] * g.pbc.s1]);
                Pbd[b][x] = __ldg(&P[g.pbd.base + b * g.pbd.s0 + dd[x] * g.pbd.s1]);
                Kbc[b][x] = 0.0; Kbd[b][x] = 0.0;
            }
        }
#pragma unroll
        for (int x = 0; x < NCX; ++x) {
            pcd4[x] = 4.0 * fabs(pcd[x]);
#pragma unroll
            for (int b = 0; b < NB; ++b) Mb[b][x] = fmax(fabs(Pbc[b][x]), fabs(Pbd[b][x]));
        }
        sfor<0, NA>([&](auto A_) {
            constexpr int a = decltype(A_)::value;
            // this row's integrals and density elements first (independent loads, one latency exposure)
            double ov[NB][NCX], pac[NCX], pad[NCX], kac[NCX], kad[NCX], ma[NCX];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int x = 0; x < NCX; ++x) ov[b][x] = __ldcg(out + (long long)((a * NB + b) * NCX + x) * ostride);
#pragma unroll
            for (int x = 0; x < NCX; ++x) {
                pac[x] = __ldg(&P[g.pac.base + a * g.pac.s0 + cc[x] * g.pac.s1]);
                pad[x] = __ldg(&P[g.pad.base + a * g.pad.s0 + dd[x] * g.pad.s1]);
                kac[x] = 0.0; kad[x] = 0.0;
            }
#pragma unroll
            for (int x = 0; x < NCX; ++x) ma[x] = fmax(pcd4[x], fmax(fabs(pac[x]), fabs(pad[x])));
            sfor<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                constexpr double s8 = 8.0 * cscale(LA, a) * cscale(LB, b);
                const long long oab = g.ab.base + a * g.ab.s0 + b * g.ab.s1;
                const double pab = pq ? __ldg(pq + a * NB + b) : __ldg(&P[oab]), qab = pq ? __ldg(pq + NA * NB + a * NB + b) : __ldg(&SQ[oab]);
                double wab = 1.0;
                if constexpr (LA == LB) wab = g.sameAB ? (a > b ? 1.0 : (a == b ? 0.5 : 0.0)) : 1.0;
                const double pab4 = 4.0 * fabs(pab);
                double jab = 0.0;
#pragma unroll
                for (int x = 0; x < NCX; ++x) {
                    const double dmax = fmax(pab4, fmax(ma[x], Mb[b][x]));
                    const double bound = (qab * qcd[x]) * dmax;
                    const double e = (bound < tol) ? 0.0 : (s8 * wab * scd[x]) * ov[b][x];
                    const double eq = -0.25 * e;
                    jab = fma(pcd[x], e, jab);
                    jcd[x] = fma(pab, e, jcd[x]);
                    kac[x] = fma(Pbd[b][x], eq, kac[x]);
                    Kbd[b][x] = fma(pac[x], eq, Kbd[b][x]);
                    kad[x] = fma(Pbc[b][x], eq, kad[x]);
                    Kbc[b][x] = fma(pad[x], eq, Kbc[b][x]);
                }
                red_add_f64(&G[oab], jab);
            });
            // consecutive ket component pairs share c (or d): their exchange contributions go to ONE element — one
            // reduction with the sum instead of NCX back-to-back reductions on the same address
            if (same_c) {
                double v = 0.0;
#pragma unroll
                for (int x = 0; x < NCX; ++x) v += kac[x];
                red_add_f64(&G[g.gac.base + a * g.gac.s0 + cc[0] * g.gac.s1], v);
            } else {
#pragma unroll
                for (int x = 0; x < NCX; ++x) red_add_f64(&G[g.gac.base + a * g.gac.s0 + cc[x] * g.gac.s1], kac[x]);
            }
            if (same_d) {
                double v = 0.0;
#pragma unroll
                for (int x = 0; x < NCX; ++x) v += kad[x];
                red_add_f64(&G[g.gad.base + a * g.gad.s0 + dd[0] * g.gad.s1], v);
            } else {
#pragma unroll
                for (int x = 0; x < NCX; ++x) red_add_f64(&G[g.gad.base + a * g.gad.s0 + dd[x] * g.gad.s1], kad[x]);
            }
        });
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (same_c) {
                double v = 0.0;
#pragma unroll
                for (int x = 0; x < NCX; ++x) v += Kbc[b][x];
                red_add_f64(&G[g.gbc.base + b * g.gbc.s0 + cc[0] * g.gbc.s1], v);
            } else {
#pragma unroll
                for (int x = 0; x < NCX; ++x) red_add_f64(&G[g.gbc.base + b * g.gbc.s0 + cc[x] * g.gbc.s1], Kbc[b][x]);
            }
            if (same_d) {
                double v = 0.0;
#pragma unroll
                for (int x = 0; x < NCX; ++x) v += Kbd[b][x];
                red_add_f64(&G[g.gbd.base + b * g.gbd.s0 + dd[0] * g.gbd.s1], v);
            } else {
#pragma unroll
                for (int x = 0; x < NCX; ++x) red_add_f64(&G[g.gbd.base + b * g.gbd.s0 + dd[x] * g.gbd.s1], Kbd[b][x]);
            }
        }
    }
#pragma unroll
    for (int x = 0; x < NCX; ++x) {
        double v = jcd[x];
        if (ket_uniform) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) red_add_f64(&G[ocd[x]], v);
        } else if (active) {
            red_add_f64(&G[ocd[x]], v);
        }
    }
}

template <bool FIXED>
static __device__ __noinline__ void digest_block_slow(const DigestArgs &dg, const PairHdr &bh, const PairHdr &kh, bool samePair,
                                               int la, int lb, int lc, int ld, int cd0, int ncdc, const double *vals)
{
    const int nb = ncomp(lb), nd = ncomp(ld), nab = ncomp(la) * nb;
    const bool sameAB = (bh.shA == bh.shB), sameCD = (kh.shA == kh.shB);
    for (int ab = 0; ab < nab; ++ab) {
        const int aa = ab / nb, bb = ab % nb;
        const double sab = ((la == 2 && (aa == 0 || aa == 3 || aa == 5)) ? 0.57735026918962576451 : 1.0) *
                           ((lb == 2 && (bb == 0 || bb == 3 || bb == 5)) ? 0.57735026918962576451 : 1.0);
        for (int cdi = 0; cdi < ncdc; ++cdi) {
            const int cd = cd0 + cdi, c = cd / nd, d = cd % nd;
            const double s = sab * ((lc == 2 && (c == 0 || c == 3 || c == 5)) ? 0.57735026918962576451 : 1.0) *
                             ((ld == 2 && (d == 0 || d == 3 || d == 5)) ? 0.57735026918962576451 : 1.0);
            digest_fn_quartet<FIXED>(dg, bh.bfA + aa, bh.bfB + bb, kh.bfA + c, kh.bfB + d, sameAB, sameCD, samePair,
                                     vals[ab * ncdc + cdi] * s);
        }
    }
}

template <int LA, int LB, int LC, int LD, int EPI, int CD0, int NCDC, bool SERIAL_CHUNKS, bool FAR>
__device__ __forceinline__ void run_chunk(const EriArgs &a, unsigned long long e, unsigned ibra, bool valid, const PairHdr &bh,
                                          const PairHdr &kh, const double *boys_tab, bool samePair, bool ket_uniform,
                                          int ib0, int ib1)
{
    constexpr int NA = ncomp(LA), NB = ncomp(LB), NC = ncomp(LC), ND = ncomp(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND;
    double out[NAB * NCDC];
    // packed {dP, sqrt(Q)} blocks of the bra pair (direct builds; nullptr otherwise)
    const double *pq = a.braPQ ? a.braPQ + (size_t)ibra * (2 * NAB) : nullptr;
    constexpr bool SCR = scratch_out<LA, LB, LC, LD>();
    // scratch column of this thread: element x at scr[x * sstride] (coalesced across the threads of the grid)
    const long long sstride = (long long)gridDim.x * blockDim.x;
    double *scr = SCR ? a.scratch + ((long long)blockIdx.x * blockDim.x + threadIdx.x) : nullptr;
    if (valid)
        eval_quartet_chunk<LA, LB, LC, LD, CD0, NCDC, r_in_smem<LA, LB, LC, LD>(), SERIAL_CHUNKS, FAR, SCR>(
            bh, BraSrc{a.braS, a.braRow, a.braN, ibra, a.braW}, kh, a.ketP, a.ketW, boys_tab,
            const_cast<double *>(boys_tab) + (FAR ? 0 : boys_rows(ltot(LA, LB, LC, LD)) * BOYS_STRIDE) + threadIdx.x,
            ka_threads<LA, LB, LC, LD, FAR>(), ib0, ib1, out, scr, sstride);
    if constexpr (SCR) {
        if constexpr (EPI == EPI_STORE) {
            if (valid) {
                double *o = a.out + e * (unsigned long long)(NAB * NCD);
                sfor<0, NAB>([&](auto ABI) {
                    constexpr int ab = decltype(ABI)::value;
                    constexpr double sab = cscale(LA, ab / NB) * cscale(LB, ab % NB);
                    sfor<0, NCDC>([&](auto CDI) {
                        constexpr int cdi = decltype(CDI)::value;
                        constexpr int cd = CD0 + cdi;
                        constexpr double sc = sab * cscale(LC, cd / ND) * cscale(LD, cd % ND);
                        o[ab * NCD + cd] = __ldcg(scr + (long long)(ab * NCDC + cdi) * sstride) * sc;
                    });
                });
            }
        } else if constexpr (EPI == EPI_DIGEST && (MMDB_DIGEST_MULTI == 2 || (MMDB_DIGEST_MULTI == 1 && LC == 2))) {
            digest_cdn_rt<LA, LB, LC, LD, NCDC>(a.dg, bh.bfA, bh.bfB, kh.bfA, kh.bfB, CD0, valid, ket_uniform, scr, sstride, pq);
        } else if constexpr (EPI == EPI_DIGEST) {
            sfor<0, NCDC>([&](auto CDI) {
                constexpr int cdi = decltype(CDI)::value;
                constexpr int cd = CD0 + cdi;
                digest_cd_rt<LA, LB, LC, LD>(a.dg, bh.bfA, bh.bfB, kh.bfA, kh.bfB, cd / ND, cd % ND,
                                             cscale(LC, cd / ND) * cscale(LD, cd % ND), valid, ket_uniform,
                                             scr + (long long)cdi * sstride, (long long)NCDC * sstride, pq);
            });
        } else {
            if (valid) {
                double tmp[NAB * NCDC];
#pragma unroll 1
                for (int x = 0; x < NAB * NCDC; ++x) tmp[x] = __ldcg(scr + (long long)x * sstride);
                if (a.dg.fixed) digest_block_slow<true>(a.dg, bh, kh, samePair, LA, LB, LC, LD, CD0, NCDC, tmp);
                else digest_block_slow<false>(a.dg, bh, kh, samePair, LA, LB, LC, LD, CD0, NCDC, tmp);
            }
        }
        return;
    }
    if constexpr (EPI == EPI_STORE) {
        if (valid) {
            double *o = a.out + e * (unsigned long long)(NAB * NCD);
            sfor<0, NAB>([&](auto ABI) {
                constexpr int ab = decltype(ABI)::value;
                constexpr double sab = cscale(LA, ab / NB) * cscale(LB, ab % NB);
                sfor<0, NCDC>([&](auto CDI) {
                    constexpr int cdi = decltype(CDI)::value;
                    constexpr int cd = CD0 + cdi;
                    constexpr double s = sab * cscale(LC, cd / ND) * cscale(LD, cd % ND);
                    o[ab * NCD + cd] = out[ab * NCDC + cdi] * s;
                });
            });
        }
    } else if constexpr (EPI == EPI_DIGEST) {
        // block addresses are rebuilt here (a few integer ops) so they are not live across the ERI evaluation;
        // the warp-level J_cd reduction inside runs for every lane (inactive lanes contribute zeros)
        if constexpr (NCD > 1 && (NCDC == 1 || r_in_smem<LA, LB, LC, LD>())) {
            sfor<0, NCDC>([&](auto CDI) {
                constexpr int cdi = decltype(CDI)::value;
                constexpr int cd = CD0 + cdi;
                double tmp[NAB];
#pragma unroll
                for (int x = 0; x < NAB; ++x) tmp[x] = out[x * NCDC + cdi];
                digest_cd_rt<LA, LB, LC, LD>(a.dg, bh.bfA, bh.bfB, kh.bfA, kh.bfB, cd / ND, cd % ND,
                                             cscale(LC, cd / ND) * cscale(LD, cd % ND), valid, ket_uniform, tmp, 1, pq);
            });
        } else {
            const DigestGeom geom = make_geom(a.dg.N, bh.bfA, bh.bfB, kh.bfA, kh.bfB);
            digest_block<LA, LB, LC, LD, CD0, NCDC>(a.dg, geom, valid, ket_uniform, out, pq);
        }
    } else {
        // second list (diagonal-type quartets, complex densities): per-function digestion, out of line
        if (valid) {
            double tmp[NAB * NCDC];
#pragma unroll
            for (int x = 0; x < NAB * NCDC; ++x) tmp[x] = out[x];
            if (a.dg.fixed) digest_block_slow<true>(a.dg, bh, kh, samePair, LA, LB, LC, LD, CD0, NCDC, tmp);
            else digest_block_slow<false>(a.dg, bh, kh, samePair, LA, LB, LC, LD, CD0, NCDC, tmp);
        }
    }
}

// FAR = true: the far-field list of a direct build — the screening kernel has proved that every primitive quartet of
// every entry is on the asymptotic Boys branch (alpha |PQ|^2 >= T_max(L)), so this variant has no Boys table in shared
// memory, no branch and no table code: fewer registers, more resident warps, branch-uniform warps.
template <int LA, int LB, int LC, int LD, int EPI, bool FAR = false>
__global__ void __launch_bounds__(ka_threads<LA, LB, LC, LD, FAR>(), min_blocks<LA, LB, LC, LD, FAR>()) eri_class_kernel(const EriArgs a)
{
    constexpr int NCD = ncomp(LC) * ncomp(LD);
    constexpr int NCDC = chunk_ncd<LA, LB, LC, LD>();
    constexpr int NCHUNK = NCD / NCDC;
    extern __shared__ double s_boys[];
    const unsigned long long n = a.count_dev ? *a.count_dev : a.n;
    {   // CTAs without work (short lists, persistent grid) leave before staging the 38 KB Boys table
        constexpr bool BC = block_chunks<LA, LB, LC, LD>();
        const unsigned long long first = (BC ? blockIdx.x / NCHUNK : blockIdx.x) * (unsigned long long)blockDim.x;
        if (first >= n) return;
    }
    if constexpr (!FAR) {
        for (int x = threadIdx.x; x < boys_rows(ltot(LA, LB, LC, LD)) * BOYS_STRIDE; x += blockDim.x) s_boys[x] = __ldg(a.boys_tab + x);
        __syncthreads();
    }
    // Block-uniform trip count: every warp stays in the loop (the digestion uses warp shuffles) and, for the
    // classes whose unrolled code exceeds the instruction cache, the warps of a CTA are kept in step with a
    // barrier per quartet so they stream through the code together (one fetch serves all of them).
    constexpr bool LOCKSTEP = (ltot(LA, LB, LC, LD) >= 4);
    // Chunked classes: the ket-component chunk is a property of the CTA (blockIdx.x % NCHUNK), not a serial
    // loop inside the thread.  Each CTA then executes the code of ONE chunk for many quartets, so its
    // instruction working set is 1/NCHUNK of the kernel and stays cache-resident (ncu: stall_no_instruction
    // was the top stall with the serial chunk loop); the price is one R build per chunk.
    constexpr bool BLOCK_CHUNKS = block_chunks<LA, LB, LC, LD>();
    const int my_chunk = BLOCK_CHUNKS ? (int)(blockIdx.x % NCHUNK) : 0;
    const unsigned long long worker = BLOCK_CHUNKS ? blockIdx.x / NCHUNK : blockIdx.x;
    const unsigned long long nworkers = BLOCK_CHUNKS ? gridDim.x / NCHUNK : gridDim.x;
    const unsigned long long wstride = nworkers * blockDim.x;
    const long long lstep = a.list_step;
    auto run_list = [&](auto CHSEL) {
        constexpr int chsel = decltype(CHSEL)::value;     // -1: all chunks serially inside the thread
        // the list entry of the NEXT quartet is read one iteration ahead
        unsigned long long base = worker * blockDim.x;
        uint2 ij_next = make_uint2(0u, 0u);
        if (base < n) ij_next = __ldg(a.list + (long long)min(base + threadIdx.x, n - 1) * lstep);
        for (; base < n; base += wstride) {
            if constexpr (LOCKSTEP) __syncthreads();
            const unsigned long long e = base + threadIdx.x;
            const bool valid = e < n;
            uint2 ij = ij_next;
            // direct builds: the top byte of the bra index is the slice of <= BRA_SLICE primitive pairs this entry covers
            const int slice = (EPI == EPI_STORE) ? 0 : (int)(ij.x >> SLICE_SHIFT);
            if constexpr (EPI != EPI_STORE) ij.x &= PAIR_MASK;
            const PairHdr bh = ld_hdr(a.braH + ij.x);
            const PairHdr kh = ld_hdr(a.ketH + ij.y);
            if (base + wstride < n) ij_next = __ldg(a.list + (long long)min(e + wstride, n - 1) * lstep);
            const bool samePair = a.same_class && (bh.pad0 == (int)ij.y);       // pad0: (parent) pair index of the bra entry
            const int ib0 = (EPI == EPI_STORE) ? 0 : slice * BRA_SLICE;
            const int ib1 = (EPI == EPI_STORE) ? bh.pnum : min(bh.pnum, ib0 + BRA_SLICE);
            bool ket_uniform = false;
            if constexpr (EPI == EPI_DIGEST) {
                const unsigned y0 = __shfl_sync(0xffffffffu, ij.y, 0);
                ket_uniform = __all_sync(0xffffffffu, ij.y == y0) && __all_sync(0xffffffffu, valid);
            }
            if constexpr (chsel >= 0) {
                run_chunk<LA, LB, LC, LD, EPI, chsel * NCDC, NCDC, false, FAR>(a, e, ij.x, valid, bh, kh, s_boys, samePair, ket_uniform, ib0, ib1);
            } else {
                sfor<0, NCHUNK>([&](auto CH) {
                    constexpr int ch = decltype(CH)::value;
                    run_chunk<LA, LB, LC, LD, EPI, ch * NCDC, NCDC, true, FAR>(a, e, ij.x, valid, bh, kh, s_boys, samePair, ket_uniform, ib0, ib1);
                });
            }
        }
    };
    if constexpr (BLOCK_CHUNKS) {
        if (worker >= nworkers) return;      // grid not a multiple of NCHUNK: surplus CTAs have no chunk
        sfor<0, NCHUNK>([&](auto CH) {
            if (my_chunk == decltype(CH)::value) run_list(CH);
        });
    } else {
        run_list(std::integral_constant<int, -1>{});
    }
}

// host launcher, defined (explicitly instantiated) in inst_*.cu
template <int LA, int LB, int LC, int LD>
cudaError_t launch_class(const EriArgs &a, int kind, int grid, cudaStream_t st);

constexpr int FAR_MAXL = 3;

// launch kinds: the three epilogues + the far-field variant of the block digestion
enum { LK_STORE = 0, LK_DIGEST = 1, LK_DIGEST_SLOW = 2, LK_DIGEST_FAR = 3, LK_COUNT = 4 };

template <int LA, int LB, int LC, int LD, int EPI, bool FAR>
static void setup_kernel(int *occ)
{
    constexpr size_t smem = class_smem_bytes<LA, LB, LC, LD, FAR>();
    constexpr int threads = ka_threads<LA, LB, LC, LD, FAR>();
    auto k = eri_class_kernel<LA, LB, LC, LD, EPI, FAR>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // persistent grid: as many CTAs as are co-resident (occupancy x SM count)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, threads, smem);
    if (*occ < 1) *occ = 1;
    // shared-memory carve-out: just what the resident CTAs need, the rest of the 256 KB stays L1
    const int pct = (int)std::min<size_t>(100, (smem * *occ + 1024 * *occ) * 100 / (228 * 1024) + 3);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

template <int LA, int LB, int LC, int LD>
cudaError_t launch_class_impl(const EriArgs &a, int kind, int grid, cudaStream_t st)
{
    static int occ[LK_COUNT] = {0, 0, 0, 0};
    // far-field variants exist for L <= FAR_MAXL only: above that Boys + R are a small part of a primitive quartet and the
    // extra launch per class pair costs more than the table branch it saves (lib.cu sends everything to the near list)
    constexpr bool HAS_FAR = (ltot(LA, LB, LC, LD) <= FAR_MAXL) && !has_s2(LA, LB, LC, LD);
    // classes with an S2 pseudo-shell exist for the direct Fock build only: no integral-storing variant
    constexpr bool HAS_STORE = !has_s2(LA, LB, LC, LD);
    if (occ[LK_DIGEST] == 0) {
        if constexpr (HAS_STORE) setup_kernel<LA, LB, LC, LD, EPI_STORE, false>(&occ[LK_STORE]);
        setup_kernel<LA, LB, LC, LD, EPI_DIGEST, false>(&occ[LK_DIGEST]);
        setup_kernel<LA, LB, LC, LD, EPI_DIGEST_SLOW, false>(&occ[LK_DIGEST_SLOW]);
        if constexpr (HAS_FAR) setup_kernel<LA, LB, LC, LD, EPI_DIGEST, true>(&occ[LK_DIGEST_FAR]);
    }
    constexpr int NCH = block_chunks<LA, LB, LC, LD>() ? ncomp(LC) * ncomp(LD) / chunk_ncd<LA, LB, LC, LD>() : 1;
    auto shape = [&](int o) { return std::max(NCH, (grid * o) / NCH * NCH); };   // multiple of the chunk count
    constexpr int threads = ka_threads<LA, LB, LC, LD, false>();
    constexpr size_t smem = class_smem_bytes<LA, LB, LC, LD, false>();
    if (kind == LK_STORE) {
        if constexpr (HAS_STORE) eri_class_kernel<LA, LB, LC, LD, EPI_STORE, false><<<shape(occ[kind]), threads, smem, st>>>(a);
        else return cudaErrorInvalidValue;
    } else if (kind == LK_DIGEST)
        eri_class_kernel<LA, LB, LC, LD, EPI_DIGEST, false><<<shape(occ[kind]), threads, smem, st>>>(a);
    else if (kind == LK_DIGEST_SLOW)
        eri_class_kernel<LA, LB, LC, LD, EPI_DIGEST_SLOW, false><<<shape(occ[kind]), threads, smem, st>>>(a);
    else if constexpr (HAS_FAR)
        eri_class_kernel<LA, LB, LC, LD, EPI_DIGEST, true><<<shape(occ[kind]), ka_threads<LA, LB, LC, LD, true>(),
                                                             class_smem_bytes<LA, LB, LC, LD, true>(), st>>>(a);
    else
        return cudaErrorInvalidValue;          // no far-field variant for this class
    return cudaGetLastError();
}

#define MMDB_INSTANTIATE_CLASS(LA, LB, LC, LD)                                                            \
    template <>                                                                                          \
    cudaError_t launch_class<LA, LB, LC, LD>(const EriArgs &a, int kind, int grid, cudaStream_t st)      \
    {                                                                                                    \
        return launch_class_impl<LA, LB, LC, LD>(a, kind, grid, st);                                     \
    }

}  // namespace mmdb

// jk_incore.cu — in-core J/K from the dense (N,N,N,N) tensor, ONE pass over TwoE.
//
// Replaces mmd/scf.py:97-98 of the reference:
//     J = einsum('pqrs,sr->pq', TwoE.astype(complex), P);  K = einsum('psqr,sr->pq', TwoE.astype(complex), P)
// which makes two passes over a 16*N^4-byte complex copy.  Here every element of TwoE is read exactly
// once (see the kernel comment); Re and Im planes of a complex density share the pass.
// HBM-bound: 8*N^4 bytes per build.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "handle.h"

namespace {

constexpr int JK_THREADS = 256;
constexpr int JK_WARPS = JK_THREADS / 32;

// One warp per (p,q): it streams the N rows T[p][x][q][:] (x = 0..N-1, N contiguous doubles each) and keeps
//     kacc   += T[p][x][q][r] * P[x][r]      -> one cross-lane reduction per warp  -> K[p][q]   (plain store)
//     jacc_r += T[p][x][q][r] * P[x][p]      -> per lane, no reduction             -> J[q][r]  (+= over p, atomics)
// J uses (pq|rs) = (rs|pq):  J[q][r] = sum_{p,x} T[p][x][q][r] P[x][p].  Two FMAs per loaded element, no shuffles
// in the streaming loop; every T element is read exactly once.
template <bool CPLX, bool VEC2, int NV>
__global__ void __launch_bounds__(JK_THREADS) jk_incore_kernel(const double *__restrict__ T, int N,
                                                               const double *__restrict__ Pre,
                                                               const double *__restrict__ Pim, double *Jre, double *Jim,
                                                               double *Kre, double *Kim)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t N2 = (size_t)N * N;
    constexpr int MAXV = NV;     // per-lane column slots: 32 * NV columns (double2 columns when VEC2)
    for (size_t pq = (size_t)blockIdx.x * JK_WARPS + warp; pq < N2; pq += (size_t)gridDim.x * JK_WARPS) {
        const int p = (int)(pq / N), q = (int)(pq % N);
        const double *base = T + (size_t)p * N2 * N + (size_t)q * N;      // T[p][0][q][0]; next x is +N2
        double kre = 0.0, kim = 0.0;
        if (VEC2) {
            const int n2 = N >> 1;
            double2 jr[MAXV], ji[MAXV];
#pragma unroll
            for (int v = 0; v < MAXV; ++v) { jr[v] = make_double2(0.0, 0.0); ji[v] = make_double2(0.0, 0.0); }
#pragma unroll 4
            for (int x = 0; x < N; ++x) {
                const double2 *row = reinterpret_cast<const double2 *>(base + (size_t)x * N2);
                const double2 *px = reinterpret_cast<const double2 *>(Pre + (size_t)x * N);
                const double pxp = Pre[(size_t)x * N + p];
                const double pxpi = CPLX ? Pim[(size_t)x * N + p] : 0.0;
#pragma unroll
                for (int v = 0; v < MAXV; ++v) {
                    const int r = lane + 32 * v;
                    if (r < n2) {
                        const double2 m = __ldcs(row + r);       // streamed once: evict-first
                        const double2 a = px[r];
                        kre = fma(m.x, a.x, fma(m.y, a.y, kre));
                        jr[v].x = fma(m.x, pxp, jr[v].x);
                        jr[v].y = fma(m.y, pxp, jr[v].y);
                        if (CPLX) {
                            const double2 ai = reinterpret_cast<const double2 *>(Pim + (size_t)x * N)[r];
                            kim = fma(m.x, ai.x, fma(m.y, ai.y, kim));
                            ji[v].x = fma(m.x, pxpi, ji[v].x);
                            ji[v].y = fma(m.y, pxpi, ji[v].y);
                        }
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < MAXV; ++v) {
                const int r = lane + 32 * v;
                if (r < n2) {
                    atomicAdd(&Jre[(size_t)q * N + 2 * r], jr[v].x);
                    atomicAdd(&Jre[(size_t)q * N + 2 * r + 1], jr[v].y);
                    if (CPLX) {
                        atomicAdd(&Jim[(size_t)q * N + 2 * r], ji[v].x);
                        atomicAdd(&Jim[(size_t)q * N + 2 * r + 1], ji[v].y);
                    }
                }
            }
        } else {
            double jr[MAXV], ji[MAXV];
#pragma unroll
            for (int v = 0; v < MAXV; ++v) { jr[v] = 0.0; ji[v] = 0.0; }
#pragma unroll 2
            for (int x = 0; x < N; ++x) {
                const double *row = base + (size_t)x * N2;
                const double *px = Pre + (size_t)x * N;
                const double pxp = px[p];
                const double pxpi = CPLX ? Pim[(size_t)x * N + p] : 0.0;
#pragma unroll
                for (int v = 0; v < MAXV; ++v) {
                    const int r = lane + 32 * v;
                    if (r < N) {
                        const double m = __ldcs(row + r);
                        kre = fma(m, px[r], kre);
                        jr[v] = fma(m, pxp, jr[v]);
                        if (CPLX) {
                            kim = fma(m, Pim[(size_t)x * N + r], kim);
                            ji[v] = fma(m, pxpi, ji[v]);
                        }
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < MAXV; ++v) {
                const int r = lane + 32 * v;
                if (r < N) {
                    atomicAdd(&Jre[(size_t)q * N + r], jr[v]);
                    if (CPLX) atomicAdd(&Jim[(size_t)q * N + r], ji[v]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kre += __shfl_xor_sync(0xffffffffu, kre, o);
            if (CPLX) kim += __shfl_xor_sync(0xffffffffu, kim, o);
        }
        if (lane == 0) {
            Kre[pq] = kre;            // K[p][q]: single owner, no atomics
            if (CPLX) Kim[pq] = kim;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Bulk-copy variant (even N): the same one-pass algorithm, but the tensor is moved by the TMA engine.
// One CTA owns (p, q0 .. q0+QB-1) and walks x = 0 .. N-1; for each x the QB rows T[p][x][q0..][:] are CONTIGUOUS
// (QB*N doubles), so one `cp.async.bulk` (SASS: UBLKCP) per stage brings them into shared memory and signals an
// mbarrier with the byte count.  STAGES copies are in flight per CTA whatever the instruction schedule of the
// consumers is — the first version kept only what 64-register warps could unroll (ncu: long-scoreboard stall 40,
// 63 % of the HBM peak).  Consumer warp w takes row w of the stage from shared memory: K[p][q0+w] (one cross-lane
// reduction at the end) and the per-lane J[q0+w][r] partial sums, exactly as in the first kernel; a dedicated producer
// warp keeps the ring full (full/empty mbarrier pairs, no block-wide barrier).
// ------------------------------------------------------------------------------------------------------------------
constexpr int JKB_QB = 8;          // rows (q values) per CTA = warps per CTA
constexpr int JKB_STAGES = 6;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised: warp JKB_QB is the PRODUCER (one lane issues the bulk copies, throttled only by the `empty`
// barriers of the ring), warps 0 .. JKB_QB-1 are CONSUMERS (row w of every stage).  No block-wide barrier in the
// stream: a consumer warp releases a stage with one mbarrier arrive, so the warps drift freely and STAGES copies of
// QB*N doubles stay in flight per CTA.
template <bool CPLX, int NV>
__global__ void __launch_bounds__((JKB_QB + 1) * 32) jk_incore_bulk_kernel(const double *__restrict__ T, int N,
                                                                           const double *__restrict__ Pre,
                                                                           const double *__restrict__ Pim, double *Jre, double *Jim,
                                                                           double *Kre, double *Kim)
{
    extern __shared__ __align__(128) unsigned char jkb_smem[];
    unsigned long long *full = reinterpret_cast<unsigned long long *>(jkb_smem);                 // [STAGES]
    unsigned long long *empty = full + JKB_STAGES;                                                // [STAGES]
    double *buf = reinterpret_cast<double *>(jkb_smem + 128);                                     // [STAGES][QB*N]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nqb = (N + JKB_QB - 1) / JKB_QB;
    const size_t N2 = (size_t)N * N;
    const int n2 = N >> 1;                                                                         // double2 columns
    // a stage = QB rows of T + the density row P[x][:] (+ its imaginary part): the consumers read everything from shared
    // memory, the L1 (mostly carved out as shared memory here) is not on the streaming path at all
    const size_t stage_elems = (size_t)JKB_QB * N + (size_t)(CPLX ? 2 : 1) * N;
    if (threadIdx.x == 0) {
        for (int s = 0; s < JKB_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], JKB_QB); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long it = 0;                                  // global stage counter: the ring runs across tasks
    if (warp == JKB_QB) {
        // ---- producer ----
        if (lane == 0) {
            for (int task = blockIdx.x; task < N * nqb; task += gridDim.x) {
                const int p = task / nqb, q0 = (task % nqb) * JKB_QB;
                const int rows = min(JKB_QB, N - q0);
                const unsigned bytes = (unsigned)(rows * N * sizeof(double));
                const double *src0 = T + (size_t)p * N2 * N + (size_t)q0 * N;         // T[p][0][q0][0]; next x is +N2
                for (int x = 0; x < N; ++x, ++it) {
                    const int s = (int)(it % JKB_STAGES);
                    if (it >= JKB_STAGES) mbar_wait(&empty[s], (unsigned)((it / JKB_STAGES - 1) & 1));
                    const unsigned pbytes = (unsigned)(N * sizeof(double));
                    mbar_expect_tx(&full[s], bytes + (CPLX ? 2u : 1u) * pbytes);
                    bulk_g2s(buf + s * stage_elems, src0 + (size_t)x * N2, bytes, &full[s]);
                    bulk_g2s(buf + s * stage_elems + (size_t)JKB_QB * N, Pre + (size_t)x * N, pbytes, &full[s]);
                    if (CPLX) bulk_g2s(buf + s * stage_elems + (size_t)JKB_QB * N + N, Pim + (size_t)x * N, pbytes, &full[s]);
                }
            }
        }
        return;
    }
    // ---- consumers ----
    for (int task = blockIdx.x; task < N * nqb; task += gridDim.x) {
        const int p = task / nqb, q0 = (task % nqb) * JKB_QB;
        const int rows = min(JKB_QB, N - q0);
        const bool active = warp < rows;
        double kre = 0.0, kim = 0.0;
        double2 jr[NV], ji[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) { jr[v] = make_double2(0.0, 0.0); ji[v] = make_double2(0.0, 0.0); }
        for (int x = 0; x < N; ++x, ++it) {
            const int s = (int)(it % JKB_STAGES);
            mbar_wait(&full[s], (unsigned)((it / JKB_STAGES) & 1));
            if (active) {
                const double *stg = buf + s * stage_elems;
                const double2 *row = reinterpret_cast<const double2 *>(stg + (size_t)warp * N);
                const double2 *px = reinterpret_cast<const double2 *>(stg + (size_t)JKB_QB * N);
                const double pxp = stg[(size_t)JKB_QB * N + p];
                const double pxpi = CPLX ? stg[(size_t)JKB_QB * N + N + p] : 0.0;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int r = lane + 32 * v;
                    if (r < n2) {
                        const double2 m = row[r];
                        const double2 a = px[r];
                        kre = fma(m.x, a.x, fma(m.y, a.y, kre));
                        jr[v].x = fma(m.x, pxp, jr[v].x);
                        jr[v].y = fma(m.y, pxp, jr[v].y);
                        if (CPLX) {
                            const double2 ai = reinterpret_cast<const double2 *>(stg + (size_t)JKB_QB * N + N)[r];
                            kim = fma(m.x, ai.x, fma(m.y, ai.y, kim));
                            ji[v].x = fma(m.x, pxpi, ji[v].x);
                            ji[v].y = fma(m.y, pxpi, ji[v].y);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);              // this warp is done with stage s
        }
        if (active) {
            const int q = q0 + warp;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int r = lane + 32 * v;
                if (r < n2) {
                    atomicAdd(&Jre[(size_t)q * N + 2 * r], jr[v].x);
                    atomicAdd(&Jre[(size_t)q * N + 2 * r + 1], jr[v].y);
                    if (CPLX) {
                        atomicAdd(&Jim[(size_t)q * N + 2 * r], ji[v].x);
                        atomicAdd(&Jim[(size_t)q * N + 2 * r + 1], ji[v].y);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kre += __shfl_xor_sync(0xffffffffu, kre, o);
                if (CPLX) kim += __shfl_xor_sync(0xffffffffu, kim, o);
            }
            if (lane == 0) {
                Kre[(size_t)p * N + q] = kre;            // K[p][q]: single owner, no atomics
                if (CPLX) Kim[(size_t)p * N + q] = kim;
            }
        }
    }
}

}  // namespace

extern "C" int mmdb_jk_incore(int device, const double *TwoE_dev, int N, const double *P_re_dev, const double *P_im_dev,
                              double *J_re_dev, double *J_im_dev, double *K_re_dev, double *K_im_dev, void *stream)
{
    if (N <= 0 || !TwoE_dev || !P_re_dev || !J_re_dev || !K_re_dev) return fail(MMDB_ERR_INVALID, "mmdb_jk_incore: bad arguments");
    if (P_im_dev && (!J_im_dev || !K_im_dev)) return fail(MMDB_ERR_INVALID, "mmdb_jk_incore: imaginary planes missing");
    CU(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N2 = (size_t)N * N;
    const bool cplx = P_im_dev != nullptr;
    if (N > 512) return fail(MMDB_ERR_UNSUPPORTED, "mmdb_jk_incore: N > 512 (the dense tensor would not fit one GPU anyway)");
    CU(cudaMemsetAsync(J_re_dev, 0, sizeof(double) * N2, st));
    if (cplx) CU(cudaMemsetAsync(J_im_dev, 0, sizeof(double) * N2, st));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    const int grid = (int)std::min<size_t>((N2 + JK_WARPS - 1) / JK_WARPS, (size_t)nsm * 16);
    const bool vec2 = (N % 2 == 0) && (((uintptr_t)TwoE_dev | (uintptr_t)P_re_dev | (uintptr_t)P_im_dev) % 16 == 0);
    const int cols = vec2 ? N / 2 : N;
    const int nv = (cols + 31) / 32;      // column slots per lane, rounded up to a compiled size
    if (vec2 && !getenv("MMDB_JK_NO_BULK")) {
        // TMA bulk-copy pipeline: STAGES x (QB rows of N doubles) in flight per CTA
        const size_t smem = 128 + (size_t)JKB_STAGES * ((size_t)JKB_QB * N + (cplx ? 2 : 1) * (size_t)N) * sizeof(double);
        const int tasks = N * ((N + JKB_QB - 1) / JKB_QB);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / smem));
        const int gridb = std::min(tasks, nsm * per_sm);
#define LAUNCHB(C, NVV)                                                                                                    \
    do {                                                                                                                   \
        cudaFuncSetAttribute(jk_incore_bulk_kernel<C, NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        jk_incore_bulk_kernel<C, NVV><<<gridb, (JKB_QB + 1) * 32, smem, st>>>(TwoE_dev, N, P_re_dev, P_im_dev, J_re_dev, J_im_dev, K_re_dev, K_im_dev); \
    } while (0)
#define LAUNCHB_NV(C)                          \
    do {                                       \
        if (nv <= 1) LAUNCHB(C, 1);            \
        else if (nv <= 2) LAUNCHB(C, 2);       \
        else if (nv <= 4) LAUNCHB(C, 4);       \
        else LAUNCHB(C, 8);                    \
    } while (0)
        if (smem <= 200 * 1024) {
            if (cplx) LAUNCHB_NV(true); else LAUNCHB_NV(false);
            CU(cudaGetLastError());
            return MMDB_OK;
        }
#undef LAUNCHB_NV
#undef LAUNCHB
    }
#define LAUNCH(C, V, NVV)                                                                                         \
    jk_incore_kernel<C, V, NVV><<<grid, JK_THREADS, 0, st>>>(TwoE_dev, N, P_re_dev, P_im_dev, J_re_dev, J_im_dev, K_re_dev, K_im_dev)
#define LAUNCH_NV(C, V)                                                         \
    do {                                                                        \
        if (nv <= 1) LAUNCH(C, V, 1);                                           \
        else if (nv <= 2) LAUNCH(C, V, 2);                                      \
        else if (nv <= 4) LAUNCH(C, V, 4);                                      \
        else if (nv <= 8) LAUNCH(C, V, 8);                                      \
        else LAUNCH(C, V, 16);                                                  \
    } while (0)
    if (cplx) {
        if (vec2) LAUNCH_NV(true, true); else LAUNCH_NV(true, false);
    } else {
        if (vec2) LAUNCH_NV(false, true); else LAUNCH_NV(false, false);
    }
#undef LAUNCH_NV
#undef LAUNCH
    CU(cudaGetLastError());
    return MMDB_OK;
}

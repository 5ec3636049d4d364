// jk_incore.cu — in-core J/K from the dense (N,N,N,N) tensor, ONE pass over TwoE.
//
// Replaces mmd/scf.py:97-98 of the reference:
//     J = einsum('pqrs,sr->pq', TwoE.astype(complex), P);  K = einsum('psqr,sr->pq', TwoE.astype(complex), P)
// which makes two passes over a 16*N^4-byte complex copy.  Here every element of TwoE is read exactly
// once (see the kernel comment); Re and Im planes of a complex density share the pass.
// HBM-bound: 8*N^4 bytes per build.
#include <algorithm>
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

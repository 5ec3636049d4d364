// jk_incore.cu — in-core J/K from the dense (N,N,N,N) tensor, ONE pass over TwoE.
//
// Replaces mmd/scf.py:97-98 of the reference:
//     J = einsum('pqrs,sr->pq', TwoE.astype(complex), P);  K = einsum('psqr,sr->pq', TwoE.astype(complex), P)
// which makes two passes over a 16*N^4-byte complex copy.  Here each CTA streams slabs
// M[q][r] = TwoE[p][x][q][r] (N*N contiguous doubles) exactly once and produces
//     J[p][x]  = sum_{q,r} M[q][r] * P[r][q]
//     K[p][q] += sum_r     M[q][r] * P[x][r]            (for all q)
// Re and Im planes of a complex density share the pass.  HBM-bound: 8*N^4 bytes per build.
#include <algorithm>
#include <string>

#include "handle.h"

namespace {

constexpr int JK_THREADS = 256;
constexpr int JK_WARPS = JK_THREADS / 32;

__global__ void transpose_kernel(const double *A, int N, double *AT)
{
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < (size_t)N * N; x += (size_t)gridDim.x * blockDim.x) {
        const size_t r = x / N, c = x % N;
        AT[c * N + r] = A[x];
    }
}

template <bool CPLX, bool VEC2>
__global__ void __launch_bounds__(JK_THREADS) jk_incore_kernel(const double *__restrict__ T, int N,
                                                               const double *__restrict__ Pre,
                                                               const double *__restrict__ Pim,
                                                               const double *__restrict__ PTre,
                                                               const double *__restrict__ PTim, double *Jre, double *Jim,
                                                               double *Kre, double *Kim)
{
    __shared__ double s_j[2][JK_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t N2 = (size_t)N * N;
    for (size_t slab = blockIdx.x; slab < N2; slab += gridDim.x) {
        const int p = (int)(slab / N), x = (int)(slab % N);
        const double *M = T + slab * N2;
        const double *px_re = Pre + (size_t)x * N;
        const double *px_im = CPLX ? Pim + (size_t)x * N : nullptr;
        double jre = 0.0, jim = 0.0;
        for (int q = warp; q < N; q += JK_WARPS) {
            const double *row = M + (size_t)q * N;
            const double *ptr = PTre + (size_t)q * N;
            const double *pti = CPLX ? PTim + (size_t)q * N : nullptr;
            double kre = 0.0, kim = 0.0;
            if (VEC2) {
                const int n2 = N >> 1;
                const double2 *row2 = reinterpret_cast<const double2 *>(row);
                const double2 *px2 = reinterpret_cast<const double2 *>(px_re);
                const double2 *pt2 = reinterpret_cast<const double2 *>(ptr);
#pragma unroll 2
                for (int r = lane; r < n2; r += 32) {
                    const double2 m = __ldcs(row2 + r);       // streamed once: evict-first
                    const double2 a = px2[r], t = pt2[r];
                    kre = fma(m.x, a.x, fma(m.y, a.y, kre));
                    jre = fma(m.x, t.x, fma(m.y, t.y, jre));
                    if (CPLX) {
                        const double2 ai = reinterpret_cast<const double2 *>(px_im)[r];
                        const double2 ti = reinterpret_cast<const double2 *>(pti)[r];
                        kim = fma(m.x, ai.x, fma(m.y, ai.y, kim));
                        jim = fma(m.x, ti.x, fma(m.y, ti.y, jim));
                    }
                }
            } else {
#pragma unroll 2
                for (int r = lane; r < N; r += 32) {
                    const double m = __ldcs(row + r);
                    kre = fma(m, px_re[r], kre);
                    jre = fma(m, ptr[r], jre);
                    if (CPLX) {
                        kim = fma(m, px_im[r], kim);
                        jim = fma(m, pti[r], jim);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kre += __shfl_xor_sync(0xffffffffu, kre, o);
                if (CPLX) kim += __shfl_xor_sync(0xffffffffu, kim, o);
            }
            if (lane == 0) {
                atomicAdd(&Kre[(size_t)p * N + q], kre);
                if (CPLX) atomicAdd(&Kim[(size_t)p * N + q], kim);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            jre += __shfl_xor_sync(0xffffffffu, jre, o);
            if (CPLX) jim += __shfl_xor_sync(0xffffffffu, jim, o);
        }
        if (lane == 0) {
            s_j[0][warp] = jre;
            s_j[1][warp] = jim;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < JK_WARPS; ++w) { a += s_j[0][w]; b += s_j[1][w]; }
            Jre[slab] = a;
            if (CPLX) Jim[slab] = b;
        }
        __syncthreads();
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
    double *PT = nullptr;
    CU(cudaMallocAsync(&PT, sizeof(double) * N2 * 2, st));
    transpose_kernel<<<(int)std::min<size_t>((N2 + 255) / 256, 4096), 256, 0, st>>>(P_re_dev, N, PT);
    if (cplx) transpose_kernel<<<(int)std::min<size_t>((N2 + 255) / 256, 4096), 256, 0, st>>>(P_im_dev, N, PT + N2);
    CU(cudaMemsetAsync(K_re_dev, 0, sizeof(double) * N2, st));
    if (cplx) CU(cudaMemsetAsync(K_im_dev, 0, sizeof(double) * N2, st));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    const int grid = (int)std::min<size_t>(N2, (size_t)nsm * 8 * 4);
    const bool vec2 = (N % 2 == 0) && (((uintptr_t)TwoE_dev | (uintptr_t)P_re_dev | (uintptr_t)P_im_dev) % 16 == 0);
#define LAUNCH(C, V)                                                                                              \
    jk_incore_kernel<C, V><<<grid, JK_THREADS, 0, st>>>(TwoE_dev, N, P_re_dev, P_im_dev, PT, PT + N2, J_re_dev, J_im_dev, \
                                                        K_re_dev, K_im_dev)
    if (cplx) {
        if (vec2) LAUNCH(true, true); else LAUNCH(true, false);
    } else {
        if (vec2) LAUNCH(false, true); else LAUNCH(false, false);
    }
#undef LAUNCH
    CU(cudaGetLastError());
    CU(cudaFreeAsync(PT, st));
    return MMDB_OK;
}

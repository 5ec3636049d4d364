// ao2mo.cu — AO->MO transformation of the dense ERI tensor and the closed-shell MP2 energy on the device.
//
// "Next" row: the consumers of mol.TwoE in the reference,
//   mmd/postscf.py:21-41  ao2mo   single_bar[P,Q,R,S] = sum_pqrs C[p,P] C[q,Q] C[r,R] C[s,S] (pq|rs)
//   mmd/postscf.py:59-70  MP2     E2 = sum_{ij occ, ab virt} (ia|jb) [2 (ia|jb) - (ib|ja)] / (e_i + e_j - e_a - e_b)
// The four quarter transformations are plain dense FP64 GEMMs (the one genuinely GEMM-shaped step of the
// path) and go to cuBLAS; the o^2 v^2 energy reduction is a small kernel.  Real orbitals only — the Python
// layer keeps the host path for complex (degenerate-subspace) orbitals.
#include <cublas_v2.h>

#include <string>

#include "handle.h"

namespace {

__global__ void __launch_bounds__(256) mp2_energy_kernel(const double *__restrict__ g, int N, int nocc,
                                                         const double *__restrict__ eps, double *out)
{
    // one CTA per (i,j); threads over (a,b)
    __shared__ double s_part[8];
    const int nv = N - nocc;
    const size_t n = (size_t)N;
    double acc = 0.0;
    for (int ij = blockIdx.x; ij < nocc * nocc; ij += gridDim.x) {
        const int i = ij / nocc, j = ij % nocc;
        const double eij = eps[i] + eps[j];
        for (int ab = threadIdx.x; ab < nv * nv; ab += blockDim.x) {
            const int a = nocc + ab / nv, b = nocc + ab % nv;
            const double iajb = g[((i * n + a) * n + j) * n + b];
            const double ibja = g[((i * n + b) * n + j) * n + a];
            acc += iajb * (2.0 * iajb - ibja) / (eij - eps[a] - eps[b]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
        atomicAdd(out, t);
    }
}

#define CB(call)                                                                                         \
    do {                                                                                                 \
        cublasStatus_t _s = (call);                                                                      \
        if (_s != CUBLAS_STATUS_SUCCESS) {                                                               \
            if (h) cublasDestroy(h);                                                                     \
            return fail(MMDB_ERR_CUDA, std::string(#call) + ": cuBLAS status " + std::to_string((int)_s)); \
        }                                                                                                \
    } while (0)

}  // namespace

// TwoE_dev: (N,N,N,N) row-major AO tensor (not modified).  C_dev: (N,N) row-major real MO coefficients
// (column P = orbital P).  MO_dev: (N,N,N,N) receives single_bar.  work_dev: N^4 doubles of scratch.
// eps_dev: N orbital energies.  e2_host receives the MP2 correlation energy (synchronises).
extern "C" int mmdb_ao2mo_mp2(int device, const double *TwoE_dev, int N, int nocc, const double *C_dev,
                              const double *eps_dev, double *MO_dev, double *work_dev, double *e2_host, void *stream)
{
    if (N <= 0 || nocc <= 0 || nocc >= N || !TwoE_dev || !C_dev || !MO_dev || !work_dev)
        return fail(MMDB_ERR_INVALID, "mmdb_ao2mo_mp2: bad arguments");
    CU(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    cublasHandle_t h = nullptr;
    CB(cublasCreate(&h));
    CB(cublasSetStream(h, st));
    const double one = 1.0, zero = 0.0;
    const long long n = N, n2 = n * n, n3 = n2 * n;
    // all matrices are row-major; row-major X = A * B is the column-major GEMM X^T = B^T * A^T
    // 1) T1[pqr,S] = sum_s T[pqr,s] C[s,S]                                  -> MO_dev
    CB(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, (int)n3, N, &one, C_dev, N, TwoE_dev, N, &zero, MO_dev, N));
    // 2) T2[pq,R,S] = sum_r C[r,R] T1[pq,r,S]        (batch over pq)        -> work_dev
    CB(cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N, &one, MO_dev, N, n2, C_dev, N, 0, &zero, work_dev, N,
                                 n2, (int)n2));
    // 3) T3[p,Q,RS] = sum_q C[q,Q] T2[p,q,RS]        (batch over p)         -> MO_dev
    CB(cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_T, (int)n2, N, N, &one, work_dev, (int)n2, n3, C_dev, N, 0, &zero,
                                 MO_dev, (int)n2, n3, N));
    // 4) T4[P,QRS] = sum_p C[p,P] T3[p,QRS]                                 -> work_dev, then copied to MO_dev
    CB(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, (int)n3, N, N, &one, MO_dev, (int)n3, C_dev, N, &zero, work_dev, (int)n3));
    cublasDestroy(h);
    h = nullptr;
    CU(cudaMemcpyAsync(MO_dev, work_dev, sizeof(double) * n3 * n, cudaMemcpyDeviceToDevice, st));
    if (e2_host) {
        double *acc = work_dev;   // scratch is free again
        CU(cudaMemsetAsync(acc, 0, sizeof(double), st));
        mp2_energy_kernel<<<nocc * nocc < 4096 ? nocc * nocc : 4096, 256, 0, st>>>(MO_dev, N, nocc, eps_dev, acc);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(e2_host, acc, sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return MMDB_OK;
}

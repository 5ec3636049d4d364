// kernels_a.cuh — class-specialised (compile-time angular momentum) shell-quartet kernels.
// One thread per contracted shell quartet; R, E, the Hermite intermediate and the contracted
// integrals live in registers.  When the (ab|cd) block is too large for the register file the ket
// component pairs are processed in NCHUNK static chunks (R is rebuilt per chunk).
#pragma once
#include "core.cuh"

namespace mmdb {

constexpr int KA_THREADS = 128;

// ket component pairs per chunk
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr int chunk_ncd()
{
    constexpr int NAB = ncart(LA) * ncart(LB), NCD = ncart(LC) * ncart(LD), NHB = nherm(LA + LB);
    int best = 1;
    for (int c = 1; c <= NCD; ++c)
        if (NCD % c == 0 && NAB * c <= 36 && NHB * c <= 40) best = c;
    return best;
}

template <int LA, int LB, int LC, int LD, int EPI, int CD0, int NCDC>
__device__ __forceinline__ void run_chunk(const EriArgs &a, unsigned long long e, const PairHdr &bh, const PairHdr &kh,
                                          const double *boys_tab, bool samePair)
{
    constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND;
    double out[NAB * NCDC];
    eval_quartet_chunk<LA, LB, LC, LD, CD0, NCDC>(bh, a.braP, kh, a.ketP, boys_tab, out);
    if constexpr (EPI == EPI_STORE) {
        double *o = a.out + e * (unsigned long long)(NAB * NCD);
        sfor<0, NAB>([&](auto ABI) {
            constexpr int ab = decltype(ABI)::value;
            constexpr double sab = comp_scale(LA, ab / NB) * comp_scale(LB, ab % NB);
            sfor<0, NCDC>([&](auto CDI) {
                constexpr int cdi = decltype(CDI)::value;
                constexpr int cd = CD0 + cdi;
                constexpr double s = sab * comp_scale(LC, cd / ND) * comp_scale(LD, cd % ND);
                o[ab * NCD + cd] = out[ab * NCDC + cdi] * s;
            });
        });
    } else {
        const bool sameAB = (bh.shA == bh.shB), sameCD = (kh.shA == kh.shB);
        sfor<0, NAB>([&](auto ABI) {
            constexpr int ab = decltype(ABI)::value;
            constexpr double sab = comp_scale(LA, ab / NB) * comp_scale(LB, ab % NB);
            const int i = bh.bfA + ab / NB, j = bh.bfB + ab % NB;
            sfor<0, NCDC>([&](auto CDI) {
                constexpr int cdi = decltype(CDI)::value;
                constexpr int cd = CD0 + cdi;
                constexpr double s = sab * comp_scale(LC, cd / ND) * comp_scale(LD, cd % ND);
                digest_fn_quartet(a.dg, i, j, kh.bfA + cd / ND, kh.bfB + cd % ND, sameAB, sameCD, samePair,
                                  out[ab * NCDC + cdi] * s);
            });
        });
    }
}

template <int LA, int LB, int LC, int LD, int EPI>
__global__ void __launch_bounds__(KA_THREADS) eri_class_kernel(const EriArgs a)
{
    constexpr int L = LA + LB + LC + LD;
    constexpr int NCD = ncart(LC) * ncart(LD);
    constexpr int NCDC = chunk_ncd<LA, LB, LC, LD>();
    constexpr int NCHUNK = NCD / NCDC;
    extern __shared__ double s_boys[];
    (void)L;
    for (int x = threadIdx.x; x < BOYS_ROWS * BOYS_STRIDE; x += blockDim.x) s_boys[x] = a.boys_tab[x];
    __syncthreads();
    const unsigned long long n = a.count_dev ? *a.count_dev : a.n;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const uint2 ij = a.list[e];
        const PairHdr bh = a.braH[ij.x];
        const PairHdr kh = a.ketH[ij.y];
        const bool samePair = a.same_class && (ij.x == ij.y);
        sfor<0, NCHUNK>([&](auto CH) {
            constexpr int ch = decltype(CH)::value;
            run_chunk<LA, LB, LC, LD, EPI, ch * NCDC, NCDC>(a, e, bh, kh, s_boys, samePair);
        });
    }
}

// host launcher, defined (explicitly instantiated) in inst_*.cu
template <int LA, int LB, int LC, int LD>
cudaError_t launch_class(const EriArgs &a, int epi, int grid, cudaStream_t st);

template <int LA, int LB, int LC, int LD>
cudaError_t launch_class_impl(const EriArgs &a, int epi, int grid, cudaStream_t st)
{
    const size_t smem = BOYS_ROWS * BOYS_STRIDE * sizeof(double);
    if (epi == EPI_STORE)
        eri_class_kernel<LA, LB, LC, LD, EPI_STORE><<<grid, KA_THREADS, smem, st>>>(a);
    else
        eri_class_kernel<LA, LB, LC, LD, EPI_DIGEST><<<grid, KA_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

#define MMDB_INSTANTIATE_CLASS(LA, LB, LC, LD)                                                            \
    template <>                                                                                          \
    cudaError_t launch_class<LA, LB, LC, LD>(const EriArgs &a, int epi, int grid, cudaStream_t st)       \
    {                                                                                                    \
        return launch_class_impl<LA, LB, LC, LD>(a, epi, grid, st);                                      \
    }

}  // namespace mmdb

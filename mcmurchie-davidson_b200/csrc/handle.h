// handle.h — the opaque mmdb_basis handle and error plumbing shared by the host translation units.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mmdb200.h"
#include "core.cuh"

extern thread_local std::string mmdb_g_err;
static inline int fail(int code, const std::string &msg)
{
    mmdb_g_err = msg;
    return code;
}
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return fail(MMDB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                                           std::to_string(__LINE__) + ")");                               \
    } while (0)
#define CHK(call)                   \
    do {                            \
        int _r = (call);            \
        if (_r != MMDB_OK) return _r; \
    } while (0)

struct ShellH {
    int am, nprim, poff, bf0;      // am: shell TYPE code (0 s, 1 p, 2 d; 3 = S2 pseudo-shell in the grouped list, core.cuh)
    double x, y, z;
    int poff2 = 0;                 // S2: offset of the SECOND contraction's coefficients (exponents are those at poff)
    int id = 0;                    // shell id the pair headers carry (index into the list's bf0 / density-block tables)
};

// Auxiliary streams of a direct build.  Small class pairs are bounded by the latency of their longest thread (a deeply
// contracted S2 quartet, a (dd|dd) block), not by throughput: on ONE extra stream their floors add up — at eight shards
// that queue, not the big classes, was the critical path (shard 0 of 8: 12.8 ms against 8.8 ideal).
constexpr int MMDB_NAUX = 8;       // streams created; MMDB_NAUX_DEFAULT of them used unless the environment says otherwise
constexpr int MMDB_NAUX_DEFAULT = 4;

// pair classes of the GROUPED shell list of the direct Fock build: the six plain ones + (S2 s), (S2 p), (S2 S2).
// An (S2, d) pair is expanded into its two plain (d s) pairs.
constexpr int MMDB_NCLASS_GC = 9;

struct PairClass {
    int la = 0, lb = 0;
    int npairs = 0;
    int64_t nprimpairs = 0;
    size_t slice_entries = 0;   // sum over pairs of ceil(pnum / BRA_SLICE): list entries one ket row can produce
    // far-field test of the screening kernel: bounding sphere of the product centres (x, y, z, radius) and smallest total
    // exponent — per SLICE of <= BRA_SLICE primitive pairs (bra side; primitives sorted by exponent, tight first) ...
    int *sbase_dev = nullptr;            // [npairs] first slice record of a pair
    double4 *sgeo_dev = nullptr;         // [slice_entries]
    double *spmin_dev = nullptr;         // [slice_entries]
    // ... and per whole pair (ket side)
    double4 *geo_dev = nullptr;          // [npairs] the same for the whole pair (ket side of the far-field test)
    double *pmin_dev = nullptr;          // [npairs]
    std::vector<mmdb::PairHdr> hdr;
    std::vector<mmdb::PrimPair> prim;
    mmdb::PairHdr *hdr_dev = nullptr;
    mmdb::PrimPair *prim_dev = nullptr;
    double2 *prim_ab_dev = nullptr;      // [nprimpairs] the two individual exponents of every primitive pair (gradient kernels)
    double *prim_soa_dev = nullptr;      // [8 fields][nprimpairs]: field f of primitive k of pair i at f*nprimpairs + row[k] + i
    long long *prim_row_dev = nullptr;   // [max pnum] row offsets of the structure-of-arrays copy
    double *Qs_dev = nullptr;   // [npairs]
    double *Qmax_dev = nullptr; // [ceil(npairs/256)] maxima of Qs over 256-pair chunks (screening early-exit)
    int *K_dev = nullptr;       // [npairs] primitive pairs per shell pair
    double *PQ_dev = nullptr;   // [npairs][2 * nab] the pair's own dP block and sqrt(Q) block, packed at the start of every direct build
    int *Kref_dev = nullptr;    // [npairs] statistics: primitive pairs summed over the member contractions | members << 24
    double *wgt_dev = nullptr;      // pairs with an S2 member: [nprimpairs][MAX_WGT] contraction weights (ket side)
    double *wgt_soa_dev = nullptr;  // the same as [MAX_WGT][nprimpairs] in the rows of prim_soa_dev (bra side)
    int2 *sh_dev = nullptr;     // [npairs] (shA, shB)
};

struct mmdb_basis {
    int device = 0;
    int nshell = 0, nbf = 0;
    int nsm = 148;
    std::vector<ShellH> sh;
    std::vector<double> exps, coefs;
    PairClass pc[MMDB_NCLASS_PAIR];
    // grouped shell list (S2 pseudo-shells) and its pair classes: what the direct Fock build runs on
    std::vector<ShellH> shg;
    int nshellg = 0;
    bool have_gc = false;                         // the list contains at least one S2 pseudo-shell
    PairClass pcg[MMDB_NCLASS_GC];
    int *shg_bf0_dev = nullptr, *shg_nf_dev = nullptr;
    double *DSg_dev = nullptr;                    // (nshellg,nshellg)
    double *boys_dev[mmdb::BOYS_MAXL + 1] = {nullptr};
    int *sh_bf0_dev = nullptr, *sh_nf_dev = nullptr;
    double *Q_dev = nullptr, *SQ_dev = nullptr;   // (N,N)
    bool have_schwarz = false;
    double *Dabs_dev = nullptr;                   // (N,N)
    double *DS_dev = nullptr;                     // (nshell,nshell)
    unsigned long long *dglob_dev = nullptr;      // max|dP| as bits
    uint2 *list_dev = nullptr;
    size_t list_cap = 0;
    unsigned long long *ctr_dev = nullptr;        // counters
    int nctr = 0;
    cudaStream_t aux_stream[MMDB_NAUX] = {nullptr};   // small class pairs run here, concurrently with the big ones and with each other
    cudaEvent_t ev_fork = nullptr, ev_join[MMDB_NAUX] = {nullptr};
    cudaStream_t main2_stream = nullptr;          // odd main-queue class pairs run here: their grids fill the SMs the previous pair's tail vacates
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
    cudaStream_t scr_stream = nullptr;            // screening of the NEXT class pair runs here, one task ahead of the ERI kernels
    cudaEvent_t ev_fork_scr = nullptr;
    std::vector<cudaEvent_t> ev_pool;             // per-task "list ready" / "list consumed" events of the screening pipeline
    double *scratch_dev = nullptr;
    size_t scratch_cap = 0;   // doubles
    double *eri_scratch_dev = nullptr;            // [2 + MMDB_NAUX regions][54 * nsm * 2048] contracted-block columns of the scratch_out classes (main / second main / auxiliary streams)
    double *stage_host = nullptr, *stage_dev = nullptr;   // mmdb_formPT_host staging: 4 planes of N^2 doubles each (page-locked / device)
    size_t stage_n = 0;
};


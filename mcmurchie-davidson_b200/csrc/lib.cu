// lib.cu — host side of libmmdb200.so: C ABI (include/mmdb200.h), shell-pair tables, screening,
// class dispatch, Schwarz table, dense fill, direct Fock build driver.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mmdb200.h"
#include "kernels_a.cuh"
#include "kernels_b.cuh"
#include "handle.h"

using namespace mmdb;

thread_local std::string mmdb_g_err;

extern "C" const char *mmdb_last_error(void) { return mmdb_g_err.c_str(); }
extern "C" int mmdb_version(void) { return 100; }
extern "C" int mmdb_device_count(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(MMDB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return MMDB_OK;
}

// ------------------------------------------------------------------------------------------
// FLOP model of SURVEY.md §8(d)
// ------------------------------------------------------------------------------------------
static int S_pair(int la, int lb)
{
    int s = 0;
    for (int a = 0; a < ncart(la); ++a)
        for (int b = 0; b < ncart(lb); ++b)
            s += (cart_pow(la, a, 0) + cart_pow(lb, b, 0) + 1) * (cart_pow(la, a, 1) + cart_pow(lb, b, 1) + 1) *
                 (cart_pow(la, a, 2) + cart_pow(lb, b, 2) + 1);
    return s;
}
extern "C" double mmdb_class_flops(int la, int lb, int lc, int ld)
{
    const int Lb = la + lb, Lk = lc + ld, L = Lb + Lk;
    (void)Lk;
    int NR = 0;
    for (int n = 0; n < L; ++n) NR += nherm(L - n) - 1;
    const int nab = ncart(la) * ncart(lb), ncd = ncart(lc) * ncart(ld);
    return 30.0 + (20 + 3 * L) + 2 * L + 3.0 * NR + 2.0 * S_pair(lc, ld) * (nherm(Lb) + 1) +
           2.0 * S_pair(la, lb) * (ncd + 1) + (double)nab * ncd;
}

// ------------------------------------------------------------------------------------------
// Boys tables (host, long double): F_m(T0) = exp(-T0) sum_k (2T0)^k / ((2m+1)(2m+3)...(2m+2k+1)),
// top m by series, the rest by downward recursion.
// ------------------------------------------------------------------------------------------
static void boys_ref_ld(int mmax, long double T, long double *F)
{
    long double term = 1.0L / (2.0L * mmax + 1.0L), sum = term;
    for (int k = 1; k < 4000; ++k) {
        term *= 2.0L * T / (2.0L * mmax + 2.0L * k + 1.0L);
        sum += term;
        if (term < 1e-22L * sum) break;
    }
    const long double eT = expl(-T);
    F[mmax] = eT * sum;
    for (int m = mmax; m > 0; --m) F[m - 1] = (2.0L * T * F[m] + eT) / (2.0L * m - 1.0L);
}

// every entry carries 2/sqrt(pi): the pair coefficients carry sqrt(pi)/2 (see PrimPair::cc), so neither Boys branch of
// the class kernels multiplies by it; boys_eval_rt (generic kernel, one-electron kernel, probes) undoes the factor
static const long double INV_SQRTPI_2 = 1.0L / 0.886226925452758013649083741670572591L;
static void make_boys_table(int L, std::vector<double> &tab);
void mmdb_make_boys_table(int L, std::vector<double> &tab) { make_boys_table(L, tab); }     // grad.cu (orders up to 9)
static void make_boys_table(int L, std::vector<double> &tab)
{
    tab.assign((size_t)BOYS_ROWS * BOYS_STRIDE, 0.0);
    long double F[BOYS_MAXL + 1 + 9 + 1];
    for (int r = 0; r < BOYS_ROWS; ++r) {
        const long double T0 = (long double)r * 0.125L;
        boys_ref_ld(L + 8, T0, F);
        long double fact = 1.0L;
        for (int k = 0; k <= 8; ++k) {
            if (k > 0) fact *= k;
            tab[(size_t)r * BOYS_STRIDE + k] = (double)(F[L + k] / fact * INV_SQRTPI_2);
        }
        tab[(size_t)r * BOYS_STRIDE + 9] = (double)(expl(-T0) * INV_SQRTPI_2);
    }
}

static inline int pc_index(int la, int lb) { return la * (la + 1) / 2 + lb; }

static int ensure_list(mmdb_basis *b, size_t entries)
{
    if (entries <= b->list_cap) return MMDB_OK;
    if (b->list_dev) cudaFree(b->list_dev);
    b->list_dev = nullptr;
    b->list_cap = 0;
    CU(cudaMalloc(&b->list_dev, entries * sizeof(uint2)));
    b->list_cap = entries;
    return MMDB_OK;
}
// scratch columns of the classes that keep their contracted block in global memory (kernels_a.cuh scratch_out):
// 18 x 3 doubles per resident thread, one region per stream that can run class kernels concurrently
static size_t eri_scratch_region(const mmdb_basis *b) { return (size_t)54 * b->nsm * 2048; }
static int ensure_eri_scratch(mmdb_basis *b)
{
    if (b->eri_scratch_dev) return MMDB_OK;
    CU(cudaMalloc(&b->eri_scratch_dev, (2 + MMDB_NAUX) * eri_scratch_region(b) * sizeof(double)));
    return MMDB_OK;
}

static int ensure_scratch(mmdb_basis *b, size_t doubles)
{
    if (doubles <= b->scratch_cap) return MMDB_OK;
    if (b->scratch_dev) cudaFree(b->scratch_dev);
    b->scratch_dev = nullptr;
    b->scratch_cap = 0;
    CU(cudaMalloc(&b->scratch_dev, doubles * sizeof(double)));
    b->scratch_cap = doubles;
    return MMDB_OK;
}

extern "C" int mmdb_basis_destroy(mmdb_basis *b)
{
    if (!b) return MMDB_OK;
    cudaSetDevice(b->device);
    auto free_class = [](PairClass &p) {
        cudaFree(p.hdr_dev); cudaFree(p.prim_dev); cudaFree(p.prim_ab_dev); cudaFree(p.prim_soa_dev); cudaFree(p.prim_row_dev); cudaFree(p.Qs_dev); cudaFree(p.Qmax_dev); cudaFree(p.K_dev); cudaFree(p.sh_dev);
        cudaFree(p.sbase_dev); cudaFree(p.sgeo_dev); cudaFree(p.spmin_dev); cudaFree(p.geo_dev); cudaFree(p.pmin_dev);
        cudaFree(p.Kref_dev); cudaFree(p.wgt_dev); cudaFree(p.wgt_soa_dev); cudaFree(p.PQ_dev);
    };
    for (auto &p : b->pc) free_class(p);
    for (auto &p : b->pcg) free_class(p);
    cudaFree(b->shg_bf0_dev); cudaFree(b->shg_nf_dev); cudaFree(b->DSg_dev);
    for (auto &t : b->boys_dev) cudaFree(t);
    cudaFree(b->sh_bf0_dev); cudaFree(b->sh_nf_dev); cudaFree(b->Q_dev); cudaFree(b->SQ_dev);
    cudaFree(b->Dabs_dev); cudaFree(b->DS_dev); cudaFree(b->dglob_dev); cudaFree(b->list_dev);
    cudaFree(b->ctr_dev); cudaFree(b->scratch_dev); cudaFree(b->eri_scratch_dev);
    if (b->stage_host) cudaFreeHost(b->stage_host);
    cudaFree(b->stage_dev);
    if (b->aux_stream[0]) {
        cudaEventDestroy(b->ev_fork);
        for (int x = 0; x < MMDB_NAUX; ++x) { cudaStreamDestroy(b->aux_stream[x]); cudaEventDestroy(b->ev_join[x]); }
    }
    if (b->scr_stream) { cudaStreamDestroy(b->scr_stream); cudaEventDestroy(b->ev_fork_scr); }
    if (b->main2_stream) { cudaStreamDestroy(b->main2_stream); cudaEventDestroy(b->ev_fork2); cudaEventDestroy(b->ev_join2); }
    for (cudaEvent_t e : b->ev_pool) cudaEventDestroy(e);
    delete b;
    return MMDB_OK;
}

static int basis_create_impl(mmdb_basis *b, int device, int nshell, const int *am, const int *nprim, const int *prim_off,
                             const double *centre, const double *exps, const double *coefs, const int *bf0, double prim_cut);

// bounding sphere (centre of the bounding box, largest distance to it) of n product centres, and their smallest exponent
static void prim_bounds(const PrimPair *pp, int n, double4 *geo, double *pmin)
{
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, pm = 1e300;
    for (int k = 0; k < n; ++k) {
        const double c[3] = {pp[k].Px, pp[k].Py, pp[k].Pz};
        for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], c[d]); hi[d] = std::max(hi[d], c[d]); }
        pm = std::min(pm, pp[k].p);
    }
    const double cx = 0.5 * (lo[0] + hi[0]), cy = 0.5 * (lo[1] + hi[1]), cz = 0.5 * (lo[2] + hi[2]);
    double r2 = 0.0;
    for (int k = 0; k < n; ++k) {
        const double dx = pp[k].Px - cx, dy = pp[k].Py - cy, dz = pp[k].Pz - cz;
        r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
    }
    *geo = make_double4(cx, cy, cz, std::sqrt(r2) * (1.0 + 1e-12) + 1e-12);
    *pmin = pm;
}

__global__ void qs_chunk_max_kernel(const double *Qs, int npairs, double *Qmax);

extern "C" int mmdb_basis_create(int device, int nshell, const int *am, const int *nprim, const int *prim_off,
                                 const double *centre, const double *exps, const double *coefs, const int *bf0,
                                 double prim_cut, mmdb_basis **out)
{
    if (!out || nshell <= 0) return fail(MMDB_ERR_INVALID, "mmdb_basis_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MMDB_ERR_CUDA, "mmdb_basis_create: no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(MMDB_ERR_INVALID, "mmdb_basis_create: device out of range");
    CU(cudaSetDevice(device));
    mmdb_basis *b = new mmdb_basis();
    b->device = device;
    const int rc = basis_create_impl(b, device, nshell, am, nprim, prim_off, centre, exps, coefs, bf0, prim_cut);
    if (rc != MMDB_OK) {               // every early return of the builder lands here: nothing leaks
        const std::string msg = mmdb_g_err;
        mmdb_basis_destroy(b);
        mmdb_g_err = msg;
        return rc;
    }
    *out = b;
    return MMDB_OK;
}

// pair classes by shell TYPE codes: the six plain ones (index la(la+1)/2 + lb) and, for the grouped shell list of the
// direct Fock build, (S2 s), (S2 p), (S2 S2)
static const int GC_CLASS[MMDB_NCLASS_GC][2] = {{0, 0}, {1, 0}, {1, 1}, {2, 0}, {2, 1}, {2, 2}, {3, 0}, {3, 1}, {3, 3}};
static int gc_class_index(int ta, int tb)
{
    for (int c = 0; c < MMDB_NCLASS_GC; ++c)
        if (GC_CLASS[c][0] == ta && GC_CLASS[c][1] == tb) return c;
    return -1;
}

// Shell pairs of one shell list, class by class, with every device table the kernels and the screen read.
static int build_pairs(mmdb_basis *b, const std::vector<ShellH> &sh, PairClass *pcs, int nclass, bool gc, double prim_cut)
{
    // type[A] >= type[B], equal types: A >= B
    const int nshell = (int)sh.size();
    const double SQRT2_PI54 = std::sqrt(2.0) * std::pow(M_PI, 1.25);
    for (int c = 0; c < nclass; ++c) {
        pcs[c].la = GC_CLASS[c][0];
        pcs[c].lb = GC_CLASS[c][1];
    }
    struct Tmp {
        PairHdr h;
        std::vector<PrimPair> pp;
        std::vector<double2> ab;       // (exponent on A, exponent on B) of every primitive pair, same order as pp
        std::vector<double> w;         // MAX_WGT contraction weights per primitive pair (pairs with an S2 member)
        int kref = 0;                  // primitive pairs summed over the member contractions (each with its own cut)
    };
    std::vector<Tmp> tmp[MMDB_NCLASS_GC];
    // (S2, d) pairs are expanded into the two plain (d, s) pairs of the members: the d classes keep their kernels
    std::vector<std::pair<ShellH, ShellH>> todo;
    for (int A = 0; A < nshell; ++A)
        for (int B = 0; B <= A; ++B) {
            int a = A, c = B;
            if (sh[a].am < sh[c].am) std::swap(a, c);
            if (sh[a].am == SH_S2 && sh[c].am == 2) {
                for (int m = 0; m < 2; ++m) {
                    ShellH mem = sh[a];
                    mem.am = 0;
                    if (m) mem.poff = mem.poff2;      // coefficients of the second contraction (exponents are equal)
                    mem.bf0 += m;
                    todo.push_back({sh[c], mem});
                }
            } else {
                todo.push_back({sh[a], sh[c]});
            }
        }
    for (const auto &pr : todo) {
        {
            const ShellH &sa = pr.first, &sb = pr.second;
            const double ABx = sa.x - sb.x, ABy = sa.y - sb.y, ABz = sa.z - sb.z;
            const double AB2 = ABx * ABx + ABy * ABy + ABz * ABz;
            Tmp t;
            std::memset(&t.h, 0, sizeof(PairHdr));
            t.h.bfA = sa.bf0; t.h.bfB = sb.bf0; t.h.shA = sa.id; t.h.shB = sb.id;
            t.h.ABx = ABx; t.h.ABy = ABy; t.h.ABz = ABz; t.h.Qs = 0.0;
            const int lab = am_of(sa.am) + am_of(sb.am);
            const int nwa = sa.am == SH_S2 ? 2 : 1, nwb = sb.am == SH_S2 ? 2 : 1, nw = nwa * nwb;
            for (int i = 0; i < sa.nprim; ++i)
                for (int j = 0; j < sb.nprim; ++j) {
                    const double ea = b->exps[sa.poff + i], eb = b->exps[sb.poff + j];
                    const double p = ea + eb, mu = ea * eb / p;
                    const double K = std::exp(-mu * AB2);
                    // contraction weights per member pair; a plain pair has one and folds it into cc
                    double w[MAX_WGT] = {0.0, 0.0, 0.0, 0.0};
                    for (int ma = 0; ma < nwa; ++ma)
                        for (int mb = 0; mb < nwb; ++mb)
                            w[ma * nwb + mb] = b->coefs[(ma ? sa.poff2 : sa.poff) + i] * b->coefs[(mb ? sb.poff2 : sb.poff) + j];
                    const double c2 = (nw == 1) ? w[0] : 1.0;
                    int nkeep = 0;
                    for (int m = 0; m < nw; ++m) {
                        // magnitude estimate: sqrt of the primitive (ss|ss)-like self repulsion, with a
                        // generous polynomial allowance for the angular factors
                        double est = std::fabs(w[m]) * K * std::pow(M_PI, 1.25) * std::pow(2.0, 0.25) / std::pow(p, 1.25);
                        est *= std::pow(1.0 + std::sqrt(AB2), lab) * std::pow(std::max(1.0, p), 0.5 * lab) * 16.0;
                        if (prim_cut > 0 && est < prim_cut) w[m] = 0.0;      // this member drops the primitive, as its plain pair would
                        else ++nkeep;
                    }
                    if (nkeep == 0) continue;
                    t.kref += nkeep;
                    PrimPair q;
                    q.p = p;
                    q.Px = (ea * sa.x + eb * sb.x) / p;
                    q.Py = (ea * sa.y + eb * sb.y) / p;
                    q.Pz = (ea * sa.z + eb * sb.z) / p;
                    q.PAx = q.Px - sa.x; q.PAy = q.Py - sa.y; q.PAz = q.Pz - sa.z;
                    // divided by sqrt(p): see prim_Fs (core.cuh); times sqrt(sqrt(pi)/2): a product of two carries the
                    // sqrt(pi)/2 of the asymptotic Boys function
                    q.cc = c2 * K * SQRT2_PI54 / (p * std::sqrt(p)) * std::sqrt(SQRTPI_2);
                    t.pp.push_back(q);
                    t.ab.push_back(make_double2(ea, eb));
                    t.w.insert(t.w.end(), w, w + MAX_WGT);
                }
            if (t.pp.empty()) continue;
            // tight primitive pairs first: the slices of a virtual bra pair then hold primitives of similar exponent, and
            // the tight slices are far-field (asymptotic Boys branch) at almost any distance
            {
                std::vector<int> ord(t.pp.size());
                for (size_t x = 0; x < ord.size(); ++x) ord[x] = (int)x;
                std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return t.pp[x].p > t.pp[y].p; });
                std::vector<PrimPair> pp2(ord.size());
                std::vector<double2> ab2(ord.size());
                std::vector<double> w2(t.w.size());
                for (size_t x = 0; x < ord.size(); ++x) {
                    pp2[x] = t.pp[ord[x]]; ab2[x] = t.ab[ord[x]];
                    for (int m = 0; m < MAX_WGT; ++m) w2[x * MAX_WGT + m] = t.w[(size_t)ord[x] * MAX_WGT + m];
                }
                t.pp.swap(pp2);
                t.ab.swap(ab2);
                t.w.swap(w2);
            }
            t.h.pnum = (int)t.pp.size();
            const int cls = gc_class_index(sa.am, sb.am);
            if (cls < 0 || cls >= nclass) return fail(MMDB_ERR_INVALID, "build_pairs: pair class out of range");
            tmp[cls].push_back(std::move(t));
        }
    }
    for (int c = 0; c < nclass; ++c) {
        PairClass &P = pcs[c];
        auto &v = tmp[c];
        const bool weighted = nwgt(P.la, P.lb) > 1;
        // homogeneous contraction depth inside a warp: order by primitive-pair count (desc); inside one depth by an ESTIMATE
        // of the Schwarz bound (desc), so the pairs that can survive a weak ket row are a prefix of every depth group and
        // the screening kernel drops the rest tile by tile on its chunk maxima (which use the exact bounds).  The estimate:
        // sqrt of the s-type self-repulsion of the pair with F_0 <= 1,  sum_kl |cc_k cc_l| sqrt(p_k p_l / (p_k + p_l)).
        const bool sort_q = getenv("MMDB_PAIR_SORT") ? atoi(getenv("MMDB_PAIR_SORT")) != 0 : true;
        for (auto &t : v) {
            double q = 0.0;
            std::vector<double> wm(t.pp.size(), 1.0);       // largest member weight of every primitive pair
            if (weighted)
                for (size_t x = 0; x < t.pp.size(); ++x) {
                    wm[x] = 0.0;
                    for (int m = 0; m < MAX_WGT; ++m) wm[x] = std::max(wm[x], std::fabs(t.w[x * MAX_WGT + m]));
                }
            for (size_t x = 0; x < t.pp.size(); ++x)
                for (size_t y = 0; y < t.pp.size(); ++y)
                    q += std::fabs(t.pp[x].cc * wm[x] * t.pp[y].cc * wm[y]) * std::sqrt(t.pp[x].p * t.pp[y].p / (t.pp[x].p + t.pp[y].p));
            t.h.Qs = sort_q ? std::sqrt(q) : 0.0;      // overwritten with the exact bound by mmdb_schwarz
        }
        std::stable_sort(v.begin(), v.end(), [](const Tmp &x, const Tmp &y) {
            if (x.h.pnum != y.h.pnum) return x.h.pnum > y.h.pnum;
            // half-decades: keeps the construction (A-major) order inside a bucket, i.e. neighbouring columns share a shell
            const int bx = x.h.Qs > 0 ? (int)std::floor(2.0 * std::log10(x.h.Qs)) : -1000, by = y.h.Qs > 0 ? (int)std::floor(2.0 * std::log10(y.h.Qs)) : -1000;
            return bx > by;
        });
        P.npairs = (int)v.size();
        std::vector<double2> prim_ab;
        std::vector<double> wgt;
        std::vector<int> Kref;
        for (auto &t : v) {
            t.h.poff = (int)P.prim.size();
            t.h.pad0 = (int)P.hdr.size();        // own index in the class (virtual pairs carry their parent's here)
            P.prim.insert(P.prim.end(), t.pp.begin(), t.pp.end());
            prim_ab.insert(prim_ab.end(), t.ab.begin(), t.ab.end());
            if (weighted) wgt.insert(wgt.end(), t.w.begin(), t.w.end());
            Kref.push_back(t.kref | (nwgt(P.la, P.lb) << 24));
            P.hdr.push_back(t.h);
        }
        P.nprimpairs = (int64_t)P.prim.size();
        if (P.npairs == 0) continue;
        if ((unsigned)P.npairs > PAIR_MASK) return fail(MMDB_ERR_UNSUPPORTED, "more than 2^24 shell pairs in one class");
        std::vector<int> K(P.npairs);
        std::vector<int2> shs(P.npairs);
        for (int i = 0; i < P.npairs; ++i) {
            K[i] = P.hdr[i].pnum;
            shs[i] = make_int2(P.hdr[i].shA, P.hdr[i].shB);
        }
        CU(cudaMalloc(&P.hdr_dev, sizeof(PairHdr) * P.npairs));
        CU(cudaMalloc(&P.prim_dev, sizeof(PrimPair) * P.prim.size()));
        CU(cudaMalloc(&P.Qs_dev, sizeof(double) * P.npairs));
        CU(cudaMalloc(&P.Qmax_dev, sizeof(double) * ((P.npairs + 255) / 256)));
        CU(cudaMalloc(&P.K_dev, sizeof(int) * P.npairs));
        CU(cudaMalloc(&P.Kref_dev, sizeof(int) * P.npairs));
        CU(cudaMalloc(&P.PQ_dev, sizeof(double) * (size_t)P.npairs * 2 * ncomp(P.la) * ncomp(P.lb)));
        CU(cudaMemcpy(P.Kref_dev, Kref.data(), sizeof(int) * P.npairs, cudaMemcpyHostToDevice));
        if (weighted) {
            CU(cudaMalloc(&P.wgt_dev, sizeof(double) * wgt.size()));
            CU(cudaMemcpy(P.wgt_dev, wgt.data(), sizeof(double) * wgt.size(), cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&P.sh_dev, sizeof(int2) * P.npairs));
        CU(cudaMemcpy(P.hdr_dev, P.hdr.data(), sizeof(PairHdr) * P.npairs, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(P.prim_dev, P.prim.data(), sizeof(PrimPair) * P.prim.size(), cudaMemcpyHostToDevice));
        if (!gc) {
            CU(cudaMalloc(&P.prim_ab_dev, sizeof(double2) * prim_ab.size()));
            CU(cudaMemcpy(P.prim_ab_dev, prim_ab.data(), sizeof(double2) * prim_ab.size(), cudaMemcpyHostToDevice));
        }
        {   // structure-of-arrays copy for the bra side (see BraSrc in core.cuh); pairs are sorted by pnum (desc),
            // so primitive k exists for the first n_k pairs and row k holds exactly those
            const int kmax = P.hdr[0].pnum;
            const long long nprim = (long long)P.prim.size();
            std::vector<long long> row(kmax + 1, 0);
            for (int k = 0; k < kmax; ++k) {
                long long nk = 0;
                while (nk < P.npairs && P.hdr[nk].pnum > k) ++nk;
                row[k + 1] = row[k] + nk;
            }
            std::vector<double> soa((size_t)8 * nprim);
            for (int i = 0; i < P.npairs; ++i)
                for (int k = 0; k < P.hdr[i].pnum; ++k) {
                    const PrimPair &q = P.prim[P.hdr[i].poff + k];
                    const double f[8] = {q.Px, q.Py, q.Pz, q.p, q.cc, q.PAx, q.PAy, q.PAz};
                    for (int x = 0; x < 8; ++x) soa[(size_t)x * nprim + row[k] + i] = f[x];
                }
            if (weighted) {
                std::vector<double> wsoa((size_t)MAX_WGT * nprim);
                for (int i = 0; i < P.npairs; ++i)
                    for (int k = 0; k < P.hdr[i].pnum; ++k)
                        for (int x = 0; x < MAX_WGT; ++x) wsoa[(size_t)x * nprim + row[k] + i] = wgt[(size_t)(P.hdr[i].poff + k) * MAX_WGT + x];
                CU(cudaMalloc(&P.wgt_soa_dev, sizeof(double) * wsoa.size()));
                CU(cudaMemcpy(P.wgt_soa_dev, wsoa.data(), sizeof(double) * wsoa.size(), cudaMemcpyHostToDevice));
            }
            CU(cudaMalloc(&P.prim_soa_dev, sizeof(double) * soa.size()));
            CU(cudaMalloc(&P.prim_row_dev, sizeof(long long) * kmax));
            CU(cudaMemcpy(P.prim_soa_dev, soa.data(), sizeof(double) * soa.size(), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(P.prim_row_dev, row.data(), sizeof(long long) * kmax, cudaMemcpyHostToDevice));
        }
        {   // bounding sphere of the product centres and the smallest total exponent of every pair (ket side of the
            // far-field test in the screening kernel)
            std::vector<double4> geo(P.npairs);
            std::vector<double> pmin(P.npairs);
            for (int i = 0; i < P.npairs; ++i) prim_bounds(&P.prim[P.hdr[i].poff], P.hdr[i].pnum, &geo[i], &pmin[i]);
            CU(cudaMalloc(&P.geo_dev, sizeof(double4) * P.npairs));
            CU(cudaMalloc(&P.pmin_dev, sizeof(double) * P.npairs));
            CU(cudaMemcpy(P.geo_dev, geo.data(), sizeof(double4) * P.npairs, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(P.pmin_dev, pmin.data(), sizeof(double) * P.npairs, cudaMemcpyHostToDevice));
            // the same per SLICE of <= BRA_SLICE primitive pairs (bra side: one list entry per slice)
            std::vector<int> sbase(P.npairs);
            std::vector<double4> sgeo;
            std::vector<double> spmin;
            for (int i = 0; i < P.npairs; ++i) {
                sbase[i] = (int)sgeo.size();
                for (int p0 = 0; p0 < P.hdr[i].pnum; p0 += BRA_SLICE) {
                    double4 gq; double pm;
                    prim_bounds(&P.prim[P.hdr[i].poff + p0], std::min(BRA_SLICE, P.hdr[i].pnum - p0), &gq, &pm);
                    sgeo.push_back(gq);
                    spmin.push_back(pm);
                }
            }
            P.slice_entries = sgeo.size();
            CU(cudaMalloc(&P.sbase_dev, sizeof(int) * P.npairs));
            CU(cudaMalloc(&P.sgeo_dev, sizeof(double4) * sgeo.size()));
            CU(cudaMalloc(&P.spmin_dev, sizeof(double) * spmin.size()));
            CU(cudaMemcpy(P.sbase_dev, sbase.data(), sizeof(int) * P.npairs, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(P.sgeo_dev, sgeo.data(), sizeof(double4) * sgeo.size(), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(P.spmin_dev, spmin.data(), sizeof(double) * spmin.size(), cudaMemcpyHostToDevice));
        }
        CU(cudaMemset(P.Qs_dev, 0, sizeof(double) * P.npairs));
        CU(cudaMemcpy(P.K_dev, K.data(), sizeof(int) * P.npairs, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(P.sh_dev, shs.data(), sizeof(int2) * P.npairs, cudaMemcpyHostToDevice));
    }
    return MMDB_OK;
}

static int basis_create_impl(mmdb_basis *b, int device, int nshell, const int *am, const int *nprim, const int *prim_off,
                             const double *centre, const double *exps, const double *coefs, const int *bf0, double prim_cut)
{
    b->nshell = nshell;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    b->nsm = prop.multiProcessorCount;
    int ntot = 0, nbf = 0;
    for (int s = 0; s < nshell; ++s) {
        if (am[s] < 0 || am[s] > MMDB_MAX_AM) {
            return fail(MMDB_ERR_UNSUPPORTED, "mmdb_basis_create: angular momentum > d is not supported on the device path");
        }
        ShellH h{am[s], nprim[s], prim_off[s], bf0[s], centre[3 * s], centre[3 * s + 1], centre[3 * s + 2]};
        b->sh.push_back(h);
        ntot = std::max(ntot, prim_off[s] + nprim[s]);
        nbf = std::max(nbf, bf0[s] + ncart(am[s]));
    }
    b->nbf = nbf;
    b->exps.assign(exps, exps + ntot);
    b->coefs.assign(coefs, coefs + ntot);

    // ---- shell pairs, by class --------------------------------------------------------------------
    for (int s = 0; s < nshell; ++s) b->sh[s].id = s;
    CHK(build_pairs(b, b->sh, b->pc, MMDB_NCLASS_PAIR, false, prim_cut));
    // ---- grouped shell list of the direct Fock build: two consecutive s shells on one centre with identical exponents
    // and adjacent functions (a generally contracted pair, e.g. the first two s functions of cc-pVDZ oxygen) become ONE
    // S2 pseudo-shell whose primitive integrals are evaluated once
    {
        std::vector<char> used(nshell, 0);
        for (int s = 0; s < nshell; ++s) {
            if (used[s]) continue;
            ShellH h = b->sh[s];
            // MMDB_GC_MODE: 0 = no grouping, 1 (default) = s shells with identical primitives, 2 = any two consecutive s
            // shells of one centre: the pseudo-shell then runs over the UNION of their primitives with zero coefficients
            // where a member lacks one (shares the pair / quartet / digestion bookkeeping but saves no primitive; measured
            // on (H2O)32/cc-pVDZ: screening 11 -> 7 ms, but the wide S2 kernels then carry everything: 73.8 -> 86.8 ms)
            const int gc_mode = getenv("MMDB_NO_GC") ? 0 : (getenv("MMDB_GC_MODE") ? atoi(getenv("MMDB_GC_MODE")) : 1);
            if (h.am == 0 && s + 1 < nshell && gc_mode > 0) {
                const ShellH &g = b->sh[s + 1];
                const bool partner = g.am == 0 && g.bf0 == h.bf0 + 1 && g.x == h.x && g.y == h.y && g.z == h.z;
                bool same = partner && g.nprim == h.nprim;
                for (int k = 0; same && k < h.nprim; ++k) same = b->exps[h.poff + k] == b->exps[g.poff + k];
                if (same && h.nprim > 1) {
                    h.am = SH_S2;
                    h.poff2 = g.poff;
                    used[s + 1] = 1;
                    b->have_gc = true;
                } else if (partner && gc_mode >= 2) {
                    std::vector<double> U, cA, cB;
                    for (int k = 0; k < h.nprim; ++k) { U.push_back(b->exps[h.poff + k]); cA.push_back(b->coefs[h.poff + k]); cB.push_back(0.0); }
                    for (int k = 0; k < g.nprim; ++k) {
                        size_t at = U.size();
                        for (size_t x = 0; x < U.size(); ++x) if (U[x] == b->exps[g.poff + k]) at = x;
                        if (at == U.size()) { U.push_back(b->exps[g.poff + k]); cA.push_back(0.0); cB.push_back(0.0); }
                        cB[at] += b->coefs[g.poff + k];
                    }
                    // two parallel blocks (exponents, coefficients of member 0 / member 1) appended to the primitive arrays
                    h.am = SH_S2;
                    h.nprim = (int)U.size();
                    h.poff = (int)b->exps.size();
                    b->exps.insert(b->exps.end(), U.begin(), U.end());
                    b->coefs.insert(b->coefs.end(), cA.begin(), cA.end());
                    h.poff2 = (int)b->exps.size();
                    b->exps.insert(b->exps.end(), U.begin(), U.end());
                    b->coefs.insert(b->coefs.end(), cB.begin(), cB.end());
                    used[s + 1] = 1;
                    b->have_gc = true;
                }
            }
            h.id = (int)b->shg.size();
            b->shg.push_back(h);
        }
        b->nshellg = (int)b->shg.size();
        if (b->have_gc) {
            CHK(build_pairs(b, b->shg, b->pcg, MMDB_NCLASS_GC, true, prim_cut));
            std::vector<int> f0(b->nshellg), nf(b->nshellg);
            for (int s = 0; s < b->nshellg; ++s) { f0[s] = b->shg[s].bf0; nf[s] = ncomp(b->shg[s].am); }
            CU(cudaMalloc(&b->shg_bf0_dev, sizeof(int) * b->nshellg));
            CU(cudaMalloc(&b->shg_nf_dev, sizeof(int) * b->nshellg));
            CU(cudaMemcpy(b->shg_bf0_dev, f0.data(), sizeof(int) * b->nshellg, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(b->shg_nf_dev, nf.data(), sizeof(int) * b->nshellg, cudaMemcpyHostToDevice));
            CU(cudaMalloc(&b->DSg_dev, (size_t)b->nshellg * b->nshellg * sizeof(double)));
        }
    }
    // ---- Boys tables --------------------------------------------------------------------------
    for (int L = 0; L <= BOYS_MAXL; ++L) {
        std::vector<double> tab;
        make_boys_table(L, tab);
        CU(cudaMalloc(&b->boys_dev[L], tab.size() * sizeof(double)));
        CU(cudaMemcpy(b->boys_dev[L], tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    // ---- shell meta, matrices, counters -------------------------------------------------------
    {
        std::vector<int> f0(nshell), nf(nshell);
        for (int s = 0; s < nshell; ++s) { f0[s] = b->sh[s].bf0; nf[s] = ncart(b->sh[s].am); }
        CU(cudaMalloc(&b->sh_bf0_dev, sizeof(int) * nshell));
        CU(cudaMalloc(&b->sh_nf_dev, sizeof(int) * nshell));
        CU(cudaMemcpy(b->sh_bf0_dev, f0.data(), sizeof(int) * nshell, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(b->sh_nf_dev, nf.data(), sizeof(int) * nshell, cudaMemcpyHostToDevice));
    }
    const size_t N2 = (size_t)nbf * nbf;
    CU(cudaMalloc(&b->Q_dev, N2 * sizeof(double)));
    CU(cudaMalloc(&b->SQ_dev, N2 * sizeof(double)));
    CU(cudaMalloc(&b->Dabs_dev, N2 * sizeof(double)));
    CU(cudaMalloc(&b->DS_dev, (size_t)nshell * nshell * sizeof(double)));
    CU(cudaMalloc(&b->dglob_dev, sizeof(unsigned long long)));
    b->nctr = 4096;
    CU(cudaMalloc(&b->ctr_dev, sizeof(unsigned long long) * b->nctr));
    return MMDB_OK;
}

extern "C" int mmdb_basis_nbf(const mmdb_basis *b, int *nbf)
{
    *nbf = b->nbf;
    return MMDB_OK;
}
extern "C" int mmdb_basis_pair_counts(const mmdb_basis *b, int64_t *npairs, int64_t *nprimpairs)
{
    for (int c = 0; c < MMDB_NCLASS_PAIR; ++c) {
        npairs[c] = b->pc[c].npairs;
        nprimpairs[c] = b->pc[c].nprimpairs;
    }
    return MMDB_OK;
}
extern "C" int mmdb_basis_pair_shells(const mmdb_basis *b, int pc, int *shA, int *shB)
{
    if (pc < 0 || pc >= MMDB_NCLASS_PAIR) return fail(MMDB_ERR_INVALID, "pair class out of range");
    const PairClass &P = b->pc[pc];
    for (int i = 0; i < P.npairs; ++i) {
        shA[i] = P.hdr[i].shA;
        shB[i] = P.hdr[i].shB;
    }
    return MMDB_OK;
}

// ------------------------------------------------------------------------------------------
// class dispatch
// ------------------------------------------------------------------------------------------
namespace mmdb {
#define DECL(LA, LB, LC, LD) \
    template <>              \
    cudaError_t launch_class<LA, LB, LC, LD>(const EriArgs &, int, int, cudaStream_t);
DECL(0, 0, 0, 0) DECL(1, 0, 0, 0) DECL(1, 0, 1, 0) DECL(1, 1, 0, 0) DECL(1, 1, 1, 0) DECL(1, 1, 1, 1)
DECL(2, 0, 0, 0) DECL(2, 0, 1, 0) DECL(2, 0, 1, 1) DECL(2, 0, 2, 0)
DECL(2, 1, 0, 0) DECL(2, 1, 1, 0) DECL(2, 1, 1, 1) DECL(2, 1, 2, 0)
DECL(2, 2, 0, 0) DECL(2, 2, 1, 0)
DECL(2, 1, 2, 1) DECL(2, 2, 1, 1) DECL(2, 2, 2, 0) DECL(2, 2, 2, 1) DECL(2, 2, 2, 2)
// classes with an S2 pseudo-shell (type code 3): direct Fock build only; ket = the more deeply contracted pair
DECL(1, 0, 2, 1) DECL(1, 1, 2, 0)
DECL(0, 0, 3, 0) DECL(1, 0, 3, 0) DECL(3, 0, 3, 0)
DECL(0, 0, 3, 1) DECL(1, 0, 3, 1) DECL(3, 1, 3, 0) DECL(3, 1, 3, 1)
DECL(0, 0, 3, 3) DECL(1, 0, 3, 3) DECL(3, 3, 3, 0) DECL(3, 3, 3, 1) DECL(3, 3, 3, 3)
#undef DECL
}  // namespace mmdb

// every class (ss|ss) ... (dd|dd) has a class-specialised kernel; the generic runtime-L kernel (impl = 1) is the
// independent cross-check
// plain classes: (la lb) >= (lc ld) in pair-class order; classes with an S2 pseudo-shell: the instantiated orientations
static int launch_eri(mmdb_basis *b, int la, int lb, int lc, int ld, EriArgs &a, int kind, int impl, cudaStream_t st)
{
    const int L = am_of(la) + am_of(lb) + am_of(lc) + am_of(ld);      // la..ld are shell type codes
    a.boys_tab = b->boys_dev[L];
    CHK(ensure_eri_scratch(b));
    int region = 0;                      // scratch columns per stream: caller's, second main, auxiliary streams
    if (b->main2_stream != nullptr && st == b->main2_stream) region = 1;
    for (int x = 0; x < MMDB_NAUX; ++x)
        if (b->aux_stream[x] != nullptr && st == b->aux_stream[x]) region = 2 + x;
    a.scratch = b->eri_scratch_dev + (size_t)region * eri_scratch_region(b);
    const int key = ((la * 4 + lb) * 4 + lc) * 4 + ld;
    cudaError_t e = cudaSuccess;
    const int gridA = b->nsm;   // x occupancy inside launch_class
    if (impl == 0) {
        switch (key) {
#define CASE(LA, LB, LC, LD)                                  \
    case ((LA * 4 + LB) * 4 + LC) * 4 + LD:                   \
        e = launch_class<LA, LB, LC, LD>(a, kind, gridA, st); \
        break;
            CASE(0, 0, 0, 0) CASE(1, 0, 0, 0) CASE(1, 0, 1, 0) CASE(1, 1, 0, 0) CASE(1, 1, 1, 0) CASE(1, 1, 1, 1)
            CASE(2, 0, 0, 0) CASE(2, 0, 1, 0) CASE(2, 0, 1, 1) CASE(2, 0, 2, 0)
            CASE(2, 1, 0, 0) CASE(2, 1, 1, 0) CASE(2, 1, 1, 1) CASE(2, 1, 2, 0)
            CASE(2, 2, 0, 0) CASE(2, 2, 1, 0)
            CASE(2, 1, 2, 1) CASE(2, 2, 1, 1) CASE(2, 2, 2, 0) CASE(2, 2, 2, 1) CASE(2, 2, 2, 2)
            CASE(1, 0, 2, 1) CASE(1, 1, 2, 0)
            CASE(0, 0, 3, 0) CASE(1, 0, 3, 0) CASE(3, 0, 3, 0)
            CASE(0, 0, 3, 1) CASE(1, 0, 3, 1) CASE(3, 1, 3, 0) CASE(3, 1, 3, 1)
            CASE(0, 0, 3, 3) CASE(1, 0, 3, 3) CASE(3, 3, 3, 0) CASE(3, 3, 3, 1) CASE(3, 3, 3, 3)
#undef CASE
            default:
                return fail(MMDB_ERR_INVALID, "launch_eri: class not instantiated");
        }
    } else {
        // generic runtime-L kernel: the independent cross-check of the class kernels (stores integrals only)
        if (kind != LK_STORE) return fail(MMDB_ERR_INVALID, "launch_eri: the generic kernel only stores integrals");
        const size_t smem = BOYS_ROWS * BOYS_STRIDE * sizeof(double);
        const int grid = b->nsm * 8;
        eri_generic_kernel<<<grid, KB_THREADS, smem, st>>>(a, la, lb, lc, ld);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return fail(MMDB_ERR_CUDA, std::string("ERI kernel launch: ") + cudaGetErrorString(e));
    return MMDB_OK;
}

// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
__global__ void zip_list_kernel(const int32_t *bi, const int32_t *ki, int64_t n, uint2 *list)
{
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (int64_t)gridDim.x * blockDim.x)
        list[x] = make_uint2((unsigned)bi[x], (unsigned)ki[x]);
}

__global__ void diag_list_kernel(int n, uint2 *list)
{
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) list[x] = make_uint2(x, x);
}

// Schwarz extraction: scratch[i][ab*nab+ab] -> Q, SQ, Qs
__global__ void schwarz_extract_kernel(const PairHdr *hdr, int npairs, int la, int lb, const double *scratch, int N,
                                       double *Q, double *SQ, double *Qs)
{
    const int na = ncart(la), nb = ncart(lb), nab = na * nb;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += gridDim.x * blockDim.x) {
        const PairHdr h = hdr[i];
        double qmax = 0.0;
        for (int ab = 0; ab < nab; ++ab) {
            const double v = scratch[(size_t)i * nab * nab + (size_t)ab * nab + ab];
            const int p = h.bfA + ab / nb, q = h.bfB + ab % nb;
            Q[(size_t)p * N + q] = v;
            Q[(size_t)q * N + p] = v;
            const double s = sqrt(v);   // NaN for (numerically) negative values, like the reference
            SQ[(size_t)p * N + q] = s;
            SQ[(size_t)q * N + p] = s;
            qmax = fmax(qmax, sqrt(fabs(v)));
        }
        Qs[i] = qmax;
    }
}

// pair bounds of the grouped pair classes from the function-level table: Qs[i] = max over the pair's function pairs of sqrt|Q|
__global__ void gc_bounds_kernel(const PairHdr *hdr, int npairs, int ta, int tb, int N, const double *Q, double *Qs)
{
    const int na = ncomp(ta), nb = ncomp(tb);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += gridDim.x * blockDim.x) {
        const PairHdr h = hdr[i];
        double qmax = 0.0;
        for (int a = 0; a < na; ++a)
            for (int c = 0; c < nb; ++c) qmax = fmax(qmax, sqrt(fabs(Q[(size_t)(h.bfA + a) * N + h.bfB + c])));
        Qs[i] = qmax;
    }
}

// {dP block, sqrt(Q) block} of every pair of a class in the component order and orientation the block digestion reads
// them (make_geom: the higher-indexed shell is the row), packed so that a thread finds them in one or two sectors
__global__ void pack_pair_blocks_kernel(const PairHdr *hdr, int npairs, int ta, int tb, int N, const double *P, const double *SQ,
                                        double *PQ)
{
    const int na = ncomp(ta), nb = ncomp(tb), nab = na * nb;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < (long long)npairs * nab; x += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(x / nab), ab = (int)(x % nab), a = ab / nb, c = ab % nb;
        const int bfA = hdr[i].bfA, bfB = hdr[i].bfB;
        const bool aHi = bfA >= bfB;
        const long long o = aHi ? (long long)(bfA + a) * N + bfB + c : (long long)(bfB + c) * N + bfA + a;
        PQ[(size_t)i * 2 * nab + ab] = P[o];
        PQ[(size_t)i * 2 * nab + nab + ab] = SQ[o];
    }
}

__global__ void dabs_kernel(const double *re, const double *im, size_t n, double *out)
{
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x)
        out[x] = im ? hypot(re[x], im[x]) : fabs(re[x]);
}

// shell-block maxima of |dP| and the global maximum
__global__ void dshell_kernel(const double *Dabs, int N, const int *bf0, const int *nf, int nshell, double *DS,
                              unsigned long long *dglob)
{
    const int total = nshell * nshell;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
        const int A = x / nshell, B = x % nshell;
        // symmetric in (A, B) also for a non-Hermitian dP: the screen may then read row C of the table where it needs
        // column C (a superset test either way; the digestion applies the reference's per-function criterion exactly)
        double m = 0.0;
        for (int a = 0; a < nf[A]; ++a)
            for (int c = 0; c < nf[B]; ++c)
                m = fmax(m, fmax(Dabs[(size_t)(bf0[A] + a) * N + bf0[B] + c], Dabs[(size_t)(bf0[B] + c) * N + bf0[A] + a]));
        DS[x] = m;
        atomicMax(dglob, (unsigned long long)__double_as_longlong(m));
    }
}

// maxima of the pair bounds over chunks of 256 consecutive pairs (one warp's columns in the screening kernel)
__global__ void qs_chunk_max_kernel(const double *Qs, int npairs, double *Qmax)
{
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (chunk * 256 >= npairs) return;
    double m = 0.0;
    for (int i = chunk * 256 + lane; i < min(npairs, chunk * 256 + 256); i += 32) m = fmax(m, Qs[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) Qmax[chunk] = m;
}

// Shell-level screen -> compact quartet lists.
// Rows are KET pairs j in [row0,row1) (j % nshards == shard), columns are BRA pairs i (i >= j when the two classes
// coincide).  One block handles one row x 1024 consecutive columns (two-phase, see the kernel), block-wide prefix sum,
// ONE atomic per list and tile to reserve list space.  Entries of a row are written in column order, so consecutive
// list entries share the ket pair (warp-uniform in the ERI kernels: the inner primitive loop and the J_cd reduction run
// on broadcast data) and walk the bra pairs.  A direct build gets one entry per SLICE of <= BRA_SLICE bra primitive
// pairs (slice id in the top byte of the bra index; the primitives of a pair are sorted by exponent, tight first) and
// THREE lists per class pair:
//   far   block-digestible entries all of whose primitive quartets are on the asymptotic Boys branch
//         (alpha_min d_min^2 >= T_max(L) from the bounding spheres of the slice's and the ket pair's product centres),
//   near  the other block-digestible entries,
//   slow  diagonal-type quartets / complex densities / deterministic mode (per-function digestion).
struct ScreenArgs {
    const double *Qs_bra, *Qs_ket, *Qmax_bra;
    const int2 *sh_bra, *sh_ket;
    const int *K_bra, *K_ket;
    const int *Kref_bra, *Kref_ket;        // statistics in units of the reference's shell quartets: primitive pairs | members << 24
    const int *sbase_bra;                  // first slice record of a bra pair
    const double4 *sgeo_bra, *geo_ket;     // bounding spheres of the product centres: per bra slice / per ket pair (or nullptr)
    const double *spmin_bra, *pmin_ket;    // smallest total exponent: per bra slice / per ket pair
    double tmax;                           // T_max(L) of the class pair (+ margin)
    int nbra, row0, row1, same_class, shard, nshards, nshell, all_pass;
    int early;                     // warp-level early exit on the chunk maxima of the bra bounds
    int split;                     // direct build: one entry per slice, classified into the far / near / slow lists
    int force_slow;                // complex density / deterministic mode: everything goes to the slow list
    int ds_cache;                  // the two density-bound rows of the ket pair's shells are staged in shared memory (2 * nshell floats)
    int want_stats;                // keep the statistics counters (candidates, quartets, primitive quartets): four more
                                   // same-address atomics per tile, only paid when the caller asked for mmdb_fock_stats
    const int *bf0;                // first function index per shell
    long long cap;                 // capacity of list_near (the slow list starts at list_near[cap-1] and grows downwards)
    const double *DS;
    const unsigned long long *dglob;
    double tol;
    uint2 *list_far, *list_near;
    unsigned long long *ctr;       // see CTR_* below
};
enum { CTR_NEAR = 0, CTR_PRIMQ = 1, CTR_CAND = 2, CTR_SLOW = 3, CTR_NQUART = 4, CTR_FAR = 5, CTR_EXECPQ = 6, CTR_WORK = 7, CTR_PER_LAUNCH = 8 };

constexpr int SCR_THREADS = 256;
#ifndef MMDB_SCR_CPT
#define MMDB_SCR_CPT 4
#endif
constexpr int SCR_CPT = MMDB_SCR_CPT;
constexpr int SCR_TILE = SCR_THREADS * SCR_CPT;
constexpr int SCR_SEG = 4;         // consecutive tiles of one row per work item
constexpr int SCR_MAXSL = 8;       // slices per pair the classification masks can hold (pairs with more go near/slow whole)

__global__ void __launch_bounds__(SCR_THREADS) screen_kernel(const ScreenArgs s)
{
    __shared__ unsigned long long s_wcnt[SCR_THREADS / 32];
    __shared__ unsigned long long s_wk[SCR_THREADS / 32];
    __shared__ unsigned long long s_wx[SCR_THREADS / 32];
    __shared__ unsigned s_wcand[SCR_THREADS / 32];
    __shared__ unsigned long long s_base[3];
    __shared__ unsigned short s_slot[SCR_THREADS / 32][SCR_CPT * 32];    // compacted phase-1 survivors per warp
    extern __shared__ float s_ds[];         // [2][nshell]: rows cd.x and cd.y of the shell-block density bounds, rounded up
    const int ntile = (s.nbra + SCR_TILE - 1) / SCR_TILE;
    // work item of a block = SCR_SEG consecutive tiles of one row (the staged density rows serve all of them)
    const int nseg = (ntile + SCR_SEG - 1) / SCR_SEG;
    const long long nblk = (long long)(s.row1 - s.row0) * nseg * SCR_SEG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double dg4 = 0.0;
    if (!s.all_pass) dg4 = 4.0 * __longlong_as_double((long long)*s.dglob);
    int staged_j = -1;
    for (long long blk = (long long)blockIdx.x * SCR_SEG; blk < nblk; blk = ((blk % SCR_SEG) == SCR_SEG - 1) ? blk + 1 + (long long)(gridDim.x - 1) * SCR_SEG : blk + 1) {
        const int j = s.row0 + (int)(blk / (nseg * SCR_SEG));
        const int tile = (int)(blk % (nseg * SCR_SEG));
        if (tile >= ntile) continue;
        const int c0 = tile * SCR_TILE;
        if (s.nshards > 1 && (j % s.nshards) != s.shard) continue;
        const int cstart = s.same_class ? j : 0;
        if (c0 + SCR_TILE <= cstart) continue;
        const double qj = s.Qs_ket[j];
        if (!s.all_pass && s.early) {
            // dead tile (block-uniform test on the chunk maxima): nothing can pass, so skip the scans, barriers
            // and atomics altogether — only the candidate count is kept for the statistics
            bool tile_live = false;
            for (int ch = c0 >> 8; ch <= ((c0 + SCR_TILE - 1) >> 8); ++ch)
                if (ch * 256 < s.nbra && !(s.Qmax_bra[ch] * qj * dg4 < s.tol)) tile_live = true;
            if (!tile_live) {
                if (threadIdx.x == 0 && s.want_stats) {
                    const int lo = max(c0, cstart), hi = min(c0 + SCR_TILE, s.nbra);
                    if (hi > lo) atomicAdd(s.ctr + CTR_CAND, (unsigned long long)(hi - lo));
                }
                continue;
            }
        }
        const int2 cd = s.sh_ket[j];
        const unsigned krj = (unsigned)s.Kref_ket[j];
        const unsigned long long kj = (unsigned long long)(krj & 0xffffffu);
        const unsigned mj = krj >> 24;
        const unsigned long long kxj = (unsigned long long)s.K_ket[j];
        // The four cross terms of the density test read DS[A or B][C or D]: rows C and D of the (symmetric) table serve
        // every column of the tile, so they are staged once — the test then costs one gather from L2 (DS[A][B]) and
        // four from shared memory instead of five from L2 (ncu: long-scoreboard stalls on exactly these loads).
        if (s.ds_cache && !s.all_pass && staged_j != j) {
            staged_j = j;
            const int ns = s.nshell;
            for (int x = threadIdx.x; x < ns; x += SCR_THREADS) {
                s_ds[x] = __double2float_ru(s.DS[(size_t)cd.x * ns + x]);
                s_ds[ns + x] = __double2float_ru(s.DS[(size_t)cd.y * ns + x]);
            }
            __syncthreads();
        }
        // Two phases per warp (128 consecutive columns).  Phase 1: the cheap density-independent bound on all columns,
        // lane-strided (coalesced), survivors compacted into a per-warp slot array in column order.  Phase 2: the
        // six-block density test, slicing and list classification on the compacted survivors only.
        const int wbase = c0 + warp * (SCR_CPT * 32);
        unsigned bits = 0, sbits = 0, ncand = 0, nquart = 0;
        unsigned nsl[SCR_CPT], fmask[SCR_CPT];       // slices of the pair; which of them are far-field
        int col[SCR_CPT];
        unsigned long long kk = 0, kx = 0;      // primitive quartets: in the reference's units / actually evaluated
        const int hiK = s.split ? max(s.bf0[cd.x], s.bf0[cd.y]) : 0;
        double4 gk = make_double4(0.0, 0.0, 0.0, 0.0);
        double qmin = 0.0;
        if (s.split && s.geo_ket) { gk = s.geo_ket[j]; qmin = s.pmin_ket[j]; }
        bool warp_live = s.all_pass || !s.early;
        {
            const int first = wbase, last = first + SCR_CPT * 32 - 1;
            for (int wchunk = first >> 8; wchunk <= (last >> 8); ++wchunk)
                if (wchunk * 256 < s.nbra && !(s.Qmax_bra[wchunk] * qj * dg4 < s.tol)) warp_live = true;
        }
        unsigned total = 0;
#pragma unroll
        for (int k = 0; k < SCR_CPT; ++k) {
            const int off = k * 32 + lane;
            const int i = wbase + off;
            const bool cand = (i < s.nbra) && (i >= cstart);
            ncand += cand ? 1u : 0u;            // candidates are counted for the statistics even when the warp exits early
            bool p1 = cand && warp_live;
            if (p1 && !s.all_pass) p1 = !(s.Qs_bra[i] * qj * dg4 < s.tol);
            const unsigned m = __ballot_sync(0xffffffffu, p1);
            if (p1) s_slot[warp][total + __popc(m & ((1u << lane) - 1u))] = (unsigned short)off;
            total += __popc(m);
        }
        __syncwarp();
        const unsigned per = (total + 31u) >> 5;      // survivors per lane (<= SCR_CPT)
        unsigned nent_far = 0, nent_near = 0, nent_slow = 0;
#pragma unroll
        for (int k = 0; k < SCR_CPT; ++k) {
            nsl[k] = 0; fmask[k] = 0;
            col[k] = 0;
            const unsigned n = lane * per + k;
            if ((unsigned)k < per && n < total) {
                const int i = wbase + s_slot[warp][n];
                col[k] = i;
                bool pass = true;
                int2 ab = make_int2(0, 0);
                if (!s.all_pass) {
                    const double qq = s.Qs_bra[i] * qj;
                    ab = s.sh_bra[i];
                    const double *DS = s.DS;
                    const int ns = s.nshell;
                    double dmax = fmax(4.0 * DS[ab.x * ns + ab.y], 4.0 * DS[cd.x * ns + cd.y]);
                    if (s.ds_cache)
                        dmax = fmax(dmax, (double)fmaxf(fmaxf(s_ds[ab.x], s_ds[ns + ab.x]), fmaxf(s_ds[ab.y], s_ds[ns + ab.y])));
                    else
                        dmax = fmax(dmax, fmax(fmax(DS[ab.x * ns + cd.x], DS[ab.x * ns + cd.y]),
                                               fmax(DS[ab.y * ns + cd.x], DS[ab.y * ns + cd.y])));
                    pass = !(qq * dmax < s.tol);
                }
                if (pass) {
                    bits |= 1u << k;
                    const int kb = s.K_bra[i];
                    const unsigned krb = (unsigned)s.Kref_bra[i];
                    kk += (unsigned long long)(krb & 0xffffffu) * kj;
                    nquart += (krb >> 24) * mj;
                    kx += (unsigned long long)kb * kxj;
                    // one list entry per slice of BRA_SLICE bra primitive pairs (direct builds only)
                    nsl[k] = s.split ? (unsigned)((kb + BRA_SLICE - 1) / BRA_SLICE) : 1u;
                    bool slow = false;
                    if (s.split)   // block digestion needs different leading shells in bra and ket (kernels_a.cuh)
                        slow = s.force_slow || max(s.bf0[ab.x], s.bf0[ab.y]) == hiK;
                    if (slow) {
                        sbits |= 1u << k;
                        nent_slow += nsl[k];
                    } else {
                        unsigned fm = 0;
                        if (s.split && s.sgeo_bra && nsl[k] <= (unsigned)SCR_MAXSL) {
                            // far field per slice: every product centre of the slice lies in the sphere gb, every one of
                            // the ket pair in gk, so |PQ| >= d for every primitive quartet; alpha >= pmin qmin / (pmin + qmin)
                            const int sb = s.sbase_bra[i];
                            for (unsigned sl = 0; sl < nsl[k]; ++sl) {
                                const double4 gb = s.sgeo_bra[sb + sl];
                                const double dx = gb.x - gk.x, dy = gb.y - gk.y, dz = gb.z - gk.z;
                                const double d = sqrt(dx * dx + dy * dy + dz * dz) - gb.w - gk.w;
                                const double pm = s.spmin_bra[sb + sl];
                                if (d > 0.0 && (pm * qmin) * (d * d) >= s.tmax * (pm + qmin)) fm |= 1u << sl;
                            }
                        }
                        fmask[k] = fm;
                        nent_far += __popc(fm);
                        nent_near += nsl[k] - __popc(fm);
                    }
                }
            }
        }
        // block-wide exclusive scan of the entry counts: far | near << 21 | slow << 42
        unsigned long long incl = (unsigned long long)nent_far | ((unsigned long long)nent_near << 21) | ((unsigned long long)nent_slow << 42);
        const unsigned long long mine = incl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        unsigned long long ks = kk, xs = kx;
        unsigned cs = ncand | (nquart << 16);      // candidates (<= SCR_CPT per thread) | shell quartets (<= 16 SCR_CPT per thread)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ks += __shfl_xor_sync(0xffffffffu, ks, o);
            xs += __shfl_xor_sync(0xffffffffu, xs, o);
            cs += __shfl_xor_sync(0xffffffffu, cs, o);
        }
        if (lane == 31) s_wcnt[warp] = incl;
        if (lane == 0) { s_wk[warp] = ks; s_wx[warp] = xs; s_wcand[warp] = cs; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0, tk = 0, tx = 0;
            unsigned tc = 0, tq = 0;
            for (int w = 0; w < SCR_THREADS / 32; ++w) {
                const unsigned long long c = s_wcnt[w];
                s_wcnt[w] = tot;
                tot += c;
                tk += s_wk[w];
                tx += s_wx[w];
                tc += s_wcand[w] & 0xffffu;
                tq += s_wcand[w] >> 16;
            }
            if (tq && s.want_stats) atomicAdd(s.ctr + CTR_NQUART, (unsigned long long)tq);
            const unsigned tf = (unsigned)(tot & 0x1fffffull), tn = (unsigned)((tot >> 21) & 0x1fffffull), tsl = (unsigned)(tot >> 42);
            s_base[0] = tf ? atomicAdd(s.ctr + CTR_FAR, (unsigned long long)tf) : 0ull;
            s_base[1] = tn ? atomicAdd(s.ctr + CTR_NEAR, (unsigned long long)tn) : 0ull;
            s_base[2] = tsl ? atomicAdd(s.ctr + CTR_SLOW, (unsigned long long)tsl) : 0ull;
            if (s.want_stats) {
                if (tk) atomicAdd(s.ctr + CTR_PRIMQ, tk);
                if (tx) atomicAdd(s.ctr + CTR_EXECPQ, tx);
                if (tc) atomicAdd(s.ctr + CTR_CAND, (unsigned long long)tc);
            }
        }
        __syncthreads();
        if (bits) {
            const unsigned long long excl = s_wcnt[warp] + (incl - mine);
            long long fpos = (long long)(s_base[0] + (excl & 0x1fffffull));
            long long npos = (long long)(s_base[1] + ((excl >> 21) & 0x1fffffull));
            long long spos = s.cap - 1 - (long long)(s_base[2] + (excl >> 42));
#pragma unroll
            for (int k = 0; k < SCR_CPT; ++k)
                if (bits & (1u << k)) {
                    for (unsigned sl = 0; sl < nsl[k]; ++sl) {
                        const uint2 ent = make_uint2((unsigned)col[k] | (sl << SLICE_SHIFT), (unsigned)j);
                        if (sbits & (1u << k)) s.list_near[spos--] = ent;
                        else if (fmask[k] & (1u << sl)) s.list_far[fpos++] = ent;
                        else s.list_near[npos++] = ent;
                    }
                }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Warp-autonomous variant of the screen for direct builds (no far lists): no block-wide barrier, no per-tile atomics.
// ncu on screen_kernel: barrier stalls were 8 of its 13 cycles per issued instruction — every 1024-column tile ends in
// a block-wide scan, a serial section of thread 0 and three to seven same-address atomics with a return value.  Here a
// warp draws chunks of consecutive (ket row, 1024 bra columns) items from a work counter, runs the same two-phase test
// on 128 columns at a time, and collects its entries in a private shared-memory buffer of SW_BUF entries that it
// flushes with ONE atomic (space reservation) and a coalesced copy.  Consecutive items share the ket row, so a flushed
// block is almost always one row: the ERI kernels still see warp-uniform ket pairs (a block boundary inside a warp of
// theirs costs that one warp the uniform fast path, about one warp in twenty).  Slow-list entries are rare and go out
// with a warp-aggregated atomic as they appear.
// ------------------------------------------------------------------------------------------
constexpr int SW_WARPS = 8;
constexpr int SW_COLS = 32 * SCR_CPT;     // columns per warp step
constexpr int SW_ITEM = 1024;             // columns per work item
constexpr int SW_CHUNK = 4;               // consecutive items per draw from the work counter
constexpr int SW_BUF = 256;                // entries per warp buffer (a step that produces more goes straight to the list)

__global__ void __launch_bounds__(SW_WARPS * 32) screen_warp_kernel(const ScreenArgs s)
{
    extern __shared__ uint2 sw_buf[];                         // [SW_WARPS][SW_BUF]
    __shared__ unsigned short s_slot[SW_WARPS][SW_COLS];      // compacted phase-1 survivors per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 *buf = sw_buf + warp * SW_BUF;
    unsigned fill = 0;
    const int ntile = (s.nbra + SW_ITEM - 1) / SW_ITEM;
    // rows of this shard: j = jfirst + r * nshards
    int jfirst = s.row0;
    while (s.nshards > 1 && (jfirst % s.nshards) != s.shard) ++jfirst;
    const int nrows = jfirst < s.row1 ? (s.row1 - jfirst + s.nshards - 1) / s.nshards : 0;
    const long long nitems = (long long)nrows * ntile;
    double dg4 = 0.0;
    if (!s.all_pass) dg4 = 4.0 * __longlong_as_double((long long)*s.dglob);
    unsigned long long st_kk = 0, st_kx = 0, st_cand = 0, st_quart = 0;      // statistics, flushed once per warp

    auto flush = [&]() {
        if (fill == 0) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(s.ctr + CTR_NEAR, (unsigned long long)fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (unsigned x = lane; x < fill; x += 32) s.list_near[base + x] = buf[x];
        __syncwarp();
        fill = 0;
    };

    for (;;) {
        long long it0 = 0;
        if (lane == 0) it0 = (long long)atomicAdd(s.ctr + CTR_WORK, (unsigned long long)SW_CHUNK);
        it0 = __shfl_sync(0xffffffffu, it0, 0);
        if (it0 >= nitems) break;
        const long long it1 = min(it0 + SW_CHUNK, nitems);
        for (long long it = it0; it < it1; ++it) {
            const int j = jfirst + (int)(it / ntile) * s.nshards;
            const int c0 = (int)(it % ntile) * SW_ITEM;
            const int cstart = s.same_class ? j : 0;
            if (c0 + SW_ITEM <= cstart) continue;
            const double qj = s.Qs_ket[j];
            const int2 cd = s.sh_ket[j];
            const unsigned krj = (unsigned)s.Kref_ket[j];
            const unsigned long long kj = (unsigned long long)(krj & 0xffffffu);
            const unsigned mj = krj >> 24;
            const unsigned long long kxj = (unsigned long long)s.K_ket[j];
            const int hiK = max(s.bf0[cd.x], s.bf0[cd.y]);
            for (int wbase = c0; wbase < min(c0 + SW_ITEM, s.nbra); wbase += SW_COLS) {
                if (wbase + SW_COLS <= cstart) continue;
                // columns [wbase, wbase + 128) lie in one 256-pair chunk of the bra bounds
                const bool live = s.all_pass || !s.early || !(s.Qmax_bra[wbase >> 8] * qj * dg4 < s.tol);
                unsigned total = 0;
#pragma unroll
                for (int k = 0; k < SCR_CPT; ++k) {
                    const int off = k * 32 + lane;
                    const int i = wbase + off;
                    const bool cand = (i < s.nbra) && (i >= cstart);
                    st_cand += cand ? 1u : 0u;
                    bool p1 = cand && live;
                    if (p1 && !s.all_pass) p1 = !(s.Qs_bra[i] * qj * dg4 < s.tol);
                    const unsigned m = __ballot_sync(0xffffffffu, p1);
                    if (p1) s_slot[warp][total + __popc(m & ((1u << lane) - 1u))] = (unsigned short)off;
                    total += __popc(m);
                }
                if (total == 0) continue;
                __syncwarp();
                const unsigned per = (total + 31u) >> 5;      // survivors per lane (<= SCR_CPT)
                unsigned nsl[SCR_CPT], slowm = 0, n_near = 0, n_slow = 0;
                int col[SCR_CPT];
#pragma unroll
                for (int k = 0; k < SCR_CPT; ++k) {
                    nsl[k] = 0;
                    col[k] = 0;
                    const unsigned n = lane * per + k;
                    if ((unsigned)k < per && n < total) {
                        const int i = wbase + s_slot[warp][n];
                        bool pass = true;
                        int2 ab = s.sh_bra[i];
                        if (!s.all_pass) {
                            const double qq = s.Qs_bra[i] * qj;
                            const double *DS = s.DS;
                            const int ns = s.nshell;
                            double dmax = fmax(4.0 * DS[ab.x * ns + ab.y], 4.0 * DS[cd.x * ns + cd.y]);
                            dmax = fmax(dmax, fmax(fmax(DS[ab.x * ns + cd.x], DS[ab.x * ns + cd.y]),
                                                   fmax(DS[ab.y * ns + cd.x], DS[ab.y * ns + cd.y])));
                            pass = !(qq * dmax < s.tol);
                        }
                        if (pass) {
                            col[k] = i;
                            const int kb = s.K_bra[i];
                            const unsigned krb = (unsigned)s.Kref_bra[i];
                            st_kk += (unsigned long long)(krb & 0xffffffu) * kj;
                            st_quart += (krb >> 24) * mj;
                            st_kx += (unsigned long long)kb * kxj;
                            nsl[k] = (unsigned)((kb + BRA_SLICE - 1) / BRA_SLICE);
                            // block digestion needs different leading shells in bra and ket (kernels_a.cuh)
                            const bool slow = s.force_slow || max(s.bf0[ab.x], s.bf0[ab.y]) == hiK;
                            if (slow) { slowm |= 1u << k; n_slow += nsl[k]; }
                            else n_near += nsl[k];
                        }
                    }
                }
                __syncwarp();          // s_slot is rewritten by the next step
                // warp-wide exclusive scan of the near-entry counts
                unsigned incl = n_near;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const unsigned tot_near = __shfl_sync(0xffffffffu, incl, 31);
                if (fill + tot_near > (unsigned)SW_BUF) flush();
                if (tot_near <= (unsigned)SW_BUF) {
                    unsigned pos = fill + incl - n_near;
#pragma unroll
                    for (int k = 0; k < SCR_CPT; ++k)
                        if (nsl[k] && !(slowm & (1u << k)))
                            for (unsigned sl = 0; sl < nsl[k]; ++sl) buf[pos++] = make_uint2((unsigned)col[k] | (sl << SLICE_SHIFT), (unsigned)j);
                    fill += tot_near;
                } else {
                    // pairs with more than SCR_MAXSL slices (contractions deeper than 64 primitive pairs): straight to the list
                    unsigned long long nbase = 0;
                    if (lane == 0) nbase = atomicAdd(s.ctr + CTR_NEAR, (unsigned long long)tot_near);
                    nbase = __shfl_sync(0xffffffffu, nbase, 0);
                    unsigned long long pos = nbase + incl - n_near;
#pragma unroll
                    for (int k = 0; k < SCR_CPT; ++k)
                        if (nsl[k] && !(slowm & (1u << k)))
                            for (unsigned sl = 0; sl < nsl[k]; ++sl) s.list_near[pos++] = make_uint2((unsigned)col[k] | (sl << SLICE_SHIFT), (unsigned)j);
                }
                // slow entries: rare (diagonal-type quartets) unless the whole build is forced onto the slow list
                if (__any_sync(0xffffffffu, n_slow != 0)) {
                    unsigned sincl = n_slow;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned v = __shfl_up_sync(0xffffffffu, sincl, o);
                        if (lane >= o) sincl += v;
                    }
                    const unsigned tot_slow = __shfl_sync(0xffffffffu, sincl, 31);
                    unsigned long long sbase = 0;
                    if (lane == 0) sbase = atomicAdd(s.ctr + CTR_SLOW, (unsigned long long)tot_slow);
                    sbase = __shfl_sync(0xffffffffu, sbase, 0);
                    long long spos = s.cap - 1 - (long long)(sbase + sincl - n_slow);
#pragma unroll
                    for (int k = 0; k < SCR_CPT; ++k)
                        if (nsl[k] && (slowm & (1u << k)))
                            for (unsigned sl = 0; sl < nsl[k]; ++sl) s.list_near[spos--] = make_uint2((unsigned)col[k] | (sl << SLICE_SHIFT), (unsigned)j);
                }
            }
        }
    }
    flush();
    if (s.want_stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_kk += __shfl_xor_sync(0xffffffffu, st_kk, o);
            st_kx += __shfl_xor_sync(0xffffffffu, st_kx, o);
            st_cand += __shfl_xor_sync(0xffffffffu, st_cand, o);
            st_quart += __shfl_xor_sync(0xffffffffu, st_quart, o);
        }
        if (lane == 0) {
            if (st_kk) atomicAdd(s.ctr + CTR_PRIMQ, st_kk);
            if (st_kx) atomicAdd(s.ctr + CTR_EXECPQ, st_kx);
            if (st_cand) atomicAdd(s.ctr + CTR_CAND, st_cand);
            if (st_quart) atomicAdd(s.ctr + CTR_NQUART, st_quart);
        }
    }
}

// scratch[entry][nfn] -> dense TwoE with all 8 images (cython/twoe.pyx:23-30)
__global__ void scatter_dense_kernel(const uint2 *list, const unsigned long long *count, const PairHdr *braH,
                                     const PairHdr *ketH, int la, int lb, int lc, int ld, int same_class,
                                     const double *scratch, int N, double *T)
{
    const int nb = ncart(lb), nc = ncart(lc), nd = ncart(ld);
    const int nfn = ncart(la) * nb * nc * nd;
    const unsigned long long total = (*count) * (unsigned long long)nfn;
    const size_t n = (size_t)N;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = w / nfn;
        int f = (int)(w % nfn);
        const int d = f % nd; f /= nd;
        const int c = f % nc; f /= nc;
        const int bb = f % nb;
        const int a = f / nb;
        const uint2 ij = list[e];
        const PairHdr bh = braH[ij.x], kh = ketH[ij.y];
        const size_t i = bh.bfA + a, j = bh.bfB + bb, k = kh.bfA + c, l = kh.bfB + d;
        // duplicates inside diagonal blocks ((a,b)/(b,a) of one shell, (ab|cd)/(cd|ab) of one pair) are
        // evaluated by separate threads and may differ in the last bit: let exactly one of them write all
        // eight images so the tensor is bit-symmetric like the reference's
        if (bh.shA == bh.shB && i < j) continue;
        if (kh.shA == kh.shB && k < l) continue;
        if (same_class && ij.x == ij.y) {
            const size_t hi1 = i > j ? i : j, lo1 = i > j ? j : i, hi2 = k > l ? k : l, lo2 = k > l ? l : k;
            if (hi1 * (hi1 + 1) / 2 + lo1 < hi2 * (hi2 + 1) / 2 + lo2) continue;
        }
        const double v = scratch[w];
        T[((i * n + j) * n + k) * n + l] = v;
        T[((k * n + l) * n + i) * n + j] = v;
        T[((j * n + i) * n + l) * n + k] = v;
        T[((l * n + k) * n + j) * n + i] = v;
        T[((j * n + i) * n + k) * n + l] = v;
        T[((l * n + k) * n + i) * n + j] = v;
        T[((i * n + j) * n + l) * n + k] = v;
        T[((k * n + l) * n + j) * n + i] = v;
    }
}

// ------------------------------------------------------------------------------------------
// ERI entry points
// ------------------------------------------------------------------------------------------
extern "C" int mmdb_eri_shell_quartets(mmdb_basis *b, int pc_bra, int pc_ket, int64_t n, const int32_t *bra_idx_dev,
                                       const int32_t *ket_idx_dev, double *out_dev, int impl, void *stream)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    if (pc_bra < pc_ket || pc_bra >= MMDB_NCLASS_PAIR || pc_ket < 0)
        return fail(MMDB_ERR_INVALID, "mmdb_eri_shell_quartets: need pc_bra >= pc_ket");
    if (n == 0) return MMDB_OK;
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    CHK(ensure_list(b, (size_t)n));
    zip_list_kernel<<<std::min<int64_t>((n + 255) / 256, 65535), 256, 0, st>>>(bra_idx_dev, ket_idx_dev, n, b->list_dev);
    PairClass &B = b->pc[pc_bra], &K = b->pc[pc_ket];
    EriArgs a;
    std::memset(&a, 0, sizeof(a));
    a.braH = B.hdr_dev; a.braP = B.prim_dev; a.braS = B.prim_soa_dev; a.braRow = B.prim_row_dev; a.braN = B.nprimpairs; a.ketH = K.hdr_dev; a.ketP = K.prim_dev;
    a.list = b->list_dev; a.list_step = 1; a.count_dev = nullptr; a.n = (unsigned long long)n; a.out = out_dev;
    a.same_class = (pc_bra == pc_ket);
    return launch_eri(b, B.la, B.lb, K.la, K.lb, a, LK_STORE, impl, st);
}

// the grouped pair classes take their bounds from the function-level table the plain classes have just filled
static int gc_refresh_bounds(mmdb_basis *b, cudaStream_t st)
{
    if (!b->have_gc) return MMDB_OK;
    for (int c = 0; c < MMDB_NCLASS_GC; ++c) {
        PairClass &P = b->pcg[c];
        if (P.npairs == 0) continue;
        gc_bounds_kernel<<<(P.npairs + 127) / 128, 128, 0, st>>>(P.hdr_dev, P.npairs, P.la, P.lb, b->nbf, b->Q_dev, P.Qs_dev);
        qs_chunk_max_kernel<<<((P.npairs + 255) / 256 + 3) / 4, 128, 0, st>>>(P.Qs_dev, P.npairs, P.Qmax_dev);
    }
    CU(cudaGetLastError());
    return MMDB_OK;
}

extern "C" int mmdb_schwarz(mmdb_basis *b, double *Q_dev, void *stream)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N2 = (size_t)b->nbf * b->nbf;
    CU(cudaMemsetAsync(b->Q_dev, 0, N2 * sizeof(double), st));
    CU(cudaMemsetAsync(b->SQ_dev, 0, N2 * sizeof(double), st));
    for (int c = 0; c < MMDB_NCLASS_PAIR; ++c) {
        PairClass &P = b->pc[c];
        if (P.npairs == 0) continue;
        const int nab = ncart(P.la) * ncart(P.lb);
        CHK(ensure_list(b, (size_t)P.npairs));
        CHK(ensure_scratch(b, (size_t)P.npairs * nab * nab));
        diag_list_kernel<<<(P.npairs + 255) / 256, 256, 0, st>>>(P.npairs, b->list_dev);
        EriArgs a;
        std::memset(&a, 0, sizeof(a));
        a.braH = P.hdr_dev; a.braP = P.prim_dev; a.braS = P.prim_soa_dev; a.braRow = P.prim_row_dev; a.braN = P.nprimpairs; a.ketH = P.hdr_dev; a.ketP = P.prim_dev;
        a.list = b->list_dev; a.list_step = 1; a.n = (unsigned long long)P.npairs; a.out = b->scratch_dev; a.same_class = 1;
        CHK(launch_eri(b, P.la, P.lb, P.la, P.lb, a, LK_STORE, 0, st));
        schwarz_extract_kernel<<<(P.npairs + 127) / 128, 128, 0, st>>>(P.hdr_dev, P.npairs, P.la, P.lb, b->scratch_dev,
                                                                       b->nbf, b->Q_dev, b->SQ_dev, P.Qs_dev);
        qs_chunk_max_kernel<<<((P.npairs + 255) / 256 + 3) / 4, 128, 0, st>>>(P.Qs_dev, P.npairs, P.Qmax_dev);
    }
    CU(cudaGetLastError());
    CHK(gc_refresh_bounds(b, st));
    if (Q_dev) CU(cudaMemcpyAsync(Q_dev, b->Q_dev, N2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    b->have_schwarz = true;
    return MMDB_OK;
}

static bool far_enabled(int L)
{
    // default: off.  Measured on (H2O)32/cc-pVDZ: the far-field kernels run at 17 TFLOP/s (model) against 7 for the near
    // kernels, but the near list keeps every entry with a single tabulated primitive quartet (44 % of the primitive work,
    // now fully divergent), so the class times barely move ((ps|ss) 12.5 -> 12.0 ms) while the per-slice classification
    // adds 2.8 ms of screening.  MMDB_FAR_MAXL=1..3 switches the lists on up to that total angular momentum.
    const int far_maxl = getenv("MMDB_NO_FAR_LIST") ? -1 : std::min(FAR_MAXL, getenv("MMDB_FAR_MAXL") ? atoi(getenv("MMDB_FAR_MAXL")) : -1);      // read per call: a run can switch
    return L <= far_maxl;
}

static int run_screen(mmdb_basis *b, PairClass &B, PairClass &K, bool same, int row0, int row1, int shard, int nshards,
                      bool all_pass, double tol, int slot, bool split, bool force_slow, long long cap, uint2 *list_far,
                      uint2 *list_near, cudaStream_t st, bool gc = false, bool want_stats = true)
{
    ScreenArgs s;
    std::memset(&s, 0, sizeof(s));
    s.Qs_bra = B.Qs_dev; s.Qs_ket = K.Qs_dev; s.Qmax_bra = B.Qmax_dev; s.sh_bra = B.sh_dev; s.sh_ket = K.sh_dev;
    s.K_bra = B.K_dev; s.K_ket = K.K_dev; s.Kref_bra = B.Kref_dev; s.Kref_ket = K.Kref_dev;
    // The far-field list pays where Boys + R dominate a primitive quartet (L <= 3); above that the extra launch per
    // class pair costs more than the table branch it saves.  MMDB_NO_FAR_LIST / MMDB_FAR_MAXL: A/B switches.
    const int Ltot = am_of(B.la) + am_of(B.lb) + am_of(K.la) + am_of(K.lb);
    if (split && !gc && far_enabled(Ltot)) {
        s.sbase_bra = B.sbase_dev; s.sgeo_bra = B.sgeo_dev; s.spmin_bra = B.spmin_dev; s.geo_ket = K.geo_dev; s.pmin_ket = K.pmin_dev;
    }
    s.tmax = (double)boys_tmax_i(Ltot) + 0.5;     // margin: rounding of the bounding-sphere distances
    s.nbra = B.npairs; s.row0 = row0; s.row1 = row1; s.same_class = same ? 1 : 0;
    s.shard = shard; s.nshards = nshards; s.nshell = gc ? b->nshellg : b->nshell; s.all_pass = all_pass ? 1 : 0;
    s.DS = gc ? b->DSg_dev : b->DS_dev; s.dglob = b->dglob_dev; s.tol = tol; s.list_far = list_far; s.list_near = list_near;
    s.ctr = b->ctr_dev + CTR_PER_LAUNCH * slot;
    s.early = getenv("MMDB_SCREEN_NO_EARLY_EXIT") ? 0 : 1;
    s.want_stats = want_stats ? 1 : 0;
    s.split = split ? 1 : 0; s.force_slow = force_slow ? 1 : 0; s.bf0 = gc ? b->shg_bf0_dev : b->sh_bf0_dev; s.cap = cap;
    const long long ntile = (B.npairs + SCR_TILE - 1) / SCR_TILE;
    const long long nblk = (long long)(row1 - row0) * ((ntile + SCR_SEG - 1) / SCR_SEG);      // work items: SCR_SEG tiles of one row
    const int grid = (int)std::min<long long>(nblk, (long long)b->nsm * 32);
    // MMDB_SCREEN_WARPS=1: the warp-autonomous kernel (measured slower, 14.1 against 10.6 ms of screening per build: it
    // removes the barriers but not the latency of the density gathers, and its buffers cost occupancy)
    if (split && s.sgeo_bra == nullptr && getenv("MMDB_SCREEN_WARPS")) {
        static bool attr_set = false;
        const size_t smem = sizeof(uint2) * SW_WARPS * SW_BUF;
        if (!attr_set) {
            cudaFuncSetAttribute(screen_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set = true;
        }
        const long long items = ((long long)(row1 - row0) / nshards + 1) * ((B.npairs + SW_ITEM - 1) / SW_ITEM);
        const int gridw = (int)std::min<long long>((items + SW_CHUNK * SW_WARPS - 1) / (SW_CHUNK * SW_WARPS), (long long)b->nsm * 4);
        if (gridw > 0) screen_warp_kernel<<<gridw, SW_WARPS * 32, smem, st>>>(s);
    } else if (grid > 0) {
        const size_t smem_ds = (s.nshell <= 4096 && !getenv("MMDB_SCREEN_NO_DS_CACHE")) ? sizeof(float) * 2 * (size_t)s.nshell : 0;
        s.ds_cache = smem_ds ? 1 : 0;
        screen_kernel<<<grid, SCR_THREADS, smem_ds, st>>>(s);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MMDB_ERR_CUDA, std::string("screen kernel: ") + cudaGetErrorString(e));
    return MMDB_OK;
}

static const size_t LIST_CAP = (size_t)1 << 27;      // entries per screening chunk (1 GiB of uint2)
static const size_t SCRATCH_CAP = (size_t)1 << 27;   // doubles (1 GiB)
// measured on (H2O)32/cc-pVDZ: (ps|dp) 6.3 -> 5.5 ms, (pp|ds) 3.9 -> 2.7 ms; the other six candidates lose ((pp|dp) 4.8 -> 6.0,
// (ds|dp) 3.0 -> 4.9, (ps|dd) 1.6 -> 3.7, (pp|dd) 1.35 -> 1.7, (ds|dd) 1.0 -> 1.8, (dp|dd) 1.3 -> 1.8: more ket-component chunks)
static const char *SWAPPED = "41,32";
static const size_t AUX_MAX_CANDIDATES = (size_t)6 << 20;   // class pairs up to this many candidates per shard go to the aux stream

extern "C" int mmdb_eri_dense(mmdb_basis *b, double *TwoE_dev, void *stream)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = b->nbf;
    CU(cudaMemsetAsync(TwoE_dev, 0, N * N * N * N * sizeof(double), st));
    CU(cudaMemsetAsync(b->ctr_dev, 0, sizeof(unsigned long long) * b->nctr, st));
    int slot = 0;
    for (int cb = 0; cb < MMDB_NCLASS_PAIR; ++cb)
        for (int ck = 0; ck <= cb; ++ck) {
            PairClass &B = b->pc[cb], &K = b->pc[ck];
            if (B.npairs == 0 || K.npairs == 0) continue;
            const size_t nfn = (size_t)ncart(B.la) * ncart(B.lb) * ncart(K.la) * ncart(K.lb);
            size_t rows_per = std::max<size_t>(1, std::min(LIST_CAP, SCRATCH_CAP / nfn) / (size_t)B.npairs);
            for (int row0 = 0; row0 < K.npairs; row0 += (int)rows_per) {
                const int row1 = (int)std::min<size_t>(K.npairs, row0 + rows_per);
                const size_t cap = (size_t)(row1 - row0) * B.npairs;
                CHK(ensure_list(b, cap));
                CHK(ensure_scratch(b, cap * nfn));
                if ((slot + 1) * CTR_PER_LAUNCH > b->nctr) return fail(MMDB_ERR_NOMEM, "mmdb_eri_dense: counter slots exhausted");
                CHK(run_screen(b, B, K, cb == ck, row0, row1, 0, 1, true, -1.0, slot, false, false, (long long)cap, b->list_dev, b->list_dev, st));
                EriArgs a;
                std::memset(&a, 0, sizeof(a));
                a.braH = B.hdr_dev; a.braP = B.prim_dev; a.braS = B.prim_soa_dev; a.braRow = B.prim_row_dev; a.braN = B.nprimpairs; a.ketH = K.hdr_dev; a.ketP = K.prim_dev;
                a.list = b->list_dev; a.list_step = 1; a.count_dev = b->ctr_dev + CTR_PER_LAUNCH * slot; a.out = b->scratch_dev;
                a.same_class = (cb == ck);
                CHK(launch_eri(b, B.la, B.lb, K.la, K.lb, a, LK_STORE, 0, st));
                scatter_dense_kernel<<<b->nsm * 16, 256, 0, st>>>(b->list_dev, b->ctr_dev + CTR_PER_LAUNCH * slot, B.hdr_dev, K.hdr_dev,
                                                                  B.la, B.lb, K.la, K.lb, cb == ck ? 1 : 0, b->scratch_dev, (int)N, TwoE_dev);
                ++slot;
            }
        }
    CU(cudaGetLastError());
    return MMDB_OK;
}

// ------------------------------------------------------------------------------------------
// direct Fock build
// ------------------------------------------------------------------------------------------
extern "C" int mmdb_fock_direct(mmdb_basis *b, const double *dP_re_dev, const double *dP_im_dev, double tol,
                                double *G_re_dev, double *G_im_dev, int shard, int nshards, int flags,
                                mmdb_fock_stats *stats, void *stream)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    if (!b->have_schwarz) return fail(MMDB_ERR_INVALID, "mmdb_fock_direct: call mmdb_schwarz first");
    if (nshards < 1 || shard < 0 || shard >= nshards) return fail(MMDB_ERR_INVALID, "mmdb_fock_direct: bad shard");
    if (dP_im_dev && !G_im_dev) return fail(MMDB_ERR_INVALID, "mmdb_fock_direct: imaginary density needs G_im");
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = b->nbf;
    const size_t N2 = (size_t)N * N;
    dabs_kernel<<<b->nsm * 4, 256, 0, st>>>(dP_re_dev, dP_im_dev, N2, b->Dabs_dev);
    CU(cudaMemsetAsync(b->dglob_dev, 0, sizeof(unsigned long long), st));
    // Grouped shell list (S2 pseudo-shells: generally contracted s functions share their primitive integrals) unless the
    // basis has none, the caller asks for the plain classes (flags bit 2: statistics in the reference's own units,
    // per-class timing of the plain kernels) or MMDB_NO_GC is set (A/B switch).
    // The grouped classes serve the class pairs whose bra AND ket are s/p-only pairs — (ss|ss), (ps|ss), (ps|ps) in the
    // reference's terms, where deep s contractions dominate the work; every class pair with a pp/ds/dp/dd bra runs on
    // the plain classes (measured: an S2 bra against a wide ket only adds registers).  As sets of function pairs the
    // grouped classes {ss, ps, S2 s, S2 p, S2 S2} and the plain classes {ss, ps} are the same, so the split is exact.
    const bool gc = b->have_gc && !(flags & 4) && !getenv("MMDB_NO_GC");
    if (gc)
        dshell_kernel<<<(b->nshellg * b->nshellg + 127) / 128, 128, 0, st>>>(b->Dabs_dev, N, b->shg_bf0_dev, b->shg_nf_dev,
                                                                             b->nshellg, b->DSg_dev, b->dglob_dev);
    dshell_kernel<<<(b->nshell * b->nshell + 127) / 128, 128, 0, st>>>(b->Dabs_dev, N, b->sh_bf0_dev, b->sh_nf_dev,
                                                                       b->nshell, b->DS_dev, b->dglob_dev);
    CU(cudaMemsetAsync(b->ctr_dev, 0, sizeof(unsigned long long) * b->nctr, st));
    struct Launch { PairClass *B, *K; bool gc; int slot; cudaEvent_t e0, em, e1; };
    std::vector<Launch> launches;
    const bool timing = (flags & 1) != 0;
    // Two queues.  Class pairs with few candidates (the d-heavy classes: a few thousand to a few million
    // quartets) cannot fill 148 SMs and are bounded by the latency of their longest thread; they run on an
    // auxiliary stream with their own list buffer, concurrently with the big classes on the caller's stream.
    // With per-class event timing everything stays on one stream.
    struct Task { PairClass *B, *K; bool gc; int row0, row1; size_t cap; bool aux; };
    std::vector<Task> tasks;
    size_t cap_main = 0, cap_aux = 0;
    std::vector<std::pair<PairClass *, PairClass *>> cpairs;
    std::vector<char> cpair_gc;
    if (gc) {
        // KET = the more deeply contracted class of the two (the ket primitive loop is warp-uniform and amortises the
        // per-bra-primitive work; measured: an S2 bra against a shallow ket costs 1.3-1.9x per primitive quartet), in
        // the order S2 S2 (<= 64 primitive pairs), S2 s, S2 p (<= 24), ss, ps (<= 9)
        // — except between two S2 classes, where the deeper one is the BRA: bra pairs are cut into slices of <= 8
        // primitive pairs, and these small class pairs need the entries more than the amortisation
        const int by_depth[5] = {8, 6, 7, 0, 1};
        for (int x = 0; x < 5; ++x)
            for (int y = x; y < 5; ++y) {
                const bool both_s2 = y < 3 && !(x == 1 && y == 2);      // (S2 p | S2 s) keeps its measured orientation
                if (both_s2) cpairs.push_back({&b->pcg[by_depth[x]], &b->pcg[by_depth[y]]});
                else cpairs.push_back({&b->pcg[by_depth[y]], &b->pcg[by_depth[x]]});
                cpair_gc.push_back(1);
            }
    }
    // Orientation of the plain class pairs: by default the pair of the higher class is the bra; for the class pairs in
    // SWAPPED the NARROW pair is the bra (fewer Hermite rows and fewer bra components per thread: smaller register
    // blocks, fewer ket-component chunks).  MMDB_SWAP="41,32,..." (bra class, ket class digits) overrides the list.
    // (Also measured: pp / ds / dp bras against the ss-type kets of the grouped list, i.e. S2 kets under a wide bra:
    // (pp|ss) 3.73 -> 4.0 ms, (ds|ss) 3.35 -> 3.7, (dp|ss) 2.0 -> 2.5 — three launches of 232-255 registers instead of one.)
    const char *swap_env = getenv("MMDB_SWAP");
    const char *swap_list = swap_env ? swap_env : SWAPPED;
    for (int cb = gc ? 2 : 0; cb < MMDB_NCLASS_PAIR; ++cb)
        for (int ck = 0; ck <= cb; ++ck) {
            const char tok[3] = {(char)('0' + cb), (char)('0' + ck), 0};
            const bool swp = strstr(swap_list, tok) != nullptr;
            if (swp) cpairs.push_back({&b->pc[ck], &b->pc[cb]});
            else cpairs.push_back({&b->pc[cb], &b->pc[ck]});
            cpair_gc.push_back(0);
        }
    if (!getenv("MMDB_NO_PACKED_BLOCKS")) {
        std::vector<PairClass *> packed;
        for (auto &cp : cpairs) {
            PairClass *B = cp.first;
            if (B->npairs == 0 || std::find(packed.begin(), packed.end(), B) != packed.end()) continue;
            packed.push_back(B);
            const long long work = (long long)B->npairs * ncomp(B->la) * ncomp(B->lb);
            pack_pair_blocks_kernel<<<(int)std::min<long long>((work + 255) / 256, 2048), 256, 0, st>>>(B->hdr_dev, B->npairs, B->la, B->lb, N, dP_re_dev,
                                                                                                  b->SQ_dev, B->PQ_dev);
        }
    }
    for (size_t cp = 0; cp < cpairs.size(); ++cp) {
        {
            PairClass &B = *cpairs[cp].first, &K = *cpairs[cp].second;
            const bool tgc = cpair_gc[cp] != 0;
            if (B.npairs == 0 || K.npairs == 0) continue;
            // rows of this shard only count towards the list capacity
            size_t rows_per = std::max<size_t>(1, LIST_CAP / (size_t)B.slice_entries * (size_t)nshards);
            const size_t aux_max = getenv("MMDB_AUX_MAX") ? (size_t)atoll(getenv("MMDB_AUX_MAX")) : AUX_MAX_CANDIDATES;
            const bool small = !timing && (size_t)B.npairs * K.npairs / nshards <= aux_max;
            for (int row0 = 0; row0 < K.npairs; row0 += (int)rows_per) {
                const int row1 = (int)std::min<size_t>(K.npairs, row0 + rows_per);
                const size_t cap = (size_t)((row1 - row0 + nshards - 1) / nshards + 1) * B.slice_entries;   // room for every slice
                tasks.push_back(Task{&B, &K, tgc, row0, row1, cap, small});
                (small ? cap_aux : cap_main) = std::max(small ? cap_aux : cap_main, cap);
            }
        }
    }
    // Screening pipeline: the main queue's lists are double-buffered and the screen of task m+1 runs on its own
    // stream while the ERI kernels of task m execute (it fills the SMs the persistent ERI grid vacates at its tail
    // instead of serialising behind it).  Per-class event timing keeps everything on one stream.
    const bool pipeline = !timing;
    // every region holds a far list [cap] followed by a near list [cap] (whose tail end is the slow list)
    CHK(ensure_list(b, 2 * ((pipeline ? 2 : 1) * cap_main + MMDB_NAUX * cap_aux)));    // [main 0 | main 1 | aux 0..] regions of one buffer
    uint2 *list_main[2] = {b->list_dev, b->list_dev + (pipeline ? 2 * cap_main : 0)};
    uint2 *list_aux = b->list_dev + (pipeline ? 2 : 1) * 2 * cap_main;
    const size_t cap_region[2] = {cap_main, cap_aux};
    if ((int)tasks.size() * CTR_PER_LAUNCH > b->nctr) return fail(MMDB_ERR_NOMEM, "mmdb_fock_direct: counter slots exhausted");
    const int naux = getenv("MMDB_NAUX") ? std::max(1, std::min(MMDB_NAUX, atoi(getenv("MMDB_NAUX")))) : MMDB_NAUX_DEFAULT;
    if (cap_aux > 0) {
        if (!b->aux_stream[0]) {
            CU(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
            for (int x = 0; x < MMDB_NAUX; ++x) {
                CU(cudaStreamCreateWithFlags(&b->aux_stream[x], cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&b->ev_join[x], cudaEventDisableTiming));
            }
        }
        CU(cudaEventRecord(b->ev_fork, st));           // density screens + zeroed counters are ready
        for (int x = 0; x < naux; ++x) CU(cudaStreamWaitEvent(b->aux_stream[x], b->ev_fork, 0));
    }
    cudaStream_t ss = st;
    size_t n_main = 0;
    for (const Task &t : tasks) n_main += t.aux ? 0 : 1;
    if (pipeline && n_main > 0) {
        if (!b->scr_stream) {
            CU(cudaStreamCreateWithFlags(&b->scr_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&b->ev_fork_scr, cudaEventDisableTiming));
        }
        while (b->ev_pool.size() < 2 * n_main) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            b->ev_pool.push_back(e);
        }
        ss = b->scr_stream;
        CU(cudaEventRecord(b->ev_fork_scr, st));       // density screens + zeroed counters are ready
        CU(cudaStreamWaitEvent(ss, b->ev_fork_scr, 0));
    }
    // Tail filling: consecutive class pairs of the main queue alternate between the caller's stream and a second one.
    // Every ERI launch is a persistent grid that claims all SMs; on ONE stream the grid of pair m+1 cannot start before
    // the last CTA of pair m has finished, so every launch ends with SMs idling behind its slowest CTA.  On two streams
    // the CTAs of pair m+1 move in as the CTAs of pair m retire (G is accumulated with atomics, the lists are
    // double-buffered per pair parity, the scratch columns per stream).  MMDB_ONE_MAIN_STREAM=1 keeps the old order.
    cudaStream_t st2 = st;
    if (pipeline && n_main > 1 && !getenv("MMDB_ONE_MAIN_STREAM")) {
        if (!b->main2_stream) {
            CU(cudaStreamCreateWithFlags(&b->main2_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&b->ev_fork2, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&b->ev_join2, cudaEventDisableTiming));
        }
        st2 = b->main2_stream;
        CU(cudaEventRecord(b->ev_fork2, st));
        CU(cudaStreamWaitEvent(st2, b->ev_fork2, 0));
    }
    int slot = 0;
    size_t m_main = 0, m_aux = 0;
    // aux tasks first: they are enqueued (and start) while the host is still launching the big classes
    for (int pass = 0; pass < 2; ++pass)
        for (const Task &t : tasks) {
            if (t.aux != (pass == 0)) continue;
            PairClass &B = *t.B, &K = *t.K;
            const bool piped = pipeline && !t.aux;
            const int aux_id = t.aux ? (int)(m_aux++ % (size_t)naux) : 0;      // tasks of one auxiliary stream share its list region, in order
            cudaStream_t s1 = t.aux ? b->aux_stream[aux_id] : ((piped && (m_main & 1)) ? st2 : st);
            uint2 *list = t.aux ? list_aux + (size_t)aux_id * 2 * cap_aux : list_main[piped ? (m_main & 1) : 0];
            cudaStream_t s_scr = piped ? ss : s1;
            cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
            if (piped) {
                ev_ready = b->ev_pool[2 * m_main];
                ev_done = b->ev_pool[2 * m_main + 1];
                if (m_main >= 2) CU(cudaStreamWaitEvent(ss, b->ev_pool[2 * (m_main - 2) + 1], 0));   // buffer consumed
            }
            Launch ln{t.B, t.K, t.gc, slot, nullptr, nullptr, nullptr};
            if (timing) {
                CU(cudaEventCreate(&ln.e0));
                CU(cudaEventCreate(&ln.em));
                CU(cudaEventCreate(&ln.e1));
                CU(cudaEventRecord(ln.e0, s1));
            }
            uint2 *list_far = list, *list_near = list + cap_region[t.aux ? 1 : 0];
            CHK(run_screen(b, B, K, t.B == t.K, t.row0, t.row1, shard, nshards, false, tol, slot, true,
                           dP_im_dev != nullptr || (flags & 2) != 0,
                           (long long)t.cap, list_far, list_near, s_scr, t.gc, stats != nullptr));
            if (piped) {
                CU(cudaEventRecord(ev_ready, ss));
                CU(cudaStreamWaitEvent(s1, ev_ready, 0));
            }
            if (timing) CU(cudaEventRecord(ln.em, s1));
            EriArgs a;
            std::memset(&a, 0, sizeof(a));
            a.braH = B.hdr_dev; a.braP = B.prim_dev; a.braS = B.prim_soa_dev; a.braRow = B.prim_row_dev; a.braN = B.nprimpairs; a.ketH = K.hdr_dev; a.ketP = K.prim_dev;
            a.braW = B.wgt_soa_dev; a.ketW = K.wgt_dev;
            a.braPQ = (dP_im_dev == nullptr && !getenv("MMDB_NO_PACKED_BLOCKS")) ? B.PQ_dev : nullptr;
            a.same_class = (t.B == t.K);
            a.dg.N = N; a.dg.tol = tol; a.dg.SQ = b->SQ_dev; a.dg.Dabs = b->Dabs_dev;
            a.dg.dPre = dP_re_dev; a.dg.dPim = dP_im_dev; a.dg.Gre = G_re_dev; a.dg.Gim = G_im_dev;
            a.dg.fixed = (flags & 2) ? 1 : 0;
            // far-field list (asymptotic Boys branch only), near list, then the slow list (diagonal-type quartets /
            // complex density / deterministic mode)
            a.list = list_far; a.list_step = 1; a.count_dev = b->ctr_dev + CTR_PER_LAUNCH * slot + CTR_FAR;
            const bool use_far = !t.gc && far_enabled(B.la + B.lb + K.la + K.lb);
            if (use_far) CHK(launch_eri(b, B.la, B.lb, K.la, K.lb, a, LK_DIGEST_FAR, 0, s1));
            a.list = list_near; a.list_step = 1; a.count_dev = b->ctr_dev + CTR_PER_LAUNCH * slot + CTR_NEAR;
            CHK(launch_eri(b, B.la, B.lb, K.la, K.lb, a, LK_DIGEST, 0, s1));
            a.list = list_near + (t.cap - 1); a.list_step = -1; a.count_dev = b->ctr_dev + CTR_PER_LAUNCH * slot + CTR_SLOW;
            CHK(launch_eri(b, B.la, B.lb, K.la, K.lb, a, LK_DIGEST_SLOW, 0, s1));
            if (piped) {
                CU(cudaEventRecord(ev_done, s1));
                ++m_main;
            }
            if (timing) CU(cudaEventRecord(ln.e1, s1));
            launches.push_back(ln);
            ++slot;
        }
    if (cap_aux > 0) {
        for (int x = 0; x < naux; ++x) {
            CU(cudaEventRecord(b->ev_join[x], b->aux_stream[x]));
            CU(cudaStreamWaitEvent(st, b->ev_join[x], 0));    // everything enqueued after this call sees the full G
        }
    }
    if (st2 != st) {
        CU(cudaEventRecord(b->ev_join2, st2));
        CU(cudaStreamWaitEvent(st, b->ev_join2, 0));
    }
    if (stats) {
        std::vector<unsigned long long> ctr(CTR_PER_LAUNCH * (size_t)slot + CTR_PER_LAUNCH, 0ull);
        CU(cudaMemcpyAsync(ctr.data(), b->ctr_dev, sizeof(unsigned long long) * CTR_PER_LAUNCH * slot, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof(*stats));
        stats->launches = 2;                                   // dabs + dshell, then per task: screen + far? + near + slow
        stats->launches += gc ? 1 : 0;
        for (auto &ln : launches) stats->launches += 3 + ((!ln.gc && far_enabled(ln.B->la + ln.B->lb + ln.K->la + ln.K->lb)) ? 1 : 0);
        for (auto &ln : launches) {
            const PairClass &B = *ln.B, &K = *ln.K;
            const int64_t nq = (int64_t)ctr[CTR_PER_LAUNCH * ln.slot + CTR_NQUART];
            const int64_t npq = (int64_t)ctr[CTR_PER_LAUNCH * ln.slot + CTR_PRIMQ];
            stats->slow_quartets += (int64_t)ctr[CTR_PER_LAUNCH * ln.slot + CTR_SLOW];
            stats->far_entries += (int64_t)ctr[CTR_PER_LAUNCH * ln.slot + CTR_FAR];
            stats->near_entries += (int64_t)ctr[CTR_PER_LAUNCH * ln.slot + CTR_NEAR];
            // statistics are kept in the reference's units: an S2 shell counts as its two s shells (quartets and
            // primitive quartets are summed over the member contractions by the screening kernel), and the class of a
            // grouped pair is the plain class of its members
            int pb = pc_index(am_of(B.la), am_of(B.lb)), pk = pc_index(am_of(K.la), am_of(K.lb));
            int fla = am_of(B.la), flb = am_of(B.lb), flc = am_of(K.la), fld = am_of(K.lb);
            if (pb < pk) { std::swap(pb, pk); std::swap(fla, flc); std::swap(flb, fld); }
            const int cidx = pb * MMDB_NCLASS_PAIR + pk;
            const int64_t nfn = (int64_t)ncart(fla) * ncart(flb) * ncart(flc) * ncart(fld);
            stats->candidates += (int64_t)ctr[CTR_PER_LAUNCH * ln.slot + CTR_CAND];
            stats->quartets += nq;
            stats->prim_quartets += npq;
            stats->exec_prim_quartets += (int64_t)ctr[CTR_PER_LAUNCH * ln.slot + CTR_EXECPQ];
            stats->fn_quartets += nq * nfn;
            stats->model_flops += (double)npq * mmdb_class_flops(fla, flb, flc, fld) + 13.0 * (double)(nq * nfn);
            stats->class_quartets[cidx] += nq;
            stats->class_prim_quartets[cidx] += npq;
            if (timing) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ln.em, ln.e1);
                stats->class_ms[cidx] += ms;
                float ms_scr = 0.f;
                cudaEventElapsedTime(&ms_scr, ln.e0, ln.em);
                stats->class_screen_ms[cidx] += ms_scr;
                if (getenv("MMDB_TRACE_LAUNCHES"))      // per-launch table (the grouped classes are folded into the plain ones in stats)
                    fprintf(stderr, "launch (%d%d|%d%d) %s entries near %llu slow %llu  quartets %lld  prim quartets %lld (evaluated %llu)  screen %.3f ms  eri %.3f ms\n",
                            B.la, B.lb, K.la, K.lb, ln.gc ? "grouped" : "plain", ctr[CTR_PER_LAUNCH * ln.slot + CTR_NEAR],
                            ctr[CTR_PER_LAUNCH * ln.slot + CTR_SLOW], (long long)nq, (long long)npq,
                            ctr[CTR_PER_LAUNCH * ln.slot + CTR_EXECPQ], ms_scr, ms);
            }
        }
    } else if (timing) {
        CU(cudaStreamSynchronize(st));
    }
    for (auto &ln : launches)
        if (ln.e0) { cudaEventDestroy(ln.e0); cudaEventDestroy(ln.em); cudaEventDestroy(ln.e1); }
    CU(cudaGetLastError());
    return MMDB_OK;
}

// deterministic mode: G holds 2^50-scaled 64-bit integers until the (integer) reductions are done
__global__ void fixed_to_double_kernel(double *G, size_t n)
{
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x)
        G[x] = (double)__double_as_longlong(G[x]) * (1.0 / FIXED_SCALE);
}

extern "C" int mmdb_fixed_to_double(int device, double *G_dev, int64_t n, void *stream)
{
    CU(cudaSetDevice(device));
    fixed_to_double_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(G_dev, (size_t)n);
    CU(cudaGetLastError());
    return MMDB_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer forms
// ------------------------------------------------------------------------------------------
extern "C" int mmdb_schwarz_host(mmdb_basis *b, double *Q_tri)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    CHK(mmdb_schwarz(b, nullptr, nullptr));
    const int N = b->nbf;
    std::vector<double> Q((size_t)N * N);
    CU(cudaMemcpy(Q.data(), b->Q_dev, Q.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int p = 0; p < N; ++p)
        for (int q = 0; q <= p; ++q) Q_tri[(size_t)p * (p + 1) / 2 + q] = Q[(size_t)p * N + q];
    return MMDB_OK;
}

// Qs[i] = max over the pair's function pairs of sqrt|Q|; SQ = sqrt(Q) from a caller-supplied table
__global__ void schwarz_from_Q_kernel(const PairHdr *hdr, int npairs, int la, int lb, int N, const double *Q,
                                      double *SQ, double *Qs)
{
    const int na = ncart(la), nb = ncart(lb);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += gridDim.x * blockDim.x) {
        const PairHdr h = hdr[i];
        double qmax = 0.0;
        for (int a = 0; a < na; ++a)
            for (int c = 0; c < nb; ++c) {
                const int p = h.bfA + a, q = h.bfB + c;
                const double v = Q[(size_t)p * N + q];
                const double s = sqrt(v);
                SQ[(size_t)p * N + q] = s;
                SQ[(size_t)q * N + p] = s;
                qmax = fmax(qmax, sqrt(fabs(v)));
            }
        Qs[i] = qmax;
    }
}

extern "C" int mmdb_set_schwarz_host(mmdb_basis *b, const double *Q_tri)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    CU(cudaSetDevice(b->device));
    const int N = b->nbf;
    std::vector<double> Q((size_t)N * N);
    for (int p = 0; p < N; ++p)
        for (int q = 0; q <= p; ++q) Q[(size_t)p * N + q] = Q[(size_t)q * N + p] = Q_tri[(size_t)p * (p + 1) / 2 + q];
    CU(cudaMemcpy(b->Q_dev, Q.data(), Q.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemset(b->SQ_dev, 0, Q.size() * sizeof(double)));
    for (int c = 0; c < MMDB_NCLASS_PAIR; ++c) {
        PairClass &P = b->pc[c];
        if (P.npairs == 0) continue;
        schwarz_from_Q_kernel<<<(P.npairs + 127) / 128, 128>>>(P.hdr_dev, P.npairs, P.la, P.lb, N, b->Q_dev, b->SQ_dev,
                                                                P.Qs_dev);
        qs_chunk_max_kernel<<<((P.npairs + 255) / 256 + 3) / 4, 128>>>(P.Qs_dev, P.npairs, P.Qmax_dev);
    }
    CU(cudaGetLastError());
    CHK(gc_refresh_bounds(b, nullptr));
    CU(cudaDeviceSynchronize());
    b->have_schwarz = true;
    return MMDB_OK;
}

extern "C" int mmdb_eri_dense_host(mmdb_basis *b, double *TwoE_host)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    CU(cudaSetDevice(b->device));
    const size_t N = b->nbf, n4 = N * N * N * N;
    double *T = nullptr;
    CU(cudaMalloc(&T, n4 * sizeof(double)));
    int r = mmdb_eri_dense(b, T, nullptr);
    if (r == MMDB_OK) {
        cudaError_t e = cudaMemcpy(TwoE_host, T, n4 * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) r = fail(MMDB_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(T);
    return r;
}

// Host passes of the reference-facing call: dP = P - P_old (cython/fock.pyx:24) split into real / imaginary planes,
// and the planes of G interleaved back into the reference's complex128 layout.  One streaming pass each.
// The two host passes run on a few threads: at N = 800 they stream 25 + 15 MB, 3.2 ms of the reference-facing call on
// one core — a third of an eight-GPU build.  (MMDB_HOST_THREADS overrides the count; small matrices stay serial.)
static int host_threads(int64_t n)
{
    if (n < (int64_t)1 << 16) return 1;
    const char *e = getenv("MMDB_HOST_THREADS");
    if (e) return std::max(1, atoi(e));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(4u, hw ? hw / 4 : 1u));
}
template <class F>
static void host_parallel(int64_t n, F &&f)       // f(begin, end, thread id)
{
    const int T = host_threads(n);
    if (T == 1) { f((int64_t)0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back([&, t]() { f(n * t / T, n * (t + 1) / T, t); });
    f((int64_t)0, n / T, 0);
    for (auto &x : th) x.join();
}

extern "C" int mmdb_c128_diff_split_host(const double *P_c128, const double *P_old_c128, int64_t n, double *re, double *im,
                                         int *has_im)
{
    int any_t[64] = {0};
    host_parallel(n, [&](int64_t x0, int64_t x1, int t) {
        int any = 0;
        for (int64_t x = x0; x < x1; ++x) {
            re[x] = P_c128[2 * x] - P_old_c128[2 * x];
            const double v = P_c128[2 * x + 1] - P_old_c128[2 * x + 1];
            im[x] = v;
            any |= (v != 0.0);
        }
        any_t[t & 63] |= any;
    });
    int any = 0;
    for (int t = 0; t < 64; ++t) any |= any_t[t];
    if (has_im) *has_im = any;
    return MMDB_OK;
}
extern "C" int mmdb_c128_join_host(const double *re, const double *im, int64_t n, double *out_c128)
{
    host_parallel(n, [&](int64_t x0, int64_t x1, int) {
        if (im)
            for (int64_t x = x0; x < x1; ++x) { out_c128[2 * x] = re[x]; out_c128[2 * x + 1] = im[x]; }
        else
            for (int64_t x = x0; x < x1; ++x) { out_c128[2 * x] = re[x]; out_c128[2 * x + 1] = 0.0; }
    });
    return MMDB_OK;
}

static int ensure_stage(mmdb_basis *b)
{
    const size_t N2 = (size_t)b->nbf * b->nbf;
    if (b->stage_n == N2) return MMDB_OK;
    if (b->stage_host) cudaFreeHost(b->stage_host);
    if (b->stage_dev) cudaFree(b->stage_dev);
    b->stage_host = nullptr; b->stage_dev = nullptr; b->stage_n = 0;
    CU(cudaMallocHost(&b->stage_host, sizeof(double) * N2 * 4));      // [dP re | dP im | G re | G im], page-locked
    CU(cudaMalloc(&b->stage_dev, sizeof(double) * N2 * 4));
    b->stage_n = N2;
    return MMDB_OK;
}

// formPT with host buffers (cython/fock.pyx:13-87 as the reference's caller sees it): complex128 (N,N) in, complex128
// un-symmetrised G out.  Staging buffers (page-locked host + device) are cached in the handle; a real density moves
// one plane each way (8 N^2 bytes), a complex one two.
extern "C" int mmdb_formPT_host(mmdb_basis *b, const double *P_c128, const double *P_old_c128, double tol,
                                double *G_c128, mmdb_fock_stats *stats)
{
    if (!b) return fail(MMDB_ERR_INVALID, "null handle");
    CU(cudaSetDevice(b->device));
    CHK(ensure_stage(b));
    const size_t N2 = b->stage_n;
    double *h_re = b->stage_host, *h_im = h_re + N2, *h_Gre = h_re + 2 * N2, *h_Gim = h_re + 3 * N2;
    double *d_re = b->stage_dev, *d_im = d_re + N2, *d_Gre = d_re + 2 * N2, *d_Gim = d_re + 3 * N2;
    int has_im = 0;
    mmdb_c128_diff_split_host(P_c128, P_old_c128, (int64_t)N2, h_re, h_im, &has_im);
    CU(cudaMemcpyAsync(d_re, h_re, sizeof(double) * N2 * (has_im ? 2 : 1), cudaMemcpyHostToDevice, nullptr));
    CU(cudaMemsetAsync(d_Gre, 0, sizeof(double) * N2 * (has_im ? 2 : 1), nullptr));
    CHK(mmdb_fock_direct(b, d_re, has_im ? d_im : nullptr, tol, d_Gre, has_im ? d_Gim : nullptr, 0, 1, 0, stats, nullptr));
    CU(cudaMemcpyAsync(h_Gre, d_Gre, sizeof(double) * N2 * (has_im ? 2 : 1), cudaMemcpyDeviceToHost, nullptr));
    CU(cudaStreamSynchronize(nullptr));
    mmdb_c128_join_host(h_Gre, has_im ? h_Gim : nullptr, (int64_t)N2, G_c128);
    return MMDB_OK;
}

// ------------------------------------------------------------------------------------------
// utilities: Boys probe, FP64 peak probe
// ------------------------------------------------------------------------------------------
__global__ void boys_probe_kernel(int mmax, int64_t n, const double *T, const double *tab, double *out)
{
    extern __shared__ double s_boys[];
    for (int x = threadIdx.x; x < BOYS_ROWS * BOYS_STRIDE; x += blockDim.x) s_boys[x] = tab[x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double F[BOYS_MAXL + 1];
        boys_eval_rt(mmax, T[i], s_boys, F);
        for (int m = 0; m <= mmax; ++m) out[i * (mmax + 1) + m] = F[m];
    }
}

extern "C" int mmdb_boys_host(int device, int mmax, int64_t n, const double *T, double *out)
{
    if (mmax < 0 || mmax > BOYS_MAXL) return fail(MMDB_ERR_INVALID, "mmdb_boys_host: mmax out of range");
    CU(cudaSetDevice(device));
    std::vector<double> tab;
    make_boys_table(mmax, tab);
    double *dtab = nullptr, *dT = nullptr, *dout = nullptr;
    CU(cudaMalloc(&dtab, tab.size() * sizeof(double)));
    CU(cudaMalloc(&dT, n * sizeof(double)));
    CU(cudaMalloc(&dout, n * (mmax + 1) * sizeof(double)));
    CU(cudaMemcpy(dtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dT, T, n * sizeof(double), cudaMemcpyHostToDevice));
    boys_probe_kernel<<<148, 128, BOYS_ROWS * BOYS_STRIDE * sizeof(double)>>>(mmax, n, dT, dtab, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout, n * (mmax + 1) * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dtab); cudaFree(dT); cudaFree(dout);
    return MMDB_OK;
}

// Probe of the TEMPLATED Boys path the class kernels run (prim_Fs<L>: boys_table<L> below T_max(L), the alpha-free
// asymptotic branch at and above it).  With p = q = 2 (alpha = 1) and unit pair coefficients, T = |PQ|^2 and
// Fs[n] = (-2)^n F_n(T).
template <int L>
__global__ void boys_class_probe_kernel(int64_t n, const double *T, const double *tab, double *out)
{
    extern __shared__ double s_boys[];
    for (int x = threadIdx.x; x < boys_rows(L) * BOYS_STRIDE; x += blockDim.x) s_boys[x] = tab[x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double Fs[L + 1];
        prim_Fs<L>(Fs, 2.0, 2.0, 1.0, 1.0, T[i], s_boys);
        double sc = 1.0;
#pragma unroll
        for (int m = 0; m <= L; ++m) { out[i * (L + 1) + m] = Fs[m] * sc * SQRTPI_2; sc *= -0.5; }      // unit coefficients carry no sqrt(pi)/2
    }
}

extern "C" int mmdb_boys_class_host(int device, int L, int64_t n, const double *T, double *out)
{
    if (L < 0 || L > BOYS_MAXL) return fail(MMDB_ERR_INVALID, "mmdb_boys_class_host: L out of range");
    CU(cudaSetDevice(device));
    std::vector<double> tab;
    make_boys_table(L, tab);
    double *dtab = nullptr, *dT = nullptr, *dout = nullptr;
    CU(cudaMalloc(&dtab, tab.size() * sizeof(double)));
    CU(cudaMalloc(&dT, n * sizeof(double)));
    CU(cudaMalloc(&dout, n * (L + 1) * sizeof(double)));
    CU(cudaMemcpy(dtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dT, T, n * sizeof(double), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)BOYS_ROWS * BOYS_STRIDE * sizeof(double);
    switch (L) {
#define PROBE(LL) case LL: boys_class_probe_kernel<LL><<<148, 128, smem>>>(n, dT, dtab, dout); break;
        PROBE(0) PROBE(1) PROBE(2) PROBE(3) PROBE(4) PROBE(5) PROBE(6) PROBE(7) PROBE(8)
#undef PROBE
    }
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout, n * (L + 1) * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dtab); cudaFree(dT); cudaFree(dout);
    return MMDB_OK;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double c)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fma(x0, a, c); x1 = fma(x1, a, c); x2 = fma(x2, a, c); x3 = fma(x3, a, c);
            x4 = fma(x4, a, c); x5 = fma(x5, a, c); x6 = fma(x6, a, c); x7 = fma(x7, a, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int mmdb_fp64_peak(int device, double *tflops, float *ms_out)
{
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    double *out = nullptr;
    CU(cudaMalloc(&out, (size_t)grid * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(e0));
        dfma_peak_kernel<<<grid, 256>>>(out, iters, 0.999999, 1e-9);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 1) best = std::min(best, ms);
    }
    const double flops = 2.0 * 8 * 16 * (double)iters * grid * 256;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaFree(out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MMDB_OK;
}

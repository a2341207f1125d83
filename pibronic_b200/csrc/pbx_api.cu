// C ABI of the pbx library (include/pbx.h): plan management, launch orchestration, reductions.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "pbx_big.cuh"
#include "pbx_fast.cuh"
#include "pbx_generic.cuh"
#include "pbx_mid.cuh"

using namespace pbx;

namespace pbx {
#define PBX_BIG_LIST(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#define PBX_BIG_DECL(A_) extern const BigLauncher big_launcher_##A_;
PBX_BIG_LIST(PBX_BIG_DECL)
#undef PBX_BIG_DECL
BigLauncher find_big_kernel(int A) {
    switch (A) {
#define PBX_BIG_CASE(A_) case A_: return big_launcher_##A_;
        PBX_BIG_LIST(PBX_BIG_CASE)
#undef PBX_BIG_CASE
    }
    return nullptr;
}
}  // namespace pbx

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(PBX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define PBX_NEED_DEVICE(p) do { if ((p)->device < 0) return fail(PBX_ERR_CUDA, "host-only plan (device -1): nothing can be launched, pbx has no CPU fallback"); } while (0)
#define PBX_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on a private stream):
// a cudaFree costs ~0.3 ms (it synchronises the device and unmaps), and a plan owns up to eight buffers -- a temperature
// x bead sweep creates and destroys a plan per point.  The pool keeps up to 512 MB of freed blocks for re-use.
// Callers free only after the work that used the block has completed (plan_destroy and ensure synchronise first).
struct PoolState {
    cudaStream_t stream = nullptr;
    bool ready = false;
};
static PoolState g_pool[64];
static std::mutex g_pool_mutex;

static cudaError_t pool_stream(cudaStream_t* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    PoolState& st = g_pool[dev];
    if (!st.ready) {
        cudaMemPool_t pool;
        if ((e = cudaDeviceGetDefaultMemPool(&pool, dev)) != cudaSuccess) return e;
        unsigned long long keep = 512ull << 20;
        if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        st.ready = true;
    }
    *out = st.stream;
    return cudaSuccess;
}

static cudaError_t dev_alloc(void** ptr, size_t bytes) {
    cudaStream_t st;
    cudaError_t e = pool_stream(&st);
    if (e != cudaSuccess) return e;
    if ((e = cudaMallocAsync(ptr, bytes, st)) != cudaSuccess) return e;
    return cudaStreamSynchronize(st);       // the block is valid on every stream from here on
}

static void dev_free(void* ptr) {
    if (!ptr) return;
    cudaStream_t st;
    if (pool_stream(&st) != cudaSuccess || cudaFreeAsync(ptr, st) != cudaSuccess) { cudaGetLastError(); cudaFree(ptr); }
}

constexpr size_t kScratchTarget = (size_t)4 << 30;  // aim for <= 4 GiB of intermediates per chunk (the sampler kernel of the large-A path needs ~1e4 samples to fill the GPU)

// ---- per-block sums ------------------------------------------------------------------------
// grid (estimator blocks, split): CTA (b, s) reduces slice s of block b with a fixed-order tree; the last CTA of a
// block to finish (atomic ticket) adds the `split` partial rows in index order -> results do not depend on timing.
// A single CTA per block left most of the GPU idle at the bench shape (100 blocks of 1e4 samples: 60 us for 32 MB).
__global__ void __launch_bounds__(256)
pbx_block_sums_kernel(const double* __restrict__ out4, long long ld, long long n_samples, long long block_size,
                      double delta_beta, int pm, int split, double* __restrict__ partials, int* __restrict__ tickets,
                      double* __restrict__ sums) {
    __shared__ double sh[PBX_NSUMS][256];
    __shared__ int is_last;
    const long long blk = blockIdx.x;
    const long long blo = blk * block_size, bhi = min(blo + block_size, n_samples);
    const long long chunk = (bhi - blo + split - 1) / split;
    const long long lo = blo + (long long)blockIdx.y * chunk, hi = min(lo + chunk, bhi);
    double acc[PBX_NSUMS];
#pragma unroll
    for (int k = 0; k < PBX_NSUMS; ++k) acc[k] = 0.0;
    const double inv2db = pm ? 1.0 / (2.0 * delta_beta) : 0.0, invdb2 = pm ? 1.0 / (delta_beta * delta_beta) : 0.0;
    for (long long x = lo + threadIdx.x; x < hi; x += 256) {
        const double rho = out4[x], g = out4[ld + x];
        const double r = g / rho;
        acc[0] += r; acc[3] = fma(r, r, acc[3]);
        if (pm) {
            const double gp = out4[2 * ld + x], gm = out4[3 * ld + x];
            const double rp = gp / rho, rm = gm / rho;
            const double d1 = (gp - gm) / rho * inv2db, d2 = (gp - 2.0 * g + gm) / rho * invdb2;
            acc[1] += rp; acc[2] += rm; acc[4] += d1; acc[5] += d2;
            acc[6] = fma(d1, d1, acc[6]); acc[7] = fma(d2, d2, acc[7]);
        }
    }
#pragma unroll
    for (int k = 0; k < PBX_NSUMS; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
#pragma unroll
            for (int k = 0; k < PBX_NSUMS; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
        __syncthreads();
    }
    if (split == 1) {
        if (threadIdx.x < PBX_NSUMS) sums[blk * PBX_NSUMS + threadIdx.x] = sh[threadIdx.x][0];
        return;
    }
    if (threadIdx.x < PBX_NSUMS) partials[(blk * split + blockIdx.y) * PBX_NSUMS + threadIdx.x] = sh[threadIdx.x][0];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&tickets[blk], 1) == split - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < PBX_NSUMS) {
        double t = 0.0;
        for (int s = 0; s < split; ++s) t += __ldcg(&partials[(blk * split + s) * PBX_NSUMS + threadIdx.x]);
        sums[blk * PBX_NSUMS + threadIdx.x] = t;
    }
}

// ---- statistics: sums for Z, E, Cv and the leave-one-out jackknife (stats.py:38-123, jackknife.py:60-105) ----
constexpr int kStatGrid = 592;   // 4 CTAs per SM; partials are added on the host in a fixed order

__device__ __forceinline__ void stat_terms(const double* __restrict__ out4, long long ld, long long x, double inv2db,
                                           double invdb2, double& r, double& d1, double& d2) {
    const double rho = out4[x], g = out4[ld + x], gp = out4[2 * ld + x], gm = out4[3 * ld + x];
    r = g / rho;
    d1 = (gp - gm) / rho * inv2db;
    d2 = (gp - 2.0 * g + gm) / rho * invdb2;
}

template <int K>
__device__ __forceinline__ void cta_reduce_store(double (&acc)[K], double* __restrict__ partials) {
    __shared__ double sh[K][256];
#pragma unroll
    for (int k = 0; k < K; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
#pragma unroll
            for (int k = 0; k < K; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < K) partials[blockIdx.x * K + threadIdx.x] = sh[threadIdx.x][0];
}

// partials[grid][4]: sum r, sum r^2, sum d1, sum d2
__global__ void __launch_bounds__(256)
pbx_stat_sums_kernel(const double* __restrict__ out4, long long ld, long long n, double inv2db, double invdb2,
                     double* __restrict__ partials) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long x = (long long)blockIdx.x * 256 + threadIdx.x; x < n; x += (long long)gridDim.x * 256) {
        double r, d1, d2;
        stat_terms(out4, ld, x, inv2db, invdb2, r, d1, d2);
        acc[0] += r; acc[1] = fma(r, r, acc[1]); acc[2] += d1; acc[3] += d2;
    }
    cta_reduce_store<4>(acc, partials);
}

// leave-one-out estimators f_E, f_C of every sample, accumulated relative to the shifts kE, kC (the full-sample
// values, within O(1/X) of every f) so that the variance does not cancel.  partials[grid][4]: sum dE, dE^2, dC, dC^2
__global__ void __launch_bounds__(256)
pbx_jackknife_kernel(const double* __restrict__ out4, long long ld, long long n, double inv2db, double invdb2,
                     double S_r, double S_1, double S_2, double inv_kbt2, double kE, double kC,
                     double* __restrict__ partials) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const double inv_nm1 = 1.0 / (double)(n - 1);
    for (long long x = (long long)blockIdx.x * 256 + threadIdx.x; x < n; x += (long long)gridDim.x * 256) {
        double r, d1, d2;
        stat_terms(out4, ld, x, inv2db, invdb2, r, d1, d2);
        const double jr = (S_r - r) * inv_nm1, j1 = (S_1 - d1) * inv_nm1, j2 = (S_2 - d2) * inv_nm1;
        const double fE = -j1 / jr;
        const double fC = (j2 / jr - fE * fE) * inv_kbt2;
        const double dE = fE - kE, dC = fC - kC;
        acc[0] += dE; acc[1] = fma(dE, dE, acc[1]); acc[2] += dC; acc[3] = fma(dC, dC, acc[3]);
    }
    cta_reduce_store<4>(acc, partials);
}

// ---- self-test of the branch-free device math (pbx_device.cuh) ------------------------------------------
__global__ void pbx_math_probe_kernel(int kind, const double* __restrict__ in, double* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = in[i];
    if (kind == 0) out[i] = log_pos(x);
    else if (kind == 1) out[i] = sqrt_pos(x);
    else if (kind == 2) out[i] = exp_fast(x);
    else if (kind == 3) { double s, c; sincos_2pi(x, s, c); out[i] = s; }
    else { double s, c; sincos_2pi(x, s, c); out[i] = c; }
}

// ---- FP64 peak probe: 8 independent dependent-FMA chains per thread ------------------------------
__global__ void __launch_bounds__(256) pbx_dfma_probe_kernel(double* out, int iters, double a, double b) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = threadIdx.x * 1e-9 + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, b);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 123.456) out[0] = s;  // never true; keeps the loop alive
}

// same measurement through the FP64 tensor path: 8 independent m8n8k4 accumulator tiles per warp.  On B200 the two share
// the FP64 units (tools/microbench/dmma_mix.cu: mixed loops never exceed the DMMA-only rate), DMMA merely reaches
// them with fewer operand reads: 37.2 vs 34.0 TFLOP/s
__global__ void __launch_bounds__(256) pbx_dmma_probe_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = threadIdx.x * 1e-9; c[k][1] = k; }
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    if (s == 123.456) out[0] = s;
}

}  // namespace

struct pbx_plan {
    HostTables H;
    int device = 0;
    uint32_t flags = 0;
    bool pm = false, jacobi = false, scale = true, mtau = false;
    double* dev_tables = nullptr;
    int* dev_int_tables = nullptr;
    double* dmma_tables = nullptr;
    DevTables D{};
    const FastKernelEntry* fast = nullptr;
    std::vector<unsigned char> fast_tables;
    BigLauncher big = nullptr;       // fused large-A kernel (pbx_big.cuh), null if the shape is not eligible
    BigParams B{};
    double* big_tab = nullptr;
    size_t big_smem = 0;
    int big_warps = BIG_WARPS;     // warps per CTA of the fused tensor-core kernel (fewer when eight do not fit the SM)
    int sms = 148;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* io = nullptr;  // device staging for the *_host entry points
    size_t io_bytes = 0;
    long long io_samples = 0;      // samples of the last *_host call still resident in `io` ([4][io_samples])
    double* stat_partials = nullptr;
    void* sums_scratch = nullptr;  // block-sum partials [blocks][split][NSUMS] followed by the tickets [blocks]
    size_t sums_scratch_bytes = 0;
    cudaStream_t own_stream = nullptr;
    long long launches = 0;
};

namespace {

int ensure(void** buf, size_t* have, size_t need) {
    if (need <= *have) return PBX_OK;
    if (*buf) { PBX_CUDA(cudaDeviceSynchronize()); dev_free(*buf); *buf = nullptr; *have = 0; }
    PBX_CUDA(dev_alloc(buf, need));
    *have = need;
    return PBX_OK;
}

int upload_tables(pbx_plan* p) {
    const HostTables& H = p->H;
    std::vector<double> flat;
    auto push = [&](const std::vector<double>& v) { size_t off = flat.size(); flat.insert(flat.end(), v.begin(), v.end()); return off; };
    // harmonic exponent in half-angle form, -1/4 [tanh(x/2) (q + q')^2 + coth(x/2) (q - q')^2]: the reference's
    // coth (q^2 + q'^2) - 2 csch q q' cancels digits for tau*omega << 1 (DESIGN.md section 2 iii); every kernel uses this form
    std::vector<double> hc(H.tanh_half.size()), cs(H.coth_half.size());
    for (size_t i = 0; i < hc.size(); ++i) { hc[i] = -0.25 * H.tanh_half[i]; cs[i] = -0.25 * H.coth_half[i]; }
    const size_t o_drs = push(H.d_rho);
    const size_t o_dv = push(H.d_vib), o_dr = push(H.d_rho_eval), o_hc = push(hc), o_cs = push(cs), o_lp = push(H.logpref),
                 o_lpr = push(H.logpref_rho), o_wc = push(H.wcum), o_e = push(H.e_off), o_l = push(H.l_off),
                 o_q = push(H.q_pack), o_s = push(H.samp);
    PBX_CUDA(dev_alloc((void**)&p->dev_tables, flat.size() * sizeof(double)));
    PBX_CUDA(cudaMemcpy(p->dev_tables, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice));
    DevTables& D = p->D;
    D.A = H.A; D.Ar = H.Ar; D.N = H.N; D.P = H.P; D.AA = H.AA; D.NN = H.NN; D.n_rho_eval = H.n_rho_eval;
    D.neg_tau = -H.tau[0];
    const double* b = p->dev_tables;
    D.d_rho_samp = b + o_drs;
    D.d_vib = b + o_dv; D.d_rho = b + o_dr; D.hc = b + o_hc; D.cs = b + o_cs; D.lpref = b + o_lp;
    D.lpref_rho = b + o_lpr; D.wcum = b + o_wc; D.e_off = b + o_e; D.l_off = b + o_l; D.q_pack = b + o_q;
    D.samp = b + o_s;
    return PBX_OK;
}

// coupling tables in mma.sync.m8n8k4.f64 fragment order (see DevTables)
int upload_dmma_tables(pbx_plan* p) {
    const HostTables& H = p->H;
    const int N = H.N, AA = H.AA, ONE = N;       // row N of the coordinate tile holds ones
    std::vector<int> fa, fb;
    std::vector<const double*> coef;
    if (H.has_quadratic)
        for (int n = 0; n < N; ++n)
            for (int m = n; m < N; ++m) { fa.push_back(n); fb.push_back(m); coef.push_back(&H.q_pack[(size_t)pair_index(n, m, N) * AA]); }
    for (int n = 0; n < N; ++n) { fa.push_back(n); fb.push_back(ONE); coef.push_back(&H.l_off[(size_t)n * AA]); }
    fa.push_back(ONE); fb.push_back(ONE); coef.push_back(H.e_off.data());
    // k-steps of 4 features, padded (zero coefficients) to whole chunks of the fused kernel's coefficient ring
    const int K = (int)fa.size(), KS = ((K + 3) / 4 + BIG_CH - 1) / BIG_CH * BIG_CH, NT = (AA + 7) / 8;
    std::vector<double> q((size_t)KS * NT * 32, 0.0);
    for (int ks = 0; ks < KS; ++ks)
        for (int j = 0; j < NT; ++j)
            for (int lane = 0; lane < 32; ++lane) {
                const int f = 4 * ks + lane % 4, k = 8 * j + lane / 4;
                if (f < K && k < AA) q[((size_t)ks * NT + j) * 32 + lane] = coef[f][k];
            }
    std::vector<int> ints((size_t)4 * KS + 8 * NT);
    for (int f = 0; f < 4 * KS; ++f) ints[f] = f < K ? (fa[f] | (fb[f] << 16)) : (ONE | (ONE << 16));
    for (int k = 0; k < 8 * NT; ++k) {
        int v = -1;
        if (k < AA) {
            int i = 0;
            while (tri(i + 1, 0) <= k) ++i;
            v = (i << 16) | (k - tri(i, 0));
        }
        ints[(size_t)4 * KS + k] = v;
    }
    // two tables: the coefficients as they are (blocked kernels: they also return V) and multiplied by -tau (fused kernel:
    // its tensor-core contraction yields X = -tau V directly)
    double* dq = nullptr;
    PBX_CUDA(dev_alloc((void**)&dq, 2 * q.size() * sizeof(double)));
    p->dmma_tables = dq;
    PBX_CUDA(cudaMemcpy(dq, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice));
    for (double& v : q) v *= -H.tau[0];
    PBX_CUDA(cudaMemcpy(dq + q.size(), q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice));
    PBX_CUDA(dev_alloc((void**)&p->dev_int_tables, ints.size() * sizeof(int)));
    PBX_CUDA(cudaMemcpy(p->dev_int_tables, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice));
    p->D.q_dmma = dq; p->D.feat = p->dev_int_tables; p->D.tri_ij = p->dev_int_tables + 4 * KS;
    p->D.KS = KS; p->D.NT = NT;
    return PBX_OK;
}

// flat constant table of the fused large-A kernel (BigParams::tab) + its launch parameters
int upload_big_tables(pbx_plan* p) {
    const HostTables& H = p->H;
    const int A = H.A, Ar = H.Ar, N = H.N;
    std::vector<double> flat;
    BigParams& B = p->B;
    B.o_al = (int)flat.size();
    for (int v = 0; v < 4; ++v) for (int n = 0; n < N; ++n) flat.push_back(-0.25 * H.tanh_half[v * N + n]);
    B.o_ga = (int)flat.size();
    for (int v = 0; v < 4; ++v) for (int n = 0; n < N; ++n) flat.push_back(-0.25 * H.coth_half[v * N + n]);
    B.o_d2v = (int)flat.size();
    for (int i = 0; i < A * N; ++i) flat.push_back(2.0 * H.d_vib[i]);
    B.o_d2r = (int)flat.size();
    for (int i = 0; i < Ar * N; ++i) flat.push_back(2.0 * H.d_rho_eval[i]);
    B.o_lpref = (int)flat.size();
    flat.insert(flat.end(), H.logpref.begin(), H.logpref.end());
    B.o_lprho = (int)flat.size();
    flat.insert(flat.end(), H.logpref_rho.begin(), H.logpref_rho.end());
    B.o_drho = (int)flat.size();
    flat.insert(flat.end(), H.d_rho.begin(), H.d_rho.end());
    B.tab_doubles = (int)flat.size();
    int max_smem = 0;
    PBX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device));
    PBX_CUDA(cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, p->device));
    // eight warps (two per scheduler) per CTA; shapes whose per-warp regions do not fit eight times (A > 12 with many
    // modes) run four, one per scheduler
    for (p->big_warps = BIG_WARPS; p->big_warps >= 4; p->big_warps /= 2) {
        p->big_smem = big_smem_bytes(A, p->pm, N, Ar, B.tab_doubles, p->D.KS, p->big_warps);
        if (p->big_smem <= (size_t)max_smem) break;
    }
    if (p->big_smem > (size_t)max_smem) { p->big = nullptr; return PBX_OK; }   // falls back to the blocked kernels
    PBX_CUDA(dev_alloc((void**)&p->big_tab, flat.size() * sizeof(double)));
    PBX_CUDA(cudaMemcpy(p->big_tab, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice));
    B.tab = p->big_tab;
    B.Ar = Ar; B.N = N; B.P = H.P; B.n_rho_eval = H.n_rho_eval; B.KS = p->D.KS; B.neg_tau = -H.tau[0];
    B.share = H.rho_shares_vib ? 1 : 0;
    B.q_dmma = p->D.q_dmma + (size_t)p->D.KS * p->D.NT * 32;      // the -tau scaled table
    B.wcum = p->D.wcum; B.feat = p->D.feat; B.tri_ij = p->D.tri_ij; B.samp = p->D.samp;
    return PBX_OK;
}

// one launch of the fused large-A kernel: R == nullptr draws the coordinates on-chip (needs N <= BIG_NMAX_SAMPLER)
int launch_big(pbx_plan* p, const double* R, uint64_t seed, long long first, long long n, double* out4, long long out_ld,
               double* mirror, long long mirror_ld, cudaStream_t st) {
    BigParams B = p->B;
    B.R = R; B.seed = seed; B.first_sample = first; B.n_samples = n; B.out4 = out4; B.out_ld = out_ld;
    B.mirror = mirror; B.mirror_ld = mirror_ld;
    PBX_CUDA(p->big(B, p->pm, R ? BIG_COORDS : BIG_SAMPLE, p->big_smem, p->sms, p->big_warps, st));
    p->launches += 1;
    return PBX_OK;
}

// doubles of intermediates per sample on the generic path (coords excluded)
size_t generic_doubles_per_sample(const HostTables& H) {
    return (size_t)H.P * ((size_t)H.A * H.A + 3 * (size_t)H.A + H.Ar);
}

// blocked kernels (pbx_mid.cuh) for A <= 16 with the default M builder and scaling
bool use_mid_path(const pbx_plan* p) { return p->H.A <= MID_AMAX && !p->jacobi && p->scale; }

template <int AT>
int launch_mid_at(pbx_plan* p, const double* R, long long n, const BeadOutputs& bo, double* out4, long long out_ld,
                  cudaStream_t st) {
    const HostTables& H = p->H;
    const long long groups = n * ((H.P + MID_IB - 1) / MID_IB);
    const size_t smem_b = MID_WARPS * mid_bead_warp_doubles(AT, H.Ar, H.N) * sizeof(double);
    auto kb = pbx_mid_bead_kernel<AT>;
    PBX_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    kb<<<(unsigned)((groups + MID_WARPS - 1) / MID_WARPS), MID_WARPS * 32, smem_b, st>>>(p->D, R, n, bo);
    PBX_CUDA(cudaGetLastError());
    if (bo.m_mat) {   // X = -tau V  ->  M = exp(X), in place
        const long long items = n * H.P;
        const size_t smem_e = MID_WARPS * mid_expm_warp_doubles(AT) * sizeof(double);
        auto ke = pbx_mid_expm_kernel<AT>;
        PBX_CUDA(cudaFuncSetAttribute(ke, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
        const long long grid_e = std::min<long long>((items + MID_WARPS - 1) / MID_WARPS, 148 * 5);   // persistent warps, 5 CTAs per SM
        ke<<<(unsigned)grid_e, MID_WARPS * 32, smem_e, st>>>(bo.m_mat, items);
        PBX_CUDA(cudaGetLastError());
        p->launches += 1;
    }
    constexpr int spw = 32 / AT;
    const unsigned grid = (unsigned)((n + (long long)MID_WARPS * spw - 1) / ((long long)MID_WARPS * spw));
    const size_t smem_c = MID_WARPS * mid_chain_warp_doubles(AT) * sizeof(double);
    if (p->pm) {
        auto kc = pbx_mid_chain_kernel<AT, true>;
        PBX_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        kc<<<grid, MID_WARPS * 32, smem_c, st>>>(p->D, bo.m_mat, bo.o_vib, bo.lr, n, out4, out4 + out_ld, out_ld);
    } else {
        auto kc = pbx_mid_chain_kernel<AT, false>;
        PBX_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        kc<<<grid, MID_WARPS * 32, smem_c, st>>>(p->D, bo.m_mat, bo.o_vib, bo.lr, n, out4, out4 + out_ld, out_ld);
    }
    PBX_CUDA(cudaGetLastError());
    p->launches += 2;
    return PBX_OK;
}

int launch_mid(pbx_plan* p, const double* R, long long n, const BeadOutputs& bo, double* out4, long long out_ld,
               cudaStream_t st) {
    switch (p->H.A) {
#define PBX_MID_CASE(A_) case A_: return launch_mid_at<A_>(p, R, n, bo, out4, out_ld, st);
        PBX_MID_CASE(1) PBX_MID_CASE(2) PBX_MID_CASE(3) PBX_MID_CASE(4) PBX_MID_CASE(5) PBX_MID_CASE(6)
        PBX_MID_CASE(7) PBX_MID_CASE(8) PBX_MID_CASE(9) PBX_MID_CASE(10) PBX_MID_CASE(11) PBX_MID_CASE(12)
        PBX_MID_CASE(13) PBX_MID_CASE(14) PBX_MID_CASE(15) PBX_MID_CASE(16)
#undef PBX_MID_CASE
    }
    return fail(PBX_ERR_UNSUPPORTED, "blocked kernels cover 1 <= A <= 16");
}

int launch_generic(pbx_plan* p, const double* R, long long n, double* out4, long long out_ld, cudaStream_t st) {
    // R: [n][N][P] device; intermediates in p->scratch (caller sized it)
    const HostTables& H = p->H;
    double* m_mat = (double*)p->scratch;
    double* o_vib = m_mat + (size_t)n * H.P * H.A * H.A;
    double* lr = o_vib + (size_t)3 * n * H.P * H.A;
    BeadOutputs bo{};
    bo.o_vib = o_vib; bo.lr = lr; bo.m_mat = m_mat; bo.n = n;
    if (use_mid_path(p)) return launch_mid(p, R, n, bo, out4, out_ld, st);
    const long long items = n * H.P;
    const unsigned grid_b = (unsigned)((items + GEN_WARPS - 1) / GEN_WARPS);
    const size_t sm_b = GEN_WARPS * bead_warp_doubles(H.A, H.Ar, H.N) * sizeof(double);
    auto kb = p->jacobi ? (p->scale ? pbx_bead_kernel<true, true> : pbx_bead_kernel<true, false>)
                        : (p->scale ? pbx_bead_kernel<false, true> : pbx_bead_kernel<false, false>);
    PBX_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
    kb<<<grid_b, GEN_WARPS * 32, sm_b, st>>>(p->D, R, n, bo);
    PBX_CUDA(cudaGetLastError());
    const unsigned grid_c = (unsigned)((n + GEN_WARPS - 1) / GEN_WARPS);
    const size_t sm_c = GEN_WARPS * chain_warp_doubles(H.A, H.Ar) * sizeof(double);
    if (p->pm) {
        PBX_CUDA(cudaFuncSetAttribute(pbx_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_c));
        pbx_chain_kernel<true><<<grid_c, GEN_WARPS * 32, sm_c, st>>>(p->D, m_mat, o_vib, lr, n, out4, out4 + out_ld, out_ld);
    } else {
        PBX_CUDA(cudaFuncSetAttribute(pbx_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_c));
        pbx_chain_kernel<false><<<grid_c, GEN_WARPS * 32, sm_c, st>>>(p->D, m_mat, o_vib, lr, n, out4, out4 + out_ld, out_ld);
    }
    PBX_CUDA(cudaGetLastError());
    p->launches += 2;
    return PBX_OK;
}

int launch_sample_coords(pbx_plan* p, unsigned long long seed, long long first, long long n, double* R, int* src,
                         cudaStream_t st) {
    const long long threads = n * ((p->H.N + 1) / 2);     // one thread per (sample, mode pair)
    pbx_mid_sample_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(p->D, seed, first, n, R, src);
    PBX_CUDA(cudaGetLastError());
    p->launches += 1;
    return PBX_OK;
}

long long generic_chunk(const pbx_plan* p, long long n, bool with_coords) {
    size_t per = generic_doubles_per_sample(p->H) + (with_coords ? (size_t)p->H.N * p->H.P : 0);
    long long c = (long long)(kScratchTarget / (per * sizeof(double)));
    c = std::max<long long>(c, 1);
    return std::min(c, n);
}

}  // namespace

extern "C" {

int pbx_abi_version(void) { return PBX_ABI_VERSION; }
int pbx_library_features(void) { return (PBX_WITH_MTAU ? PBX_FEATURE_MTAU : 0); }
const char* pbx_last_error(void) { return g_last_error.c_str(); }

int pbx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pbx_plan_create(const pbx_model* vib, const pbx_rho* rho, int32_t beads, double beta, double delta_beta,
                    uint32_t flags, int32_t device, pbx_plan** out) {
    if (!out) return fail(PBX_ERR_ARG, "out is null");
    *out = nullptr;
    pbx_plan* p = new (std::nothrow) pbx_plan();
    if (!p) return fail(PBX_ERR_ARG, "out of host memory");
    std::string err;
    int rc = build_tables(vib, rho, beads, beta, delta_beta, flags, p->H, err);
    if (rc != PBX_OK) { delete p; return fail(rc, err); }
    p->device = device; p->flags = flags;
    p->pm = flags & PBX_FLAG_PM; p->jacobi = flags & PBX_FLAG_EIG_JACOBI; p->scale = !(flags & PBX_FLAG_NO_SCALING);
    const bool want_fast = !(flags & PBX_FLAG_FORCE_GENERIC) && p->scale;
    if (want_fast) p->fast = find_fast_kernel(p->H.A, p->H.N, p->H.Ar);
    p->mtau = flags & PBX_FLAG_M_TAU_PM;
    // a run-time compiled shape carries the default variants only
    if (p->fast && ((p->jacobi && !(p->fast->caps & FAST_CAP_JACOBI)) || (p->mtau && !(p->fast->caps & FAST_CAP_MTAU)))) p->fast = nullptr;
    if (p->mtau && (!p->pm || p->jacobi || !p->fast)) {
        delete p;
        return fail(PBX_ERR_UNSUPPORTED, "PBX_FLAG_M_TAU_PM needs PBX_FLAG_PM, the default exp(-tau V) builder, a model shape "
                                         "with a register-resident kernel (csrc/shapes.def) and a library built with PBX_WITH_MTAU=1");
    }
    if (p->fast) {
        p->fast_tables.assign(p->fast->table_bytes, 0);
        p->fast->fill(p->H, p->fast_tables.data());
    }
    if (device == -1) {  // host-only plan: tables can be inspected (pbx_plan_table), nothing can be launched
        *out = p;
        return PBX_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        delete p;
        cudaGetLastError();
        return fail(PBX_ERR_CUDA, "no CUDA device available: pbx has no CPU fallback");
    }
    if (device < 0 || device >= ndev) { delete p; return fail(PBX_ERR_ARG, "device index out of range"); }
    DeviceGuard guard(device);
    if (!guard.ok) { delete p; return fail(PBX_ERR_CUDA, "cudaSetDevice failed"); }
    // shared-memory footprint of the generic kernels must fit an SM
    size_t sm_need = GEN_WARPS * std::max(bead_warp_doubles(p->H.A, p->H.Ar, p->H.N),
                                          chain_warp_doubles(p->H.A, p->H.Ar)) * sizeof(double);
    if (p->H.A <= MID_AMAX)
        sm_need = std::max({sm_need, MID_WARPS * mid_bead_warp_doubles(p->H.A, p->H.Ar, p->H.N) * sizeof(double),
                            MID_WARPS * mid_expm_warp_doubles(p->H.A) * sizeof(double)});
    if (sm_need > 200 * 1024) { delete p; return fail(PBX_ERR_UNSUPPORTED, "A too large for the generic kernels"); }
    rc = upload_tables(p);
    if (rc != PBX_OK) { pbx_plan_destroy(p); return rc; }
    if (p->H.A <= MID_AMAX) {
        rc = upload_dmma_tables(p);
        if (rc != PBX_OK) { pbx_plan_destroy(p); return rc; }
        // fused large-A kernel: shapes without a register-resident kernel (or on request), default M builder, scaled
        const bool want_big = (!p->fast || (flags & PBX_FLAG_PREFER_DMMA)) && !(flags & PBX_FLAG_FORCE_GENERIC) &&
                              !(flags & PBX_FLAG_NO_FUSED_DMMA) && !p->jacobi && p->scale && !p->mtau && p->H.Ar <= BIG_ARMAX;
        if (want_big) p->big = find_big_kernel(p->H.A);
        if (p->big) {
            rc = upload_big_tables(p);
            if (rc != PBX_OK) { pbx_plan_destroy(p); return rc; }
        }
        if (p->big) p->fast = nullptr;
    }
    e = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { pbx_plan_destroy(p); return cuda_fail(e, "cudaStreamCreate"); }
    *out = p;
    return PBX_OK;
}

int pbx_plan_destroy(pbx_plan* p) {
    if (!p) return PBX_OK;
    if (p->device < 0) { delete p; return PBX_OK; }
    DeviceGuard guard(p->device);
    cudaDeviceSynchronize();
    dev_free(p->dev_tables);
    dev_free(p->dev_int_tables);
    dev_free(p->dmma_tables);
    dev_free(p->big_tab);
    dev_free(p->scratch);
    dev_free(p->io);
    dev_free(p->stat_partials);
    dev_free(p->sums_scratch);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    delete p;
    return PBX_OK;
}

int64_t pbx_plan_table(const pbx_plan* p, const char* name, double* out, int64_t count) {
    if (!p || !name) return fail(PBX_ERR_ARG, "null argument");
    const HostTables& H = p->H;
    const std::vector<double>* v = nullptr;
    std::vector<double> tmp;
    const std::string s(name);
    if (s == "d_vib") v = &H.d_vib; else if (s == "d_rho") v = &H.d_rho; else if (s == "d_rho_eval") v = &H.d_rho_eval;
    else if (s == "delta_vib") v = &H.delta_vib; else if (s == "delta_rho") v = &H.delta_rho;
    else if (s == "weights") v = &H.weights; else if (s == "coth") v = &H.coth; else if (s == "csch") v = &H.csch;
    else if (s == "logpref") v = &H.logpref; else if (s == "logpref_rho") v = &H.logpref_rho;
    else if (s == "e_off") v = &H.e_off; else if (s == "l_off") v = &H.l_off; else if (s == "q_pack") v = &H.q_pack;
    else if (s == "samp") v = &H.samp;
    else if (s == "tau") { tmp.assign(H.tau, H.tau + 3); v = &tmp; }
    else return fail(PBX_ERR_ARG, "unknown table name: " + s);
    const int64_t avail = (int64_t)v->size();
    if (out && count > 0) std::memcpy(out, v->data(), (size_t)std::min(count, avail) * sizeof(double));
    return avail;
}

int pbx_has_register_kernel(int32_t A, int32_t N, int32_t A_rho) { return find_fast_kernel(A, N, A_rho) ? 1 : 0; }

int pbx_register_shape_library(const char* path) {
    if (!path) return fail(PBX_ERR_ARG, "null path");
    void* handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!handle) return fail(PBX_ERR_ARG, std::string("dlopen failed: ") + dlerror());
    typedef const FastKernelEntry* (*entry_fn)(void);
    entry_fn fn = (entry_fn)dlsym(handle, "pbx_jit_entry");
    if (!fn) { dlclose(handle); return fail(PBX_ERR_ARG, "pbx_jit_entry not found in the library"); }
    const FastKernelEntry* e = fn();
    if (!e || e->A < 1 || e->N < 1 || e->AR < 1) { dlclose(handle); return fail(PBX_ERR_ARG, "bad kernel entry"); }
    register_fast_kernel(e);       // already known: keep the first one (the handle stays open either way)
    return PBX_OK;
}

int pbx_plan_is_fast(const pbx_plan* p) { return (p && p->fast) ? 1 : 0; }
int pbx_plan_kernel_path(const pbx_plan* p) {
    if (!p) return PBX_ERR_ARG;
    if (p->fast) return PBX_PATH_REGISTER;
    if (p->big) return PBX_PATH_FUSED_DMMA;
    return use_mid_path(p) ? PBX_PATH_BLOCKED : PBX_PATH_GENERIC;
}
int64_t pbx_plan_launch_count(const pbx_plan* p) { return p ? p->launches : 0; }
int64_t pbx_plan_launch_param_bytes(const pbx_plan* p) {
    if (!p) return 0;
    if (p->big) return (int64_t)sizeof(BigParams);
    return p->fast ? (int64_t)(p->fast->table_bytes + sizeof(FastLaunch)) : (int64_t)sizeof(DevTables);
}

}  // extern "C"

namespace {
// fused sampler + estimator into out4 [4][n] (device); `mirror` (nullable, register-resident kernels only) is a
// device-accessible HOST copy [4][mirror_ld] the kernel writes as well
int sample_eval_impl(pbx_plan* p, uint64_t seed, int64_t first_sample, int64_t n, double* out4, double* mirror,
                     int64_t mirror_ld, cudaStream_t st) {
    if (p->fast) {
        FastLaunch L{};
        L.samp = p->D.samp; L.seed = seed; L.first_sample = first_sample; L.n_samples = n; L.out4 = out4; L.out_ld = n;
        L.mirror = mirror; L.mirror_ld = mirror_ld;
        if (!p->jacobi && !(p->flags & PBX_FLAG_NO_WARPSPEC)) {   // producer/consumer warps (pbx_fast_ws.cuh)
            PBX_CUDA(p->fast->launch_ws(p->fast_tables.data(), L, p->pm, p->H.rho_shares_vib, p->mtau, st));
            p->launches += p->fast->ws_kernels(p->pm, p->mtau);
            return PBX_OK;
        }
        PBX_CUDA(p->fast->launch(p->fast_tables.data(), L, MODE_SAMPLE, p->pm, p->jacobi, p->H.rho_shares_vib, p->mtau, st));
        p->launches += 1;
        return PBX_OK;
    }
    if (p->big && p->H.N <= BIG_NMAX_SAMPLER)      // sampler fused in: no scratch at all
        return launch_big(p, nullptr, seed, first_sample, n, out4, n, mirror, mirror_ld, st);
    if (p->big) {                                  // more modes than lanes: coordinates through a scratch buffer, in chunks
        const size_t coords = (size_t)p->H.N * p->H.P;
        const long long chunk = std::min<long long>(n, std::max<long long>(1, (long long)(kScratchTarget / (coords * sizeof(double)))));
        int rc = ensure(&p->scratch, &p->scratch_bytes, (size_t)chunk * coords * sizeof(double));
        if (rc != PBX_OK) return rc;
        for (long long off = 0; off < n; off += chunk) {
            const long long m = std::min<long long>(chunk, n - off);
            rc = launch_sample_coords(p, seed, first_sample + off, m, (double*)p->scratch, nullptr, st);
            if (rc != PBX_OK) return rc;
            rc = launch_big(p, (const double*)p->scratch, 0, 0, m, out4 + off, n, mirror ? mirror + off : nullptr, mirror_ld, st);
            if (rc != PBX_OK) return rc;
        }
        return PBX_OK;
    }
    const long long chunk = generic_chunk(p, n, true);
    const size_t per = generic_doubles_per_sample(p->H), coords = (size_t)p->H.N * p->H.P;
    int rc = ensure(&p->scratch, &p->scratch_bytes, (size_t)chunk * (per + coords) * sizeof(double));
    if (rc != PBX_OK) return rc;
    double* R = (double*)p->scratch + (size_t)chunk * per;
    for (long long off = 0; off < n; off += chunk) {
        const long long m = std::min<long long>(chunk, n - off);
        rc = launch_sample_coords(p, seed, first_sample + off, m, R, nullptr, st);
        if (rc != PBX_OK) return rc;
        rc = launch_generic(p, R, m, out4 + off, n, st);
        if (rc != PBX_OK) return rc;
    }
    return PBX_OK;
}

// device-visible alias of a host pointer if the range is pinned and mapped (cudaHostAlloc / cudaHostRegister), else null
double* mapped_alias(double* host, size_t bytes) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    cudaPointerAttributes end{};
    if (cudaPointerGetAttributes(&end, (char*)host + bytes - 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (end.type != cudaMemoryTypeHost || !end.devicePointer) return nullptr;
    return (double*)at.devicePointer;
}
}  // namespace

extern "C" {

int pbx_sample_eval_dev(pbx_plan* p, uint64_t seed, int64_t first_sample, int64_t n, double* out4, void* stream) {
    if (!p || !out4) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0) return PBX_OK;
    if (first_sample < 0) return fail(PBX_ERR_ARG, "first_sample < 0");
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    return sample_eval_impl(p, seed, first_sample, n, out4, nullptr, 0, (cudaStream_t)stream);
}

int pbx_eval_coords_dev(pbx_plan* p, const double* R, int64_t n, double* out4, void* stream) {
    if (!p || !R || !out4) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0) return PBX_OK;
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const HostTables& H = p->H;
    if (p->fast) {      // one launch, the caller's R read in place
        FastLaunch L{};
        L.samp = p->D.samp; L.coords = R; L.n_samples = n;
        L.out4 = out4; L.out_ld = n;
        PBX_CUDA(p->fast->launch(p->fast_tables.data(), L, MODE_COORDS, p->pm, p->jacobi, p->H.rho_shares_vib, p->mtau, st));
        p->launches += 1;
        return PBX_OK;
    }
    if (p->big) return launch_big(p, R, 0, 0, n, out4, n, nullptr, 0, st);      // reads the caller's R in place
    const long long chunk = generic_chunk(p, n, false);
    int rc = ensure(&p->scratch, &p->scratch_bytes, (size_t)chunk * generic_doubles_per_sample(H) * sizeof(double));
    if (rc != PBX_OK) return rc;
    for (long long off = 0; off < n; off += chunk) {
        const long long m = std::min<long long>(chunk, n - off);
        rc = launch_generic(p, R + (size_t)off * H.N * H.P, m, out4 + off, n, st);
        if (rc != PBX_OK) return rc;
    }
    return PBX_OK;
}

int pbx_sample_coords_dev(pbx_plan* p, uint64_t seed, int64_t first_sample, int64_t n, double* R, int32_t* src,
                          void* stream) {
    if (!p || !R) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0) return PBX_OK;
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    return launch_sample_coords(p, seed, first_sample, n, R, src, (cudaStream_t)stream);
}

int pbx_eval_stages_dev(pbx_plan* p, const double* R, int64_t n, double* o_rho, double* o_vib, double* scale,
                        double* v_mat, double* m_mat, void* stream) {
    if (!p || !R) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0) return PBX_OK;
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const HostTables& H = p->H;
    BeadOutputs bo{};
    bo.o_rho = o_rho; bo.o_vib = o_vib; bo.scale = scale; bo.v_mat = v_mat; bo.m_mat = m_mat; bo.n = n;
    const long long items = n * H.P;
    const size_t sm_b = GEN_WARPS * bead_warp_doubles(H.A, H.Ar, H.N) * sizeof(double);
    auto kb = p->jacobi ? (p->scale ? pbx_bead_kernel<true, true> : pbx_bead_kernel<true, false>)
                        : (p->scale ? pbx_bead_kernel<false, true> : pbx_bead_kernel<false, false>);
    PBX_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
    kb<<<(unsigned)((items + GEN_WARPS - 1) / GEN_WARPS), GEN_WARPS * 32, sm_b, st>>>(p->D, R, n, bo);
    PBX_CUDA(cudaGetLastError());
    p->launches += 1;
    return PBX_OK;
}

int pbx_chain_trace_dev(pbx_plan* p, const double* m_mat, const double* o_diag, int64_t n, double* g_out,
                        void* stream) {
    if (!p || !m_mat || !o_diag || !g_out) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0) return PBX_OK;
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    const size_t sm_c = GEN_WARPS * chain_warp_doubles(p->H.A, p->H.Ar) * sizeof(double);
    PBX_CUDA(cudaFuncSetAttribute(pbx_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_c));
    pbx_chain_kernel<false><<<(unsigned)((n + GEN_WARPS - 1) / GEN_WARPS), GEN_WARPS * 32, sm_c, (cudaStream_t)stream>>>(
        p->D, m_mat, o_diag, nullptr, n, nullptr, g_out, n);
    PBX_CUDA(cudaGetLastError());
    p->launches += 1;
    return PBX_OK;
}

int pbx_block_sums_dev(pbx_plan* p, const double* out4, int64_t n, int64_t block_size, double* sums, void* stream) {
    if (!p || !out4 || !sums) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0 || block_size <= 0) return fail(PBX_ERR_ARG, "n_samples and block_size must be positive");
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    const long long blocks = (n + block_size - 1) / block_size;
    // enough CTAs to fill the GPU (4 per SM), slices of at least 256 samples
    long long split = std::min<long long>({32, (4 * 148 + blocks - 1) / blocks, std::max<long long>(1, block_size / 256)});
    double* partials = nullptr;
    int* tickets = nullptr;
    if (split > 1) {
        const size_t part_bytes = (size_t)blocks * split * PBX_NSUMS * sizeof(double), need = part_bytes + (size_t)blocks * sizeof(int);
        int rc = ensure(&p->sums_scratch, &p->sums_scratch_bytes, need);
        if (rc != PBX_OK) return rc;
        PBX_CUDA(cudaMemsetAsync((char*)p->sums_scratch + part_bytes, 0, (size_t)blocks * sizeof(int), (cudaStream_t)stream));
        partials = (double*)p->sums_scratch;
        tickets = (int*)((char*)p->sums_scratch + part_bytes);
    }
    pbx_block_sums_kernel<<<dim3((unsigned)blocks, (unsigned)split), 256, 0, (cudaStream_t)stream>>>(
        out4, n, n, block_size, p->H.delta_beta, p->pm ? 1 : 0, (int)split, partials, tickets, sums);
    PBX_CUDA(cudaGetLastError());
    p->launches += 1;
    return PBX_OK;
}

int pbx_sample_eval_host(pbx_plan* p, uint64_t seed, int64_t first_sample, int64_t n, double* out4_host,
                         int64_t ld_host, int64_t block_size, double* sums_host) {
    if (!p || !out4_host) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0) return PBX_OK;
    if (ld_host < n) return fail(PBX_ERR_ARG, "ld_host < n_samples");
    if (sums_host && block_size <= 0) return fail(PBX_ERR_ARG, "block_size must be positive");
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    const int64_t blocks = sums_host ? (n + block_size - 1) / block_size : 0;
    int rc = ensure(&p->io, &p->io_bytes, ((size_t)4 * n + (size_t)blocks * PBX_NSUMS) * sizeof(double));
    if (rc != PBX_OK) return rc;
    double* out_dev = (double*)p->io;
    double* sums_dev = out_dev + (size_t)4 * n;
    p->io_samples = n;
    if (first_sample < 0) return fail(PBX_ERR_ARG, "first_sample < 0");
    const size_t rows = p->pm ? 4 : 2;
    // pinned + mapped host buffer and a register-resident kernel: the kernel writes the host rows itself
    double* mirror = (p->fast || p->big) ? mapped_alias(out4_host, ((rows - 1) * (size_t)ld_host + (size_t)n) * sizeof(double)) : nullptr;
    rc = sample_eval_impl(p, seed, first_sample, n, out_dev, mirror, ld_host, p->own_stream);
    if (rc != PBX_OK) return rc;
    if (sums_host) {
        rc = pbx_block_sums_dev(p, out_dev, n, block_size, sums_dev, p->own_stream);
        if (rc != PBX_OK) return rc;
    }
    if (!mirror)
        PBX_CUDA(cudaMemcpy2DAsync(out4_host, ld_host * sizeof(double), out_dev, n * sizeof(double), n * sizeof(double), rows,
                                   cudaMemcpyDeviceToHost, p->own_stream));
    if (sums_host)
        PBX_CUDA(cudaMemcpyAsync(sums_host, sums_dev, (size_t)blocks * PBX_NSUMS * sizeof(double), cudaMemcpyDeviceToHost,
                                 p->own_stream));
    PBX_CUDA(cudaStreamSynchronize(p->own_stream));
    return PBX_OK;
}

int pbx_eval_coords_host(pbx_plan* p, const double* R_host, int64_t n, double* out4_host, int64_t ld_host) {
    if (!p || !R_host || !out4_host) return fail(PBX_ERR_ARG, "null argument");
    if (n <= 0) return PBX_OK;
    if (ld_host < n) return fail(PBX_ERR_ARG, "ld_host < n_samples");
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    const size_t np = (size_t)p->H.N * p->H.P;
    const size_t out_bytes = (size_t)4 * n * sizeof(double);
    int rc = ensure(&p->io, &p->io_bytes, out_bytes + n * np * sizeof(double));
    if (rc != PBX_OK) return rc;
    double* out_dev = (double*)p->io;
    double* R_dev = out_dev + (size_t)4 * n;
    p->io_samples = n;
    PBX_CUDA(cudaMemcpyAsync(R_dev, R_host, n * np * sizeof(double), cudaMemcpyHostToDevice, p->own_stream));
    rc = pbx_eval_coords_dev(p, R_dev, n, out_dev, p->own_stream);
    if (rc != PBX_OK) return rc;
    const size_t rows = p->pm ? 4 : 2;
    PBX_CUDA(cudaMemcpy2DAsync(out4_host, ld_host * sizeof(double), out_dev, n * sizeof(double), n * sizeof(double), rows,
                               cudaMemcpyDeviceToHost, p->own_stream));
    PBX_CUDA(cudaStreamSynchronize(p->own_stream));
    return PBX_OK;
}

}  // extern "C"

namespace {
// Z, E, Cv and their leave-one-out jackknife from a device [4][n] array; needs only (beta, delta_beta).
// `partials` is a device buffer of kStatGrid * 4 doubles.  Synchronises `st`.
int stats_core(const double* out4, int64_t n, double beta, double db, double* partials, double* stats_host, cudaStream_t st) {
    const double inv2db = 1.0 / (2.0 * db), invdb2 = 1.0 / (db * db);
    const double kB = 1.38064852e-23 / 1.6021766208e-19;        // pibronic/constants.py:12,24
    const double T = 1.0 / (kB * beta), kbt2 = kB * T * T;
    double part[kStatGrid * 4];
    auto total = [&](int col) {
        long double acc = 0.0L;
        for (int b = 0; b < kStatGrid; ++b) acc += part[b * 4 + col];
        return acc;
    };
    pbx_stat_sums_kernel<<<kStatGrid, 256, 0, st>>>(out4, n, n, inv2db, invdb2, partials);
    PBX_CUDA(cudaGetLastError());
    PBX_CUDA(cudaMemcpyAsync(part, partials, sizeof(part), cudaMemcpyDeviceToHost, st));
    PBX_CUDA(cudaStreamSynchronize(st));
    const long double X = (long double)n, S_r = total(0), S_rr = total(1), S_1 = total(2), S_2 = total(3);
    const long double Z = S_r / X;
    long double var = S_rr / X - Z * Z;
    if (var < 0) var = 0;
    const long double E = -(S_1 / X) / Z;
    const long double Cv = ((S_2 / X) / Z - E * E) / kbt2;
    pbx_jackknife_kernel<<<kStatGrid, 256, 0, st>>>(out4, n, n, inv2db, invdb2, (double)S_r, (double)S_1, (double)S_2,
                                                   1.0 / kbt2, (double)E, (double)Cv, partials);
    PBX_CUDA(cudaGetLastError());
    PBX_CUDA(cudaMemcpyAsync(part, partials, sizeof(part), cudaMemcpyDeviceToHost, st));
    PBX_CUDA(cudaStreamSynchronize(st));
    const long double sE = total(0), sEE = total(1), sC = total(2), sCC = total(3);
    const long double mean_fE = (long double)(double)E + sE / X, mean_fC = (long double)(double)Cv + sC / X;
    long double var_fE = (sEE - sE * sE / X) / (X - 1), var_fC = (sCC - sC * sC / X) / (X - 1);
    if (var_fE < 0) var_fE = 0;
    if (var_fC < 0) var_fC = 0;
    stats_host[0] = (double)Z;
    stats_host[1] = (double)(std::sqrt(var) / std::sqrt(X - 1));
    stats_host[2] = (double)E;  stats_host[3] = 0.0;
    stats_host[4] = (double)Cv; stats_host[5] = 0.0;
    stats_host[6] = (double)(X * E - (X - 1) * mean_fE);
    stats_host[7] = (double)(std::sqrt(X - 1) * std::sqrt(var_fE));
    stats_host[8] = (double)(X * Cv - (X - 1) * mean_fC);
    stats_host[9] = (double)(std::sqrt(X - 1) * std::sqrt(var_fC));
    return PBX_OK;
}
}  // namespace

extern "C" {

int pbx_stats_dev(pbx_plan* p, const double* out4, int64_t n, double* stats_host, void* stream) {
    if (!p || !out4 || !stats_host) return fail(PBX_ERR_ARG, "null argument");
    if (n < 2) return fail(PBX_ERR_ARG, "statistics need at least 2 samples");
    if (!p->pm) return fail(PBX_ERR_ARG, "statistics need a PBX_FLAG_PM plan (g+ and g-)");
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    if (!p->stat_partials) PBX_CUDA(dev_alloc((void**)&p->stat_partials, kStatGrid * 4 * sizeof(double)));
    p->launches += 2;
    return stats_core(out4, n, p->H.beta, p->H.delta_beta, p->stat_partials, stats_host, (cudaStream_t)stream);
}

// plan-free forms: statistics need (beta, delta_beta) only
int pbx_stats_arrays_dev(const double* out4, int64_t n, double beta, double delta_beta, int32_t device, double* stats_host,
                         void* stream) {
    if (!out4 || !stats_host) return fail(PBX_ERR_ARG, "null argument");
    if (n < 2) return fail(PBX_ERR_ARG, "statistics need at least 2 samples");
    if (!(beta > 0) || !(delta_beta > 0)) return fail(PBX_ERR_ARG, "need beta > 0 and delta_beta > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(PBX_ERR_CUDA, "no CUDA device available: pbx has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(PBX_ERR_ARG, "device index out of range");
    DeviceGuard guard(device);
    double* partials = nullptr;
    PBX_CUDA(dev_alloc((void**)&partials, kStatGrid * 4 * sizeof(double)));
    const int rc = stats_core(out4, n, beta, delta_beta, partials, stats_host, (cudaStream_t)stream);   // returns after its D2H copy
    dev_free(partials);
    return rc;
}

int pbx_stats_arrays_host(const double* out4_host, int64_t ld_host, int64_t n, double beta, double delta_beta, int32_t device,
                          double* stats_host) {
    if (!out4_host || !stats_host) return fail(PBX_ERR_ARG, "null argument");
    if (n < 2 || ld_host < n) return fail(PBX_ERR_ARG, "need n >= 2 and ld_host >= n");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(PBX_ERR_CUDA, "no CUDA device available: pbx has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(PBX_ERR_ARG, "device index out of range");
    DeviceGuard guard(device);
    double* dev = nullptr;
    PBX_CUDA(dev_alloc((void**)&dev, (size_t)4 * n * sizeof(double)));
    cudaError_t e = cudaMemcpy2D(dev, n * sizeof(double), out4_host, ld_host * sizeof(double), n * sizeof(double), 4, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? pbx_stats_arrays_dev(dev, n, beta, delta_beta, device, stats_host, nullptr) : cuda_fail(e, "cudaMemcpy2D");
    dev_free(dev);
    return rc;
}

int pbx_stats_host(pbx_plan* p, const double* out4_host, int64_t ld_host, int64_t n, double* stats_host) {
    if (!p || !out4_host || !stats_host) return fail(PBX_ERR_ARG, "null argument");
    if (n < 2 || ld_host < n) return fail(PBX_ERR_ARG, "need n >= 2 and ld_host >= n");
    PBX_NEED_DEVICE(p);
    DeviceGuard guard(p->device);
    int rc = ensure(&p->io, &p->io_bytes, (size_t)4 * n * sizeof(double));
    if (rc != PBX_OK) return rc;
    PBX_CUDA(cudaMemcpy2DAsync(p->io, n * sizeof(double), out4_host, ld_host * sizeof(double), n * sizeof(double), 4,
                               cudaMemcpyHostToDevice, p->own_stream));
    p->io_samples = n;
    return pbx_stats_dev(p, (const double*)p->io, n, stats_host, p->own_stream);
}

int pbx_stats_last(pbx_plan* p, double* stats_host) {
    if (!p || !stats_host) return fail(PBX_ERR_ARG, "null argument");
    PBX_NEED_DEVICE(p);
    if (p->io_samples < 2) return fail(PBX_ERR_ARG, "no device-resident results: call a *_host entry point first");
    return pbx_stats_dev(p, (const double*)p->io, p->io_samples, stats_host, p->own_stream);
}

int pbx_math_probe_dev(int32_t kind, const double* in_dev, double* out_dev, int64_t n, void* stream) {
    if (!in_dev || !out_dev || n <= 0 || kind < 0 || kind > 4) return fail(PBX_ERR_ARG, "bad argument");
    pbx_math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, in_dev, out_dev, n);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // extern "C"

namespace {
int fp64_peak(int32_t device, int which, double* tflops_out) {
    if (!tflops_out) return fail(PBX_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(PBX_ERR_CUDA, "no such CUDA device");
    }
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    PBX_CUDA(cudaGetDeviceProperties(&prop, device));
    double* dummy = nullptr;
    PBX_CUDA(cudaMalloc((void**)&dummy, 8));
    const int grid = prop.multiProcessorCount * 8, iters = 1 << 15;
    cudaEvent_t e0, e1;
    PBX_CUDA(cudaEventCreate(&e0)); PBX_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int kind = 0; kind < 2; ++kind) {
        if (kind != which && which != 2) continue;
        for (int rep = 0; rep < 5; ++rep) {
            PBX_CUDA(cudaEventRecord(e0));
            if (kind == 0) pbx_dfma_probe_kernel<<<grid, 256>>>(dummy, iters, 0.999999, 1e-7);
            else pbx_dmma_probe_kernel<<<grid, 256>>>(dummy, iters / 8, 0.999999, 1e-7);
            PBX_CUDA(cudaEventRecord(e1));
            PBX_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            PBX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            // per thread and iteration: 8 FMAs (vector) or 8 tiles x 256 FMAs / 32 lanes (tensor), over iters/8 iterations
            const double flops = 2.0 * 8.0 * iters * 256.0 * grid;
            if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dummy);
    *tflops_out = best;
    return PBX_OK;
}
}  // namespace

extern "C" {

int pbx_fp64_peak_tflops(int32_t device, double* tflops_out) { return fp64_peak(device, 0, tflops_out); }
int pbx_fp64_peak_tflops_kind(int32_t device, int32_t kind, double* tflops_out) {
    if (kind < 0 || kind > 2) return fail(PBX_ERR_ARG, "kind must be 0 (vector DFMA), 1 (tensor DMMA) or 2 (the larger)");
    return fp64_peak(device, kind, tflops_out);
}

}  // extern "C"

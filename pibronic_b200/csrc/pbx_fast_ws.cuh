// Warp-specialised form of the register-resident small-A kernel (fused sampler + estimator).
//
// In the one-role kernel (pbx_fast.cuh) every thread alternates between a sampler phase (Philox,
// Box-Muller: integer and conversion heavy, long dependent chains) and an estimator phase (dense
// FP64).  With 248 registers per thread only two warps fit a scheduler, and whenever both sit in the
// sampler phase the FP64 pipe idles: ncu showed the sampler phase taking 37 % of the time for 15 % of
// the FP64 instructions (profiles/r01_summary.md).  Here the two phases run in different warps of
// one 512-thread CTA per SM (launched at 128 registers per thread, then redistributed):
//
//   warpgroups 0,1 (8 producer warps, setmaxnreg.dec -> PBX_WS_REG_PROD registers)
//                 standard normals of one sample per thread, WS_TB beads per stage, into a
//                 shared-memory ring [stage][bead][mode][sample];
//   warpgroups 2,3 (8 consumer warps, setmaxnreg.inc -> PBX_WS_REG_CONS registers)
//                 one sample per thread: ring recurrence z -> bead coordinates (O(N) state), the
//                 harmonic factors, V, exp(-tau V), the three chained products -- bead_step of
//                 pbx_fast.cuh, unchanged.
//
// Hand-over: mbarrier pairs full/empty per (group of 64 samples, stage); producer warps 2g, 2g+1 feed exactly
// consumer warps 2g, 2g+1, so warps only ever wait for their own partners, not for the CTA.
// Philox counters and every floating-point expression are those of the one-role kernel: results are
// bit-identical (tests/test_gpu_parity.py::test_warp_specialised_kernel_is_bit_identical).
//
// Samples whose tau+- exponent differences leave the range of the short exp series are flagged
// with rho = NaN and recomputed by pbx_fast_kernel<MODE_REDO> (full exponentials) right after.
#pragma once
#include "pbx_fast.cuh"

#ifndef PBX_WS_TB
#define PBX_WS_TB 4          // beads per ring stage
#endif
#ifndef PBX_WS_STAGES
#define PBX_WS_STAGES 2
#endif
#ifndef PBX_WS_PROD_WGS
#define PBX_WS_PROD_WGS 2    // producer warpgroups: 1 = two samples per producer thread, 2 = one
#endif
#ifndef PBX_WS_CTAS
#define PBX_WS_CTAS 0        // CTAs per SM; 0 = chosen per shape (ws_ctas)
#endif
#ifndef PBX_WS_PARK_NS
#define PBX_WS_PARK_NS 400   // sleep of a waiting producer between two polls of its "empty" barrier, ns
#endif
#ifndef PBX_WS_YP_SMEM
#define PBX_WS_YP_SMEM 0     // 1: previous bead (recurrence state) in shared memory instead of registers
#endif

namespace pbx {

constexpr int WS_TB = PBX_WS_TB, WS_STAGES = PBX_WS_STAGES;
constexpr int WS_PROD = 128 * PBX_WS_PROD_WGS, WS_CONS = 256, WS_THREADS = WS_PROD + WS_CONS;
constexpr int WS_SPT = WS_CONS / WS_PROD;   // samples per producer thread
constexpr int WS_GROUPS = 4;                // barrier groups: 64 consumers (2 warps) and the producer warps feeding them
constexpr int WS_GROUP_PROD = WS_PROD / WS_GROUPS;   // producer threads per group

// Register budget per shape.  The 64 K registers of an SM are split between the CTAs resident on it; inside a CTA the
// producer warpgroups hand registers to the consumer warpgroups (setmaxnreg).  A = 2 needs < 96 registers per consumer
// thread: two CTAs per SM (16 consumer warps) hide the latency of its short dependent chains.
template <int A, int N, int AR>
constexpr int ws_ctas() { return PBX_WS_CTAS ? PBX_WS_CTAS : (A <= 2 ? 2 : 1); }
template <int A, int N, int AR>
constexpr int ws_reg_prod() {
#ifdef PBX_WS_REG_PROD
    return PBX_WS_REG_PROD;
#else
    return PBX_WS_PROD_WGS == 1 ? 56 : 32;
#endif
}
// registers per thread the CTA is launched with (what __launch_bounds__(WS_THREADS, ctas) makes ptxas use): its pool
template <int A, int N, int AR>
constexpr int ws_reg_launch() { return 65536 / (WS_THREADS * ws_ctas<A, N, AR>()) / 8 * 8; }
template <int A, int N, int AR>
constexpr int ws_reg_cons() {
#ifdef PBX_WS_REG_CONS
    return PBX_WS_REG_CONS;
#else
    // what the producers give up goes to the consumers, in the allocation unit of 8 registers; never more than the pool
    // of the CTA holds (setmaxnreg.inc would wait for ever)
    return (WS_THREADS * ws_reg_launch<A, N, AR>() - WS_PROD * ws_reg_prod<A, N, AR>()) / WS_CONS / 8 * 8;
#endif
}
static_assert(WS_PROD * ws_reg_prod<4, 6, 4>() + WS_CONS * ws_reg_cons<4, 6, 4>() <= WS_THREADS * ws_reg_launch<4, 6, 4>(), "register pool");
static_assert(WS_PROD * ws_reg_prod<2, 2, 2>() + WS_CONS * ws_reg_cons<2, 2, 2>() <= WS_THREADS * ws_reg_launch<2, 2, 2>(), "register pool");

template <int N>
constexpr size_t ws_smem_bytes() {
    // ring of normals + per-sample columns y0 and dsrc + the barriers
    return ((size_t)WS_STAGES * WS_TB * N + (2 + PBX_WS_YP_SMEM) * N) * WS_CONS * sizeof(double) +
           (size_t)WS_GROUPS * WS_STAGES * 2 * 8;
}

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(ws_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(ws_smem_u32(bar)) : "memory");
}
// producer side: between two polls the warp sleeps, so that idle producers (they run ahead of the consumers) do not
// take issue slots from the consumer warps of their scheduler (ncu counted 1e9 try_wait instructions per launch)
__device__ __forceinline__ void ws_mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WSP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra WSP_DONE;\n"
        "nanosleep.u32 %3;\n"
        "bra WSP_WAIT;\n"
        "WSP_DONE:\n"
        "}\n" ::"r"(ws_smem_u32(bar)), "r"(parity), "r"(PBX_WS_PARK_NS), "r"(PBX_WS_PARK_NS) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WS_DONE;\n"
        "bra WS_WAIT;\n"
        "WS_DONE:\n"
        "}\n" ::"r"(ws_smem_u32(bar)), "r"(parity) : "memory");
}

template <int A, int N, int AR, bool PM, bool SHARE, bool MTAU = false>
__global__ void __launch_bounds__(WS_THREADS, (ws_ctas<A, N, AR>()))
pbx_fast_ws_kernel(const __grid_constant__ FastTables<A, N, AR> T, const FastLaunch L) {
    constexpr int NV = PM ? 3 : 1;
    constexpr int H = (N + 1) / 2;
    extern __shared__ __align__(16) double smem[];
    double* ring = smem;                                              // [STAGES][TB][N][CONS] standard normals
    double* y0s = ring + (size_t)WS_STAGES * WS_TB * N * WS_CONS;     // [N][CONS] first bead, relative to the shift
    double* dss = y0s + (size_t)N * WS_CONS;                          // [N][CONS] shift of the drawn mixture component
    double* yps = dss + (size_t)N * WS_CONS;                          // [N][CONS] previous bead (PBX_WS_YP_SMEM)
    uint64_t* bars = reinterpret_cast<uint64_t*>(yps + (size_t)PBX_WS_YP_SMEM * N * WS_CONS);
    auto full_bar = [&](int g, int s) { return bars + (g * WS_STAGES + s) * 2; };
    auto empty_bar = [&](int g, int s) { return bars + (g * WS_STAGES + s) * 2 + 1; };

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int g = 0; g < WS_GROUPS; ++g)
            for (int s = 0; s < WS_STAGES; ++s) {
                ws_mbar_init(full_bar(g, s), WS_GROUP_PROD);   // every producer thread of the group arrives
                ws_mbar_init(empty_bar(g, s), 64);             // every lane of its two consumer warps arrives
            }
    }
    __syncthreads();
    const int P = T.P;
    const int n_stage = (P + WS_TB - 1) / WS_TB;
    const uint2 key = make_uint2((uint32_t)L.seed, (uint32_t)(L.seed >> 32));
    // samples travel in groups of 64 (two producer warps + two consumer warps with their own barriers).  CTAs of the
    // complete waves take four groups; the last, partial wave is spread over all SMs with fewer groups per CTA -- the
    // warps of an absent group leave at once and the remaining ones have the schedulers to themselves
    const long long cta = blockIdx.x;
    const bool tail = cta >= L.ws_full_ctas;
    const long long first_group = tail ? 4 * L.ws_full_ctas + (cta - L.ws_full_ctas) * L.ws_tail_groups : 4 * cta;
    const int n_groups = tail ? L.ws_tail_groups : WS_GROUPS;
    const long long cta_first = 64 * first_group;

    if (tid < WS_PROD) {
        // ------------------------------------------------------------------ producer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(ws_reg_prod<A, N, AR>()));
        const int g = tid / WS_GROUP_PROD, t = tid - g * WS_GROUP_PROD;
        if (g >= n_groups || cta_first + 64 * g >= L.n_samples) return;
        int col[WS_SPT];
        unsigned long long gidx[WS_SPT];
#pragma unroll
        for (int q = 0; q < WS_SPT; ++q) {
            col[q] = 64 * g + WS_GROUP_PROD * q + t;
            long long x = cta_first + col[q];
            if (x >= L.n_samples) x = L.n_samples - 1;
            gidx[q] = (unsigned long long)(L.first_sample + x);
        }
        for (int k = 0; k < n_stage; ++k) {
            const int s = k % WS_STAGES;
            if (k >= WS_STAGES) ws_mbar_wait_parked(empty_bar(g, s), (uint32_t)((k / WS_STAGES - 1) & 1));
            double* st = ring + (size_t)s * WS_TB * N * WS_CONS;
#pragma unroll 1
            for (int jj = 0; jj < WS_TB; ++jj) {
                const int j = k * WS_TB + jj;
                if (j < P) {
#pragma unroll
                    for (int h = 0; h < H; ++h)
#pragma unroll
                        for (int q = 0; q < WS_SPT; ++q) {     // the thread's samples interleave
                            const uint4 r = philox4x32_10(make_uint4((uint32_t)gidx[q], (uint32_t)(gidx[q] >> 32),
                                                                     (uint32_t)(j * H + h), STREAM_NORMALS), key);
                            double z0, z1;
                            normal_pair(r, z0, z1);
                            st[(size_t)(jj * N + 2 * h) * WS_CONS + col[q]] = z0;
                            if (2 * h + 1 < N) st[(size_t)(jj * N + 2 * h + 1) * WS_CONS + col[q]] = z1;
                        }
                }
            }
            ws_mbar_arrive(full_bar(g, s));
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(ws_reg_cons<A, N, AR>()));
    const int c = tid - WS_PROD, g = c >> 6;
    if (g >= n_groups || cta_first + 64 * g >= L.n_samples) return;
    long long x = cta_first + c;
    const bool live = x < L.n_samples;
    if (!live) x = L.n_samples - 1;
    const unsigned long long gidx = (unsigned long long)(L.first_sample + x);
    double* y0 = y0s + c;
    double* dsrc = dss + c;
    {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, STREAM_SOURCE), key);
        const int src = pick_source<AR>(u01_half_open(r.x, r.y), T.wcum);
#pragma unroll
        for (int n = 0; n < N; ++n) {
            double d = T.d_rho[0][n];
#pragma unroll
            for (int a = 1; a < AR; ++a) d = (src == a) ? T.d_rho[a][n] : d;
            dsrc[n * WS_CONS] = d;
        }
    }
    bool bad = false;
    double Tm[NV][A][A];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < A; ++i)
#pragma unroll
            for (int j = 0; j < A; ++j) Tm[v][i][j] = (i == j) ? 1.0 : 0.0;
    double lrho[AR];
#pragma unroll
    for (int a = 0; a < AR; ++a) lrho[a] = 0.0;

#if PBX_WS_YP_SMEM
    double* yp = yps + c;
#define PBX_YP(n) yp[(n) * WS_CONS]
#else
    double yp[N];      // y_{j-1}: previous bead relative to the shift
#define PBX_YP(n) yp[n]
#endif
    ws_mbar_wait(full_bar(g, 0), 0u);
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const double y = __ldg(L.samp + n * 3) * ring[(size_t)n * WS_CONS + c];
        PBX_YP(n) = y;
        y0[n * WS_CONS] = y;
    }
    for (int k = 0; k < n_stage; ++k) {
        const int s = k % WS_STAGES;
        if (k > 0) ws_mbar_wait(full_bar(g, s), (uint32_t)((k / WS_STAGES) & 1));
        const double* st = ring + (size_t)s * WS_TB * N * WS_CONS + c;
#pragma unroll 1
        for (int jj = (k == 0) ? 1 : 0; jj < WS_TB; ++jj) {
            const int j = k * WS_TB + jj;
            if (j >= P) break;
            // bead j from its normals (cyclic-tridiagonal Cholesky recurrence), then the estimator step of bead j-1
            const double* tab = L.samp + (size_t)j * N * 3;
            double Rc[N], Rn[N];
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const double a = __ldg(tab + n * 3 + 0), b = __ldg(tab + n * 3 + 1), e = __ldg(tab + n * 3 + 2);
                const double ds = dsrc[n * WS_CONS], yprev = PBX_YP(n);
                Rc[n] = yprev + ds;
                double y = a * st[(size_t)(jj * N + n) * WS_CONS];
                y = fma(b, yprev, fma(e, y0[n * WS_CONS], y));
                PBX_YP(n) = y;
                Rn[n] = y + ds;
            }
            bead_step<A, N, AR, PM, false, SHARE, false, MTAU>(T, Rc, Rn, Tm, lrho, bad);
        }
        ws_mbar_arrive(empty_bar(g, s));
    }
    {   // bead P-1 closes the ring on bead 0
        double Rc[N], Rn[N];
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const double ds = dsrc[n * WS_CONS];
            Rc[n] = PBX_YP(n) + ds;
            Rn[n] = y0[n * WS_CONS] + ds;
        }
        bead_step<A, N, AR, PM, false, SHARE, false, MTAU>(T, Rc, Rn, Tm, lrho, bad);
    }
#undef PBX_YP
    if (!live) return;
    double rho = 0.0;
#pragma unroll
    for (int a = 0; a < AR; ++a) rho += exp_fast(lrho[a]);
    if (bad) rho = __longlong_as_double(0x7ff8000000000000LL);     // NaN: picked up by the MODE_REDO launch
    L.out4[x] = rho;
    if (L.mirror) L.mirror[x] = rho;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < A; ++i) tr += Tm[v][i][i];
        L.out4[(size_t)(1 + v) * L.out_ld + x] = tr;
        if (L.mirror) L.mirror[(size_t)(1 + v) * L.mirror_ld + x] = tr;
    }
}

template <int A, int N, int AR, bool PM, bool SHARE, bool MTAU = false>
cudaError_t launch_ws_one(const FastTables<A, N, AR>& T, const FastLaunch& L_in, cudaStream_t stream) {
    FastLaunch L = L_in;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long groups = (L.n_samples + 63) / 64;
    const long long slots = (long long)sms * ws_ctas<A, N, AR>();            // CTAs resident at a time: one wave
    L.ws_full_ctas = (groups / WS_GROUPS) / slots * slots;                  // whole waves
    const long long rest = groups - WS_GROUPS * L.ws_full_ctas;
    L.ws_tail_groups = (int)std::min<long long>(WS_GROUPS, std::max<long long>(1, (rest + slots - 1) / slots));
    const long long blocks = L.ws_full_ctas + (rest + L.ws_tail_groups - 1) / L.ws_tail_groups;
    constexpr size_t smem = ws_smem_bytes<N>();
    auto kernel = pbx_fast_ws_kernel<A, N, AR, PM, SHARE, MTAU>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)blocks, WS_THREADS, smem, stream>>>(T, L);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if constexpr (PM && (PBX_DELTA_EXP || MTAU))   // recompute the (normally zero) flagged samples with full exponentials
        return launch_one<A, N, AR, MODE_REDO, PM, false, SHARE, MTAU>(T, L, stream);
    return cudaSuccess;
}

// number of kernels one warp-specialised launch enqueues
template <int N>
int ws_launches(bool pm, bool mtau) { return (ws_smem_bytes<N>() <= 227 * 1024 && pm && (PBX_DELTA_EXP || mtau)) ? 2 : 1; }

template <int A, int N, int AR>
cudaError_t launch_fast_ws(const void* tables, const FastLaunch& L, bool pm, bool share, bool mtau, cudaStream_t stream) {
    if constexpr (ws_smem_bytes<N>() > 227 * 1024)      // ring does not fit an SM: one-role kernel
        return launch_fast<A, N, AR>(tables, L, MODE_SAMPLE, pm, false, share, mtau, stream);
    const auto& T = *reinterpret_cast<const FastTables<A, N, AR>*>(tables);
#if !PBX_WITH_MTAU
    if (mtau) return cudaErrorNotSupported;
#endif
    if constexpr (A == AR) {
        if (share) {
#if PBX_WITH_MTAU
            if (pm && mtau) return launch_ws_one<A, N, AR, true, true, true>(T, L, stream);
#endif
            return pm ? launch_ws_one<A, N, AR, true, true>(T, L, stream) : launch_ws_one<A, N, AR, false, true>(T, L, stream);
        }
    }
#if PBX_WITH_MTAU
    if (pm && mtau) return launch_ws_one<A, N, AR, true, false, true>(T, L, stream);
#endif
    return pm ? launch_ws_one<A, N, AR, true, false>(T, L, stream) : launch_ws_one<A, N, AR, false, false>(T, L, stream);
}

template <int A, int N, int AR>
constexpr FastKernelEntry make_entry() {
    return FastKernelEntry{A, N, AR, (PBX_WITH_JACOBI ? FAST_CAP_JACOBI : 0) | (PBX_WITH_MTAU ? FAST_CAP_MTAU : 0), sizeof(FastTables<A, N, AR>), &fill_fast_tables<A, N, AR>, &launch_fast<A, N, AR>,
                           &launch_fast_ws<A, N, AR>, &ws_launches<N>};
}

}  // namespace pbx

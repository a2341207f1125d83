// Register-resident small-A kernel: ONE SAMPLE PER THREAD, beads streamed in tiles.
//
// For each sample the thread walks the ring polymer in tiles of PBX_TILE beads:
//   phase S  (small rolled code): draw the tile's bead coordinates -- Philox4x32-10, FP64 Box-Muller,
//            cyclic-tridiagonal ring recurrence -- into a per-thread column of shared memory
//            (or, MODE_COORDS, copy caller supplied coordinates there);
//   phase E  (fully unrolled over surfaces and modes): per bead, the four sets of harmonic factors O
//            (rho; vib at tau, tau+, tau-) in log space, the scaling S, the packed symmetric coupling
//            matrix V, M = exp(-tau V), and the three chained products T_v <- (T_v M) diag(O_v).
// Splitting the phases keeps each loop body inside the instruction cache (the single fused body of
// round 1 was ~50 KB and stalled on instruction fetch, profiles/r01_summary.md).
// Nothing but the 32-byte result leaves the SM.
//
// Model constants arrive as a __grid_constant__ kernel parameter, i.e. in the constant bank: after
// unrolling every table element is an LDCU'd uniform-register operand of a DFMA, so the coupling
// tables cost no per-thread loads and no registers.
//
// Reference path reproduced per sample: /root/reference/pibronic/pimc/pimc.py:1420-1449
// (block_compute_pm body) with the helpers at 1062-1213; SURVEY.md App. A.
#pragma once
#include <algorithm>
#include <type_traits>

#include "pbx_device.cuh"

#ifndef PBX_BLOCK
#define PBX_BLOCK 128        // threads per CTA
#endif
#ifndef PBX_MIN_BLOCKS
#define PBX_MIN_BLOCKS 1     // __launch_bounds__ minimum CTAs per SM
#endif
#ifndef PBX_TILE
#define PBX_TILE 8           // beads per tile
#endif
#ifndef PBX_UNROLL_H
#define PBX_UNROLL_H 1       // 1: unroll the normal-pair loop of the sampler phase (independent chains interleave)
#endif
#ifndef PBX_WITH_MTAU
#define PBX_WITH_MTAU 0      // 1: also instantiate the PBX_FLAG_M_TAU_PM kernels (consistent estimator; not in the reference)
#endif
#ifndef PBX_WITH_JACOBI
#define PBX_WITH_JACOBI 1    // 0: no Jacobi-eigensolve variants (run-time compiled shapes, pibronic_b200/jit.py)
#endif
#ifndef PBX_DELTA_EXP
#define PBX_DELTA_EXP 1      // 1: O(tau+-) = O(tau) * exp(delta), delta from analytic difference tables
#endif

namespace pbx {

template <int A, int N, int AR>
struct FastTables {
    static constexpr int AA = A * (A + 1) / 2;
    static constexpr int NN = N * (N + 1) / 2;
    // harmonic exponent in half-angle form: l = logpref - 1/4 sum_n [tanh(x/2) (q+q')^2 + coth(x/2) (q-q')^2],
    // x = tau*omega_n; with q = R - d the second term does not depend on the surface
    double d2_vib[A][N];  // 2 * d
    double d_rho[AR][N];  // d   (sampler shift)
    double d2_rho[AR][N]; // 2 * d
    double al[4][N];      // -1/4 tanh(x/2)  (rows: vib tau, tau+, tau-, rho)
    double ga[4][N];      // -1/4 coth(x/2)
    double dal[2][N];     // al(tau+-) - al(tau), formed analytically
    double dga[2][N];     // ga(tau+-) - ga(tau)
    double lpref[3][A];
    double dlpref[2][A];
    double lpref_rho[AR];
    double wcum[AR];
    double e_off[AA];     // the three coupling tables are pre-multiplied by -tau: the kernel builds X = -tau V directly
    double l_off[N][AA];  // (M always uses tau, also for g+-: reference quirk Q2, pimc.py:1183)
    double q_pack[NN][AA];
    double kappa;         // delta_beta / beta = tau+/tau - 1 (PBX_FLAG_M_TAU_PM)
    int P;
    int n_rho_eval;
};

struct FastLaunch {
    const double* samp;      // [P][N][3] recurrence table (global)
    const double* coords;    // MODE_COORDS: the caller's coordinates R[x][N][P], read in place
    unsigned long long seed;
    long long first_sample;  // global index of sample 0 of this launch (Philox counter)
    long long n_samples;
    double* out4;            // [4][out_ld]
    long long out_ld;
    double* mirror;          // optional second copy of the results, [4][mirror_ld]: mapped pinned HOST memory written
    long long mirror_ld;     // straight from the kernel (posted PCIe writes overlap the compute; no D2H pass afterwards)
    long long ws_full_ctas;  // warp-specialised kernel: CTAs [0, ws_full_ctas) take 4 groups of 64 samples (whole waves),
    int ws_tail_groups;      // the CTAs of the last, partial wave ws_tail_groups each (filled in by launch_ws_one)
};

// MODE_REDO: MODE_SAMPLE restricted to the samples the warp-specialised kernel (pbx_fast_ws.cuh) flagged with
// rho = NaN (|log O(tau+-) - log O(tau)| beyond the short-series range); evaluated with full exponentials
enum { MODE_SAMPLE = 0, MODE_COORDS = 1, MODE_REDO = 2 };

// exp(d) for |d| <= 2^-8 by a degree-5 Taylor polynomial (remainder < 4.9e-18)
__device__ __forceinline__ double exp_small(double d) {
    double p = 1.0 / 120;
    p = fma(p, d, 1.0 / 24);
    p = fma(p, d, 1.0 / 6);
    p = fma(p, d, 0.5);
    p = fma(p, d, 1.0);
    return fma(p, d, 1.0);
}
// |d| >= 2^-8, NaN or inf, decided on the exponent field (integer pipe: the FP64 pipe is the bottleneck)
__device__ __forceinline__ bool exp_small_out_of_range(double d) {
    return (__double2hiint(d) & 0x7fffffff) >= 0x3f700000;
}

// one bead of the estimator: updates the chained products Tm and the log-accumulators of rho
// SAFE = false: O(tau+-) = O(tau) * exp_small(delta) unconditionally, `bad` records |delta| >= 2^-8 (the caller then
// redoes the sample with SAFE = true, full exponentials) -- keeps the hot loop free of a data dependent branch
// MTAU (PBX_FLAG_M_TAU_PM): g+- use M(tau+-) = exp(-tau+- V) = M exp(+-kappa X), X = -tau V, instead of the reference's
// M(tau) for all three (quirk Q2); the correction exp(+-Y) - I = +-Y + Y^2/2, Y = kappa X, is exact to rounding for
// ||Y|| < 2^-14 (next term ||Y||^3/6 < 4e-14 at the bound, 1e-19 at the reference's delta_beta), else `bad`
template <int A, int N, int AR, bool PM, bool JACOBI, bool SHARE, bool SAFE, bool MTAU = false>
__device__ __forceinline__ void bead_step(const FastTables<A, N, AR>& T, const double (&Rc)[N], const double (&Rn)[N],
                                          double (&Tm)[PM ? 3 : 1][A][A], double (&lrho)[AR], bool& bad) {
    constexpr int AA = A * (A + 1) / 2;
    constexpr int NV = PM ? 3 : 1;
    // ---- harmonic factors, log space (half-angle form, see FastTables)
    double lv[NV][A], lr[AR];
    double rs[N], w2[N];          // R_p + R_{p+1} and (R_p - R_{p+1})^2: surface independent
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, gr = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const double w = Rc[n] - Rn[n];
        rs[n] = Rc[n] + Rn[n];
        w2[n] = w * w;
        g0 = fma(T.ga[0][n], w2[n], g0);
        if (PM) {
#if PBX_DELTA_EXP
            g1 = fma(T.dga[0][n], w2[n], g1);
            g2 = fma(T.dga[1][n], w2[n], g2);
#else
            g1 = fma(T.ga[1][n], w2[n], g1);
            g2 = fma(T.ga[2][n], w2[n], g2);
#endif
        }
        if (!SHARE) gr = fma(T.ga[3][n], w2[n], gr);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) {
        double e0 = g0, e1 = g1, e2 = g2;
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const double u = rs[n] - T.d2_vib[a][n];
            const double u2 = u * u;
            e0 = fma(T.al[0][n], u2, e0);
            if (PM) {
#if PBX_DELTA_EXP
                e1 = fma(T.dal[0][n], u2, e1);
                e2 = fma(T.dal[1][n], u2, e2);
#else
                e1 = fma(T.al[1][n], u2, e1);
                e2 = fma(T.al[2][n], u2, e2);
#endif
            }
        }
        lv[0][a] = T.lpref[0][a] + e0;
        if (PM) {
#if PBX_DELTA_EXP
            lv[1][a] = T.dlpref[0][a] + e1;   // log O(tau+) - log O(tau)
            lv[2][a] = T.dlpref[1][a] + e2;
#else
            lv[1][a] = T.lpref[1][a] + e1;
            lv[2][a] = T.lpref[2][a] + e2;
#endif
        }
        if (SHARE) lr[a < AR ? a : 0] = (a < T.n_rho_eval) ? T.lpref_rho[a < AR ? a : 0] + e0 : -INFINITY;
    }
    if (!SHARE) {
#pragma unroll
        for (int a = 0; a < AR; ++a) {
            double acc = gr;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const double u = rs[n] - T.d2_rho[a][n];
                acc = fma(T.al[3][n], u * u, acc);
            }
            lr[a] = (a < T.n_rho_eval) ? T.lpref_rho[a] + acc : -INFINITY;
        }
    }
    // S = max over both models' factors (pimc.py:1076-1084), kept as a logarithm
    double logS = lv[0][0];
#pragma unroll
    for (int a = 1; a < A; ++a) logS = fmax(logS, lv[0][a]);
#pragma unroll
    for (int a = 0; a < AR; ++a) logS = fmax(logS, lr[a]);
#pragma unroll
    for (int a = 0; a < AR; ++a) lrho[a] += lr[a] - logS;
    double O[NV][A];
#pragma unroll
    for (int a = 0; a < A; ++a) O[0][a] = exp_fast(lv[0][a] - logS);
    if (PM) {
#if PBX_DELTA_EXP
        if (!SAFE) {
#pragma unroll
            for (int a = 0; a < A; ++a) {
                bad = bad || exp_small_out_of_range(lv[1][a]) || exp_small_out_of_range(lv[2][a]);
                O[1][a] = O[0][a] * exp_small(lv[1][a]);
                O[2][a] = O[0][a] * exp_small(lv[2][a]);
            }
        } else {   // huge delta_beta or far-out coordinates: full exponentials
#pragma unroll
            for (int a = 0; a < A; ++a) {
                O[1][a] = exp_fast(lv[0][a] + lv[1][a] - logS);
                O[2][a] = exp_fast(lv[0][a] + lv[2][a] - logS);
            }
        }
#else
#pragma unroll
        for (int a = 0; a < A; ++a) { O[1][a] = exp_fast(lv[1][a] - logS); O[2][a] = exp_fast(lv[2][a] - logS); }
#endif
    }

    // ---- X = -tau * V(R_p), packed symmetric (the tables carry the factor -tau)
    double X[AA];
#pragma unroll
    for (int k = 0; k < AA; ++k) X[k] = 0.0;
#pragma unroll
    for (int i = 1; i < A; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) {
            double acc = T.e_off[tri(i, j)];
#pragma unroll
            for (int n = 0; n < N; ++n) acc = fma(T.l_off[n][tri(i, j)], Rc[n], acc);
            X[tri(i, j)] = acc;
        }
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int m = n; m < N; ++m) {
            const double rr = Rc[n] * Rc[m];
#pragma unroll
            for (int k = 0; k < AA; ++k) X[k] = fma(T.q_pack[pair_index(n, m, N)][k], rr, X[k]);
        }
    double M[AA], Mpm[(PM && MTAU) ? 2 : 1][AA];
    if (PM && MTAU) {
        if (SAFE) {                 // full exponentials of -tau+- V
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                double Xw[AA];
                const double f = (w == 0) ? 1.0 + T.kappa : 1.0 - T.kappa;
#pragma unroll
                for (int k = 0; k < AA; ++k) Xw[k] = f * X[k];
                if (JACOBI) sym_exp_jacobi<A>(Xw, Mpm[w]);
                else sym_expm<A>(Xw, Mpm[w]);
            }
            if (JACOBI) sym_exp_jacobi<A>(X, M);
            else sym_expm<A>(X, M);
        } else {
            double Y[AA], Yh[AA], mass = 0.0;
#pragma unroll
            for (int k = 0; k < AA; ++k) { Y[k] = T.kappa * X[k]; mass += fabs(Y[k]); }
            bad = bad || (__double2hiint(2.0 * mass) & 0x7fffffff) >= 0x3f100000;     // ||Y||_1 <= 2 mass >= 2^-14
            if (JACOBI) sym_exp_jacobi<A>(X, M);
            else sym_expm<A>(X, M);                       // scales X in place: Y was taken before
            sym_mul<A>(Y, Y, Yh);
            double Ep[AA], Em[AA];
#pragma unroll
            for (int k = 0; k < AA; ++k) { Ep[k] = fma(0.5, Yh[k], Y[k]); Em[k] = fma(0.5, Yh[k], -Y[k]); }
            sym_mul<A>(M, Ep, Mpm[0]);
            sym_mul<A>(M, Em, Mpm[1]);
#pragma unroll
            for (int k = 0; k < AA; ++k) { Mpm[0][k] += M[k]; Mpm[1][k] += M[k]; }
        }
    } else {
        if (JACOBI) sym_exp_jacobi<A>(X, M);
        else sym_expm<A>(X, M);
    }

    // ---- chain: T_v <- (T_v M_v) diag(O_v)
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const double (&Mv)[AA] = (PM && MTAU && v > 0) ? Mpm[v > 0 ? v - 1 : 0] : M;
            double row[A];
#pragma unroll
            for (int j = 0; j < A; ++j) {
                double acc = Tm[v][i][0] * Mv[sym(0, j)];
#pragma unroll
                for (int k = 1; k < A; ++k) acc = fma(Tm[v][i][k], Mv[sym(k, j)], acc);
                row[j] = acc * O[v][j];
            }
#pragma unroll
            for (int j = 0; j < A; ++j) Tm[v][i][j] = row[j];
        }
}

// shared memory (doubles) per thread: (TILE+1) coordinate slots + sampler state y0, yprev, dsrc
template <int N> constexpr int fast_smem_doubles_per_thread() { return (PBX_TILE + 1) * N + 3 * N; }

template <int A, int N, int AR, int MODE, bool PM, bool JACOBI, bool SHARE, bool MTAU = false>
__global__ void __launch_bounds__(PBX_BLOCK, PBX_MIN_BLOCKS)
pbx_fast_kernel(const __grid_constant__ FastTables<A, N, AR> T, const FastLaunch L) {
    constexpr int NV = PM ? 3 : 1;
    constexpr int TB = PBX_TILE;
    extern __shared__ double smem[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int P = T.P;
    // per-thread columns: element i of this thread lives at base[i * nt]
    double* tile = smem + tid;                               // [(TB+1)][N]
    double* y0 = smem + (size_t)(TB + 1) * N * nt + tid;     // [N] first bead relative to the source shift
    double* yprev = y0 + (size_t)N * nt;                     // [N]
    double* dsrc = yprev + (size_t)N * nt;                   // [N] shift of the mixture component drawn

    // everything that depends on the sample index; MODE_REDO calls it for every flagged sample of a grid-stride scan
    auto process = [&](const long long x, const bool live) {
    const unsigned long long gidx = (unsigned long long)(L.first_sample + x);
    const uint2 key = make_uint2((uint32_t)L.seed, (uint32_t)(L.seed >> 32));
    if (MODE != MODE_COORDS) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, STREAM_SOURCE), key);
        const int src = pick_source<AR>(u01_half_open(r.x, r.y), T.wcum);
#pragma unroll
        for (int n = 0; n < N; ++n) {
            double d = T.d_rho[0][n];
#pragma unroll
            for (int a = 1; a < AR; ++a) d = (src == a) ? T.d_rho[a][n] : d;
            dsrc[n * nt] = d;
        }
    }
    // coordinates of bead j (j == P closes the ring: bead 0 again) into tile slot `slot`
    auto put_bead = [&](int j, int slot) {
        double* dst = tile + (size_t)slot * N * nt;
        if (MODE != MODE_COORDS) {
            if (j == P) {
#pragma unroll 1
                for (int n = 0; n < N; ++n) dst[n * nt] = y0[n * nt] + dsrc[n * nt];
                return;
            }
            const double* tab = L.samp + (size_t)j * N * 3;
#if PBX_UNROLL_H
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int h = 0; h < (N + 1) / 2; ++h) {   // branch-free body: unrolled pairs interleave
                const uint4 r = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32),
                                                         (uint32_t)(j * ((N + 1) / 2) + h), STREAM_NORMALS), key);
                double z[2];
                normal_pair(r, z[0], z[1]);
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int n = 2 * h + w;
                    if (n < N) {
                        const double a = __ldg(tab + n * 3 + 0), b = __ldg(tab + n * 3 + 1), e = __ldg(tab + n * 3 + 2);
                        double y = a * z[w];
                        if (j > 0) y = fma(b, yprev[n * nt], fma(e, y0[n * nt], y));
                        if (j == 0) y0[n * nt] = y;
                        yprev[n * nt] = y;
                        dst[n * nt] = y + dsrc[n * nt];
                    }
                }
            }
        } else {
            // the caller's R[x][n][p] read in place: the lanes of a warp are N P doubles apart, but a lane's 32-byte sector
            // serves four consecutive beads from L1 (DRAM traffic 1.2x the array; a transposed copy cost 3x and a pass)
            const double* src = L.coords + (size_t)x * N * P + (j == P ? 0 : j);
            double v[N];
#pragma unroll
            for (int n = 0; n < N; ++n) v[n] = __ldg(src + (size_t)n * P);   // all loads in flight together
#pragma unroll
            for (int n = 0; n < N; ++n) dst[n * nt] = v[n];
        }
    };

    double rho, tr[NV];
    // the whole sample as a function of SAFE: state is rebuilt from the Philox counters, so a redo is exact
    auto run_sample = [&](auto safe_tag) -> bool {
        constexpr bool SAFE = decltype(safe_tag)::value;
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

        put_bead(0, 0);
        for (int t0 = 0; t0 < P; t0 += TB) {
            // ---- phase S: beads t0+1 .. t0+TB into slots 1..TB (slot 0 holds bead t0)
            if (MODE != MODE_COORDS) {
#pragma unroll 1
                for (int jj = 1; jj <= TB; ++jj)
                    if (t0 + jj <= P) put_bead(t0 + jj, jj);
            } else {
#pragma unroll
                for (int jj = 1; jj <= TB; ++jj)
                    if (t0 + jj <= P) put_bead(t0 + jj, jj);
            }
            // ---- phase E
#pragma unroll 1
            for (int jj = 0; jj < TB; ++jj) {
                if (t0 + jj >= P) break;
                double Rc[N], Rn[N];
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    Rc[n] = tile[(size_t)(jj * N + n) * nt];
                    Rn[n] = tile[(size_t)((jj + 1) * N + n) * nt];
                }
                bead_step<A, N, AR, PM, JACOBI, SHARE, SAFE, MTAU>(T, Rc, Rn, Tm, lrho, bad);
            }
            // bead t0+TB becomes slot 0 of the next tile
#pragma unroll
            for (int n = 0; n < N; ++n) tile[(size_t)n * nt] = tile[(size_t)(TB * N + n) * nt];
        }
        rho = 0.0;
#pragma unroll
        for (int a = 0; a < AR; ++a) rho += exp_fast(lrho[a]);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            tr[v] = 0.0;
#pragma unroll
            for (int i = 0; i < A; ++i) tr[v] += Tm[v][i][i];
        }
        return bad;
    };
    if constexpr (MODE == MODE_REDO) {
        run_sample(std::true_type{});
    } else {
        bool redo = run_sample(std::false_type{});
        if (PM && (PBX_DELTA_EXP || MTAU)) {
            if (__any_sync(__activemask(), redo)) {     // cold path: never taken at delta_beta/beta ~ 5e-6
                if (redo) run_sample(std::true_type{});
            }
        }
    }
    if (!live) return;
    L.out4[x] = rho;
#pragma unroll
    for (int v = 0; v < NV; ++v) L.out4[(size_t)(1 + v) * L.out_ld + x] = tr[v];
    if (L.mirror) {
        L.mirror[x] = rho;
#pragma unroll
        for (int v = 0; v < NV; ++v) L.mirror[(size_t)(1 + v) * L.mirror_ld + x] = tr[v];
    }
    };   // process

    if constexpr (MODE == MODE_REDO) {
        // no block-wide barrier in this kernel: every thread scans its own stride of the result array, eight loads in flight
        // (the scan of 1e6 results is latency bound: 16 us with one load at a time)
        const long long stride = (long long)gridDim.x * nt;
        for (long long x0 = (long long)blockIdx.x * nt + tid; x0 < L.n_samples; x0 += 8 * stride) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (x0 + k * stride < L.n_samples) ? __ldcg(L.out4 + x0 + k * stride) : 0.0;
            unsigned flagged = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) flagged |= isnan(v[k]) ? (1u << k) : 0u;
            while (flagged) {
                const int k = __ffs(flagged) - 1;
                flagged &= flagged - 1;
                process(x0 + k * stride, true);
            }
        }
    } else {
        const long long x = (long long)blockIdx.x * nt + tid;
        const bool live = x < L.n_samples;
        process(live ? x : L.n_samples - 1, live);   // a dead thread stays in step with its CTA; its result is dropped
    }
}

// type-erased launcher stored in the plan
enum { FAST_CAP_JACOBI = 1, FAST_CAP_MTAU = 2 };

struct FastKernelEntry {
    int A, N, AR;
    int caps;                // FAST_CAP_*: which optional variants this build carries
    size_t table_bytes;
    void (*fill)(const HostTables&, void* dst);
    cudaError_t (*launch)(const void* tables, const FastLaunch& L, int mode, bool pm, bool jacobi, bool share, bool mtau,
                          cudaStream_t stream);
    // warp-specialised sampler + estimator (pbx_fast_ws.cuh; scaling-and-squaring M only); enqueues `ws_kernels(pm)` kernels
    cudaError_t (*launch_ws)(const void* tables, const FastLaunch& L, bool pm, bool share, bool mtau, cudaStream_t stream);
    int (*ws_kernels)(bool pm, bool mtau);
};

template <int A, int N, int AR>
void fill_fast_tables(const HostTables& H, void* dst) {
    auto& T = *reinterpret_cast<FastTables<A, N, AR>*>(dst);
    constexpr int AA = A * (A + 1) / 2, NN = N * (N + 1) / 2;
    for (int a = 0; a < A; ++a) for (int n = 0; n < N; ++n) T.d2_vib[a][n] = 2.0 * H.d_vib[a * N + n];
    for (int a = 0; a < AR; ++a) for (int n = 0; n < N; ++n) { T.d_rho[a][n] = H.d_rho[a * N + n]; T.d2_rho[a][n] = 2.0 * H.d_rho_eval[a * N + n]; }
    for (int v = 0; v < 4; ++v) for (int n = 0; n < N; ++n) { T.al[v][n] = -0.25 * H.tanh_half[v * N + n]; T.ga[v][n] = -0.25 * H.coth_half[v * N + n]; }
    for (int v = 0; v < 2; ++v) for (int n = 0; n < N; ++n) { T.dal[v][n] = -0.25 * H.dtanh_half[v * N + n]; T.dga[v][n] = -0.25 * H.dcoth_half[v * N + n]; }
    for (int v = 0; v < 3; ++v) for (int a = 0; a < A; ++a) T.lpref[v][a] = H.logpref[v * A + a];
    for (int v = 0; v < 2; ++v) for (int a = 0; a < A; ++a) T.dlpref[v][a] = H.dlogpref[v * A + a];
    for (int a = 0; a < AR; ++a) { T.lpref_rho[a] = H.logpref_rho[a]; T.wcum[a] = H.wcum[a]; }
    const double mt = -H.tau[0];
    for (int k = 0; k < AA; ++k) T.e_off[k] = mt * H.e_off[k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < AA; ++k) T.l_off[n][k] = mt * H.l_off[(size_t)n * AA + k];
    for (int q = 0; q < NN; ++q) for (int k = 0; k < AA; ++k) T.q_pack[q][k] = mt * H.q_pack[(size_t)q * AA + k];
    T.kappa = H.delta_beta / H.beta;
    T.P = H.P;
    T.n_rho_eval = H.n_rho_eval;
}

template <int A, int N, int AR, int MODE, bool PM, bool JACOBI, bool SHARE, bool MTAU = false>
cudaError_t launch_one(const FastTables<A, N, AR>& T, const FastLaunch& L, cudaStream_t stream) {
    const int threads = PBX_BLOCK;
    long long blocks = (L.n_samples + threads - 1) / threads;
    if (MODE == MODE_REDO) blocks = std::min<long long>(blocks, 2 * 148);   // grid-stride scan
    const size_t smem = (size_t)fast_smem_doubles_per_thread<N>() * threads * sizeof(double);
    auto kernel = pbx_fast_kernel<A, N, AR, MODE, PM, JACOBI, SHARE, MTAU>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<(unsigned)blocks, threads, smem, stream>>>(T, L);
    return cudaGetLastError();
}

template <int A, int N, int AR, int MODE, bool PM, bool JACOBI, bool MTAU = false>
cudaError_t launch_share(const FastTables<A, N, AR>& T, const FastLaunch& L, bool share, cudaStream_t stream) {
    if constexpr (A == AR) {
        if (share) return launch_one<A, N, AR, MODE, PM, JACOBI, true, MTAU>(T, L, stream);
    }
    return launch_one<A, N, AR, MODE, PM, JACOBI, false, MTAU>(T, L, stream);
}

// mtau (PBX_FLAG_M_TAU_PM) exists for the PM, scaling-and-squaring kernels only; pbx_plan_create rejects other combinations
template <int A, int N, int AR>
cudaError_t launch_fast(const void* tables, const FastLaunch& L, int mode, bool pm, bool jacobi, bool share, bool mtau,
                        cudaStream_t stream) {
    const auto& T = *reinterpret_cast<const FastTables<A, N, AR>*>(tables);
#if PBX_WITH_MTAU
#define PBX_DISPATCH_MTAU(MODE_) if (pm && mtau) return launch_share<A, N, AR, MODE_, true, false, true>(T, L, share, stream);
#else
#define PBX_DISPATCH_MTAU(MODE_) if (mtau) return cudaErrorNotSupported;
#endif
#if PBX_WITH_JACOBI
#define PBX_DISPATCH_JACOBI(MODE_)                                                                                       \
    if (jacobi) return pm ? launch_share<A, N, AR, MODE_, true, true>(T, L, share, stream)                              \
                          : launch_share<A, N, AR, MODE_, false, true>(T, L, share, stream);
#else
#define PBX_DISPATCH_JACOBI(MODE_) if (jacobi) return cudaErrorNotSupported;
#endif
#define PBX_DISPATCH(MODE_)                                                                               \
    PBX_DISPATCH_MTAU(MODE_)                                                                              \
    PBX_DISPATCH_JACOBI(MODE_)                                                                            \
    if (pm) return launch_share<A, N, AR, MODE_, true, false>(T, L, share, stream);                      \
    return launch_share<A, N, AR, MODE_, false, false>(T, L, share, stream);
    if (mode == MODE_SAMPLE) { PBX_DISPATCH(MODE_SAMPLE) }
    PBX_DISPATCH(MODE_COORDS)
#undef PBX_DISPATCH
#undef PBX_DISPATCH_MTAU
#undef PBX_DISPATCH_JACOBI
}

// defined in pbx_fast_registry.cu: shapes compiled into the library, then shapes registered at run time
const FastKernelEntry* find_fast_kernel(int A, int N, int AR);
bool register_fast_kernel(const FastKernelEntry* entry);

}  // namespace pbx

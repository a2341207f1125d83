// Register-resident small-A kernel: ONE SAMPLE PER THREAD, beads streamed.
//
// For each sample the thread walks the ring polymer bead by bead.  Per bead it (optionally) draws
// the next bead's coordinates (Philox + Box-Muller + ring recurrence), forms the four sets of
// harmonic factors O (rho; vib at tau, tau+, tau-) in log space, the scaling S, the packed
// symmetric coupling matrix V, M = exp(-tau V), and advances the three chained products
// T_v <- (T_v M) diag(O_v).  Nothing but the 32-byte result leaves the SM.
//
// Model constants arrive as a __grid_constant__ kernel parameter, i.e. in the constant bank:
// after full unrolling every table element is an immediate c[0x0][..] operand of a DFMA, so the
// coupling tables cost no load instructions and no registers.
//
// Reference path reproduced per sample: /root/reference/pibronic/pimc/pimc.py:1420-1449
// (block_compute_pm body) with the helpers at 1062-1213; SURVEY.md App. A.
#pragma once
#include "pbx_device.cuh"

namespace pbx {

template <int A, int N, int AR>
struct FastTables {
    static constexpr int AA = A * (A + 1) / 2;
    static constexpr int NN = N * (N + 1) / 2;
    double d_vib[A][N];
    double d_rho[AR][N];
    double hc[4][N];      // -0.5 * coth  (rows: vib tau, tau+, tau-, rho)
    double cs[4][N];      // csch
    double lpref[3][A];
    double lpref_rho[AR];
    double wcum[AR];
    double e_off[AA];
    double l_off[N][AA];
    double q_pack[NN][AA];
    double neg_tau;       // -tau (M always uses tau: reference quirk Q2, pimc.py:1183)
    int P;
    int n_rho_eval;
};

struct FastLaunch {
    const double* samp;      // [P][N][3] recurrence table (global)
    const double* coords_t;  // MODE_COORDS: transposed coordinates [N][P][ld]
    long long ld;            // leading dimension (samples) of coords_t
    unsigned long long seed;
    long long first_sample;  // global index of sample 0 of this launch (Philox counter)
    long long n_samples;
    double* out4;            // [4][out_ld]
    long long out_ld;
};

enum { MODE_SAMPLE = 0, MODE_COORDS = 1 };

template <int A, int N, int AR, int MODE, bool PM, bool JACOBI>
__global__ void __launch_bounds__(128)
pbx_fast_kernel(const __grid_constant__ FastTables<A, N, AR> T, const FastLaunch L) {
    constexpr int AA = A * (A + 1) / 2;
    constexpr int NV = PM ? 3 : 1;
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= L.n_samples) return;
    const int P = T.P;

    // ---------------- sampler state
    const unsigned long long gidx = (unsigned long long)(L.first_sample + x);
    const uint2 key = make_uint2((uint32_t)L.seed, (uint32_t)(L.seed >> 32));
    double dsrc[N];   // shift of the mixture component this sample is drawn from
    double y0[N];     // first bead, relative to dsrc (ring recurrence anchor)
    double yprev[N];
    double R0[N], Rc[N], Rn[N];
    if (MODE == MODE_SAMPLE) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, STREAM_SOURCE), key);
        const int src = pick_source<AR>(u01_half_open(r.x, r.y), T.wcum);
#pragma unroll
        for (int n = 0; n < N; ++n) {
            double d = T.d_rho[0][n];
#pragma unroll
            for (int a = 1; a < AR; ++a) d = (src == a) ? T.d_rho[a][n] : d;
            dsrc[n] = d;
        }
    }
    auto next_bead = [&](int j, double (&R)[N]) {
        if (MODE == MODE_SAMPLE) {
            const double* tab = L.samp + (size_t)j * N * 3;
#pragma unroll
            for (int h = 0; h < (N + 1) / 2; ++h) {
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
                        if (j > 0) y = fma(b, yprev[n], fma(e, y0[n], y));
                        if (j == 0) y0[n] = y;
                        yprev[n] = y;
                        R[n] = y + dsrc[n];
                    }
                }
            }
        } else {
            const double* src = L.coords_t + (size_t)j * L.ld + x;
#pragma unroll
            for (int n = 0; n < N; ++n) R[n] = __ldg(src + (size_t)n * T.P * L.ld);
        }
    };

    // ---------------- accumulators
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

    next_bead(0, R0);
#pragma unroll
    for (int n = 0; n < N; ++n) Rc[n] = R0[n];

    for (int p = 0; p < P; ++p) {
        if (p + 1 < P) {
            next_bead(p + 1, Rn);
        } else {
#pragma unroll
            for (int n = 0; n < N; ++n) Rn[n] = R0[n];
        }
        // ---- harmonic factors, log space: l = logpref - 1/2 sum_n [coth (q^2+q'^2) - 2 csch q q']
        double lv[NV][A], lr[AR];
#pragma unroll
        for (int a = 0; a < A; ++a) {
#pragma unroll
            for (int v = 0; v < NV; ++v) lv[v][a] = T.lpref[v][a];
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const double q = Rc[n] - T.d_vib[a][n], qn = Rn[n] - T.d_vib[a][n];
                const double s2 = fma(q, q, qn * qn), pr = q * qn;
#pragma unroll
                for (int v = 0; v < NV; ++v) lv[v][a] = fma(T.hc[v][n], s2, fma(T.cs[v][n], pr, lv[v][a]));
            }
        }
#pragma unroll
        for (int a = 0; a < AR; ++a) {
            double acc = T.lpref_rho[a];
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const double q = Rc[n] - T.d_rho[a][n], qn = Rn[n] - T.d_rho[a][n];
                acc = fma(T.hc[3][n], fma(q, q, qn * qn), fma(T.cs[3][n], q * qn, acc));
            }
            lr[a] = (a < T.n_rho_eval) ? acc : -INFINITY;
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
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int a = 0; a < A; ++a) O[v][a] = exp(lv[v][a] - logS);

        // ---- X = -tau * V(R_p), packed symmetric
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
#pragma unroll
        for (int k = 0; k < AA; ++k) X[k] *= T.neg_tau;

        double M[AA];
        if (JACOBI) sym_exp_jacobi<A>(X, M);
        else sym_expm<A>(X, M);

        // ---- chain: T_v <- (T_v M) diag(O_v)
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int i = 0; i < A; ++i) {
                double row[A];
#pragma unroll
                for (int j = 0; j < A; ++j) {
                    double acc = Tm[v][i][0] * M[sym(0, j)];
#pragma unroll
                    for (int k = 1; k < A; ++k) acc = fma(Tm[v][i][k], M[sym(k, j)], acc);
                    row[j] = acc * O[v][j];
                }
#pragma unroll
                for (int j = 0; j < A; ++j) Tm[v][i][j] = row[j];
            }
#pragma unroll
        for (int n = 0; n < N; ++n) Rc[n] = Rn[n];
    }

    double rho = 0.0;
#pragma unroll
    for (int a = 0; a < AR; ++a) rho += exp(lrho[a]);
    L.out4[x] = rho;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < A; ++i) tr += Tm[v][i][i];
        L.out4[(size_t)(1 + v) * L.out_ld + x] = tr;
    }
}

// type-erased launcher stored in the plan
struct FastKernelEntry {
    int A, N, AR;
    size_t table_bytes;
    void (*fill)(const HostTables&, void* dst);
    cudaError_t (*launch)(const void* tables, const FastLaunch& L, int mode, bool pm, bool jacobi,
                          cudaStream_t stream);
};

template <int A, int N, int AR>
void fill_fast_tables(const HostTables& H, void* dst) {
    auto& T = *reinterpret_cast<FastTables<A, N, AR>*>(dst);
    constexpr int AA = A * (A + 1) / 2, NN = N * (N + 1) / 2;
    for (int a = 0; a < A; ++a) for (int n = 0; n < N; ++n) T.d_vib[a][n] = H.d_vib[a * N + n];
    for (int a = 0; a < AR; ++a) for (int n = 0; n < N; ++n) T.d_rho[a][n] = H.d_rho[a * N + n];
    for (int v = 0; v < 4; ++v) for (int n = 0; n < N; ++n) { T.hc[v][n] = -0.5 * H.coth[v * N + n]; T.cs[v][n] = H.csch[v * N + n]; }
    for (int v = 0; v < 3; ++v) for (int a = 0; a < A; ++a) T.lpref[v][a] = H.logpref[v * A + a];
    for (int a = 0; a < AR; ++a) { T.lpref_rho[a] = H.logpref_rho[a]; T.wcum[a] = H.wcum[a]; }
    for (int k = 0; k < AA; ++k) T.e_off[k] = H.e_off[k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < AA; ++k) T.l_off[n][k] = H.l_off[(size_t)n * AA + k];
    for (int q = 0; q < NN; ++q) for (int k = 0; k < AA; ++k) T.q_pack[q][k] = H.q_pack[(size_t)q * AA + k];
    T.neg_tau = -H.tau[0];
    T.P = H.P;
    T.n_rho_eval = H.n_rho_eval;
}

template <int A, int N, int AR, int MODE, bool PM, bool JACOBI>
cudaError_t launch_one(const FastTables<A, N, AR>& T, const FastLaunch& L, cudaStream_t stream) {
    const int threads = 128;
    const long long blocks = (L.n_samples + threads - 1) / threads;
    pbx_fast_kernel<A, N, AR, MODE, PM, JACOBI><<<(unsigned)blocks, threads, 0, stream>>>(T, L);
    return cudaGetLastError();
}

template <int A, int N, int AR>
cudaError_t launch_fast(const void* tables, const FastLaunch& L, int mode, bool pm, bool jacobi,
                        cudaStream_t stream) {
    const auto& T = *reinterpret_cast<const FastTables<A, N, AR>*>(tables);
#define PBX_DISPATCH(MODE_)                                                                        \
    if (pm) return jacobi ? launch_one<A, N, AR, MODE_, true, true>(T, L, stream)                  \
                          : launch_one<A, N, AR, MODE_, true, false>(T, L, stream);                \
    return jacobi ? launch_one<A, N, AR, MODE_, false, true>(T, L, stream)                         \
                  : launch_one<A, N, AR, MODE_, false, false>(T, L, stream);
    if (mode == MODE_SAMPLE) { PBX_DISPATCH(MODE_SAMPLE) }
    PBX_DISPATCH(MODE_COORDS)
#undef PBX_DISPATCH
}

template <int A, int N, int AR>
constexpr FastKernelEntry make_entry() {
    return FastKernelEntry{A, N, AR, sizeof(FastTables<A, N, AR>), &fill_fast_tables<A, N, AR>, &launch_fast<A, N, AR>};
}

// defined in pbx_fast_*.cu
const FastKernelEntry* find_fast_kernel(int A, int N, int AR);

}  // namespace pbx

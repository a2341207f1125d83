// Blocked kernels for 1 <= A <= 16 surfaces (compile-time A) (BASELINE config c4: A=12, N=24, P=256), any N, A_rho.
// Same three steps through HBM scratch as pbx_generic.cuh (coords -> per-bead stage -> chain) but
// organised for reuse instead of one item at a time:
//
//   pbx_mid_sample_kernel : thread per (sample, MODE PAIR) -- the N ring recurrences are independent,
//                           so a 2400-sample chunk yields N/2 times more threads than thread-per-sample
//   pbx_mid_bead_kernel   : one warp per MID_IB = 8 consecutive beads of a sample: harmonic factors O (log space) and
//                           X = -tau V.  The coupling matrix
//                           V[bead][k] = e_off[k] + sum_n R_n l_off[n][k] + sum_{n<=m} q[n,m][k] R_n R_m is a
//                           dense (8 beads x 325 features) x (325 x 78) contraction at c4: it runs on the FP64
//                           tensor cores (mma.sync.m8n8k4.f64), 10 MMAs per 4 features, the table pre-tiled in
//                           fragment order (210 KB, read as coalesced 256-byte rows from L1/L2).
//   pbx_mid_expm_kernel   : one warp per (sample, bead): M = exp(X) in place, the four A x A products (+ squarings)
//                           as mma.sync.m8n8k4.f64 tiles fed from shared memory (12 MMAs and 12 8-byte loads per
//                           lane and product at A = 12).  Its own kernel so that neither part carries the other's
//                           shared memory.
//   pbx_mid_chain_kernel  : lane = (sample, row i): the rows of the three chained products live in
//                           registers, M_p is broadcast from shared memory, floor(32/A) samples per warp.
//
// Reference: /root/reference/pibronic/pimc/pimc.py:1087-1129 (O), 1076-1084 (S), 1139-1187 (V, M),
// 1194-1209 (chain), 1132-1136 (rho).
#pragma once
#include "pbx_dmma.cuh"
#include "pbx_generic.cuh"

namespace pbx {

constexpr int MID_IB = 8;        // beads per warp in the bead kernel
constexpr int MID_RS = MID_IB + 2;  // row stride of the coordinate tile (even: 16-byte aligned rows)
constexpr int MID_WARPS = 4;
constexpr int MID_AMAX = 16;

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N_)); }

// ---------------------------------------------------------------------------------------------
// sampler: thread per (sample, mode pair); same Philox counters and arithmetic as the fused kernel's sampler phase
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pbx_mid_sample_kernel(DevTables T, unsigned long long seed, long long first_sample, long long n_samples,
                      double* __restrict__ R, int* __restrict__ src_out) {
    const int N = T.N, P = T.P, half = (N + 1) / 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_samples * half) return;
    const long long x = idx / half;
    const int h = (int)(idx - x * half);
    const unsigned long long gidx = (unsigned long long)(first_sample + x);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 rs = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, STREAM_SOURCE), key);
    const double u = u01_half_open(rs.x, rs.y);
    int src = 0;
    for (int a = 0; a < T.Ar - 1; ++a) src += (u >= T.wcum[a]) ? 1 : 0;
    if (src_out && h == 0) src_out[x] = src;
    double y0[2] = {0.0, 0.0}, yprev[2] = {0.0, 0.0}, shift[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) shift[w] = (2 * h + w < N) ? T.d_rho_samp[src * N + 2 * h + w] : 0.0;
    double* Rx = R + (size_t)x * N * P;
    for (int j = 0; j < P; ++j) {
        const double* tab = T.samp + ((size_t)j * N + 2 * h) * 3;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)(j * half + h),
                                                 STREAM_NORMALS), key);
        double z[2];
        normal_pair(r, z[0], z[1]);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int n = 2 * h + w;
            if (n < N) {
                double y = tab[w * 3 + 0] * z[w];
                if (j > 0) y = fma(tab[w * 3 + 1], yprev[w], fma(tab[w * 3 + 2], y0[w], y));
                if (j == 0) y0[w] = y;
                yprev[w] = y;
                Rx[(size_t)n * P + j] = y + shift[w];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// warp product of AT x AT matrices in shared memory (row-major, leading dimension AT) on the FP64 tensor cores:
// mma.sync.m8n8k4 tiles, ceil(AT/8)^2 output tiles x ceil(AT/4) k-steps (12 MMAs at AT = 12).  With g = lane/4,
// c = lane%4 a lane holds A[8 mt + g][4 ks + c], B[4 ks + c][8 nt + g] and C[8 mt + g][8 nt + 2c + {0,1}]; rows and
// columns beyond AT are fed as zeros.  Per product a lane issues 2 ceil(AT/8) ceil(AT/4) 8-byte shared loads (12 at
// AT = 12) -- the register-blocked vector form needed 60 and ran into the shared-memory bandwidth (ncu: L1 99 % busy).
// ---------------------------------------------------------------------------------------------

template <int AT>
struct MmaFrag {   // the entries of a matrix owned by one lane, accumulator layout
    double v[MidShape<AT>::MT][MidShape<AT>::MT][2];
};


template <int AT>
__device__ __forceinline__ void mma_matmul(const double* __restrict__ X, const double* __restrict__ Y, MmaFrag<AT>& C, int lane) {
    constexpr int MT = MidShape<AT>::MT, KS = MidShape<AT>::KS;
    const int g = lane >> 2, c = lane & 3;
    double a[MT][KS], b[MT][KS];
#pragma unroll
    for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int i = 8 * t + g, k = 4 * ks + c;
            const bool ok = i < AT && k < AT;
            a[t][ks] = ok ? X[i * AT + k] : 0.0;          // A[row i][k]
            b[t][ks] = ok ? Y[k * AT + i] : 0.0;          // B[k][column i]
        }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
            C.v[mt][nt][0] = 0.0; C.v[mt][nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) dmma_884(C.v[mt][nt][0], C.v[mt][nt][1], a[mt][ks], b[nt][ks]);
        }
}

template <int AT>
__device__ __forceinline__ void mma_store(double* __restrict__ dst, const MmaFrag<AT>& C, int lane) {
    const int g = lane >> 2, c = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MidShape<AT>::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MidShape<AT>::MT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * mt + g, j = 8 * nt + 2 * c + e;
                if (i < AT && j < AT) dst[i * AT + j] = C.v[mt][nt][e];
            }
}

template <int AT>
__device__ __forceinline__ void mma_load(const double* __restrict__ src, MmaFrag<AT>& C, int lane) {
    const int g = lane >> 2, c = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MidShape<AT>::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MidShape<AT>::MT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * mt + g, j = 8 * nt + 2 * c + e;
                C.v[mt][nt][e] = (i < AT && j < AT) ? src[i * AT + j] : 0.0;
            }
}

// M = exp(X) (degree-12 Taylor in four products, scaling and squaring as in sym_expm); X in shared
// memory is overwritten by its scaled copy; W0..W3 are AT*AT work matrices; the result is left in the
// lane fragments `out` AND in shared memory at the returned pointer.
template <int AT>
__device__ __forceinline__ const double* warp_expm_at(double* X, double* W0, double* W1, double* W2, double* W3,
                                                      MmaFrag<AT>& out, int lane) {
    constexpr int MT = MidShape<AT>::MT, AA2 = AT * AT;
    double norm = 0.0;
    if (lane < AT) {
#pragma unroll
        for (int j = 0; j < AT; ++j) norm += fabs(X[lane * AT + j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) norm = fmax(norm, __shfl_xor_sync(0xffffffffu, norm, o));
    const int s = expm_squarings(norm);
    const double scale = __hiloint2double((1023 - s) << 20, 0);
    for (int e = lane; e < AA2; e += 32) X[e] *= scale;
    __syncwarp();
    // four-product degree-12 Taylor polynomial (pbx_device.cuh, namespace t12)
    MmaFrag<AT> x1, x2, x3, y0, b;
    const int g = lane >> 2, c = lane & 3;
    auto diag = [&](int mt, int nt, int e) { return (8 * mt + g) == (8 * nt + 2 * c + e); };
#define PBX_FRAG_LOOP                                  \
    _Pragma("unroll") for (int mt = 0; mt < MT; ++mt)  \
    _Pragma("unroll") for (int nt = 0; nt < MT; ++nt)  \
    _Pragma("unroll") for (int e = 0; e < 2; ++e)
    mma_load<AT>(X, x1, lane);
    mma_matmul<AT>(X, X, x2, lane);
    mma_store<AT>(W0, x2, lane);            // W0 = X^2
    __syncwarp();
    mma_matmul<AT>(X, W0, x3, lane);
    mma_store<AT>(W1, x3, lane);            // W1 = X^3
    PBX_FRAG_LOOP b.v[mt][nt][e] = fma(t12::c1, x3.v[mt][nt][e], fma(t12::c2, x2.v[mt][nt][e], t12::c3 * x1.v[mt][nt][e]));
    mma_store<AT>(W2, b, lane);
    __syncwarp();
    mma_matmul<AT>(W1, W2, y0, lane);       // Y0 = X^3 (c1 X^3 + c2 X^2 + c3 X)
    __syncwarp();                           // W2 is rewritten below
    PBX_FRAG_LOOP {
        b.v[mt][nt][e] = y0.v[mt][nt][e] + fma(t12::c4, x3.v[mt][nt][e], fma(t12::c5, x2.v[mt][nt][e], t12::c6 * x1.v[mt][nt][e]));
        out.v[mt][nt][e] = y0.v[mt][nt][e] + fma(t12::c7, x3.v[mt][nt][e], t12::c8 * x2.v[mt][nt][e]);
    }
    mma_store<AT>(W2, b, lane);
    mma_store<AT>(W3, out, lane);
    __syncwarp();
    mma_matmul<AT>(W2, W3, out, lane);
    PBX_FRAG_LOOP out.v[mt][nt][e] = out.v[mt][nt][e] +
        fma(t12::c9, y0.v[mt][nt][e], fma(t12::c10, x3.v[mt][nt][e], fma(0.5, x2.v[mt][nt][e], x1.v[mt][nt][e]))) +
        (diag(mt, nt, e) ? 1.0 : 0.0);
#undef PBX_FRAG_LOOP
    __syncwarp();                           // all lanes are done reading W2, W3
    double* cur = W2;
    double* nxt = W3;
    mma_store<AT>(cur, out, lane);
    __syncwarp();
    for (int q = 0; q < s; ++q) {
        mma_matmul<AT>(cur, cur, out, lane);
        mma_store<AT>(nxt, out, lane);
        __syncwarp();
        double* t = cur; cur = nxt; nxt = t;
    }
    return cur;
}

// shared memory (doubles) of one warp of the V/O kernel: coordinate tile, a 2-bead staging tile, the log factors
__host__ __device__ inline size_t mid_bead_warp_doubles(int A, int Ar, int N) {
    size_t n = (size_t)(N + 1) * MID_RS + 2 * (size_t)A * A + (size_t)MID_IB * (3 * A + Ar);
    return (n + 1) & ~(size_t)1;
}
// ... and of the exp kernel: X double-buffered and four work matrices
__host__ __device__ inline size_t mid_expm_warp_doubles(int A) { return 6 * (size_t)A * A; }

// ---------------------------------------------------------------------------------------------
// per-bead stage, part 1: one warp per MID_IB consecutive beads of one sample; AT = number of surfaces.
// Harmonic factors O (log space) and X = -tau V (written to out.m_mat, where part 2 turns it into exp(X) in place).
// Small per-warp footprint (7 KB of shared memory at c4, no work matrices) -> 20 warps per SM.
// ---------------------------------------------------------------------------------------------
template <int AT>
__global__ void __launch_bounds__(MID_WARPS * 32)
pbx_mid_bead_kernel(DevTables T, const double* __restrict__ R, long long n_samples, BeadOutputs out) {
    extern __shared__ __align__(16) double sm[];
    constexpr int AA2 = AT * AT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Ar = T.Ar, N = T.N, P = T.P;
    const int groups = (P + MID_IB - 1) / MID_IB;
    const long long gid = (long long)blockIdx.x * MID_WARPS + warp;
    if (gid >= n_samples * groups) return;
    const long long x = gid / groups;
    const int p0 = (int)(gid - x * groups) * MID_IB;
    const int nb = min(MID_IB, P - p0);
    double* w = sm + (size_t)warp * mid_bead_warp_doubles(AT, Ar, N);
    double* Rt = w;                               // [N + 1][MID_RS]: beads p0 .. p0+MID_IB (ring closed); row N = ones
    double* Xst = Rt + (size_t)(N + 1) * MID_RS;  // [2][AT][AT] staging for coalesced stores
    double* lall = Xst + 2 * (size_t)AA2;         // [MID_IB][3*AT + Ar] log factors
    const int nl = 3 * AT + Ar;
    const double* Rx = R + (size_t)x * N * P;
    for (int e = lane; e < N * (MID_IB + 1); e += 32) {
        const int n = e / (MID_IB + 1), jj = e - n * (MID_IB + 1);
        int p = p0 + jj;
        if (p >= P) p -= P;                        // bead P is bead 0; beyond that only padding
        Rt[n * MID_RS + jj] = (p0 + jj <= P) ? Rx[(size_t)n * P + p] : 0.0;
    }
    if (lane < MID_RS) Rt[N * MID_RS + lane] = 1.0;
    __syncwarp();
    const size_t xp0 = (size_t)x * P + p0;

    // ---- O factors in log space for the beads of the group at once; work items (set, surface)
    for (int it = lane; it < nl; it += 32) {
        const bool is_rho = it >= 3 * AT;
        const int v = is_rho ? 3 : it / AT;
        const int a = is_rho ? it - 3 * AT : it - v * AT;
        const double* d = is_rho ? T.d_rho + (size_t)a * N : T.d_vib + (size_t)a * N;
        const double start = is_rho ? T.lpref_rho[a] : T.lpref[v * AT + a];
        double acc[MID_IB];
#pragma unroll
        for (int jj = 0; jj < MID_IB; ++jj) acc[jj] = start;
        for (int n = 0; n < N; ++n) {
            const double dn = d[n], hc = T.hc[v * N + n], cs = T.cs[v * N + n];
            double r[MID_IB + 1];
#pragma unroll
            for (int jj = 0; jj <= MID_IB; ++jj) r[jj] = Rt[n * MID_RS + jj];
#pragma unroll
            for (int jj = 0; jj < MID_IB; ++jj) {
                const double sp = (r[jj] - dn) + (r[jj + 1] - dn), sm = r[jj] - r[jj + 1];
                acc[jj] = fma(hc, sp * sp, fma(cs, sm * sm, acc[jj]));
            }
        }
        const bool dead = is_rho && a >= T.n_rho_eval;
#pragma unroll
        for (int jj = 0; jj < MID_IB; ++jj) lall[jj * nl + it] = dead ? -INFINITY : acc[jj];
    }
    __syncwarp();
    for (int jj = 0; jj < nb; ++jj) {
        const size_t xp = xp0 + jj;
        const double* lv = lall + jj * nl;          // [3][AT] then [Ar]
        double logS = -INFINITY;
        for (int a = lane; a < AT; a += 32) logS = fmax(logS, lv[a]);
        for (int a = lane; a < Ar; a += 32) logS = fmax(logS, lv[3 * AT + a]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) logS = fmax(logS, __shfl_xor_sync(0xffffffffu, logS, o));
        if (out.scale && lane == 0) out.scale[xp] = exp(logS);
        for (int it = lane; it < 3 * AT; it += 32) {
            const int v = it / AT, a = it - v * AT;
            if (out.o_vib) out.o_vib[((size_t)v * out.n * P + xp) * AT + a] = exp(lv[it] - logS);
        }
        for (int a = lane; a < Ar; a += 32) {
            if (out.lr) out.lr[xp * Ar + a] = lv[3 * AT + a] - logS;
            if (out.o_rho) out.o_rho[xp * Ar + a] = exp(lv[3 * AT + a] - logS);
        }
    }
    if (!out.v_mat && !out.m_mat) return;

    // ---- V for the 8 beads of the group on the FP64 tensor cores: V[bead][k] = sum_f feat_f(bead) coef[f][k] is an
    //      (8 x K) x (K x AA) product, K = N(N+1)/2 + N + 1 features (R_n R_m, R_n, 1).  One mma.sync.m8n8k4 per
    //      (4 features, 8 packed entries): lane (r = lane/4, c = lane%4) forms its A element -- feature 4 ks + c of
    //      bead r -- with one multiply, the B fragments are coalesced 256-byte rows of the pre-tiled table.
    constexpr int NT = MidShape<AT>::NT;
    double acc[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
    const int r = lane >> 2, c = lane & 3;
    const double* qb = T.q_dmma + lane;
    const double* Rr = Rt + r;
#pragma unroll 4
    for (int ks = 0; ks < T.KS; ++ks) {
        const int f = __ldg(T.feat + 4 * ks + c);
        const double a = Rr[(f & 0xffff) * MID_RS] * Rr[(f >> 16) * MID_RS];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double b = __ldg(qb + ((size_t)ks * NT + j) * 32);
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(acc[j][0]), "+d"(acc[j][1]) : "d"(a), "d"(b));
        }
    }
    // accumulator layout: lane holds V[bead r][k = 8 j + 2 c + e], e = 0, 1.  Two beads at a time go through the staging
    // tile as full symmetric matrices and leave as coalesced rows.
    int ij[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) ij[j][e] = __ldg(T.tri_ij + 8 * j + 2 * c + e);
    for (int pair = 0; 2 * pair < nb; ++pair) {
        if ((r >> 1) == pair) {
            double* dst = Xst + (size_t)(r & 1) * AA2;
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (ij[j][e] < 0) continue;
                    const int i = ij[j][e] >> 16, jj2 = ij[j][e] & 0xffff;
                    dst[i * AT + jj2] = acc[j][e];
                    dst[jj2 * AT + i] = acc[j][e];
                }
        }
        __syncwarp();
        const int nbp = min(2, nb - 2 * pair);
        for (int e = lane; e < nbp * AA2; e += 32) {
            const double v = Xst[e];
            if (out.v_mat) out.v_mat[(xp0 + 2 * pair) * AA2 + e] = v;
            if (out.m_mat) out.m_mat[(xp0 + 2 * pair) * AA2 + e] = v * T.neg_tau;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// per-bead stage, part 2: m_mat[item] <- exp(m_mat[item]) in place, one warp per (sample, bead) matrix.
// Persistent warps with a grid stride: the next matrix is fetched with cp.async while the current one is
// exponentiated (a warp that loaded its own matrix first sat on the HBM latency: long_scoreboard 7 per issue).
// ---------------------------------------------------------------------------------------------
template <int AT>
__global__ void __launch_bounds__(MID_WARPS * 32)
pbx_mid_expm_kernel(double* __restrict__ m_mat, long long n_items) {
    extern __shared__ __align__(16) double sm[];
    constexpr int AA2 = AT * AT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = (long long)gridDim.x * MID_WARPS;
    long long item = (long long)blockIdx.x * MID_WARPS + warp;
    if (item >= n_items) return;
    double* base = sm + (size_t)warp * mid_expm_warp_doubles(AT);
    double* Xb[2] = {base, base + AA2};
    double *W0 = base + 2 * AA2, *W1 = W0 + AA2, *W2 = W1 + AA2, *W3 = W2 + AA2;
    auto fetch = [&](long long it, double* dst) {
        const double* g = m_mat + (size_t)it * AA2;
        for (int e = lane; e < AA2; e += 32) cp_async8(dst + e, g + e);
        cp_async_commit();
    };
    fetch(item, Xb[0]);
    for (int b = 0; item < n_items; item += stride, b ^= 1) {
        cp_async_wait<0>();
        __syncwarp();
        const long long nxt = item + stride;
        if (nxt < n_items) fetch(nxt, Xb[b ^ 1]);
        MmaFrag<AT> m;
        warp_expm_at<AT>(Xb[b], W0, W1, W2, W3, m, lane);
        mma_store<AT>(m_mat + (size_t)item * AA2, m, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// chain: lane = (sample slot s, row i); floor(32/AT) samples per warp; M_p prefetched with cp.async
// ---------------------------------------------------------------------------------------------
constexpr int MID_DEPTH = 4;   // beads in flight per warp

__host__ __device__ inline size_t mid_chain_warp_doubles(int A) {
    return (size_t)MID_DEPTH * (32 / A) * ((size_t)A * A + 3 * A) + 32 * 3;
}

template <int AT, bool PM>
__global__ void __launch_bounds__(MID_WARPS * 32)
pbx_mid_chain_kernel(DevTables T, const double* __restrict__ m_mat, const double* __restrict__ o_vib,
                     const double* __restrict__ lr, long long n_samples, double* __restrict__ rho_out,
                     double* __restrict__ g_out, long long g_ld) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NV = PM ? 3 : 1, AA2 = AT * AT, SPW = 32 / AT, SLOT = SPW * (AA2 + 3 * AT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Ar = T.Ar, P = T.P;
    const long long xbase = ((long long)blockIdx.x * MID_WARPS + warp) * SPW;
    if (xbase >= n_samples) return;
    const int s = lane / AT, i = lane - s * AT;
    const bool owner = s < SPW && xbase + s < n_samples;
    const int ns = (int)min((long long)SPW, n_samples - xbase);     // samples of this warp
    double* w = sm + (size_t)warp * mid_chain_warp_doubles(AT);
    double* ring = w;                             // [MID_DEPTH][SPW][AA2 + 3*AT]
    double* tr = ring + (size_t)MID_DEPTH * SLOT; // [3][32]
    // ---- rho(x) = sum_a exp(sum_p lr[x][p][a]) for the warp's samples
    if (lr && rho_out) {
        for (int q = 0; q < ns; ++q) {
            const double* base = lr + (size_t)(xbase + q) * P * Ar;
            double rho = 0.0;
            for (int a = 0; a < Ar; ++a) {
                double part = 0.0;
                for (int p = lane; p < P; p += 32) part += base[(size_t)p * Ar + a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                rho += exp(part);
            }
            if (lane == 0) rho_out[xbase + q] = rho;
        }
    }
    auto prefetch = [&](int p) {
        if (p < P) {
            double* dst = ring + (size_t)(p % MID_DEPTH) * SLOT;
            for (int q = 0; q < ns; ++q) {
                const size_t xp = (size_t)(xbase + q) * P + p;
                const double* msrc = m_mat + xp * AA2;
                for (int e = lane; e < AA2; e += 32) cp_async8(dst + q * (AA2 + 3 * AT) + e, msrc + e);
                for (int e = lane; e < NV * AT; e += 32) {
                    const int v = e / AT, a = e - v * AT;
                    cp_async8(dst + q * (AA2 + 3 * AT) + AA2 + e, o_vib + ((size_t)v * n_samples * P + xp) * AT + a);
                }
            }
        }
        cp_async_commit();
    };
    double Trow[NV][AT];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < AT; ++k) Trow[v][k] = (k == i) ? 1.0 : 0.0;
#pragma unroll
    for (int d = 0; d < MID_DEPTH - 1; ++d) prefetch(d);
    for (int p = 0; p < P; ++p) {
        prefetch(p + MID_DEPTH - 1);
        cp_async_wait<MID_DEPTH - 1>();
        __syncwarp();
        if (owner) {
            const double* Mq = ring + (size_t)(p % MID_DEPTH) * SLOT + s * (AA2 + 3 * AT);
            const double* Oq = Mq + AA2;
            double acc[NV][AT];
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int j = 0; j < AT; ++j) acc[v][j] = 0.0;
            if constexpr (AT % 2 == 0) {      // every slot starts 16-byte aligned: two elements of M per shared load
#pragma unroll
                for (int k = 0; k < AT; ++k)
#pragma unroll
                    for (int j = 0; j < AT; j += 2) {
                        const double2 m = *reinterpret_cast<const double2*>(Mq + k * AT + j);
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            acc[v][j] = fma(Trow[v][k], m.x, acc[v][j]);
                            acc[v][j + 1] = fma(Trow[v][k], m.y, acc[v][j + 1]);
                        }
                    }
            } else {
#pragma unroll
                for (int k = 0; k < AT; ++k)
#pragma unroll
                    for (int j = 0; j < AT; ++j) {
                        const double m = Mq[k * AT + j];
#pragma unroll
                        for (int v = 0; v < NV; ++v) acc[v][j] = fma(Trow[v][k], m, acc[v][j]);
                    }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int j = 0; j < AT; ++j) Trow[v][j] = acc[v][j] * Oq[v * AT + j];
        }
        __syncwarp();
    }
    // ---- traces: lane (s, i) holds T_v[i][i]
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double d = 0.0;
#pragma unroll
        for (int k = 0; k < AT; ++k) d = (k == i) ? Trow[v][k] : d;
        tr[v * 32 + lane] = owner ? d : 0.0;
    }
    __syncwarp();
    if (owner && i == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double t = 0.0;
            for (int k = 0; k < AT; ++k) t += tr[v * 32 + s * AT + k];
            g_out[(size_t)v * g_ld + xbase + s] = t;
        }
    }
}

}  // namespace pbx

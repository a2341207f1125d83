// Blocked kernels for 1 <= A <= 16 surfaces (compile-time A) (BASELINE config c4: A=12, N=24, P=256), any N, A_rho.
// Same three steps through HBM scratch as pbx_generic.cuh (coords -> per-bead stage -> chain) but
// organised for reuse instead of one item at a time:
//
//   pbx_mid_sample_kernel : thread per (sample, MODE PAIR) -- the N ring recurrences are independent,
//                           so a 2400-sample chunk yields N/2 times more threads than thread-per-sample
//   pbx_mid_bead_kernel   : one warp per MID_IB = 8 consecutive beads of a sample.  The coupling matrix
//                           V[bead][k] = e_off[k] + sum_n R_n l_off[n][k] + sum_{n<=m} q[n,m][k] R_n R_m is a
//                           dense (8 beads x 325 features) x (325 x 78) contraction at c4: it runs on the FP64
//                           tensor cores (mma.sync.m8n8k4.f64), 10 MMAs per 4 features, the table pre-tiled in
//                           fragment order (210 KB, read as coalesced 256-byte rows from L1/L2).
//                           exp(-tau V) uses a register-blocked 4x8-lane product (5 shared loads per
//                           6 FMAs at A=12 instead of 2 per FMA).
//   pbx_mid_chain_kernel  : lane = (sample, row i): the rows of the three chained products live in
//                           registers, M_p is broadcast from shared memory, floor(32/A) samples per warp.
//
// Reference: /root/reference/pibronic/pimc/pimc.py:1087-1129 (O), 1076-1084 (S), 1139-1187 (V, M),
// 1194-1209 (chain), 1132-1136 (rho).
#pragma once
#include "pbx_generic.cuh"

namespace pbx {

constexpr int MID_IB = 8;        // beads per warp in the bead kernel
constexpr int MID_RS = MID_IB + 2;  // row stride of the coordinate tile (even: 16-byte aligned rows)
constexpr int MID_WARPS = 4;
constexpr int MID_AMAX = 16;

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
    for (int w = 0; w < 2; ++w) shift[w] = (2 * h + w < N) ? T.d_rho[src * N + 2 * h + w] : 0.0;
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
// register-blocked warp product of AT x AT matrices in shared memory (AT compile time), lanes as a
// 4 x 8 grid: lane (li, lj) owns rows li + 4*ri (ri < RI) and columns lj + 8*cj (cj < CJ).
// Out-of-range rows/columns are clamped (computed, never stored) so the inner loop has no predicates.
// ---------------------------------------------------------------------------------------------
template <int AT> struct MidShape {
    static constexpr int RI = (AT + 3) / 4, CJ = (AT + 7) / 8;
    static constexpr int AA = AT * (AT + 1) / 2, NT = (AA + 7) / 8, AA2 = AT * AT;   // NT: 8-wide mma tiles over the packed entries
};

template <int AT>
struct LaneBlock {   // the entries of a matrix owned by one lane
    double v[MidShape<AT>::RI][MidShape<AT>::CJ];
};

template <int AT>
__device__ __forceinline__ void blk_matmul(const double* __restrict__ X, const double* __restrict__ Y, LaneBlock<AT>& C,
                                           int li, int lj) {
    constexpr int RI = MidShape<AT>::RI, CJ = MidShape<AT>::CJ;
    int row[RI], col[CJ];
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) row[ri] = min(li + 4 * ri, AT - 1) * AT;
#pragma unroll
    for (int cj = 0; cj < CJ; ++cj) col[cj] = min(lj + 8 * cj, AT - 1);
#pragma unroll
    for (int ri = 0; ri < RI; ++ri)
#pragma unroll
        for (int cj = 0; cj < CJ; ++cj) C.v[ri][cj] = 0.0;
#pragma unroll
    for (int k = 0; k < AT; ++k) {
        double xv[RI], yv[CJ];
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) xv[ri] = X[row[ri] + k];
#pragma unroll
        for (int cj = 0; cj < CJ; ++cj) yv[cj] = Y[k * AT + col[cj]];
#pragma unroll
        for (int ri = 0; ri < RI; ++ri)
#pragma unroll
            for (int cj = 0; cj < CJ; ++cj) C.v[ri][cj] = fma(xv[ri], yv[cj], C.v[ri][cj]);
    }
}

template <int AT>
__device__ __forceinline__ void blk_store(double* __restrict__ dst, const LaneBlock<AT>& C, int li, int lj) {
#pragma unroll
    for (int ri = 0; ri < MidShape<AT>::RI; ++ri)
#pragma unroll
        for (int cj = 0; cj < MidShape<AT>::CJ; ++cj) {
            const int i = li + 4 * ri, j = lj + 8 * cj;
            if (i < AT && j < AT) dst[i * AT + j] = C.v[ri][cj];
        }
}

template <int AT>
__device__ __forceinline__ void blk_load(const double* __restrict__ src, LaneBlock<AT>& C, int li, int lj) {
#pragma unroll
    for (int ri = 0; ri < MidShape<AT>::RI; ++ri)
#pragma unroll
        for (int cj = 0; cj < MidShape<AT>::CJ; ++cj)
            C.v[ri][cj] = src[min(li + 4 * ri, AT - 1) * AT + min(lj + 8 * cj, AT - 1)];
}

// M = exp(X) (degree-12 Taylor in four products, scaling and squaring as in sym_expm); X in shared
// memory is overwritten by its scaled copy; W0..W3 are AT*AT work matrices; the result is left in the
// lane blocks `out` AND in shared memory at the returned pointer.
template <int AT>
__device__ __forceinline__ const double* warp_expm_at(double* X, double* W0, double* W1, double* W2, double* W3,
                                                      LaneBlock<AT>& out, int lane) {
    constexpr int RI = MidShape<AT>::RI, CJ = MidShape<AT>::CJ, AA2 = AT * AT;
    const int li = lane >> 3, lj = lane & 7;
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
    LaneBlock<AT> x1, x2, x3, y0, b;
    auto diag = [&](int ri, int cj) { return (li + 4 * ri) == (lj + 8 * cj); };
    blk_load<AT>(X, x1, li, lj);
    blk_matmul<AT>(X, X, x2, li, lj);
    blk_store<AT>(W0, x2, li, lj);          // W0 = X^2
    __syncwarp();
    blk_matmul<AT>(X, W0, x3, li, lj);
    blk_store<AT>(W1, x3, li, lj);          // W1 = X^3
#pragma unroll
    for (int ri = 0; ri < RI; ++ri)
#pragma unroll
        for (int cj = 0; cj < CJ; ++cj)
            b.v[ri][cj] = fma(t12::c1, x3.v[ri][cj], fma(t12::c2, x2.v[ri][cj], t12::c3 * x1.v[ri][cj]));
    blk_store<AT>(W2, b, li, lj);
    __syncwarp();
    blk_matmul<AT>(W1, W2, y0, li, lj);     // Y0 = X^3 (c1 X^3 + c2 X^2 + c3 X)
    __syncwarp();                           // W2 is rewritten below
#pragma unroll
    for (int ri = 0; ri < RI; ++ri)
#pragma unroll
        for (int cj = 0; cj < CJ; ++cj) {
            b.v[ri][cj] = y0.v[ri][cj] + fma(t12::c4, x3.v[ri][cj], fma(t12::c5, x2.v[ri][cj], t12::c6 * x1.v[ri][cj]));
            out.v[ri][cj] = y0.v[ri][cj] + fma(t12::c7, x3.v[ri][cj], t12::c8 * x2.v[ri][cj]);
        }
    blk_store<AT>(W2, b, li, lj);
    blk_store<AT>(W3, out, li, lj);
    __syncwarp();
    blk_matmul<AT>(W2, W3, out, li, lj);
#pragma unroll
    for (int ri = 0; ri < RI; ++ri)
#pragma unroll
        for (int cj = 0; cj < CJ; ++cj)
            out.v[ri][cj] = out.v[ri][cj] + fma(t12::c9, y0.v[ri][cj], fma(t12::c10, x3.v[ri][cj], fma(0.5, x2.v[ri][cj], x1.v[ri][cj]))) +
                            (diag(ri, cj) ? 1.0 : 0.0);
    __syncwarp();                           // all lanes are done reading W2, W3
    double* cur = W2;
    double* nxt = W3;
    blk_store<AT>(cur, out, li, lj);
    __syncwarp();
    for (int q = 0; q < s; ++q) {
        blk_matmul<AT>(cur, cur, out, li, lj);
        blk_store<AT>(nxt, out, li, lj);
        __syncwarp();
        double* t = cur; cur = nxt; nxt = t;
    }
    return cur;
}

// shared memory (doubles) of one warp of the bead kernel
__host__ __device__ inline size_t mid_bead_warp_doubles(int A, int Ar, int N) {
    size_t n = (size_t)(N + 1) * MID_RS + (size_t)MID_IB * A * A + 4 * (size_t)A * A + (size_t)MID_IB * (3 * A + Ar);
    return (n + 1) & ~(size_t)1;
}

// ---------------------------------------------------------------------------------------------
// per-bead stage: one warp per MID_IB consecutive beads of one sample; AT = number of surfaces
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
    double* Xs = Rt + (size_t)(N + 1) * MID_RS;   // [MID_IB][AT][AT]
    double* W0 = Xs + (size_t)MID_IB * AA2;       // 4 work matrices
    double *W1 = W0 + AA2, *W2 = W1 + AA2, *W3 = W2 + AA2;
    double* lall = W3 + AA2;                      // [MID_IB][3*AT + Ar] log factors
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

    // ---- V for the 8 beads of the group on the FP64 tensor cores: V[bead][k] = sum_f feat_f(bead) coef[f][k] is an
    //      (8 x K) x (K x AA) product, K = N(N+1)/2 + N + 1 features (R_n R_m, R_n, 1).  One mma.sync.m8n8k4 per
    //      (4 features, 8 packed entries): lane (r = lane/4, c = lane%4) forms its A element -- feature 4 ks + c of
    //      bead r -- with one multiply, the B fragments are coalesced 256-byte rows of the pre-tiled table.
    if (out.v_mat || out.m_mat) {
        constexpr int NT = MidShape<AT>::NT;
        double acc[NT][2];
#pragma unroll
        for (int j = 0; j < NT; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
        const int r = lane >> 2, c = lane & 3;
        const double* qb = T.q_dmma + lane;
        const double* Rr = Rt + r;
#pragma unroll 2
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
        // accumulator layout: lane holds V[bead r][k = 8 j + 2 c + e], e = 0, 1
        if (r < nb) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ij = __ldg(T.tri_ij + 8 * j + 2 * c + e);
                    if (ij < 0) continue;
                    const int i = ij >> 16, jj2 = ij & 0xffff;
                    const double v = acc[j][e];
                    if (out.v_mat) {
                        out.v_mat[(xp0 + r) * AA2 + i * AT + jj2] = v;
                        out.v_mat[(xp0 + r) * AA2 + jj2 * AT + i] = v;
                    }
                    Xs[(size_t)r * AA2 + i * AT + jj2] = v * T.neg_tau;
                    Xs[(size_t)r * AA2 + jj2 * AT + i] = v * T.neg_tau;
                }
        }
        __syncwarp();
    }

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
            double q[MID_IB + 1];
#pragma unroll
            for (int jj = 0; jj <= MID_IB; ++jj) q[jj] = Rt[n * MID_RS + jj] - dn;
#pragma unroll
            for (int jj = 0; jj < MID_IB; ++jj)
                acc[jj] = fma(hc, fma(q[jj], q[jj], q[jj + 1] * q[jj + 1]), fma(cs, q[jj] * q[jj + 1], acc[jj]));
        }
        const bool dead = is_rho && a >= T.n_rho_eval;
#pragma unroll
        for (int jj = 0; jj < MID_IB; ++jj) lall[jj * nl + it] = dead ? -INFINITY : acc[jj];
    }
    __syncwarp();

    const int li = lane >> 3, lj = lane & 7;
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
        if (out.m_mat) {
            LaneBlock<AT> m;
            warp_expm_at<AT>(Xs + (size_t)jj * AA2, W0, W1, W2, W3, m, lane);
            blk_store<AT>(out.m_mat + xp * AA2, m, li, lj);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// chain: lane = (sample slot s, row i); floor(32/AT) samples per warp; M_p prefetched with cp.async
// ---------------------------------------------------------------------------------------------
constexpr int MID_DEPTH = 4;   // beads in flight per warp

__host__ __device__ inline size_t mid_chain_warp_doubles(int A) {
    return (size_t)MID_DEPTH * (32 / A) * ((size_t)A * A + 3 * A) + 32 * 3;
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N_)); }

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
#pragma unroll
            for (int k = 0; k < AT; ++k)
#pragma unroll
                for (int j = 0; j < AT; ++j) {
                    const double m = Mq[k * AT + j];
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v][j] = fma(Trow[v][k], m, acc[v][j]);
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

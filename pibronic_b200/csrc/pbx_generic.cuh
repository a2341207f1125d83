// Generic (any A, N, A_rho) kernels: the universal fallback (A > 16, Jacobi M builder, unscaled
// stage-by-stage introspection API); 1 <= A <= 16 normally runs on the blocked kernels of pbx_mid.cuh.  Three steps through HBM scratch:
//
//   coords  : R[x][n][p]                    (caller supplied, or pbx_mid_sample_kernel in pbx_mid.cuh)
//   beads   : one WARP per (sample, bead):  O factors (log space), scale S, V, M = exp(-tau V)
//             -> o_vib[3][x][P][A], lr[x][P][Ar], scale[x][P], v_mat / m_mat[x][P][A][A]
//   chain   : one WARP per sample:          T_v <- (T_v M_p) diag(O_v[p]),  rho from the lr sums
//
// Reference: /root/reference/pibronic/pimc/pimc.py:1087-1129 (O), 1076-1084 (S), 1139-1187 (V, M),
// 1194-1209 (chain), 1132-1136 (rho).
#pragma once
#include "pbx_device.cuh"

namespace pbx {

// device copy of the tables (pointers into one allocation)
struct DevTables {
    int A, Ar, N, P, AA, NN, n_rho_eval;
    double neg_tau;
    const double *d_vib, *d_rho;     // [A][N], [Ar][N]: shifts in the harmonic exponents
    const double *d_rho_samp;        // [Ar][N]: shift of the drawn mixture component (== d_rho unless PBX_QUIRK_RHO_DOUBLE_SHIFT)
    const double *hc, *cs;           // [4][N]  (-1/4 tanh(x/2), -1/4 coth(x/2)), x = tau omega
    const double *lpref, *lpref_rho; // [3][A], [Ar]
    const double *wcum;              // [Ar]
    const double *e_off, *l_off, *q_pack;  // [AA], [N][AA], [NN][AA]
    const double *samp;              // [P][N][3]
    // coupling tables in FP64 tensor-core (mma.sync m8n8k4) fragment order, pbx_mid_bead_kernel:
    // V[bead][k] = sum_f feature_f(bead) * coef[f][k], features = R_n R_m (n <= m), R_n, 1
    const double *q_dmma;            // [KS][NT][32]: coef[4 ks + lane%4][8 j + lane/4]
    const int *feat;                 // [4 KS]: rows (ra | rb << 16) of the coordinate tile whose product is feature f
    const int *tri_ij;               // [8 NT]: (i << 16 | j) of packed entry k, -1 beyond AA
    int KS, NT;
};

struct BeadOutputs {
    double* o_vib;   // [3][n][P][A] (rows tau, tau+, tau-), divided by S; may be null
    double* o_rho;   // [n][P][Ar] divided by S; may be null
    double* lr;      // [n][P][Ar] log(O_rho/S); may be null
    double* scale;   // [n][P]; may be null
    double* v_mat;   // [n][P][A][A]; may be null
    double* m_mat;   // [n][P][A][A]; may be null
    long long n;     // samples in this launch (stride of the variant axis of o_vib)
};

constexpr int GEN_WARPS = 4;  // warps per CTA in the bead / chain kernels

// shared memory (doubles) needed by one warp of the bead kernel
__host__ __device__ inline size_t bead_warp_doubles(int A, int Ar, int N) {
    return 2 * (size_t)N + 3 * (size_t)A + Ar + 6 * (size_t)A * A;
}

// ---------------------------------------------------------------------------------------------
// warp-cooperative dense product of small square matrices in shared memory: C = X * Y
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_matmul(const double* X, const double* Y, double* C, int A, int lane) {
    for (int e = lane; e < A * A; e += 32) {
        const int i = e / A, j = e - i * A;
        double acc = 0.0;
        for (int k = 0; k < A; ++k) acc = fma(X[i * A + k], Y[k * A + j], acc);
        C[e] = acc;
    }
    __syncwarp();
}

// M = exp(X) in shared memory, same algorithm as sym_expm (degree-12 Taylor, scaling and squaring)
__device__ __forceinline__ void warp_expm(double* X, double* X2, double* X3, double* X4, double* B, double* M,
                                          int A, int lane) {
    double norm = 0.0;
    for (int i = lane; i < A; i += 32) {
        double row = 0.0;
        for (int j = 0; j < A; ++j) row += fabs(X[i * A + j]);
        norm = fmax(norm, row);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) norm = fmax(norm, __shfl_xor_sync(0xffffffffu, norm, o));
    const int s = expm_squarings(norm);
    const double scale = __hiloint2double((1023 - s) << 20, 0);
    const int AA2 = A * A;
    for (int e = lane; e < AA2; e += 32) X[e] *= scale;
    __syncwarp();
    warp_matmul(X, X, X2, A, lane);
    warp_matmul(X, X2, X3, A, lane);
    warp_matmul(X2, X2, X4, A, lane);
    constexpr double c2 = 1.0 / 2, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720, c7 = 1.0 / 5040,
                     c8 = 1.0 / 40320, c9 = 1.0 / 362880, c10 = 1.0 / 3628800, c11 = 1.0 / 39916800,
                     c12 = 1.0 / 479001600;
    for (int e = lane; e < AA2; e += 32) {
        const int i = e / A, j = e - i * A;
        B[e] = fma(c12, X4[e], fma(c11, X3[e], fma(c10, X2[e], c9 * X[e]))) + (i == j ? c8 : 0.0);
    }
    __syncwarp();
    warp_matmul(X4, B, M, A, lane);
    for (int e = lane; e < AA2; e += 32) {
        const int i = e / A, j = e - i * A;
        B[e] = M[e] + fma(c7, X3[e], fma(c6, X2[e], c5 * X[e])) + (i == j ? c4 : 0.0);
    }
    __syncwarp();
    warp_matmul(X4, B, M, A, lane);
    for (int e = lane; e < AA2; e += 32) {
        const int i = e / A, j = e - i * A;
        M[e] = M[e] + fma(c3, X3[e], fma(c2, X2[e], X[e])) + (i == j ? 1.0 : 0.0);
    }
    __syncwarp();
    // squarings ping-pong between M and B
    double* cur = M;
    double* nxt = B;
    for (int q = 0; q < s; ++q) {
        warp_matmul(cur, cur, nxt, A, lane);
        double* t = cur; cur = nxt; nxt = t;
    }
    if (cur != M) {
        for (int e = lane; e < AA2; e += 32) M[e] = cur[e];
        __syncwarp();
    }
}

// M = U exp(lambda) U^T by cyclic Jacobi in shared memory: rotations sequential, the row/column
// updates of each rotation spread over the lanes.  Slow; kept as the reference-formulation check.
__device__ __forceinline__ void warp_exp_jacobi(double* S, double* U, double* M, int A, int lane) {
    const int AA2 = A * A;
    for (int e = lane; e < AA2; e += 32) U[e] = (e / A == e % A) ? 1.0 : 0.0;
    double fro = 0.0;
    for (int e = lane; e < AA2; e += 32) fro = fma(S[e], S[e], fro);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fro += __shfl_xor_sync(0xffffffffu, fro, o);
    const double tol = 1e-33 * fro;
    __syncwarp();
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int e = lane; e < AA2; e += 32) { const int i = e / A, j = e - i * A; if (i != j) off = fma(S[e], S[e], off); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
        if (off <= 2.0 * tol) break;
        for (int p = 0; p < A - 1; ++p)
            for (int q = p + 1; q < A; ++q) {
                const double apq = S[p * A + q], app = S[p * A + p], aqq = S[q * A + q];
                __syncwarp();
                if (fabs(apq) <= 1e-300) continue;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                const double c = rsqrt(fma(t, t, 1.0)), sn = t * c;
                for (int k = lane; k < A; k += 32) {
                    if (k != p && k != q) {
                        const double akp = S[k * A + p], akq = S[k * A + q];
                        const double np_ = fma(c, akp, -sn * akq), nq_ = fma(sn, akp, c * akq);
                        S[k * A + p] = np_; S[p * A + k] = np_;
                        S[k * A + q] = nq_; S[q * A + k] = nq_;
                    }
                    const double ukp = U[k * A + p], ukq = U[k * A + q];
                    U[k * A + p] = fma(c, ukp, -sn * ukq);
                    U[k * A + q] = fma(sn, ukp, c * ukq);
                }
                if (lane == 0) {
                    S[p * A + p] = fma(-t, apq, app);
                    S[q * A + q] = fma(t, apq, aqq);
                    S[p * A + q] = 0.0; S[q * A + p] = 0.0;
                }
                __syncwarp();
            }
    }
    __syncwarp();
    for (int e = lane; e < AA2; e += 32) {
        const int i = e / A, j = e - i * A;
        double acc = 0.0;
        for (int k = 0; k < A; ++k) acc = fma(U[i * A + k] * exp(S[k * A + k]), U[j * A + k], acc);
        M[e] = acc;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// per-bead stage: one warp per (sample, bead)
// ---------------------------------------------------------------------------------------------
template <bool JACOBI, bool SCALE>
__global__ void __launch_bounds__(GEN_WARPS * 32)
pbx_bead_kernel(DevTables T, const double* __restrict__ R, long long n_samples, BeadOutputs out) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int A = T.A, Ar = T.Ar, N = T.N, P = T.P, AA2 = A * A;
    const long long item = (long long)blockIdx.x * GEN_WARPS + warp;
    if (item >= n_samples * P) return;
    const long long x = item / P;
    const int p = (int)(item - x * P), pn = (p + 1 == P) ? 0 : p + 1;
    double* w = sm + (size_t)warp * bead_warp_doubles(A, Ar, N);
    double* Rc = w; double* Rn = Rc + N;
    double* lv = Rn + N;            // [3][A]
    double* lr = lv + 3 * A;        // [Ar]
    double* X = lr + Ar;            // 6 matrices
    double *X2 = X + AA2, *X3 = X2 + AA2, *X4 = X3 + AA2, *B = X4 + AA2, *M = B + AA2;
    const double* Rx = R + (size_t)x * N * P;
    for (int n = lane; n < N; n += 32) { Rc[n] = Rx[(size_t)n * P + p]; Rn[n] = Rx[(size_t)n * P + pn]; }
    __syncwarp();
    // ---- O factors in log space; work items (set, surface): 3*A vib + Ar rho
    for (int it = lane; it < 3 * A + Ar; it += 32) {
        const bool is_rho = it >= 3 * A;
        const int v = is_rho ? 3 : it / A;
        const int a = is_rho ? it - 3 * A : it - v * A;
        const double* d = is_rho ? T.d_rho + (size_t)a * N : T.d_vib + (size_t)a * N;
        double acc = is_rho ? T.lpref_rho[a] : T.lpref[v * A + a];
        for (int n = 0; n < N; ++n) {
            const double q = Rc[n] - d[n], qn = Rn[n] - d[n];
            const double sp = q + qn, sm = Rc[n] - Rn[n];
            acc = fma(T.hc[v * N + n], sp * sp, fma(T.cs[v * N + n], sm * sm, acc));
        }
        if (is_rho) lr[a] = (a < T.n_rho_eval) ? acc : -INFINITY;
        else lv[it] = acc;
    }
    __syncwarp();
    double logS = 0.0;
    if (SCALE) {
        logS = -INFINITY;
        for (int a = lane; a < A; a += 32) logS = fmax(logS, lv[a]);
        for (int a = lane; a < Ar; a += 32) logS = fmax(logS, lr[a]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) logS = fmax(logS, __shfl_xor_sync(0xffffffffu, logS, o));
    }
    const size_t xp = (size_t)x * P + p;
    if (out.scale && lane == 0) out.scale[xp] = exp(logS);
    for (int it = lane; it < 3 * A; it += 32) {
        const int v = it / A, a = it - v * A;
        if (out.o_vib) out.o_vib[((size_t)v * out.n * P + xp) * A + a] = exp(lv[it] - logS);
    }
    for (int a = lane; a < Ar; a += 32) {
        if (out.lr) out.lr[xp * Ar + a] = lr[a] - logS;
        if (out.o_rho) out.o_rho[xp * Ar + a] = exp(lr[a] - logS);
    }
    if (!out.v_mat && !out.m_mat) return;
    // ---- V (packed entries over lanes), written to X as a full symmetric matrix scaled by -tau
    const int AA = T.AA;
    for (int k = lane; k < AA; k += 32) {
        // invert k -> (i, j), i >= j
        int i = (int)((sqrt(8.0 * k + 1.0) - 1.0) * 0.5);
        while (tri(i + 1, 0) <= k) ++i;
        while (tri(i, 0) > k) --i;
        const int j = k - tri(i, 0);
        double acc = T.e_off[k];
        for (int n = 0; n < N; ++n) acc = fma(T.l_off[(size_t)n * AA + k], Rc[n], acc);
        const double* qp = T.q_pack + k;
        for (int n = 0; n < N; ++n) {
            double inner = 0.0;
            for (int m = n; m < N; ++m) { inner = fma(qp[0], Rc[m], inner); qp += AA; }
            acc = fma(inner, Rc[n], acc);
        }
        if (out.v_mat) { out.v_mat[xp * AA2 + i * A + j] = acc; out.v_mat[xp * AA2 + j * A + i] = acc; }
        X[i * A + j] = acc * T.neg_tau; X[j * A + i] = acc * T.neg_tau;
    }
    __syncwarp();
    if (!out.m_mat) return;
    if (JACOBI) warp_exp_jacobi(X, X2, M, A, lane);
    else warp_expm(X, X2, X3, X4, B, M, A, lane);
    for (int e = lane; e < AA2; e += 32) out.m_mat[xp * AA2 + e] = M[e];
}

// ---------------------------------------------------------------------------------------------
// chain: one warp per sample; T_v (NV of them) double-buffered in shared memory
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t chain_warp_doubles(int A, int Ar) { return 7 * (size_t)A * A + 3 * (size_t)A + Ar; }

template <bool PM>
__global__ void __launch_bounds__(GEN_WARPS * 32)
pbx_chain_kernel(DevTables T, const double* __restrict__ m_mat, const double* __restrict__ o_vib,
                 const double* __restrict__ lr, long long n_samples, double* __restrict__ rho_out,
                 double* __restrict__ g_out, long long g_ld) {
    extern __shared__ double sm[];
    constexpr int NV = PM ? 3 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int A = T.A, Ar = T.Ar, P = T.P, AA2 = A * A;
    const long long x = (long long)blockIdx.x * GEN_WARPS + warp;
    if (x >= n_samples) return;
    double* w = sm + (size_t)warp * chain_warp_doubles(A, Ar);
    double* Tcur = w;                 // [3][A][A]
    double* Tnew = Tcur + 3 * AA2;    // [3][A][A]
    double* Mp = Tnew + 3 * AA2;      // [A][A]
    double* Op = Mp + AA2;            // [3][A]
    double* lsum = Op + 3 * A;        // [Ar]
    for (int e = lane; e < NV * AA2; e += 32) { const int r = e % AA2; Tcur[e] = (r / A == r % A) ? 1.0 : 0.0; }
    for (int a = lane; a < Ar; a += 32) lsum[a] = 0.0;
    __syncwarp();
    for (int p = 0; p < P; ++p) {
        const size_t xp = (size_t)x * P + p;
        for (int e = lane; e < AA2; e += 32) Mp[e] = m_mat[xp * AA2 + e];
        for (int it = lane; it < NV * A; it += 32) {
            const int v = it / A, a = it - v * A;
            Op[it] = o_vib[((size_t)v * n_samples * P + xp) * A + a];
        }
        if (lr) for (int a = lane; a < Ar; a += 32) lsum[a] += lr[xp * Ar + a];
        __syncwarp();
        for (int e = lane; e < NV * AA2; e += 32) {
            const int v = e / AA2, r = e - v * AA2, i = r / A, j = r - i * A;
            const double* Trow = Tcur + v * AA2 + i * A;
            double acc = 0.0;
            for (int k = 0; k < A; ++k) acc = fma(Trow[k], Mp[k * A + j], acc);
            Tnew[e] = acc * Op[v * A + j];
        }
        __syncwarp();
        double* t = Tcur; Tcur = Tnew; Tnew = t;
    }
    if (rho_out) {
        double rho = 0.0;
        for (int a = lane; a < Ar; a += 32) rho += exp(lsum[a]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rho += __shfl_xor_sync(0xffffffffu, rho, o);
        if (lane == 0) rho_out[x] = rho;
    }
    for (int v = 0; v < NV; ++v) {
        double tr = 0.0;
        for (int i = lane; i < A; i += 32) tr += Tcur[v * AA2 + i * A + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
        if (lane == 0) g_out[(size_t)v * g_ld + x] = tr;
    }
}

}  // namespace pbx

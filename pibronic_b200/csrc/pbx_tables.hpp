// Host-side precompute of every temperature/model dependent constant of the PIMC estimator.
//
// Follows (without copying) the reference's set-up phase, /root/reference/pibronic/pimc/pimc.py:
//   TemperatureDependentClass.__init__   59-89    coth, csch, O-matrix prefactor for one tau
//   ModelVibronic.precompute             172-177, 208-235   Delta_a, shift d_a, zeroing of diag(E), diag(L)
//   ModelVibronicPM.precompute           261-269  tau+ / tau-
//   ModelSampling.precompute             336-351, 372-406   Delta, d, mixture weights, sigma
// Differences by design (DESIGN.md "sampler"): prefactors are kept as logarithms, the coupling
// tables are packed symmetric, and instead of sigma[n][k] in the ring normal-mode basis the
// sampler uses a sequential (cyclic-tridiagonal Cholesky) recurrence with the same covariance.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pbx.h"

namespace pbx {

// packed lower-triangle index of a symmetric A x A matrix, i >= j
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }
__host__ __device__ constexpr int sym(int i, int j) { return i >= j ? tri(i, j) : tri(j, i); }
// index of the mode pair (n <= m) in row-major upper-triangle order
__host__ __device__ constexpr int pair_index(int n, int m, int N) { return n * N - n * (n - 1) / 2 + (m - n); }

struct HostTables {
    int A = 0, Ar = 0, N = 0, P = 0, AA = 0, NN = 0;
    int n_rho_eval = 0;  // sampling surfaces that enter rho(R) (quirk Q1 => min(A, Ar))
    uint32_t flags = 0;
    double beta = 0, delta_beta = 0;
    double tau[3] = {0, 0, 0};  // tau, tau+, tau-
    std::vector<double> d_vib, d_rho, delta_vib, delta_rho, weights, wcum;
    std::vector<double> d_rho_eval;       // shift used in rho's harmonic exponent: d_rho (+ d_vib under PBX_QUIRK_RHO_DOUBLE_SHIFT)
    std::vector<double> coth, csch;       // [4][N]: vib tau, vib tau+, vib tau-, rho tau (rho's own omega)
    std::vector<double> logpref;          // [3][A]
    // differences of the tau+ / tau- tables from the tau ones, formed analytically (not by subtracting
    // rounded doubles): rows 0 = tau+ - tau, 1 = tau- - tau.  Used to get O(tau+-) = O(tau) * exp(delta).
    std::vector<double> dcoth, dcsch;     // [2][N]
    // half-angle form of the harmonic exponent (no cancellation between the two terms):
    //   coth(x)(q^2+q'^2) - 2 csch(x) q q' = 1/2 [ tanh(x/2) (q+q')^2 + coth(x/2) (q-q')^2 ]
    std::vector<double> tanh_half, coth_half;     // [4][N] rows like coth/csch
    std::vector<double> dtanh_half, dcoth_half;   // [2][N] analytic differences tau+- minus tau
    std::vector<double> dlogpref;         // [2][A]
    std::vector<double> logpref_rho;      // [Ar]
    std::vector<double> e_off;            // [AA]   diagonal entries zero
    std::vector<double> l_off;            // [N][AA] diagonal entries zero
    std::vector<double> q_pack;           // [NN][AA] 0.5*(Q[n,m]+Q[m,n]) for n<m, 0.5*Q[n,n] for n==m
    std::vector<double> samp;             // [P][N][3]: a, b, e of y_j = a z + b y_{j-1} + e y_0
    bool has_quadratic = false;
    // true when sampling surface a has the same shift and frequencies as system surface a for every
    // a < min(A, Ar) -- e.g. rho = diagonal of the model (vIO.create_basic_diagonal_model): the harmonic
    // exponent sums of rho and of the tau variant are then the same numbers
    bool rho_shares_vib = false;
};

inline bool close_enough(double x, double y) { return std::fabs(x - y) <= 1e-8 + 1e-5 * std::fabs(y); }

// Cyclic-tridiagonal Cholesky of Lambda = alpha*I - s*C (C = ring adjacency), O(P), long double.
// Output recurrence for beads generated in order j = 0..P-1:
//   y_0 = a_0 z_0;  y_1 = a_1 z_1 + b_1 y_0;  y_j = a_j z_j + b_j y_{j-1} + e_j y_0
// so that cov(y) = Lambda^{-1} exactly.
inline void ring_recurrence(long double alpha, long double s, int P, std::vector<long double>& a,
                            std::vector<long double>& b, std::vector<long double>& e) {
    const long double beta = -s;
    std::vector<long double> diag(P), sub(P, 0.0L), corner(P, 0.0L);  // L[i][i], L[i+1][i], L[P-1][i]
    // columns 0 .. P-3
    long double prev_sub = 0.0L;  // L[i][i-1]
    for (int i = 0; i <= P - 3; ++i) {
        diag[i] = std::sqrt(alpha - prev_sub * prev_sub);
        sub[i] = beta / diag[i];
        corner[i] = (i == 0) ? beta / diag[0] : (-corner[i - 1] * prev_sub) / diag[i];
        prev_sub = sub[i];
    }
    // column P-2: its sub-diagonal element IS the corner row element
    diag[P - 2] = std::sqrt(alpha - prev_sub * prev_sub);
    const long double last_sub = (beta - (P >= 3 ? corner[P - 3] * prev_sub : 0.0L)) / diag[P - 2];
    // P == 3: corner[0] and sub[0]... handled since for P==3 column 0 has sub[0]=L[1][0], corner[0]=L[2][0]
    long double acc = alpha - last_sub * last_sub;
    for (int i = 0; i <= P - 3; ++i) acc -= corner[i] * corner[i];
    diag[P - 1] = std::sqrt(acc);
    sub[P - 2] = last_sub;
    a.assign(P, 0.0L); b.assign(P, 0.0L); e.assign(P, 0.0L);
    for (int j = 0; j < P; ++j) {
        const int i = P - 1 - j;  // row of L^T x = z solved at step j
        a[j] = 1.0L / diag[i];
        if (j == 0) continue;
        b[j] = -sub[i] / diag[i];              // coefficient of x_{i+1} = y_{j-1}
        if (j >= 2) e[j] = -corner[i] / diag[i];  // coefficient of x_{P-1} = y_0
    }
}

inline int build_tables(const pbx_model* vib, const pbx_rho* rho, int P, double beta, double delta_beta,
                        uint32_t flags, HostTables& T, std::string& err) {
    if (!vib || !rho || !vib->energy || !vib->omega || !rho->energy || !rho->omega) {
        err = "null model pointer"; return PBX_ERR_ARG;
    }
    const int A = vib->A, N = vib->N, Ar = rho->A;
    if (A < 1 || N < 1 || Ar < 1) { err = "A, N, A_rho must be >= 1"; return PBX_ERR_ARG; }
    if (P < 3) { err = "circulant matrix requires 3 or more beads"; return PBX_ERR_ARG; }
    if (rho->N != N) { err = "sampling model has a different number of modes"; return PBX_ERR_MODEL; }
    if (!(beta > 0) || !(delta_beta >= 0) || !(delta_beta < beta)) { err = "need 0 <= delta_beta < beta"; return PBX_ERR_ARG; }
    for (int n = 0; n < N; ++n)
        if (!(vib->omega[n] > 0) || !(rho->omega[n] > 0)) { err = "frequencies must be positive"; return PBX_ERR_MODEL; }
    T = HostTables();
    T.A = A; T.Ar = Ar; T.N = N; T.P = P; T.AA = A * (A + 1) / 2; T.NN = N * (N + 1) / 2;
    T.flags = flags; T.beta = beta; T.delta_beta = delta_beta;
    T.n_rho_eval = (flags & PBX_QUIRK_RHO_TRUNC) ? (A < Ar ? A : Ar) : Ar;
    T.tau[0] = beta / P; T.tau[1] = (beta + delta_beta) / P; T.tau[2] = (beta - delta_beta) / P;
    auto E = [&](int i, int j) { return vib->energy[i * A + j]; };
    auto L = [&](int n, int i, int j) { return vib->linear ? vib->linear[(n * A + i) * A + j] : 0.0; };
    auto Q = [&](int n, int m, int i, int j) {
        return vib->quadratic ? vib->quadratic[((n * N + m) * A + i) * A + j] : 0.0; };
    // the reference asserts symmetry of the coupling matrix every block (pimc.py:1169); check once here
    for (int i = 0; i < A; ++i)
        for (int j = 0; j < i; ++j) {
            if (!close_enough(E(i, j), E(j, i))) { err = "energies not symmetric in the surfaces"; return PBX_ERR_MODEL; }
            for (int n = 0; n < N; ++n) {
                if (!close_enough(L(n, i, j), L(n, j, i))) { err = "linear couplings not symmetric in the surfaces"; return PBX_ERR_MODEL; }
                for (int m = 0; m < N; ++m)
                    if (!close_enough(Q(n, m, i, j), Q(n, m, j, i))) { err = "quadratic couplings not symmetric in the surfaces"; return PBX_ERR_MODEL; }
            }
        }
    // ---- vibronic model: fold the diagonal linear terms (pimc.py:172-177, 208-218)
    T.d_vib.assign(A * N, 0.0); T.delta_vib.assign(A, 0.0);
    std::vector<long double> tilde(A);
    for (int a = 0; a < A; ++a) {
        long double delta = 0.0L;
        for (int n = 0; n < N; ++n) {
            const long double l = L(n, a, a), w = vib->omega[n];
            delta += l * l / w;
            T.d_vib[a * N + n] = (double)(-l / w);
        }
        T.delta_vib[a] = (double)(-0.5L * delta);
        tilde[a] = (long double)E(a, a) + (long double)T.delta_vib[a];
    }
    // ---- sampling model (pimc.py:336-351, 372-381)
    T.d_rho.assign(Ar * N, 0.0); T.delta_rho.assign(Ar, 0.0); T.weights.assign(Ar, 0.0); T.wcum.assign(Ar, 0.0);
    std::vector<long double> tilde_r(Ar);
    auto Lr = [&](int n, int a) { return rho->linear ? rho->linear[n * Ar + a] : 0.0; };
    long double tmin = 0.0L;
    for (int a = 0; a < Ar; ++a) {
        long double delta = 0.0L;
        for (int n = 0; n < N; ++n) {
            const long double l = Lr(n, a), w = rho->omega[n];
            delta += l * l / w;
            T.d_rho[a * N + n] = (double)(-l / w);
        }
        T.delta_rho[a] = (double)(-0.5L * delta);
        tilde_r[a] = (long double)rho->energy[a] + (long double)T.delta_rho[a];
        if (a == 0 || tilde_r[a] < tmin) tmin = tilde_r[a];
    }
    T.d_rho_eval = T.d_rho;
    if (flags & PBX_QUIRK_RHO_DOUBLE_SHIFT) {
        if (Ar != A) { err = "PBX_QUIRK_RHO_DOUBLE_SHIFT needs a sampling model with as many surfaces as the system"; return PBX_ERR_ARG; }
        for (int i = 0; i < A * N; ++i) T.d_rho_eval[i] += T.d_vib[i];
    }
    {   // weights ~ exp(-beta*tilde) (the common 1/prod sinh factor cancels in the normalisation)
        long double total = 0.0L;
        std::vector<long double> w(Ar);
        for (int a = 0; a < Ar; ++a) { w[a] = std::exp(-(long double)beta * (tilde_r[a] - tmin)); total += w[a]; }
        long double run = 0.0L;
        for (int a = 0; a < Ar; ++a) { T.weights[a] = (double)(w[a] / total); run += w[a] / total; T.wcum[a] = (double)run; }
        T.wcum[Ar - 1] = 2.0;  // any uniform in [0,1) falls below the last edge
    }
    // ---- coth / csch / log prefactors (pimc.py:59-89)
    T.coth.assign(4 * N, 0.0); T.csch.assign(4 * N, 0.0);
    T.logpref.assign(3 * A, 0.0); T.logpref_rho.assign(Ar, 0.0);
    T.tanh_half.assign(4 * N, 0.0); T.coth_half.assign(4 * N, 0.0);
    for (int v = 0; v < 4; ++v) {
        const long double t = (v < 3) ? T.tau[v] : T.tau[0];
        const double* om = (v < 3) ? vib->omega : rho->omega;
        long double half_log_csch = 0.0L;
        for (int n = 0; n < N; ++n) {
            const long double x = t * (long double)om[n];
            T.coth[v * N + n] = (double)(1.0L / std::tanh(x));
            T.csch[v * N + n] = (double)(1.0L / std::sinh(x));
            half_log_csch += -0.5L * std::log(std::sinh(x));
        }
        for (int n = 0; n < N; ++n) {
            const long double hx = 0.5L * t * (long double)om[n];
            T.tanh_half[v * N + n] = (double)std::tanh(hx);
            T.coth_half[v * N + n] = (double)(1.0L / std::tanh(hx));
        }
        if (v < 3) for (int a = 0; a < A; ++a) T.logpref[v * A + a] = (double)(-t * tilde[a] + half_log_csch);
        else for (int a = 0; a < Ar; ++a) T.logpref_rho[a] = (double)(-t * tilde_r[a] + half_log_csch);
    }
    // ---- analytic differences between the tau+/- and tau tables of the vibronic model
    T.dcoth.assign(2 * N, 0.0); T.dcsch.assign(2 * N, 0.0); T.dlogpref.assign(2 * A, 0.0);
    T.dtanh_half.assign(2 * N, 0.0); T.dcoth_half.assign(2 * N, 0.0);
    for (int v = 1; v < 3; ++v) {
        const long double t0 = T.tau[0], tv = T.tau[v];
        long double half_dlog = 0.0L;
        for (int n = 0; n < N; ++n) {
            const long double a = t0 * (long double)vib->omega[n], b = tv * (long double)vib->omega[n];
            const long double sa = std::sinh(a), sb = std::sinh(b);
            // coth(b) - coth(a) = sinh(a - b) / (sinh a sinh b);  csch(b) - csch(a) = (sinh a - sinh b) / (sinh a sinh b)
            const long double dsinh = 2.0L * std::cosh(0.5L * (a + b)) * std::sinh(0.5L * (b - a));  // sinh b - sinh a
            T.dcoth[(v - 1) * N + n] = (double)(std::sinh(a - b) / (sa * sb));
            T.dcsch[(v - 1) * N + n] = (double)(-dsinh / (sa * sb));
            half_dlog += -0.5L * std::log1p(dsinh / sa);   // -1/2 log(sinh b / sinh a)
            // tanh(b/2) - tanh(a/2) = sinh((b-a)/2) / (cosh(a/2) cosh(b/2));  coth(b/2) - coth(a/2) = -sinh((b-a)/2) / (sinh(a/2) sinh(b/2))
            const long double sh = std::sinh(0.5L * (b - a));
            T.dtanh_half[(v - 1) * N + n] = (double)(sh / (std::cosh(0.5L * a) * std::cosh(0.5L * b)));
            T.dcoth_half[(v - 1) * N + n] = (double)(-sh / (std::sinh(0.5L * a) * std::sinh(0.5L * b)));
        }
        for (int a = 0; a < A; ++a) T.dlogpref[(v - 1) * A + a] = (double)(-(tv - t0) * tilde[a] + half_dlog);
    }
    // ---- packed coupling tables: what is left in V after the fold (pimc.py:208-218, 1147-1160);
    //      lower triangle, like np.linalg.eigh(UPLO='L') reads it
    T.e_off.assign(T.AA, 0.0); T.l_off.assign((size_t)N * T.AA, 0.0); T.q_pack.assign((size_t)T.NN * T.AA, 0.0);
    for (int i = 0; i < A; ++i)
        for (int j = 0; j <= i; ++j) {
            const int k = tri(i, j);
            if (i != j) {
                T.e_off[k] = E(i, j);
                for (int n = 0; n < N; ++n) T.l_off[(size_t)n * T.AA + k] = L(n, i, j);
            }
            for (int n = 0; n < N; ++n)
                for (int m = n; m < N; ++m) {
                    const double q = (n == m) ? 0.5 * Q(n, n, i, j) : 0.5 * (Q(n, m, i, j) + Q(m, n, i, j));
                    T.q_pack[(size_t)pair_index(n, m, N) * T.AA + k] = q;
                    if (q != 0.0) T.has_quadratic = true;
                }
        }
    T.rho_shares_vib = (Ar == A);
    for (int n = 0; n < N && T.rho_shares_vib; ++n) {
        if (vib->omega[n] != rho->omega[n]) T.rho_shares_vib = false;
        for (int a = 0; a < A && a < Ar; ++a)
            if (T.d_vib[a * N + n] != T.d_rho_eval[a * N + n]) T.rho_shares_vib = false;
    }
    // ---- sampler recurrence, one (a,b,e) triple per (bead, mode); precision = 2coth - csch*C
    T.samp.assign((size_t)P * N * 3, 0.0);
    std::vector<long double> a, b, e;
    for (int n = 0; n < N; ++n) {
        const long double x = (long double)T.tau[0] * (long double)rho->omega[n];
        ring_recurrence(2.0L / std::tanh(x), 1.0L / std::sinh(x), P, a, b, e);
        for (int j = 0; j < P; ++j) {
            T.samp[((size_t)j * N + n) * 3 + 0] = (double)a[j];
            T.samp[((size_t)j * N + n) * 3 + 1] = (double)b[j];
            T.samp[((size_t)j * N + n) * 3 + 2] = (double)e[j];
        }
    }
    return PBX_OK;
}

}  // namespace pbx

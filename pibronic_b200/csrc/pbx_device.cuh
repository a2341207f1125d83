// Device building blocks shared by all pbx kernels: Philox4x32-10, FP64 Box-Muller, and the two
// ways of forming M = exp(-tau V) for a small symmetric V held in registers
// (reference: np.linalg.eigh + einsum, /root/reference/pibronic/pimc/pimc.py:1171-1187).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pbx_math_tables.h"
#include "pbx_tables.hpp"

namespace pbx {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (sample_lo, sample_hi, draw, stream), key = seed
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}

enum : uint32_t { STREAM_NORMALS = 0u, STREAM_SOURCE = 1u };

// 64 random bits -> double in (0, 1]: (n + 1) 2^-53 with n the top 53 bits.  n + 1 <= 2^53 converts exactly and the
// power of two is an integer subtraction on the exponent field (the FP64 pipe is the bottleneck, the integer pipe is not)
__device__ __forceinline__ double u01_open_low(uint32_t hi, uint32_t lo) {
    const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
    const double d = (double)((bits >> 11) + 1ull);
    return __hiloint2double(__double2hiint(d) - (53 << 20), __double2loint(d));
}
// 64 random bits -> double in [0, 1)
__device__ __forceinline__ double u01_half_open(uint32_t hi, uint32_t lo) {
    const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
    return (double)(bits >> 11) * 0x1.0p-53;
}

// ------------------------------------------------------------------------------------------
// Branch-free FP64 elementary functions for the argument ranges this path needs.  The CUDA math
// library versions carry slow-path branches (denormals, huge arguments, special values); a branch
// ends the basic block, so ptxas cannot interleave independent evaluations -- in the sampler phase
// that left one dependent chain at a time in flight (profiles/r01_summary.md).  Every FP64 instruction
// of the sampler competes with the estimator for the FP64 units, so the functions are table driven
// (pbx_math_tables.h, 10 KB, L1 resident) with short polynomials: a Box-Muller pair costs 28 FP64
// instructions (round 1: 47) and 5 conversions.  Coefficients sit in __constant__ memory: one LDCU.128
// fetches two of them (an FP64 immediate costs two UMOVs).
// Accuracy of each: a few ulp (checked against numpy in tests/test_gpu_parity.py::test_device_math).
// ------------------------------------------------------------------------------------------
static __constant__ double kNeg2Log1pC[4] = {-2.0, 1.0, -2.0 / 3, 1.0 / 2};   // -2 log1p(t) = t (-2 + t - 2/3 t^2 + t^3/2 - 2/5 t^4)
static __constant__ double kExpC[14] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
constexpr double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;

// single MUFU instruction (the rounded intrinsic rsqrtf carries a slow-path branch)
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// -2 ln(u) for a normal, positive u in [2^-63, 1]: u = 2^k m with m in [0.75, 1.5),
// -2 ln u = -2 k ln 2 + 2 ln(rc) - 2 log1p(m rc - 1) with rc ~ 1/m from a 385-entry table (the pair of u's lane is one
// 16-byte L1 load) and -2 k ln 2 from a second one; |m rc - 1| <= 1/768, so the series ends at t^5 (remainder 1e-18).
// The centre of the cell around m = 1 is exactly 1 and both table terms vanish there: the result keeps its RELATIVE
// accuracy for u -> 1 (6e-16 at the edge of that cell).  7 FP64 instructions.
__device__ __forceinline__ double neg2_log_pos(double u) {
    const int hi = __double2hiint(u), lo = __double2loint(u);
    const int frac = hi & 0x000fffff;
    const bool big = frac >= 0x00080000;                       // mantissa >= 1.5: use m / 2 in [0.75, 1)
    const int j = 1023 - (hi >> 20) - (big ? 1 : 0);           // -k
    const int idx = big ? ((frac + 0x800) >> 12) - 128 : ((frac + 0x400) >> 11) + 128;   // round(512 m) - 384
    const double m = __hiloint2double(frac | (big ? 0x3fe00000 : 0x3ff00000), lo);
    const double2 tb = __ldg(reinterpret_cast<const double2*>(kLogTab) + idx);
    const double base = __ldg(kLogExpTab + min(max(j, 0), 63)) + tb.y;
    const double t = fma(m, tb.x, -1.0);
    double q = fma(t, -2.0 / 5, kNeg2Log1pC[3]);
#pragma unroll
    for (int i = 2; i >= 0; --i) q = fma(q, t, kNeg2Log1pC[i]);
    return fma(t, q, base);
}
__device__ __forceinline__ double log_pos(double u) { return -0.5 * neg2_log_pos(u); }

// sqrt(x) for x in [0, 1e6] (x is -2 ln u; sqrt(0) = 0).  MUFU seed r ~ x^-1/2 (22 bits), e = 1 - x r^2 exactly rounded,
// sqrt x = x r (1 - e)^-1/2 = x r (1 + e/2 + 3/8 e^2) with the next term 5/16 e^3 < 1e-19.  6 FP64 instructions.
__device__ __forceinline__ double sqrt_pos(double x) {
    const double r = (double)rsqrt_approx(fmaxf((float)x, 1e-30f));
    const double e = fma(-x, r * r, 1.0);
    const double w = x * r;
    const double pe = fma(e, 0.375, 0.5) * e;
    return fma(w, pe, w);
}

// sin(2 pi u), cos(2 pi u) for u = b 2^-53, b a 53-bit integer: u = i/256 + f 2^-53 with i = round(256 u) and the signed
// remainder f formed in integers (exact), {sin, cos}(2 pi i/256) from a 4 KB table, h = 2 pi f 2^-53 in [-pi/256, pi/256]:
// sin h to h^5 (remainder 8e-18), cos h - 1 to h^6 (1e-20), then the angle-sum formulas.  13 FP64 instructions.
__device__ __forceinline__ void sincos_2pi_bits(unsigned long long b, double& sn, double& cs) {
    const unsigned long long t = b + (1ull << 44);
    const int idx = (int)(t >> 45) & 255;
    const long long f = (long long)(t & ((1ull << 45) - 1)) - (1ll << 44);
    const double2 sc = __ldg(reinterpret_cast<const double2*>(kSinCosTab) + idx);
    const double h = (double)f * (6.283185307179586476925 * 0x1.0p-53);
    const double h2 = h * h;
    const double sh = fma(h * h2, fma(h2, 1.0 / 120, -1.0 / 6), h);
    const double cm1 = h2 * fma(h2, fma(h2, -1.0 / 720, 1.0 / 24), -0.5);
    sn = fma(sc.x, cm1, fma(sc.y, sh, sc.x));
    cs = fma(sc.y, cm1, fma(-sc.x, sh, sc.y));
}
// the same for a double u in [0, 1) that is a multiple of 2^-53 (self-test entry)
__device__ __forceinline__ void sincos_2pi(double u, double& sn, double& cs) {
    sincos_2pi_bits((unsigned long long)(u * 0x1.0p53), sn, cs);
}

// exp(x) for x <= 700 (arguments here are <= 0 up to rounding); exp(x < -708) = 0, exp(-inf) = 0.
// x = (32 k + j) ln2/32 + r, |r| <= ln2/64: exp(x) = 2^k 2^(j/32) p(r) with a 32-entry table and a degree-6
// polynomial (remainder 3.5e-18); the integer n = 32 k + j is read off the low word of x*32/ln2 + 1.5*2^52, and
// 2^k is an integer add on the exponent field (the product stays normal for x >= -708).  10 FP64 instructions.
__device__ __forceinline__ double exp_fast(double x) {
    constexpr double kMagic = 6755399441055744.0;              // 1.5 * 2^52
    const double t = fma(x, 46.166241308446828, kMagic);       // 32 / ln 2
    const int n = __double2loint(t);
    const double dn = t - kMagic;
    double r = fma(dn, -kLn2Hi * 0.03125, x);
    r = fma(dn, -kLn2Lo * 0.03125, r);
    double p = kExpC[6];
#pragma unroll
    for (int i = 5; i >= 0; --i) p = fma(p, r, kExpC[i]);
    const double v = p * __ldg(kExp2Tab + (n & 31));
    const double scaled = __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
    return (x >= -708.0) ? scaled : 0.0;                       // also x = -inf / NaN garbage -> 0
}

// two independent N(0,1) variates from one Philox block (Box-Muller, all FP64):
// sqrt(-2 ln u1) {cos, sin}(2 pi u2), u1 = (top 53 bits of (x, y) + 1) 2^-53 in (0, 1], u2 = (top 53 bits of (z, w)) 2^-53
__device__ __forceinline__ void normal_pair(uint4 r, double& z0, double& z1) {
    const double rad = sqrt_pos(neg2_log_pos(u01_open_low(r.x, r.y)));
    double s, c;
    sincos_2pi_bits((((unsigned long long)r.z << 32) | r.w) >> 11, s, c);
    z0 = rad * c;
    z1 = rad * s;
}

// mixture component: first a with u < wcum[a]
template <int AR, typename WC>
__device__ __forceinline__ int pick_source(double u, const WC& wcum) {
    int src = 0;
#pragma unroll
    for (int a = 0; a < AR - 1; ++a) src += (u >= wcum[a]) ? 1 : 0;
    return src;
}

// ------------------------------------------------------------------------------------------
// small symmetric matrices in registers, packed lower triangle (index tri(i,j), i >= j)
// ------------------------------------------------------------------------------------------
template <int A> struct SymMat { double v[A * (A + 1) / 2]; };

// C = X*Y for COMMUTING symmetric X, Y (so C is symmetric): A*(A+1)/2 * A FMAs
template <int A>
__device__ __forceinline__ void sym_mul(const double (&X)[A * (A + 1) / 2], const double (&Y)[A * (A + 1) / 2],
                                        double (&C)[A * (A + 1) / 2]) {
#pragma unroll
    for (int i = 0; i < A; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double acc = X[sym(i, 0)] * Y[sym(0, j)];
#pragma unroll
            for (int k = 1; k < A; ++k) acc = fma(X[sym(i, k)], Y[sym(k, j)], acc);
            C[tri(i, j)] = acc;
        }
}

// C = X*Y + Z for commuting symmetric X, Y and symmetric Z: the leading product of every entry becomes an FMA
template <int A>
__device__ __forceinline__ void sym_mul_add(const double (&X)[A * (A + 1) / 2], const double (&Y)[A * (A + 1) / 2],
                                            const double (&Z)[A * (A + 1) / 2], double (&C)[A * (A + 1) / 2]) {
#pragma unroll
    for (int i = 0; i < A; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double acc = Z[tri(i, j)];
#pragma unroll
            for (int k = 0; k < A; ++k) acc = fma(X[sym(i, k)], Y[sym(k, j)], acc);
            C[tri(i, j)] = acc;
        }
}

// Number of squarings s for exp(X) = (T12(X / 2^s))^(2^s): the smallest s >= 0 with ||X||_1 / 2^s < theta.
// theta = 1/3 bounds the truncation error of the degree-12 Taylor polynomial by
// theta^13/13! * e^theta = 1.4e-16 (in the 2-norm, which the 1-norm overestimates), i.e. half an ulp of the O(1) entries of M;
// theta = 1/4 (PBX_EXPM_THETA_INV 4) gives 2.4e-18 but costs a squaring whenever 1/4 <= ||X|| < 1/3 -- on the c2
// workload that is 3 % of the beads but 68 % of the WARPS (the loop runs to the largest s of the 32 lanes).
#ifndef PBX_EXPM_THETA_INV
#define PBX_EXPM_THETA_INV 3
#endif
__device__ __forceinline__ int expm_squarings(double norm) {
    // y = norm / theta = m * 2^e with m in [0.5, 1)  ->  s = max(0, e)
    int s = ((__double2hiint(norm * (double)PBX_EXPM_THETA_INV) >> 20) & 0x7ff) - 1022;
    return s < 0 ? 0 : (s > 60 ? 60 : s);
}

// Degree-12 Taylor polynomial of exp in FOUR matrix products (Paterson-Stockmeyer needs five):
//   X2 = X X,  X3 = X X2,  Y0 = X3 (c1 X3 + c2 X2 + c3 X),
//   T12(X) = (Y0 + c4 X3 + c5 X2 + c6 X)(Y0 + c7 X3 + c8 X2) + c9 Y0 + c10 X3 + X2/2 + X + I
// (the evaluation scheme of Sastre, Ibanez & Defez, "Boosting the computation of the matrix exponential", 2019).
// The c_i solve the 12 coefficient-matching equations exactly (mpmath, 60 digits; all positive, the largest is 5,
// so no cancellation is introduced); in float64 the result is as close to exp(X) as the five-product form
// (max abs error 2.2e-16 vs 4.4e-16 over random symmetric X with ||X||_1 <= 1/3, A = 2..12).
namespace t12 {
constexpr double c1 = 0x1.7f48de54a68e8p-15, c2 = 0x1.1f76a6bf7ceaep-12, c3 = 0x1.1f76a6bf7ceaep-9,
                 c4 = 0x1.0a6e06477c9d5p-6, c5 = 0x1.90673e3a44edap-3, c6 = 0x1.4f2fd96e4727ep+0,
                 c7 = 0x1.2287dccb8569cp-6, c8 = 0x1.37d0cd0183c4dp-5, c9 = 0x1.4134deeb04fc8p+2,
                 c10 = 0x1.de886872333aep-4;
}

// M = exp(X) for symmetric X by scaling and squaring: s = expm_squarings(||X||_1) halvings, the four-product
// degree-12 polynomial above on packed symmetric storage (all factors are polynomials in X: they commute and their
// products are symmetric), s squarings.
template <int A>
__device__ __forceinline__ void sym_expm(double (&X)[A * (A + 1) / 2], double (&M)[A * (A + 1) / 2]) {
    constexpr int AA = A * (A + 1) / 2;
    double norm = 0.0;
#pragma unroll
    for (int i = 0; i < A; ++i) {
        double row = 0.0;
#pragma unroll
        for (int j = 0; j < A; ++j) row += fabs(X[sym(i, j)]);
        norm = fmax(norm, row);
    }
    const int s = expm_squarings(norm);
    const double scale = __hiloint2double((1023 - s) << 20, 0);
#pragma unroll
    for (int k = 0; k < AA; ++k) X[k] *= scale;
    if constexpr (A == 2) {
        // two surfaces (every model of the reference's examples/paper_1.5025058): closed form instead of four products.
        // X = m I + Y, Y = [[h, b], [b, -h]], Y^2 = q I with q = h^2 + b^2:  exp X = e^m (cosh(sqrt q) I + sinh(sqrt q)/sqrt q Y),
        // both even series in sqrt q, cut after q^6 (|X| < 1/3 after scaling: the same degree-12 truncation as T12, all terms
        // positive).  35 FP64 instructions instead of 66.
        const double hd = 0.5 * X[2];
        const double m = fma(0.5, X[0], hd), h = fma(0.5, X[0], -hd);
        const double q = fma(h, h, X[1] * X[1]);
        double ch = 1.0 / 479001600, sh = 1.0 / 6227020800.0;
        ch = fma(ch, q, 1.0 / 3628800);  sh = fma(sh, q, 1.0 / 39916800);
        ch = fma(ch, q, 1.0 / 40320);    sh = fma(sh, q, 1.0 / 362880);
        ch = fma(ch, q, 1.0 / 720);      sh = fma(sh, q, 1.0 / 5040);
        ch = fma(ch, q, 1.0 / 24);       sh = fma(sh, q, 1.0 / 120);
        ch = fma(ch, q, 0.5);            sh = fma(sh, q, 1.0 / 6);
        ch = fma(ch, q, 1.0);            sh = fma(sh, q, 1.0);
        const double em = exp_fast(m);
        const double es = em * sh, ec = em * ch;
        M[0] = fma(es, h, ec);
        M[1] = es * X[1];
        M[2] = fma(-es, h, ec);
        for (int r = 0; r < s; ++r) {
            double W2[AA];
            sym_mul<A>(M, M, W2);
#pragma unroll
            for (int k = 0; k < AA; ++k) M[k] = W2[k];
        }
        return;
    }
    double X2[AA], X3[AA], W[AA], Y0[AA], Bm[AA];
    sym_mul<A>(X, X, X2);
    sym_mul<A>(X, X2, X3);
#pragma unroll
    for (int k = 0; k < AA; ++k) W[k] = fma(t12::c1, X3[k], fma(t12::c2, X2[k], t12::c3 * X[k]));
    sym_mul<A>(X3, W, Y0);
    // every sum as a chain of FMAs (a lone DMUL or DADD costs the FP64 pipe a full slot)
    double Z[AA];
#pragma unroll
    for (int k = 0; k < AA; ++k) {
        W[k] = fma(t12::c4, X3[k], fma(t12::c5, X2[k], fma(t12::c6, X[k], Y0[k])));
        Bm[k] = fma(t12::c7, X3[k], fma(t12::c8, X2[k], Y0[k]));
        Z[k] = fma(t12::c9, Y0[k], fma(t12::c10, X3[k], fma(0.5, X2[k], X[k])));
    }
#pragma unroll
    for (int i = 0; i < A; ++i) Z[tri(i, i)] += 1.0;
    sym_mul_add<A>(W, Bm, Z, M);
    for (int q = 0; q < s; ++q) {
        sym_mul<A>(M, M, W);
#pragma unroll
        for (int k = 0; k < AA; ++k) M[k] = W[k];
    }
}

// M = U exp(lambda) U^T with (lambda, U) from a cyclic Jacobi eigensolve of symmetric X, all in
// registers (indices are compile-time constants after unrolling).  Sweeps run until the
// off-diagonal mass is below 1e-33 * ||X||_F^2 for every lane of the warp (max 12 sweeps).
template <int A>
__device__ __forceinline__ void sym_exp_jacobi(const double (&X)[A * (A + 1) / 2], double (&M)[A * (A + 1) / 2]) {
    double S[A][A], U[A][A];
#pragma unroll
    for (int i = 0; i < A; ++i)
#pragma unroll
        for (int j = 0; j < A; ++j) { S[i][j] = X[sym(i, j)]; U[i][j] = (i == j) ? 1.0 : 0.0; }
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < A; ++i)
#pragma unroll
        for (int j = 0; j < A; ++j) fro = fma(S[i][j], S[i][j], fro);
    const double tol = 1e-33 * fro;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int i = 1; i < A; ++i)
#pragma unroll
            for (int j = 0; j < i; ++j) off = fma(S[i][j], S[i][j], off);
        if (__all_sync(__activemask(), off <= tol)) break;
#pragma unroll
        for (int p = 0; p < A - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < A; ++q) {
                const double apq = S[p][q];
                double c = 1.0, sn = 0.0;
                if (fabs(apq) > 1e-300) {
                    const double theta = (S[q][q] - S[p][p]) / (2.0 * apq);
                    const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                    c = rsqrt(fma(t, t, 1.0));
                    sn = t * c;
                    S[p][p] = fma(-t, apq, S[p][p]);
                    S[q][q] = fma(t, apq, S[q][q]);
                    S[p][q] = 0.0; S[q][p] = 0.0;
                }
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    if (k != p && k != q) {
                        const double akp = S[k][p], akq = S[k][q];
                        const double np_ = fma(c, akp, -sn * akq), nq_ = fma(sn, akp, c * akq);
                        S[k][p] = np_; S[p][k] = np_;
                        S[k][q] = nq_; S[q][k] = nq_;
                    }
                    const double ukp = U[k][p], ukq = U[k][q];
                    U[k][p] = fma(c, ukp, -sn * ukq);
                    U[k][q] = fma(sn, ukp, c * ukq);
                }
            }
    }
    double ev[A];
#pragma unroll
    for (int k = 0; k < A; ++k) ev[k] = exp(S[k][k]);
#pragma unroll
    for (int i = 0; i < A; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < A; ++k) acc = fma(U[i][k] * ev[k], U[j][k], acc);
            M[tri(i, j)] = acc;
        }
}

}  // namespace pbx

// Fused large-A kernel (1 <= A <= 16 surfaces, any N, A_rho <= 32; BASELINE config c4: A=12, N=24, P=256).
//
// ONE WARP PER SAMPLE, the whole estimator in one launch: nothing but the 32-byte result leaves the SM (the blocked
// kernels of pbx_mid.cuh wrote, re-read, re-wrote and re-read M through HBM: ~1.4 MB per c4 sample).  The warp walks
// the ring polymer in groups of BIG_G = 16 beads:
//
//   sampler    (MODE_SAMPLE) Philox4x32-10 + FP64 Box-Muller for the group's (bead, mode pair) items spread over the
//              lanes, then the cyclic-tridiagonal ring recurrence with lane = mode -> coordinate tile Rt[n][bead] in
//              shared memory; (MODE_COORDS) the tile is copied from the caller's R[x][n][p].
//   O factors  lane = (bead, half of the surfaces): harmonic exponents of all surfaces in the half-angle form of
//              pbx_fast.cuh (surface independent (R - R')^2 part hoisted), scale S = max, O/S = exp(l - log S) into
//              shared memory, log(O_rho/S) summed per sampling surface (lane = surface).
//   V build    V[bead][k] = sum_f feature_f(bead) coef[f][k], a (16 beads x 325 features) x (325 x 78 packed entries)
//              contraction at c4, on the FP64 tensor cores (mma.sync.m8n8k4.f64): two 8-bead row tiles share every
//              coefficient fragment (the fragment-ordered table is read once per 16 beads, prefetched two k-steps
//              ahead in registers), accumulators = 40 doubles per lane.  X = -tau V goes to shared memory packed,
//              with its Frobenius norm.
//   per bead   M = exp(X): degree-12 Taylor polynomial in four products (pbx_device.cuh, t12) + squarings chosen from
//              ||X||_F, the 12 x 12 products as DMMA tiles on fragments that stay in registers (symmetric factors: the
//              A- and B-operand fragments of a matrix are the same registers); then the chain
//              [T_0; T_1; T_2] <- ([T_0; T_1; T_2] M) diag(O_v), a (3A x A)(A x A) DMMA product, T in shared memory.
//
// Reference: /root/reference/pibronic/pimc/pimc.py:326-334, 613-631 (sampler), 1087-1129 (O), 1076-1084 (S),
// 1132-1136 (rho), 1139-1187 (V, M), 1194-1209 (chain), 1413-1449 (order of operations in block_compute_pm).
#pragma once
#include <atomic>
#include <type_traits>

#include "pbx_dmma.cuh"

#ifndef PBX_BIG_WARPS
#define PBX_BIG_WARPS 8
#endif
#ifndef PBX_BIG_CH
#define PBX_BIG_CH 4           // k-steps per stage of the coefficient ring
#endif
#ifndef PBX_BIG_COPY_SPLIT
#define PBX_BIG_COPY_SPLIT 1   // bulk copies per stage (must divide BIG_CH * NT * 2: every copy a multiple of 128 bytes)
#endif
#ifndef PBX_BIG_STAGES
#define PBX_BIG_STAGES 3
#endif

namespace pbx {

constexpr int BIG_G = 16;            // beads per group: two m8 row tiles of the coupling contraction
constexpr int BIG_RS = BIG_G + 2;    // row stride of the coordinate tile (17 beads: the group and the next one)
constexpr int BIG_WARPS = PBX_BIG_WARPS, BIG_CH = PBX_BIG_CH, BIG_STAGES = PBX_BIG_STAGES, BIG_COPY_SPLIT = PBX_BIG_COPY_SPLIT;
constexpr int BIG_AMAX = 16, BIG_ARMAX = 32, BIG_NMAX_SAMPLER = 32;

enum { BIG_COORDS = 0, BIG_SAMPLE = 1 };

struct BigParams {
    int Ar, N, P, n_rho_eval, KS;   // KS: k-steps (4 features each) of the coupling contraction
    int share;                      // 1: rho has the model's diagonal shifts and frequencies (HostTables::rho_shares_vib)
    double neg_tau;
    // flat table staged into shared memory at kernel start: al[4][N], ga[4][N] (-1/4 tanh(x/2), -1/4 coth(x/2); rows vib
    // tau, tau+, tau-, rho), d2v[A][N], d2r[Ar][N] (2 d), lpref[3][A], lprho[Ar], drho[Ar][N]
    const double* tab;
    int tab_doubles;
    int o_al, o_ga, o_d2v, o_d2r, o_lpref, o_lprho, o_drho;
    const double* wcum;      // [Ar]
    const double* q_dmma;    // [KS][NT][32] coupling coefficients times -tau in mma fragment order (cf. DevTables::q_dmma)
    const int* feat;         // [4 KS]
    const int* tri_ij;       // [8 NT]
    const double* samp;      // [P][N][3]
    const double* R;         // BIG_COORDS: [n][N][P]
    unsigned long long seed;
    long long first_sample, n_samples;
    double* out4;
    long long out_ld;
    double* mirror;          // optional mapped host copy of the results
    long long mirror_ld;
};

// row stride of the A x A matrices in shared memory: == 4 (mod 8) so that the 8-byte operand-fragment loads of a
// half-warp (rows g = 0..3, columns c = 0..3) fall into 16 different banks
__host__ __device__ constexpr int big_ldm(int A) { return A <= 4 ? 4 : (A <= 12 ? 12 : 20); }
__host__ __device__ constexpr int big_even(int n) { return (n + 1) & ~1; }

// per-warp shared memory, in doubles.  Regions time-shared within a group:
//   A: coordinate tile Rt [N+1][RS] (sampler, O factors, V build)   |  three A x LDM work matrices (per-bead stage)
//   B: normals Zt [N][17] (sampler) -> T = R + R' [N][16], lrs [16][Ar], logS [16] (O factors)  |  packed X [16][8 NT]
struct BigLayout {
    int regA, regB, ovib, nrm, sst, total;   // offsets of the regions after A, and the total
    int ov;                                  // row stride of ovib
};
__host__ __device__ inline BigLayout big_layout(int A, int NV, int N, int Ar) {
    const int ldm = big_ldm(A), NT = (A * (A + 1) / 2 + 7) / 8;
    BigLayout L;
    int a = (N + 1) * BIG_RS, a2 = 3 * A * ldm;
    const int szA = big_even(a > a2 ? a : a2);
    int b = N * (BIG_G + 1), b2 = N * BIG_G + BIG_G * Ar + BIG_G, b3 = BIG_G * 8 * NT;
    b = b > b2 ? b : b2;
    const int szB = big_even(b > b3 ? b : b3);
    L.ov = big_even(NV * A);
    L.regA = 0; L.regB = szA; L.ovib = L.regB + szB; L.nrm = L.ovib + BIG_G * L.ov; L.sst = L.nrm + BIG_G;
    L.total = big_even(L.sst + NV * A * ldm);
    return L;
}

// doubles of CTA-shared tables in front of the per-warp regions: the constant table, the feature offsets [4 KS] (int2),
// the coefficient ring and its barriers
__host__ __device__ inline int big_cta_doubles(int tab_doubles, int KS, int NT) {
    return big_even(tab_doubles) + 4 * KS + BIG_STAGES * BIG_CH * NT * 32 + 2 * BIG_STAGES;
}

// ---- mbarrier / TMA bulk-copy primitives (shared::cta)
__device__ __forceinline__ uint32_t big_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void big_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(big_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void big_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(big_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void big_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(big_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void big_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "BIG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra BIG_DONE;\n"
        "bra BIG_WAIT;\n"
        "BIG_DONE:\n"
        "}\n" ::"r"(big_smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`; 16-byte aligned addresses and size
__device__ __forceinline__ void big_bulk_copy(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(big_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(big_smem_u32(bar)) : "memory");
}

template <int AT> struct BigFrag { double v[MidShape<AT>::MT][MidShape<AT>::MT][2]; };   // accumulator layout
template <int AT> struct BigOp { double v[MidShape<AT>::MT][MidShape<AT>::KS]; };        // operand layout (A == B^T for symmetric matrices)

template <int AT>
__device__ __forceinline__ void big_prod(const BigOp<AT>& X, const BigOp<AT>& Y, BigFrag<AT>& C) {
    constexpr int MT = MidShape<AT>::MT, KS = MidShape<AT>::KS;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) { C.v[mt][nt][0] = 0.0; C.v[mt][nt][1] = 0.0; }
    // k-step outermost: consecutive tensor instructions write different accumulator tiles (a dependent DMMA cannot
    // issue until the previous one has left the pipe).  X and Y commute (polynomials of one matrix), the product is
    // symmetric: only the tiles on and below the diagonal are formed (3 of 4 at A = 12), big_frag_store mirrors them.
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt <= mt; ++nt) dmma_884(C.v[mt][nt][0], C.v[mt][nt][1], X.v[mt][ks], Y.v[nt][ks]);
}

// C += X * Y, same tiles
template <int AT>
__device__ __forceinline__ void big_prod_acc(const BigOp<AT>& X, const BigOp<AT>& Y, BigFrag<AT>& C) {
    constexpr int MT = MidShape<AT>::MT, KS = MidShape<AT>::KS;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt <= mt; ++nt) dmma_884(C.v[mt][nt][0], C.v[mt][nt][1], X.v[mt][ks], Y.v[nt][ks]);
}

// operand fragment of a matrix in shared memory (row stride LDM): lane (g, c) holds M[8 t + g][4 ks + c]
template <int AT>
__device__ __forceinline__ void big_op_load(const double* __restrict__ buf, BigOp<AT>& O, int g, int c) {
    constexpr int LDM = big_ldm(AT);
#pragma unroll
    for (int t = 0; t < MidShape<AT>::MT; ++t)
#pragma unroll
        for (int ks = 0; ks < MidShape<AT>::KS; ++ks) {
            const int i = 8 * t + g, k = 4 * ks + c;
            O.v[t][ks] = (i < AT && k < AT) ? buf[i * LDM + k] : 0.0;
        }
}

// accumulator fragment of a SYMMETRIC matrix -> shared memory (full matrix): lane (g, c) holds C[8 mt + g][8 nt + 2c + {0,1}]
// for the tiles nt <= mt; one 16-byte store per tile (a column AT of an odd-AT matrix lands in the row padding; its value
// is zero) plus the transposed elements of the tiles below the diagonal
template <int AT>
__device__ __forceinline__ void big_frag_store(double* __restrict__ buf, const BigFrag<AT>& F, int g, int c) {
    constexpr int LDM = big_ldm(AT);
#pragma unroll
    for (int mt = 0; mt < MidShape<AT>::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt <= mt; ++nt) {
            const int i = 8 * mt + g, j = 8 * nt + 2 * c;
            if (i < AT && j < AT) {
                *reinterpret_cast<double2*>(buf + i * LDM + j) = make_double2(F.v[mt][nt][0], F.v[mt][nt][1]);
                if (nt < mt) {      // the mirror image of an off-diagonal tile
                    buf[j * LDM + i] = F.v[mt][nt][0];
                    if (j + 1 < AT) buf[(j + 1) * LDM + i] = F.v[mt][nt][1];
                }
            }
        }
}

// CTAs per SM the kernel is compiled for: up to 8 surfaces every matrix is a single 8 x 8 tile and 128 registers hold
// the fragments, so two CTAs (16 warps, four per scheduler) share an SM when their shared memory fits as well -- the
// per-bead stage of the small shapes is latency bound (short dependent DMMA chains); above, one CTA with 255 registers
#ifndef PBX_BIG_CTAS_SMALL
#define PBX_BIG_CTAS_SMALL 2
#endif
__host__ __device__ constexpr int big_min_ctas(int AT) { return AT <= 8 ? PBX_BIG_CTAS_SMALL : 1; }

template <int AT, bool PM, int MODE, int CTAS = 1>
__global__ void __launch_bounds__(BIG_WARPS * 32, CTAS)
pbx_big_kernel(const BigParams Q) {
    extern __shared__ __align__(16) double sm[];
    using Sh = MidShape<AT>;
    constexpr int NV = PM ? 3 : 1, MT = Sh::MT, KSA = Sh::KS, NT = Sh::NT;
    constexpr int LDM = big_ldm(AT), ROWS = NV * AT, MTS = (ROWS + 7) / 8, XSTR = 8 * NT;
    constexpr int SH = (AT + 1) / 2;     // surfaces per lane in the O-factor stage
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, c = lane & 3, jb = lane & 15, hf = lane >> 4;
    const int N = Q.N, P = Q.P, Ar = Q.Ar;

    // ---- model tables -> shared memory (once per CTA)
    double* tabs = sm;
    for (int i = threadIdx.x; i < Q.tab_doubles; i += blockDim.x) tabs[i] = Q.tab[i];
    const double* al = tabs + Q.o_al;        // [4][N]
    const double* ga = tabs + Q.o_ga;        // [4][N]
    const double* d2v = tabs + Q.o_d2v;      // [AT][N]
    const double* d2r = tabs + Q.o_d2r;      // [Ar][N]
    const double* lpref = tabs + Q.o_lpref;  // [3][AT]
    const double* lprho = tabs + Q.o_lprho;  // [Ar]
    const double* drho = tabs + Q.o_drho;    // [Ar][N]
    // feature table of the coupling contraction as byte offsets into a row of the coordinate tile
    int2* feat_s = reinterpret_cast<int2*>(sm + big_even(Q.tab_doubles));     // [4 KS]
    for (int i = threadIdx.x; i < 4 * Q.KS; i += blockDim.x) {
        const int f = __ldg(Q.feat + i);
        feat_s[i] = make_int2((f & 0xffff) * BIG_RS * 8, (f >> 16) * BIG_RS * 8);
    }

    // ---- coefficient ring: BIG_STAGES stages of BIG_CH k-steps in fragment order + a full/empty mbarrier pair per stage
    const int stage_doubles = BIG_CH * NT * 32;
    double* ring = sm + big_even(Q.tab_doubles) + 4 * Q.KS;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(ring + (size_t)BIG_STAGES * stage_doubles);
    uint64_t* bar_empty = bar_full + BIG_STAGES;
    // warps per CTA: BIG_WARPS, or fewer when the per-warp regions of a large shape do not fit the SM eight times
    const int cta_warps = (int)(blockDim.x >> 5);
    const long long nwarps = (long long)gridDim.x * cta_warps;
    const long long iters = (Q.n_samples - blockIdx.x + nwarps - 1) / nwarps;
    const int n_chunks = Q.KS / BIG_CH;
    // the CTA consumes the ring in one sequence of chunks, the same in every warp; slot and phase advance incrementally
    // (no divisions on the critical path: thread 0, the producer, must not fall behind the other warps)
    int chunks_left = (int)(iters * ((Q.P + BIG_G - 1) / BIG_G) * n_chunks);      // chunks this CTA has not consumed yet
    const double* q_src = Q.q_dmma;
    auto issue_chunk = [&](int i, int slot) {      // chunk i of the table -> ring slot (one thread)
        const int k0 = i * BIG_CH;                          // KS is a multiple of BIG_CH (zero padded)
        const uint32_t bytes = (uint32_t)(BIG_CH * NT * 32 * sizeof(double));
        big_mbar_expect_tx(&bar_full[slot], bytes);
        // BIG_COPY_SPLIT copies per stage (1 is best: a 10 KB copy lands in ~540 cycles, tools/microbench/bulk_copy_latency.cu)
#pragma unroll
        for (int q = 0; q < BIG_COPY_SPLIT; ++q)
            big_bulk_copy(ring + (size_t)slot * stage_doubles + q * (stage_doubles / BIG_COPY_SPLIT),
                          q_src + (size_t)k0 * NT * 32 + q * (stage_doubles / BIG_COPY_SPLIT), bytes / BIG_COPY_SPLIT, &bar_full[slot]);
    };
    if (threadIdx.x == 0) {
        for (int st = 0; st < BIG_STAGES; ++st) { big_mbar_init(&bar_full[st], 1); big_mbar_init(&bar_empty[st], cta_warps); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int st = 0; st < BIG_STAGES; ++st)
            if (st < chunks_left) issue_chunk(st % n_chunks, st);
    int r_slot = 0;               // ring slot and phase parity of the next chunk to consume
    uint32_t r_phase = 0;
    bool r_first = true;          // nothing consumed yet: no slot to refill
    const BigLayout L = big_layout(AT, NV, N, Ar);
    double* w = sm + big_cta_doubles(Q.tab_doubles, Q.KS, NT) + (size_t)warp * L.total;
    double* Rt = w + L.regA;                 // [N+1][RS]; row N = ones
    double* bufs = w + L.regA;               // 3 x [AT][LDM]   (per-bead stage; Rt is dead then)
    double* Zt = w + L.regB;                 // [N][17] standard normals of the group (sampler)
    double* Tt = w + L.regB;                 // [N][16] R + R'
    double* lrs = Tt + N * BIG_G;            // [16][Ar] log O_rho
    double* Ls = lrs + BIG_G * Ar;           // [16] log S
    double* Xs = w + L.regB;                 // [16][XSTR] packed X = -tau V
    double* ovib = w + L.ovib;               // [16][ov] O_v / S
    double* nrm = w + L.nrm;                 // [16] ||X||_F^2
    double* Sst = w + L.sst;                 // [NV*AT][LDM] stacked chain products

    // ---- per-lane gather indices into a packed symmetric matrix: operand layout and accumulator layout
    int opi[MT][KSA], aci[MT][MT][2];
#pragma unroll
    for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int ks = 0; ks < KSA; ++ks) {
            const int i = 8 * t + g, k = 4 * ks + c;
            opi[t][ks] = (i < AT && k < AT) ? sym(i, k) : -1;
        }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * mt + g, j = 8 * nt + 2 * c + e;
                aci[mt][nt][e] = (i < AT && j < AT) ? sym(i, j) : -1;
            }
    // weights of the lane's packed entries in the Frobenius norm (diagonal 1, off-diagonal 2, padding 0) as bit masks
    unsigned long long fro_valid = 0ull, fro_diag = 0ull;     // two bits per tile: 34 bits at 16 surfaces (17 tiles)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int ij = __ldg(Q.tri_ij + 8 * j + 2 * c + e);
            if (ij >= 0) fro_valid |= 1ull << (2 * j + e);
            if (ij >= 0 && (ij >> 16) == (ij & 0xffff)) fro_diag |= 1ull << (2 * j + e);
        }
    // variant of the stacked-chain rows this lane holds
    int vrow[MTS];
#pragma unroll
    for (int mt = 0; mt < MTS; ++mt) { const int r = 8 * mt + g; vrow[mt] = r < ROWS ? r / AT : 0; }

    const uint2 key = make_uint2((uint32_t)Q.seed, (uint32_t)(Q.seed >> 32));
    const int H = (N + 1) / 2;

    // consecutive samples go to different SMs first, then to the next warp slot
    // every warp of the CTA makes the same number of passes (the ring protocol counts on it); a warp without a sample
    // in the last pass recomputes the last sample and drops the result
    for (long long it = 0; it < iters; ++it) {
        long long x = (long long)warp * gridDim.x + blockIdx.x + it * nwarps;
        const bool live = x < Q.n_samples;
        if (!live) x = Q.n_samples - 1;
        const unsigned long long gidx = (unsigned long long)(Q.first_sample + x);
        // ---- sampler state (lane = mode): previous bead, first bead, shift of the drawn mixture component
        double yprev = 0.0, y0 = 0.0, shift = 0.0, rcarry = 0.0;
        if (MODE == BIG_SAMPLE) {
            const uint4 rs = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, STREAM_SOURCE), key);
            const double u = u01_half_open(rs.x, rs.y);
            int src = 0;
            for (int a = 0; a < Ar - 1; ++a) src += (u >= __ldg(Q.wcum + a)) ? 1 : 0;
            if (lane < N) shift = drho[src * N + lane];
        }
        // chain state: stacked identities
        for (int e = lane; e < ROWS * LDM; e += 32) {
            const int r = e / LDM, j = e - r * LDM;
            Sst[e] = (j == r % AT) ? 1.0 : 0.0;
        }
        double racc = 0.0;      // lane a < Ar: sum over beads of log(O_rho[a] / S)

        for (int p0 = 0; p0 < P; p0 += BIG_G) {
            const int nb = min(BIG_G, P - p0);
            __syncwarp();
            // ================================================================ coordinates of beads p0 .. p0 + 16
            if (MODE == BIG_SAMPLE) {
                // new beads of this group: j in [jlo, jhi]; column of bead j in the tile = j - p0
                const int jlo = (p0 == 0) ? 0 : p0 + 1, jhi = min(p0 + BIG_G, P - 1);
                const int nnew = jhi - jlo + 1;
                for (int it = lane; it < nnew * H; it += 32) {
                    const int dj = it / H, hh = it - dj * H, j = jlo + dj;
                    const uint4 r = philox4x32_10(make_uint4((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)(j * H + hh),
                                                             STREAM_NORMALS), key);
                    double z0, z1;
                    normal_pair(r, z0, z1);
                    Zt[(2 * hh) * (BIG_G + 1) + (j - p0)] = z0;
                    if (2 * hh + 1 < N) Zt[(2 * hh + 1) * (BIG_G + 1) + (j - p0)] = z1;
                }
                __syncwarp();
                if (lane < N) {
                    double* row = Rt + lane * BIG_RS;
                    if (p0 > 0) row[0] = rcarry;
                    for (int j4 = jlo; j4 <= jhi; j4 += 4) {        // recurrence coefficients of four beads at a time: one latency
                        double ca[4], cb[4], ce[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double* tb = Q.samp + ((size_t)min(j4 + q, jhi) * N + lane) * 3;
                            ca[q] = __ldg(tb); cb[q] = __ldg(tb + 1); ce[q] = __ldg(tb + 2);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int j = j4 + q;
                            if (j <= jhi) {
                                double y = ca[q] * Zt[lane * (BIG_G + 1) + (j - p0)];
                                if (j > 0) y = fma(cb[q], yprev, fma(ce[q], y0, y));
                                if (j == 0) y0 = y;
                                yprev = y;
                                row[j - p0] = y + shift;
                            }
                        }
                    }
                    if (p0 + BIG_G >= P) row[P - p0] = y0 + shift;            // the ring closes on bead 0
                    for (int col = min(BIG_G, P - p0) + 1; col <= BIG_G; ++col) row[col] = 0.0;
                    rcarry = row[BIG_G];
                }
            } else {
                const double* Rx = Q.R + (size_t)x * N * P;
                for (int e = lane; e < N * (BIG_G + 1); e += 32) {
                    const int n = e / (BIG_G + 1), col = e - n * (BIG_G + 1);
                    const int p = p0 + col;
                    Rt[n * BIG_RS + col] = (p <= P) ? __ldg(Rx + (size_t)n * P + (p == P ? 0 : p)) : 0.0;
                }
            }
            if (lane < BIG_RS) Rt[N * BIG_RS + lane] = 1.0;
            __syncwarp();

            // ================================================================ O factors: lane = (bead jb, half hf)
            {
                double gsum[NV + 1];
#pragma unroll
                for (int v = 0; v <= NV; ++v) gsum[v] = 0.0;
                for (int n = hf; n < N; n += 2) {          // the two halves split the modes, then add up
                    const double r0 = Rt[n * BIG_RS + jb], r1 = Rt[n * BIG_RS + jb + 1];
                    const double dw = r0 - r1, w2 = dw * dw;
                    Tt[n * BIG_G + jb] = r0 + r1;
#pragma unroll
                    for (int v = 0; v < NV; ++v) gsum[v] = fma(ga[v * N + n], w2, gsum[v]);
                    gsum[NV] = fma(ga[3 * N + n], w2, gsum[NV]);
                }
#pragma unroll
                for (int v = 0; v <= NV; ++v) gsum[v] += __shfl_xor_sync(0xffffffffu, gsum[v], 16);
                __syncwarp();
                // vib surfaces a = hf, hf + 2, ...
                double ev[NV][SH];
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int s = 0; s < SH; ++s) ev[v][s] = 0.0;
                for (int n = 0; n < N; ++n) {
                    const double t = Tt[n * BIG_G + jb];
                    double alv[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) alv[v] = al[v * N + n];
#pragma unroll
                    for (int s = 0; s < SH; ++s) {
                        const int a = min(hf + 2 * s, AT - 1);
                        const double u = t - d2v[a * N + n], u2 = u * u;
#pragma unroll
                        for (int v = 0; v < NV; ++v) ev[v][s] = fma(alv[v], u2, ev[v][s]);
                    }
                }
                double lmax = -INFINITY;
                if (Q.share) {
                    // rho is the diagonal of the model (same shifts and frequencies): its exponent sums are those of the tau set
#pragma unroll
                    for (int s = 0; s < SH; ++s) {
                        const int a = hf + 2 * s;
                        if (a < AT) {
                            const double lr = (a < Q.n_rho_eval) ? lprho[a] + (gsum[0] + ev[0][s]) : -INFINITY;
                            lrs[jb * Ar + a] = lr;
                            lmax = fmax(lmax, lr);
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < SH; ++s) {
                    const int a = min(hf + 2 * s, AT - 1);
#pragma unroll
                    for (int v = 0; v < NV; ++v) ev[v][s] += lpref[v * AT + a] + gsum[v];
                    lmax = fmax(lmax, ev[0][s]);
                }
                // sampling surfaces a = hf, hf + 2, ... < Ar
                for (int a = hf; a < (Q.share ? 0 : Ar); a += 2) {
                    double acc = gsum[NV];
                    const double* d2 = d2r + a * N;
                    for (int n = 0; n < N; ++n) {
                        const double u = Tt[n * BIG_G + jb] - d2[n];
                        acc = fma(al[3 * N + n], u * u, acc);
                    }
                    const double lr = (a < Q.n_rho_eval) ? lprho[a] + acc : -INFINITY;
                    lrs[jb * Ar + a] = lr;
                    lmax = fmax(lmax, lr);
                }
                lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 16));     // log S of bead jb (pimc.py:1076-1084)
                if (hf == 0) Ls[jb] = lmax;
                if (jb < nb) {
#pragma unroll
                    for (int s = 0; s < SH; ++s) {
                        const int a = hf + 2 * s;
                        if (a < AT) {
#pragma unroll
                            for (int v = 0; v < NV; ++v) ovib[jb * L.ov + v * AT + a] = exp_fast(ev[v][s] - lmax);
                        }
                    }
                }
                __syncwarp();
                if (lane < Ar) {
                    double t = 0.0;
                    for (int j = 0; j < nb; ++j) t += lrs[j * Ar + lane] - Ls[j];
                    racc += t;
                }
            }

            // ================================================================ V for the 16 beads on the FP64 tensor cores
            double acc[2][NT][2];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < NT; ++j) { acc[m][j][0] = 0.0; acc[m][j][1] = 0.0; }
            {
                // The coefficient fragments arrive through a shared-memory ring that one thread keeps filled with TMA bulk
                // copies (cp.async.bulk + mbarrier): BIG_STAGES chunks of BIG_CH k-steps, read by all warps of the CTA --
                // the table crosses L2 -> SM once per CTA and group instead of once per warp, and its latency is off the
                // warps' critical path (with per-warp __ldg two k-steps ahead the first DMMA of every k-step waited ~150
                // cycles on the long scoreboard; without the loads the kernel ran 15 % faster).
                // A operands: three-stage software pipeline -- feature offsets of step k+3 and coordinate values of step
                // k+2 are loaded before the tensor instructions of step k are issued, the products for step k+2 formed after.
                // Per k-step a warp issues 20 tensor instructions (16 cycles of the FP64 pipe each) and ~20 others: the B
                // fragments are refilled IN PLACE right after their last use (the register dependence keeps each shared-memory
                // load between two tensor instructions, where the issue slot is free anyway), the A operands run two steps
                // ahead.  A chunk (BIG_CH steps) is fully unrolled, without bounds checks: KS is padded to whole chunks.
                const char* Rb = reinterpret_cast<const char*>(Rt + g);
                const int last = 4 * Q.KS - 1;
                auto rd = [&](int off) { return *reinterpret_cast<const double*>(Rb + off); };
                int2 f = feat_s[c];
                double a0 = rd(f.x) * rd(f.y), a1 = rd(f.x + 64) * rd(f.y + 64);      // step 0
                f = feat_s[4 + c];
                double b0 = rd(f.x) * rd(f.y), b1 = rd(f.x + 64) * rd(f.y + 64);      // step 1
                int2 fn = feat_s[min(8 + c, last)];                                    // offsets of step 2
                big_mbar_wait(&bar_full[r_slot], r_phase);
                const double* bs = ring + (size_t)r_slot * stage_doubles + lane;
                double bq[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) bq[j] = bs[j * 32];
                for (int i = 0; i < n_chunks; ++i) {
                    const int slot = r_slot;
                    const int n_slot = (r_slot + 1 == BIG_STAGES) ? 0 : r_slot + 1;       // where the next chunk lives
                    const uint32_t n_phase = r_phase ^ (n_slot == 0 ? 1u : 0u);
                    if (threadIdx.x == 0 && !r_first && chunks_left >= BIG_STAGES) {
                        // producer: the slot of the previous chunk gets the chunk BIG_STAGES - 1 ahead of this one
                        const int ps = (r_slot == 0) ? BIG_STAGES - 1 : r_slot - 1;
                        const uint32_t pp = r_phase ^ (r_slot == 0 ? 1u : 0u);
                        int ti = i + BIG_STAGES - 1;
                        while (ti >= n_chunks) ti -= n_chunks;
                        big_mbar_wait(&bar_empty[ps], pp);
                        issue_chunk(ti, ps);
                    }
                    r_first = false;
                    __syncwarp();
                    const double* bnext = bs;          // fragments of the first step of the next chunk
#pragma unroll
                    for (int u = 0; u < BIG_CH; ++u) {
                        const int k = i * BIG_CH + u;
                        // coordinate values of step k+2, feature offsets of step k+3
                        const double ra0 = rd(fn.x), rb0 = rd(fn.y), ra1 = rd(fn.x + 64), rb1 = rd(fn.y + 64);
                        fn = feat_s[min(4 * (k + 3) + c, last)];
                        const double* src = bs + (u + 1) * NT * 32;
                        if (u == BIG_CH - 1) {          // the refills of the last step read the next chunk: wait for it here
                            const bool more = i + 1 < n_chunks;
                            if (more) big_mbar_wait(&bar_full[n_slot], n_phase);
                            bnext = ring + (size_t)(more ? n_slot : slot) * stage_doubles + lane;
                            src = bnext;
                        }
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            dmma_884(acc[0][j][0], acc[0][j][1], a0, bq[j]);
                            dmma_884(acc[1][j][0], acc[1][j][1], a1, bq[j]);
                            bq[j] = src[j * 32];
                        }
                        a0 = b0; a1 = b1;
                        b0 = ra0 * rb0; b1 = ra1 * rb1;
                    }
                    __syncwarp();
                    if (lane == 0) big_mbar_arrive(&bar_empty[slot]);
                    bs = bnext;
                    r_slot = n_slot; r_phase = n_phase; --chunks_left;
                }
            }
            __syncwarp();      // Tt / lrs (region B) are dead: the packed X tile takes their place
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                double fro = 0.0;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double x0 = acc[m][j][0], x1 = acc[m][j][1];          // the table carries the factor -tau
                    *reinterpret_cast<double2*>(Xs + (8 * m + g) * XSTR + 8 * j + 2 * c) = make_double2(x0, x1);
                    const double w0 = ((fro_valid >> (2 * j)) & 1ull) ? (((fro_diag >> (2 * j)) & 1ull) ? 1.0 : 2.0) : 0.0;
                    const double w1 = ((fro_valid >> (2 * j + 1)) & 1ull) ? (((fro_diag >> (2 * j + 1)) & 1ull) ? 1.0 : 2.0) : 0.0;
                    fro = fma(w0 * x0, x0, fma(w1 * x1, x1, fro));
                }
                fro += __shfl_xor_sync(0xffffffffu, fro, 1);
                fro += __shfl_xor_sync(0xffffffffu, fro, 2);
                if (c == 0) nrm[8 * m + g] = fro;
            }
            __syncwarp();

            // ================================================================ per bead: M = exp(X), chain step
            // Software pipeline over the beads of the group: while the chain consumes M(jc) -- a (3A x A)(A x A) product
            // of 10 independent accumulator tiles per k-step -- the exponential of the NEXT bead is formed; its four
            // dependent products and their shared-memory round trips hide behind the chain's tensor instructions (done one
            // after the other the two left the FP64 pipe idle half of the time).  step<chain, expm>(jc, je): the straight-
            // line segments between two warp barriers hold a piece of each.
            double* B0 = bufs;
            double* B1 = bufs + AT * LDM;
            double* B2 = bufs + 2 * AT * LDM;
            BigOp<AT> oM;                                       // M(jc) in operand layout
#define PBX_FRAG_LOOP                                  \
    _Pragma("unroll") for (int mt = 0; mt < MT; ++mt)  \
    _Pragma("unroll") for (int nt = 0; nt <= mt; ++nt) \
    _Pragma("unroll") for (int e = 0; e < 2; ++e)
            auto step = [&](auto chain_tag, auto expm_tag, const int jc, const int je) {
                constexpr bool CHN = decltype(chain_tag)::value, EXP = decltype(expm_tag)::value;
                BigOp<AT> oX, oA, oB;
                BigFrag<AT> x1, x2, x3, y0f, fb, fo;
                double cs[MTS][MT][2];
                int s = 0;
                if constexpr (CHN) {
#pragma unroll
                    for (int mt = 0; mt < MTS; ++mt)
#pragma unroll
                        for (int nt = 0; nt < MT; ++nt) { cs[mt][nt][0] = 0.0; cs[mt][nt][1] = 0.0; }
                }
                // k-steps [lo, hi) of [T_0; T_1; T_2] M      (pimc.py:1198-1206)
                auto chain_k = [&](const int lo, const int hi) {
#pragma unroll
                    for (int ks = 0; ks < KSA; ++ks)
                        if (ks >= lo && ks < hi) {
                            double as[MTS];
#pragma unroll
                            for (int mt = 0; mt < MTS; ++mt) {
                                const int r = 8 * mt + g, k = 4 * ks + c;
                                as[mt] = (r < ROWS && k < AT) ? Sst[r * LDM + k] : 0.0;
                            }
#pragma unroll
                            for (int mt = 0; mt < MTS; ++mt)
#pragma unroll
                                for (int nt = 0; nt < MT; ++nt) dmma_884(cs[mt][nt][0], cs[mt][nt][1], as[mt], oM.v[nt][ks]);
                        }
                };
                // ---- segment 0: X^2
                if constexpr (EXP) {
                    const double* Xp = Xs + je * XSTR;
                    // squarings: smallest s >= 0 with ||X||_F / 2^s < theta  <=>  theta^-2 ||X||_F^2 < 4^s
                    const double y = nrm[je] * (double)(PBX_EXPM_THETA_INV * PBX_EXPM_THETA_INV);
                    if (y >= 1.0) s = ((((__double2hiint(y) >> 20) & 0x7ff) - 1023) >> 1) + 1;
                    s = s > 60 ? 60 : s;
#pragma unroll
                    for (int t = 0; t < MT; ++t)
#pragma unroll
                        for (int ks = 0; ks < KSA; ++ks) oX.v[t][ks] = opi[t][ks] >= 0 ? Xp[opi[t][ks]] : 0.0;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int nt = 0; nt <= mt; ++nt)
#pragma unroll
                            for (int e = 0; e < 2; ++e) x1.v[mt][nt][e] = aci[mt][nt][e] >= 0 ? Xp[aci[mt][nt][e]] : 0.0;
                    if (s > 0) {          // warp-uniform; no squarings (hence no scaling) on most workloads
                        const double scale = __hiloint2double((1023 - s) << 20, 0);
#pragma unroll
                        for (int t = 0; t < MT; ++t)
#pragma unroll
                            for (int ks = 0; ks < KSA; ++ks) oX.v[t][ks] *= scale;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                            for (int nt = 0; nt <= mt; ++nt)
#pragma unroll
                                for (int e = 0; e < 2; ++e) x1.v[mt][nt][e] *= scale;
                    }
                }
                if constexpr (CHN) chain_k(0, 1);
                if constexpr (EXP) {
                    big_prod<AT>(oX, oX, x2);                       // X^2
                    big_frag_store<AT>(B0, x2, g, c);
                }
                __syncwarp();
                // ---- segment 1: X^3 and the first factor of Y0
                if constexpr (CHN) chain_k(1, 2);
                if constexpr (EXP) {
                    big_op_load<AT>(B0, oA, g, c);
                    big_prod<AT>(oX, oA, x3);                       // X^3
                    PBX_FRAG_LOOP fb.v[mt][nt][e] = fma(t12::c1, x3.v[mt][nt][e], fma(t12::c2, x2.v[mt][nt][e], t12::c3 * x1.v[mt][nt][e]));
                    big_frag_store<AT>(B1, x3, g, c);
                    big_frag_store<AT>(B2, fb, g, c);
                }
                __syncwarp();
                // ---- segment 2: Y0 = X^3 (c1 X^3 + c2 X^2 + c3 X)
                if constexpr (CHN) chain_k(2, KSA);
                if constexpr (EXP) {
                    big_op_load<AT>(B1, oA, g, c);
                    big_op_load<AT>(B2, oB, g, c);
                    big_prod<AT>(oA, oB, y0f);
                    PBX_FRAG_LOOP {      // sums as FMA chains: a lone DMUL or DADD costs the FP64 units a full slot
                        fb.v[mt][nt][e] = fma(t12::c4, x3.v[mt][nt][e], fma(t12::c5, x2.v[mt][nt][e], fma(t12::c6, x1.v[mt][nt][e], y0f.v[mt][nt][e])));
                        fo.v[mt][nt][e] = fma(t12::c7, x3.v[mt][nt][e], fma(t12::c8, x2.v[mt][nt][e], y0f.v[mt][nt][e]));
                    }
                }
                __syncwarp();                                   // every lane has read B1, B2 and its rows of T
                // ---- segment 3: T <- (T M) diag(O_v); the two factors of the last product
                if constexpr (CHN) {
                    const double* Op = ovib + jc * L.ov;
#pragma unroll
                    for (int mt = 0; mt < MTS; ++mt)
#pragma unroll
                        for (int nt = 0; nt < MT; ++nt) {
                            const int r = 8 * mt + g, j = 8 * nt + 2 * c;
                            if (r < ROWS && j < AT) {
                                const double o0 = Op[vrow[mt] * AT + j], o1 = (j + 1 < AT) ? Op[vrow[mt] * AT + j + 1] : 0.0;
                                *reinterpret_cast<double2*>(Sst + r * LDM + j) = make_double2(cs[mt][nt][0] * o0, cs[mt][nt][1] * o1);
                            }
                        }
                }
                if constexpr (EXP) {
                    big_frag_store<AT>(B0, fb, g, c);
                    big_frag_store<AT>(B1, fo, g, c);
                }
                __syncwarp();
                // ---- segment 4: T12 = (Y0 + c4 X^3 + c5 X^2 + c6 X)(Y0 + c7 X^3 + c8 X^2) + c9 Y0 + c10 X^3 + X^2/2 + X + I
                if constexpr (EXP) {
                    big_op_load<AT>(B0, oA, g, c);
                    big_op_load<AT>(B1, oB, g, c);
                    // the additive terms are the start value of the product's accumulators
                    PBX_FRAG_LOOP fo.v[mt][nt][e] =
                        fma(t12::c9, y0f.v[mt][nt][e], fma(t12::c10, x3.v[mt][nt][e], fma(0.5, x2.v[mt][nt][e], x1.v[mt][nt][e]))) +
                        (((8 * mt + g) == (8 * nt + 2 * c + e)) ? 1.0 : 0.0);
                    big_prod_acc<AT>(oA, oB, fo);
                    double* cur = B2;
                    double* nxt = B0;
                    big_frag_store<AT>(cur, fo, g, c);
                    __syncwarp();
                    big_op_load<AT>(cur, oM, g, c);                 // M (or its 2^s-th root) in operand layout
                    for (int q = 0; q < s; ++q) {
                        big_prod<AT>(oM, oM, fo);
                        big_frag_store<AT>(nxt, fo, g, c);
                        __syncwarp();
                        big_op_load<AT>(nxt, oM, g, c);
                        double* t = cur; cur = nxt; nxt = t;          // the old buffer was last read before the barrier above
                    }
                    __syncwarp();                               // B0..B2 are rewritten by the next step
                }
            };
            step(std::false_type{}, std::true_type{}, 0, 0);                          // M(0)
            for (int jj = 0; jj + 1 < nb; ++jj) step(std::true_type{}, std::true_type{}, jj, jj + 1);
            step(std::true_type{}, std::false_type{}, nb - 1, 0);                     // last bead of the group
#undef PBX_FRAG_LOOP
        }

        // ---- rho(x) = sum_a exp(sum_p log(O_rho[a]/S)); g_v(x) = tr T_v
        double rho = (lane < Ar) ? exp_fast(racc) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rho += __shfl_xor_sync(0xffffffffu, rho, o);
        double tr[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double t = (lane < AT) ? Sst[(v * AT + lane) * LDM + lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            tr[v] = t;
        }
        if (lane == 0 && live) {
            Q.out4[x] = rho;
#pragma unroll
            for (int v = 0; v < NV; ++v) Q.out4[(size_t)(1 + v) * Q.out_ld + x] = tr[v];
            if (Q.mirror) {
                Q.mirror[x] = rho;
#pragma unroll
                for (int v = 0; v < NV; ++v) Q.mirror[(size_t)(1 + v) * Q.mirror_ld + x] = tr[v];
            }
        }
        __syncwarp();
    }
}

// shared memory of one CTA, bytes
inline size_t big_smem_bytes(int A, bool pm, int N, int Ar, int tab_doubles, int KS, int warps = BIG_WARPS) {
    return ((size_t)big_cta_doubles(tab_doubles, KS, (A * (A + 1) / 2 + 7) / 8) + (size_t)warps * big_layout(A, pm ? 3 : 1, N, Ar).total) * sizeof(double);
}

template <int AT>
cudaError_t launch_big_at(const BigParams& Q, bool pm, int mode, size_t smem, int sms, int warps, cudaStream_t st) {
#define PBX_BIG_GO(PM_, MODE_)                                                                                       \
    {                                                                                                                \
        /* the 128-register build when two of its CTAs fit an SM (shared memory decides: N, A_rho), else the */      \
        /* one-CTA build; persistent grid: as many CTAs as are resident at a time */                                 \
        if constexpr (big_min_ctas(AT) > 1) {                                                                        \
            auto k2 = pbx_big_kernel<AT, PM_, MODE_, big_min_ctas(AT)>;                                              \
            cudaError_t e2 = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
            if (e2 != cudaSuccess) return e2;                                                                        \
            static std::atomic<unsigned long long> memo{0};      /* (smem << 8 | CTAs per SM) of the last query */   \
            const unsigned long long seen = memo.load(std::memory_order_relaxed);                                    \
            int per_sm = (int)(seen & 0xff);                                                                         \
            if ((seen >> 8) != ((unsigned long long)smem << 4 | (unsigned)warps) || per_sm == 0) {                                            \
                e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2, warps * 32, smem);               \
                if (e2 != cudaSuccess) return e2;                                                                    \
                per_sm = std::max(1, std::min(per_sm, 255));                                                         \
                memo.store((((unsigned long long)smem << 4 | (unsigned)warps) << 8) | (unsigned)per_sm, std::memory_order_relaxed);           \
            }                                                                                                        \
            if (per_sm >= 2) {                                                                                       \
                const long long ctas = std::min<long long>(Q.n_samples > 0 ? Q.n_samples : 1, (long long)sms * per_sm); \
                k2<<<(unsigned)ctas, warps * 32, smem, st>>>(Q);                                                 \
                return cudaGetLastError();                                                                           \
            }                                                                                                        \
        }                                                                                                            \
        auto k = pbx_big_kernel<AT, PM_, MODE_, 1>;                                                                  \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
        if (e != cudaSuccess) return e;                                                                              \
        const long long ctas = std::min<long long>(Q.n_samples > 0 ? Q.n_samples : 1, sms);                          \
        k<<<(unsigned)ctas, warps * 32, smem, st>>>(Q);                                                          \
        return cudaGetLastError();                                                                                   \
    }
    if (pm) { if (mode == BIG_SAMPLE) PBX_BIG_GO(true, BIG_SAMPLE) else PBX_BIG_GO(true, BIG_COORDS) }
    if (mode == BIG_SAMPLE) PBX_BIG_GO(false, BIG_SAMPLE) else PBX_BIG_GO(false, BIG_COORDS)
#undef PBX_BIG_GO
}

// defined in pbx_big_inst.cu, one translation unit per A
typedef cudaError_t (*BigLauncher)(const BigParams&, bool, int, size_t, int, int, cudaStream_t);
BigLauncher find_big_kernel(int A);

}  // namespace pbx

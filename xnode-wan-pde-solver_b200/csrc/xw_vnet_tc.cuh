// xw_vnet_tc.cuh -- test-function net on the 5th-generation tensor cores (generation 3 of the v-net kernels).
//
// The hidden layers of v_phi (reference src/model.py:37-47: the SAME Hv x Hv matrix nv times) are
// [rows x Hv] x [Hv x Hv] contractions.  Here they run as tcgen05.mma kind::tf32 with 3xTF32 error
// compensation (x = x_hi + x_lo, both exactly representable in TF32 up to the hardware's truncation of
// x_lo;  a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, accumulated in fp32): measured 1e-6 against fp64 on
// B200 (tools/umma_probe.py), i.e. the fp32 tolerance of the path is kept.
//
//   A operand (activations): TENSOR MEMORY, lane = row, one 32-bit column per hidden unit, written by
//                            the thread that owns the row (tcgen05.st) -- no shared-memory round trip,
//                            no bank conflicts, no proxy fence;
//   B operand (weights)    : shared memory, canonical K-major no-swizzle image [k/4][n][4], staged once
//                            per CTA as a hi and a lo image; the bias rides in input column 50, which the
//                            value rows set to 1;
//   D accumulator          : tensor memory, read back with tcgen05.ld by the owner of the row.
//
// One elected thread issues the 21 MMAs of a layer (3 terms x 7 k-steps of 8) and commits them to an
// mbarrier; the 128 row owners wait on it, apply relu (and the relu mask to the tangent row), split into
// hi/lo and store the next layer's A operand.  Every wait is bounded (trap instead of hang).
#pragma once
#ifndef XW_EMU
#include "xw_umma.cuh"

namespace xw {
namespace tc {

constexpr int HV = 50;          // compiled capacity of the hidden width
constexpr int KP = 56;          // padded contraction length (units 0..49, bias column 50, zeros)
constexpr int NP = 56;          // MMA N (output units, padded)
constexpr int BIASC = 50;

XW_HD constexpr int kin_of(int d) { return (d + 2 + 7) / 8 * 8; }      // (t, x, 1) padded to k-steps of 8

// B-operand image (K-major, no swizzle): element (k, n) of a [K x NP] operand
__device__ __forceinline__ int b_off(int k, int n) { return (k >> 2) * (NP * 4) + (n >> 3) * 32 + (n & 7) * 4 + (k & 3); }
__device__ __forceinline__ void put_split(float* hi, float* lo, int off, float v) {
    const float h = umma::tf32_hi(v);
    hi[off] = h;
    lo[off] = v - h;
}
// Waits of the production kernels are bounded in TIME (about 10 s of SM clock, xw_mbar_wait in xw_platform.cuh), not in
// probes: a wrong descriptor or a protocol bug surfaces as a CUDA error instead of a hung GPU, while a phase that is merely
// slow (contention, MPS, a debugger, throttled clocks) can no longer expire the wait spuriously.
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* mbar, uint32_t& parity) {
    xw_mbar_wait(mbar, parity);
    parity ^= 1u;
}

struct WImages {
    float *wh_hi, *wh_lo;       // [KP x NP]  B[n = o][k = i] = Wh[o][i], k = 50: bh[o]
    float *wht_hi, *wht_lo;     // [KP x NP]  B[n = i][k = o] = Wh[o][i]            (backward only, else NULL)
    float *wi_hi, *wi_lo;       // [kin x NP] B[n = o][k = c] = Wi[o][c], k = C: bi[o]
    float* wz;                  // [64]: Wz[o] (zero padded), wz[KP] = bz
};

__device__ void stage_images(const WImages& w, const float* XW_RESTRICT th, int d, int Hvr, int kin) {
    const VLayout g(d, Hvr);
    const int C = d + 1;
    for (int i = threadIdx.x; i < KP * NP; i += blockDim.x) {
        w.wh_hi[i] = 0.f; w.wh_lo[i] = 0.f;
        if (w.wht_hi) { w.wht_hi[i] = 0.f; w.wht_lo[i] = 0.f; }
    }
    for (int i = threadIdx.x; i < kin * NP; i += blockDim.x) { w.wi_hi[i] = 0.f; w.wi_lo[i] = 0.f; }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) w.wz[i] = i < Hvr ? th[g.Wz + i] : (i == KP ? th[g.bz] : 0.f);
    __syncthreads();
    for (int e = threadIdx.x; e < Hvr * Hvr; e += blockDim.x) {
        const int o = e / Hvr, i = e - o * Hvr;
        const float v = th[g.Wh + e];
        put_split(w.wh_hi, w.wh_lo, b_off(i, o), v);
        if (w.wht_hi) put_split(w.wht_hi, w.wht_lo, b_off(o, i), v);
    }
    for (int e = threadIdx.x; e < Hvr * C; e += blockDim.x) {
        const int o = e / C, c = e - o * C;
        put_split(w.wi_hi, w.wi_lo, b_off(c, o), th[g.Wi + e]);
    }
    for (int o = threadIdx.x; o < Hvr; o += blockDim.x) {
        put_split(w.wh_hi, w.wh_lo, b_off(BIASC, o), th[g.bh + o]);
        put_split(w.wi_hi, w.wi_lo, b_off(C, o), th[g.bi + o]);
    }
    umma::fence_smem_to_async();
    __syncthreads();
}


// relu masks as SIGN bits, one funnel shift per element: after the loop bit (31 - o) of s0 is the sign of v[o]
// (o < 32), bit (31 - (o - 32)) of s1 the sign of v[o] (o >= 32).  A set bit means "relu output is 0" (-0.0 counts
// as negative, +0.0 as positive: both give 0 either way).
__device__ __forceinline__ void sign_masks(const float (&v)[KP], uint32_t& s0, uint32_t& s1) {
    s0 = 0u; s1 = 0u;
#pragma unroll
    for (int o = 0; o < 32; ++o) s0 = __funnelshift_l(__float_as_uint(v[o]), s0, 1);
#pragma unroll
    for (int o = 32; o < HV; ++o) s1 = __funnelshift_l(__float_as_uint(v[o]), s1, 1);
    s1 <<= (64 - HV);                       // left-align: unit o >= 32 sits at bit 31 - (o - 32)
}
__device__ __forceinline__ void apply_sign_masks(float (&v)[KP], uint32_t s0, uint32_t s1) {
#pragma unroll
    for (int o = 0; o < 32; ++o) v[o] = ((s0 >> (31 - o)) & 1u) ? 0.f : v[o];
#pragma unroll
    for (int o = 32; o < HV; ++o) v[o] = ((s1 >> (31 - (o - 32))) & 1u) ? 0.f : v[o];
}

// D[128 x NP] = A[128 x 8*ksteps] * B^T, A from tensor memory (hi / lo column blocks), B images in smem.
// KS > 0: compile-time k-step count (straight-line issue: one IADD on the descriptor per MMA)
template <int KS>
__device__ __forceinline__ void issue_3xtf32(uint32_t tD, uint32_t tA_hi, uint32_t tA_lo, const float* b_hi, const float* b_lo,
                                             int ksteps_rt, uint32_t idesc) {
    const uint64_t dh = umma::smem_desc(b_hi, NP * 16, 128), dl = umma::smem_desc(b_lo, NP * 16, 128);
    constexpr uint64_t kStep = (2 * NP * 16) >> 4;        // two 16-byte chunks of k per MMA
    if (KS > 0) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {                     // small terms first
#pragma unroll
            for (int ks = 0; ks < (KS > 0 ? KS : 1); ++ks)
                umma::mma_tf32_ts(tD, (t == 0 ? tA_lo : tA_hi) + 8 * ks, (t == 1 ? dl : dh) + kStep * ks, idesc, (t | ks) ? 1u : 0u);
        }
    } else {
        uint32_t acc = 0;
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {
            const uint32_t ta = t == 0 ? tA_lo : tA_hi;
            const uint64_t db = t == 1 ? dl : dh;
#pragma unroll 1
            for (int ks = 0; ks < ksteps_rt; ++ks) {
                umma::mma_tf32_ts(tD, ta + 8 * ks, db + kStep * ks, idesc, acc);
                acc = 1;
            }
        }
    }
}

// The same product issued by THREE warps, one 3xTF32 term each.  One thread cannot issue tcgen05.mma faster than one per ~46
// cycles whatever the shape (measured, profiles/r02y_mma_issue_bench.txt: 45.6 cycles per MMA at N = 24, 56; 28.4 at N = 56
// with two or more issuing warps = the pipe's own 128 N / 256), so a layer's 21 MMAs take 960 cycles from one thread against
// 590 of tensor-pipe time.  Warp iw of the three calls this (lane 0 issues): iw = 0: A_lo x B_hi, 1: A_hi x B_lo,
// 2: A_hi x B_hi.  MMAs of DIFFERENT threads are not ordered among each other -- a first "overwrite" MMA of one warp
// behind tcgen05.fence::before_thread_sync + a named barrier was overtaken by the accumulating ones of the other warps in
// about one tile of a few thousand (caught by the N = 4096 golden) -- so every MMA ACCUMULATES and the row owners clear
// the accumulator with tcgen05.st next to their operand stores (ordered before all three issuers by the usual
// wait::st / fence / barrier).  Every issuer commits to `mbar`, which must have been initialised with a count of 3.
template <int KS>
__device__ __forceinline__ void issue_3xtf32_split(int iw, uint32_t tD, uint32_t tA_hi, uint32_t tA_lo, const float* b_hi,
                                                   const float* b_lo, int ksteps_rt, uint32_t idesc, uint64_t* mbar) {
    constexpr uint64_t kStep = (2 * NP * 16) >> 4;        // two 16-byte chunks of k per MMA
    const uint64_t db = umma::smem_desc(iw == 1 ? b_lo : b_hi, NP * 16, 128);
    const uint32_t ta = iw == 0 ? tA_lo : tA_hi;
    umma::fence_after();
    if (KS > 0) {
#pragma unroll
        for (int ks = 0; ks < (KS > 0 ? KS : 1); ++ks) umma::mma_tf32_ts(tD, ta + 8 * ks, db + kStep * ks, idesc, 1u);
    } else {
#pragma unroll 1
        for (int ks = 0; ks < ksteps_rt; ++ks) umma::mma_tf32_ts(tD, ta + 8 * ks, db + kStep * ks, idesc, 1u);
    }
    umma::commit(mbar);
}
__device__ __forceinline__ void tmem_zero56(uint32_t addr) {
    uint32_t z[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) z[i] = 0u;
    umma::tmem_st56(addr, z);
}

// split 56 fp32 values into the hi / lo A operand of this thread's row
__device__ __forceinline__ void store_a_row_at(uint32_t addr_hi, uint32_t addr_lo, const float (&v)[KP]) {
    uint32_t r[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) r[i] = __float_as_uint(v[i]) & 0xFFFFE000u;
    umma::tmem_st56(addr_hi, r);
#pragma unroll
    for (int i = 0; i < KP; ++i) r[i] = __float_as_uint(v[i] - __uint_as_float(r[i]));
    umma::tmem_st56(addr_lo, r);
}

// =============================================================================================
// interior forward over all points: v, dv/dt (forward-mode tangent), weak-form integrands, seeds.
// A tile is 64 points = 128 rows: lanes 0..15 of warp w own the VALUE rows of points 16w..16w+15,
// lanes 16..31 the TANGENT rows of the same points (relu masks travel by warp shuffle).
// =============================================================================================
template <bool SPLIT>          // SPLIT: a layer's MMAs issued by three warps (issue_3xtf32_split) -- faster, but the order in
                               // which the three terms reach the accumulator, hence the last bits of v, varies from run to run
__global__ void __launch_bounds__(384) k_vnet_tc_fwd(VtileFwdArgs a, int ngroups) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = a.d + 1, kin = kin_of(a.d), KA = kin > KP ? kin : KP;
    WImages w;
    w.wh_hi = reinterpret_cast<float*>(smem_raw);
    w.wh_lo = w.wh_hi + KP * NP;
    w.wht_hi = nullptr; w.wht_lo = nullptr;
    w.wi_hi = w.wh_lo + KP * NP;
    w.wi_lo = w.wi_hi + kin * NP;
    w.wz = w.wi_lo + kin * NP;
    double* red = reinterpret_cast<double*>(w.wz + 64);
    uint64_t* mbars = reinterpret_cast<uint64_t*>(red + 4 * 32);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = tid >> 7, wq = warp & 3;
    const bool is_tan = lane >= 16;
    stage_images(w, a.theta, a.d, a.Hvr, kin);
    if (tid < 4) umma::mbar_init(mbars + tid, SPLIT ? 3 : 1);
    if (warp == 0) umma::tmem_alloc(slot, 512);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    // every warpgroup runs its own stream of tiles in its own tensor-memory columns (D: 56, A hi/lo: 2 KA):
    // up to three independent layer chains keep the tensor pipe fed while the others are in their epilogues
    const uint32_t tbase = *slot + (uint32_t)(wg * (NP + 2 * KA));
    const uint32_t colD = 0, colA = NP;
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * wq) << 16);
    uint64_t* mbar = mbars + wg;
    const bool issuer = (tid & 127) == 0;
    const uint32_t idesc = umma::idesc_tf32(128, NP);
    uint32_t parity = 0;
    const long long npts = (long long)a.n * a.L;
    const long long ntiles = (npts + 63) / 64;
    const int L = a.L;
    double accs[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long tix = (long long)blockIdx.x * ngroups + wg; tix < ntiles; tix += (long long)gridDim.x * ngroups) {
        const long long p = tix * 64 + 16 * wq + (lane & 15);
        const bool valid = p < npts;
        const long long n = valid ? p / L : 0;
        const int l = (int)(p - n * L);
        const float* xr = a.p.x + n * a.p.x_sn + (long long)l * a.p.x_sl;
        const float tval = valid ? a.p.t[n * a.p.t_sn + (long long)l * a.p.t_sl] : 0.f;
        // ---- input rows: value (t, x, 1), tangent d/dt = e_0 ----------------------------------
#pragma unroll 1
        for (int c8 = 0; c8 < kin; c8 += 8) {
            float hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int idx = c8 + e;
                float v = 0.f;
                if (is_tan) v = idx == 0 ? 1.f : 0.f;
                else if (valid) v = idx == 0 ? tval : (idx <= a.d ? xr[idx - 1] : (idx == C ? 1.f : 0.f));
                hi[e] = umma::tf32_hi(v);
                lo[e] = v - hi[e];
            }
            umma::tmem_st8(lane_addr + colA + c8, hi);
            umma::tmem_st8(lane_addr + colA + KA + c8, lo);
        }
        if (SPLIT) tmem_zero56(lane_addr + colD);
        umma::tmem_wait_st();
        umma::fence_before();
        umma::group_sync(1 + wg);
        if (SPLIT) {
            if (wq < 3 && lane == 0) issue_3xtf32_split<0>(wq, tbase + colD, tbase + colA, tbase + colA + KA, w.wi_hi, w.wi_lo, kin / 8, idesc, mbar);
        } else if (issuer) {
            umma::fence_after();
            issue_3xtf32<0>(tbase + colD, tbase + colA, tbase + colA + KA, w.wi_hi, w.wi_lo, kin / 8, idesc);
            umma::commit(mbar);
        }
        float h[KP];
        mbar_wait_or_trap(mbar, parity);
        umma::fence_after();
        umma::tmem_ld56(lane_addr + colD, h);
        // ---- hidden layers -------------------------------------------------------------------
#pragma unroll 1
        for (int layer = 0; layer < a.nv; ++layer) {
            uint32_t m0, m1;
            sign_masks(h, m0, m1);
            m0 = __shfl_sync(0xffffffffu, m0, lane & 15);      // the value row's mask, for both rows of the point
            m1 = __shfl_sync(0xffffffffu, m1, lane & 15);
            apply_sign_masks(h, m0, m1);
            h[BIASC] = is_tan ? 0.f : 1.f;
#pragma unroll
            for (int o = BIASC + 1; o < KP; ++o) h[o] = 0.f;
            store_a_row_at(lane_addr + colA, lane_addr + colA + KA, h);
            if (SPLIT) tmem_zero56(lane_addr + colD);
            umma::tmem_wait_st();
            umma::fence_before();
            umma::group_sync(1 + wg);
            if (SPLIT) {
                if (wq < 3 && lane == 0) issue_3xtf32_split<KP / 8>(wq, tbase + colD, tbase + colA, tbase + colA + KA, w.wh_hi, w.wh_lo, 0, idesc, mbar);
            } else if (issuer) {
                umma::fence_after();
                issue_3xtf32<KP / 8>(tbase + colD, tbase + colA, tbase + colA + KA, w.wh_hi, w.wh_lo, 0, idesc);
                umma::commit(mbar);
            }
            mbar_wait_or_trap(mbar, parity);
            umma::fence_after();
            umma::tmem_ld56(lane_addr + colD, h);
        }
        // ---- tanh + output layer ---------------------------------------------------------------
        float pv = 0.f, pt = 0.f;
#pragma unroll
        for (int o = 0; o < HV; ++o) {
            const float wzo = w.wz[o];
            const float y = tanh_fast(h[o]);
            const float s = wzo * (1.f - y * y);
            const float sv = __shfl_sync(0xffffffffu, s, lane & 15);
            pv = fmaf(wzo, y, pv);
            pt = fmaf(sv, h[o], pt);
        }
        const float dv_t = __shfl_sync(0xffffffffu, pt, (lane & 15) + 16);
        if (!is_tan && valid) {
            const float v = pv + w.wz[KP];
            DomW W;
            if (a.wbuf) { W.w = a.wbuf[p]; W.dw_t = a.dwtbuf[p]; }     // (virtual input: the coordinates are not spatial)
            else W = domain_w(a.dom_kind, a.dp0, a.dp1, a.dp2, tval, xr, a.d);
            float cu, cv;
            float Au, Ap;
            weak_A(a.c0, a.c1, a.Aval, a.Ader, p, a.u[p], Au, Ap);
            weak_point_terms(v, dv_t, W.w, W.dw_t, a.u[p], a.f[p], l == 0 ? a.h[n] : 0.f, l, L, Au, Ap, accs, cu, cv);
            a.cot_u[p] = cu;
            a.cot_v[p] = cv;
            if (a.vcache) { f4 cch; cch.x = v; cch.y = dv_t; cch.z = W.w; cch.w = W.dw_t; st4(a.vcache + 4 * p, cch); }
        }
    }
    const int idx[4] = {0, 1, 2, 3};
    block_sum_to_global<4>(accs, red, a.sums, idx);
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free(*slot, 512);
}

// =============================================================================================
// v net backward on the tensor cores:  parameter gradients of sum_p G[p] v[p],  G = k0*cot_v + k1*v + k2*w
//
// A tile is 128 points, thread = point = TMEM lane; a three-stage warp-specialised pipeline over tiles (512 threads,
// one CTA per SM):
//   F (warps 0-7)  : forward recompute of tile t+1 (input layer + nv hidden layers, value rows only; two warps per lane
//                    quadrant, 28 columns each), post-relu activations r_0..r_{nv-1} and relu masks -> a double-buffered
//                    per-CTA scratch (global memory, L2 resident), output layer, cotangent G, dWz | dbz, delta_nv -> mailbox;
//   R (warps 8-11) : the delta chain of tile t:  R-op  delta_{k-1} = relu'(r_{k-1}) . (delta_k Wh)   [128 x 56] x [56 x 56]
//                    (A from tensor memory), and the transposed hi/lo image of delta_k in shared memory;
//   P (warps 12-15): transposed hi/lo image of (r_{k-1} | 1) from the scratch and the
//                    P-op  dWh | dbh += delta_k^T (r_{k-1} | 1)   [56 x 128] x [128 x 56]  (both operands from shared memory,
//                    K = the 128 points, two 64-point halves with their own images / barriers / issuers);
//                    input layer: dWi | dbi += delta_0^T (t, x, 1) the same way; the accumulators live in tensor memory
//                    and are flushed into an fp32 image in shared memory every few tiles (xw_capi.cu, tc_flush_tiles).
// Every cross-role hand-off is an mbarrier with a time-bounded wait.
// =============================================================================================
// Tensor-memory columns of k_vnet_tc_bwd3 (512 in all): accumulators D_f (56), D_r (56), dWh (112: the r_hi | r_lo blocks),
// dWi (kin) and the TS-form A operands A_f, A_r (hi | lo, 112 each) = 448 + kin.  Up to kin = 48 every accumulator starts on a
// multiple of 16 columns; the widest input (kin = 56: d = 47..54 and the virtual net of 4.1c) only fits fully packed, on
// multiples of 8.
struct P3Cols { int DF, AF, DR, AR, WH, WI; };
__device__ __forceinline__ P3Cols p3_cols(int kin, int packed) {
    P3Cols c;
    if (kin > 48 || packed) { c.WH = 0; c.WI = 112; c.DF = 168; c.DR = 224; c.AF = 280; c.AR = 392; }
    else { c.WH = 0; c.DF = 112; c.DR = 176; c.AF = 240; c.AR = 352; c.WI = 464; }
    return c;
}

// Stacked transposed images of the weight-gradient MMAs (P-op).  dWh[o][i] = sum_p delta[p][o] r[p][i] has only 56 rows and 56
// columns, both operands are per-point data (shared memory, SS form) and the 3xTF32 scheme needs the hi/lo cross terms.
// Instead of three M = 64 MMAs per k-step (each re-reading one image of A and one of B: 3 x 3840 B of operand traffic, and
// M = 64 occupies the tensor pipe as long as M = 128 does) the hi and lo rows are STACKED: A = [delta_hi ; delta_lo]
// (M = 128: rows 0..55 | 56..111), B = [r_hi ; r_lo] (N = 112).  ONE MMA per k-step of 8 points then yields all four
// hi/lo products in four accumulator blocks (7680 B of operand traffic, 56 instead of 84 pipe cycles, 16 instead of 48
// instructions per layer and tile); the blocks are summed when a tile's accumulators are flushed.
// Image element (row, point r): 16-byte chunks of 4 consecutive points, 8-row core matrices (128 B), 14 row groups per
// chunk, chunk stride padded to 452 floats (= 4 mod 32: the 32 lanes of a warp hit 32 banks with their scalar stores).
constexpr int TCS2 = 14 * 32 + 4;            // chunk stride of a stacked image (floats)
constexpr int THALF2 = 16 * TCS2;            // one half image: 64 points = 16 chunks
constexpr int TIMG2 = 2 * THALF2;
__device__ __forceinline__ int ts_off(int row, int r) { return (r >> 6) * THALF2 + ((r & 63) >> 2) * TCS2 + (row >> 3) * 32 + (row & 7) * 4 + (r & 3); }

// one HALF of a P-op (64 points = 8 k-steps).  The two halves have their own image halves, barriers and issuing threads:
// while one half's MMAs run, the other half's images are being rewritten.
__device__ __forceinline__ void issue_pop_half_stacked(uint32_t tD, const float* a_img, const float* b_img, uint32_t idesc) {
    constexpr uint64_t kStep = (2 * TCS2 * 4) >> 4;
    const uint64_t ad = umma::smem_desc(a_img, TCS2 * 4, 128), bd = umma::smem_desc(b_img, TCS2 * 4, 128);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) umma::mma_tf32(tD, ad + kStep * ks, bd + kStep * ks, idesc, 1u);
}
// input layer: B = (t, x, 1) with hi rows [0, kin) and lo rows [56, 56 + kin); N = kin per MMA, both into the same columns
__device__ __forceinline__ void issue_pop_half_in(uint32_t tD, const float* a_img, const float* b_img, uint32_t idesc) {
    constexpr uint64_t kStep = (2 * TCS2 * 4) >> 4;
    const uint64_t ad = umma::smem_desc(a_img, TCS2 * 4, 128);
    const uint64_t bh = umma::smem_desc(b_img, TCS2 * 4, 128), bl = umma::smem_desc(b_img + 7 * 32, TCS2 * 4, 128);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        umma::mma_tf32(tD, ad + kStep * ks, bh + kStep * ks, idesc, 1u);
        umma::mma_tf32(tD, ad + kStep * ks, bl + kStep * ks, idesc, 1u);
    }
}
// clears this thread's lane of the dWh / dWi accumulator columns
__device__ __forceinline__ void zero_acc(uint32_t lane_wh, uint32_t lane_wi, int kin) {
    uint32_t z[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) z[i] = 0u;
    umma::tmem_st56(lane_wh, z);
    umma::tmem_st56(lane_wh + KP, z);
    float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c8 = 0; c8 < kin; c8 += 8) umma::tmem_st8(lane_wi + c8, z8);
}

// relu masks of half a row (28 units) as sign bits: bit (31 - i) = sign of v[i]
__device__ __forceinline__ uint32_t sign_mask28(const float (&v)[28]) {
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < 28; ++i) m = __funnelshift_l(__float_as_uint(v[i]), m, 1);
    return m << 4;
}
// w0: units 0..27, w1: units 28..55 (bit 31 - (o mod 28)); a set bit zeroes the unit
__device__ __forceinline__ void apply_sign_masks_2x28(float (&v)[KP], uint32_t w0, uint32_t w1) {
#pragma unroll
    for (int o = 0; o < 28; ++o) v[o] = ((w0 >> (31 - o)) & 1u) ? 0.f : v[o];
#pragma unroll
    for (int o = 28; o < HV; ++o) v[o] = ((w1 >> (31 - (o - 28))) & 1u) ? 0.f : v[o];
}
// split a 56-wide row into the hi / lo A operand, 8 columns at a time (few live registers)
__device__ __forceinline__ void store_a_row_chunked(uint32_t addr_hi, uint32_t addr_lo, const float (&v)[KP]) {
#pragma unroll
    for (int c8 = 0; c8 < KP; c8 += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { hi[e] = umma::tf32_hi(v[c8 + e]); lo[e] = v[c8 + e] - hi[e]; }
        umma::tmem_st8(addr_hi + c8, hi);
        umma::tmem_st8(addr_lo + c8, lo);
    }
}


// L2 residency hints for the activation scratch of k_vnet_tc_bwd3: the scratch (2 x nv x 28 KB per CTA, ~76 MB per launch)
// is written by the F role and read back by the R / P roles one or two tiles later, then overwritten in place -- it never
// needs to reach HBM.  With default policies the streaming inputs evicted it (28.6 GB of DRAM write-back per launch in
// the r01h capture); scratch accesses carry an evict_last policy and the once-read inputs are loaded evict_first.
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_scratch(f4* ptr, const f4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_scratch1(float* ptr, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ f4 ld_scratch(const f4* ptr, uint64_t pol) {
    f4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol)
                 : "memory");
    return v;
}

// XW_TC_PROF (tools/tc_prof.py builds a separate library with it): SM cycles each role of k_vnet_tc_bwd3 spends in each of
// its waits, summed over the CTAs -- which hand-off a role is actually blocked on.  Not compiled into the product library.
#ifdef XW_TC_PROF
__device__ unsigned long long g_tc_prof[8][12];
#define XW_PF_DECL long long pf_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long pf_t0_ = clock64();
#define XW_PF(slot, stmt) { const long long t_ = clock64(); stmt; pf_[slot] += clock64() - t_; }
#define XW_PF_END(role) { pf_[7] = clock64() - pf_t0_; for (int s_ = 0; s_ < 12; ++s_) atomicAdd(&g_tc_prof[role][s_], (unsigned long long)pf_[s_]); }
#else
#define XW_PF_DECL
#define XW_PF(slot, stmt) { stmt; }
#define XW_PF_END(role) {}
#endif

__device__ __forceinline__ void tmem_zero28(uint32_t addr) {
    uint32_t z[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) z[i] = 0u;
    umma::tmem_st28(addr, z);
}

#ifndef XW_TC_SPLIT_MASK
#define XW_TC_SPLIT_MASK 3      // development switch: bit 0 = F-op, bit 1 = R-op issued by three warps
#endif
template <bool kBwdSplit>      // the F-op / R-op of a layer issued by three warps each (issue_3xtf32_split): 25.2 -> 24.5 ms
__global__ void __launch_bounds__(512, 1) k_vnet_tc_bwd3(VtileBwdArgs a) {
    constexpr bool kSplitF = kBwdSplit && (XW_TC_SPLIT_MASK & 1), kSplitR = kBwdSplit && (XW_TC_SPLIT_MASK & 2);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = a.d + 1, kin = kin_of(a.d), GS = KP + kin + 1;     // (odd row stride: the per-row flushes hit 32 banks)
    const VLayout g(a.d, a.Hvr);
    const P3Cols tc = p3_cols(kin, a.tm_packed);
    WImages w;
    w.wh_hi = reinterpret_cast<float*>(smem_raw);
    w.wh_lo = w.wh_hi + KP * NP;
    w.wht_hi = w.wh_lo + KP * NP;
    w.wht_lo = w.wht_hi + KP * NP;
    w.wi_hi = w.wht_lo + KP * NP;
    w.wi_lo = w.wi_hi + kin * NP;
    w.wz = w.wi_lo + kin * NP;
    float* dT = w.wz + 64;                                       // stacked image of delta_k   (A operand of the P-op)
    float* rT = dT + TIMG2;                                      // stacked image of (r_{k-1} | 1) / (t, x, 1)   (B operand)
    float* gimg = rT + TIMG2;                                    // [56][GS] (+512: overrun pad of the M = 128 reads)
    float* zacc = gimg + KP * GS + 512;                          // [64] dWz | dbz
    float* vex = zacc + 64;                                      // [2][128] partial output dot products of the two F column halves
    uint64_t* mb = reinterpret_cast<uint64_t*>(vex + 256);
    uint64_t *mF = mb, *mR = mb + 1, *mFD = mb + 3, *mFC = mb + 4, *mPC = mb + 6;
    uint64_t* mPh = mb + 8;          // [2] P-op of half h complete
    uint64_t* mDPh = mb + 10;        // [2] delta image of half h written (64 arrivals)
    uint32_t* slot = reinterpret_cast<uint32_t*>(mb + 12);
    // roles: F = warps 0-7 (two warps per lane quadrant, 28 columns each), R = warps 8-11, P = warps 12-15
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = tid < 256 ? 0 : (tid >> 7) - 1, j = tid & 127;
    for (int i = tid; i < 2 * TIMG2 + KP * GS + 512 + 64; i += blockDim.x) dT[i] = 0.f;
    stage_images(w, a.theta, a.d, a.Hvr, kin);
    if (tid == 0) {
        umma::mbar_init(mF, kSplitF ? 3 : 1); umma::mbar_init(mR, kSplitR ? 3 : 1); umma::mbar_init(mPh, 1); umma::mbar_init(mPh + 1, 1);
        umma::mbar_init(mFD, 256); umma::mbar_init(mFC, 128); umma::mbar_init(mDPh, 64); umma::mbar_init(mDPh + 1, 64); umma::mbar_init(mPC, 128);
    }
    if (warp == 0) umma::tmem_alloc(slot, 512);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tbase = *slot;
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t idesc = umma::idesc_tf32(128, NP);
    const int hh_ = j >> 6;                                      // which 64-point half this thread's row belongs to
    // weight-gradient MMAs on the stacked images: M = 128 (delta_hi rows | delta_lo rows), N = 112 (r_hi | r_lo); accumulator
    // row m lands in tensor-memory lane m
    const uint32_t idesc_p = umma::idesc_tf32(128, 2 * NP), idesc_pin = umma::idesc_tf32(128, kin);
    const long long npts = (long long)a.n * a.L;
    const long long ntiles = (npts + 127) / 128;
    const int L = a.L, nv = a.nv, nvs = nv > 0 ? nv : 1;
    f4* scr = reinterpret_cast<f4*>(a.scratch) + (size_t)blockIdx.x * 2 * nvs * 14 * 128 + j;
    const uint64_t pol = l2_evict_last_policy();

    if (wg == 0) {
        // ======================================================================== F: forward of every tile
        // 256 threads: thread (ch, j) owns columns [28 ch, 28 ch + 28) of row j
        const int ch = (tid >> 7) & 1, cb = 28 * ch;
        const float k0 = (float)a.coefs[0], k1 = (float)a.coefs[1], k2 = (float)a.coefs[2];
        uint32_t pF = 0, pFC = 0, pPC = 0;
        int it = 0;
        XW_PF_DECL
        float gwz[28], gbz = 0.f;                                 // dWz | dbz of this thread's row and units, over all its tiles
#pragma unroll
        for (int i = 0; i < 28; ++i) gwz[i] = 0.f;
        for (long long tix = blockIdx.x; tix < ntiles; tix += gridDim.x, ++it) {
            const long long p = tix * 128 + j;
            const bool valid = p < npts;
            const long long n = valid ? p / L : 0;
            const int l = (int)(p - n * L);
            const float* xr = a.p.x + n * a.p.x_sn + (long long)l * a.p.x_sl;
            const float tval = valid ? a.p.t[n * a.p.t_sn + (long long)l * a.p.t_sl] : 0.f;
            f4* sb = scr + (size_t)(it & 1) * nvs * 14 * 128;
#pragma unroll 1
            for (int c8 = 8 * ch; c8 < kin; c8 += 16) {           // the two halves alternate over the 8-column chunks
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int idx = c8 + e;
                    float v = 0.f;
                    if (valid) v = idx == 0 ? tval : (idx <= a.d ? __ldcs(xr + idx - 1) : (idx == C ? 1.f : 0.f));
                    hi[e] = umma::tf32_hi(v);
                    lo[e] = v - hi[e];
                }
                umma::tmem_st8(lane_addr + tc.AF + c8, hi);
                umma::tmem_st8(lane_addr + tc.AF + KP + c8, lo);
            }
            umma::tmem_wait_st();
            umma::fence_before();
            if (it > 0) {                                         // R has taken the previous tile out of the mailbox,
                XW_PF(0, mbar_wait_or_trap(mFC, pFC))             // and P has seen its activations land
                XW_PF(0, mbar_wait_or_trap(mPC, pPC))
            }
            if (kSplitF) { umma::fence_after(); tmem_zero28(lane_addr + tc.DF + cb); umma::tmem_wait_st(); umma::fence_before(); }
            XW_PF(1, asm volatile("bar.sync 1, 256;" ::: "memory"))
            if (kSplitF) {
                if (warp < 3 && lane == 0) issue_3xtf32_split<0>(warp, tbase + tc.DF, tbase + tc.AF, tbase + tc.AF + KP, w.wi_hi, w.wi_lo, kin / 8, idesc, mF);
            } else if (tid == 0) {
                umma::fence_after();
                issue_3xtf32<0>(tbase + tc.DF, tbase + tc.AF, tbase + tc.AF + KP, w.wi_hi, w.wi_lo, kin / 8, idesc);
                umma::commit(mF);
            }
            float h[28];
            XW_PF(2, mbar_wait_or_trap(mF, pF))
            umma::fence_after();
            umma::tmem_ld28(lane_addr + tc.DF + cb, h);
#pragma unroll 1
            for (int layer = 0; layer < nv; ++layer) {
                const uint32_t m = sign_mask28(h);
#pragma unroll
                for (int i = 0; i < 28; ++i) h[i] = fmaxf(h[i], 0.f);     // (accumulator columns 50..55 are exactly 0)
                if (ch == 0) {
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        f4 v; v.x = h[4 * c]; v.y = h[4 * c + 1]; v.z = h[4 * c + 2]; v.w = h[4 * c + 3];
                        st_scratch(sb + (size_t)(layer * 14 + c) * 128, v, pol);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 6; ++c) {                         // units 28..51
                        f4 v; v.x = h[4 * c]; v.y = h[4 * c + 1]; v.z = h[4 * c + 2]; v.w = h[4 * c + 3];
                        st_scratch(sb + (size_t)(layer * 14 + 7 + c) * 128, v, pol);
                    }
                }
                st_scratch1(reinterpret_cast<float*>(sb + (size_t)(layer * 14 + 13) * 128) + ch, __uint_as_float(m), pol);
                uint32_t rh[28], rl[28];
#pragma unroll
                for (int i = 0; i < 28; ++i) {
                    float v = h[i];
                    if (ch == 1 && i == BIASC - 28) v = 1.f;              // the bias column
                    if (ch == 1 && i > BIASC - 28) v = 0.f;
                    rh[i] = __float_as_uint(v) & 0xFFFFE000u;
                    rl[i] = __float_as_uint(v - __uint_as_float(rh[i]));
                }
                XW_PF(9, umma::tmem_st28(lane_addr + tc.AF + cb, rh);
                umma::tmem_st28(lane_addr + tc.AF + KP + cb, rl);
                if (kSplitF) tmem_zero28(lane_addr + tc.DF + cb);
                umma::tmem_wait_st();
                umma::fence_before())
                XW_PF(3, asm volatile("bar.sync 1, 256;" ::: "memory"))
                if (kSplitF) {
                    if (warp < 3 && lane == 0) XW_PF(6, issue_3xtf32_split<KP / 8>(warp, tbase + tc.DF, tbase + tc.AF, tbase + tc.AF + KP, w.wh_hi, w.wh_lo, 0, idesc, mF))
                } else if (tid == 0) {
                    umma::fence_after();
                    XW_PF(6, issue_3xtf32<KP / 8>(tbase + tc.DF, tbase + tc.AF, tbase + tc.AF + KP, w.wh_hi, w.wh_lo, 0, idesc); umma::commit(mF))
                }
                XW_PF(4, mbar_wait_or_trap(mF, pF))
                umma::fence_after();
                XW_PF(8, umma::tmem_ld28(lane_addr + tc.DF + cb, h))
            }
            // output layer, cotangent G, dWz | dbz, delta_nv -> mailbox
            {
                float vp = 0.f;
#pragma unroll
                for (int i = 0; i < 28; ++i) {
                    if (cb + i < HV) {
                        h[i] = tanh_fast(h[i]);
                        vp = fmaf(w.wz[cb + i], h[i], vp);
                    }
                }
                vex[ch * 128 + j] = vp;
                XW_PF(5, asm volatile("bar.sync 1, 256;" ::: "memory"))
                const float v = vex[j] + vex[128 + j] + w.wz[KP];
                float G = 0.f;
                if (valid) {
                    const float wdom = a.wbuf ? __ldcs(a.wbuf + p) : domain_w(a.dom_kind, a.dp0, a.dp1, a.dp2, tval, xr, a.d).w;
                    G = fmaf(k0, __ldcs(a.cot + p), fmaf(k1, v, k2 * wdom));
                }
                gbz += G;
                uint32_t r[28];
#pragma unroll
                for (int i = 0; i < 28; ++i) {
                    r[i] = 0u;
                    if (cb + i < HV) {
                        const float t = h[i];
                        gwz[i] = fmaf(G, t, gwz[i]);
                        r[i] = __float_as_uint(G * w.wz[cb + i] * (1.f - t * t));
                    }
                }
                umma::tmem_st28(lane_addr + tc.DF + cb, r);
                umma::tmem_wait_st();
                umma::fence_before();
                umma::mbar_arrive(mFD);
            }
        }
        if (tid == 0) XW_PF_END(0)
        if (tid == 32) XW_PF_END(4)
        // per-thread dWz | dbz -> the CTA's accumulators (once per kernel)
#pragma unroll
        for (int i = 0; i < 28; ++i) {
            if (cb + i < HV) {
                const float sz = warp_sum(gwz[i]);
                if (lane == 0) atomicAdd(zacc + cb + i, sz);
            }
        }
        if (ch == 0) {
            const float sb2 = warp_sum(gbz);
            if (lane == 0) atomicAdd(zacc + KP, sb2);
        }
    } else if (wg == 1) {
        // ======================================================================== R: the delta chain
        // R also writes the transposed hi/lo images of delta_k (it holds the split values anyway); they may
        // only be overwritten once the P-op that read the previous delta has completed (mP)
        uint32_t pR = 0, pFD = 0, pPr = 0;
        bool published = false;
        int it = 0;
        XW_PF_DECL
        for (long long tix = blockIdx.x; tix < ntiles; tix += gridDim.x, ++it) {
            const f4* sb = scr + (size_t)(it & 1) * nvs * 14 * 128;
            float h[KP];
            XW_PF(0, mbar_wait_or_trap(mFD, pFD))
            umma::fence_after();
            umma::tmem_ld56(lane_addr + tc.DF, h);
            umma::fence_before();
            umma::mbar_arrive(mFC);
#pragma unroll 1
            for (int k = nv; k >= 0; --k) {
                if (k > 0) {                                     // R-op first: it does not depend on the images
                    XW_PF(8, store_a_row_chunked(lane_addr + tc.AR, lane_addr + tc.AR + KP, h);
                    if (kSplitR) tmem_zero56(lane_addr + tc.DR);
                    umma::tmem_wait_st();
                    umma::fence_before())
                    XW_PF(1, umma::group_sync(2))
                    if (kSplitR) {
                        if ((warp & 3) < 3 && lane == 0) XW_PF(6, issue_3xtf32_split<KP / 8>(warp & 3, tbase + tc.DR, tbase + tc.AR, tbase + tc.AR + KP, w.wht_hi, w.wht_lo, 0, idesc, mR))
                    } else if (j == 0) {
                        umma::fence_after();
                        XW_PF(6, issue_3xtf32<KP / 8>(tbase + tc.DR, tbase + tc.AR, tbase + tc.AR + KP, w.wht_hi, w.wht_lo, 0, idesc); umma::commit(mR))
                    }
                }
                if (published) XW_PF(2, mbar_wait_or_trap(mPh + hh_, pPr))   // this half's P-op on the previous delta image is done
#ifdef XW_TC_PROF
                const long long tim_ = clock64();
#endif
#pragma unroll
                for (int o = 0; o < HV; ++o) {
                    const float hi = umma::tf32_hi(h[o]);
                    dT[ts_off(o, j)] = hi;
                    dT[ts_off(KP + o, j)] = h[o] - hi;
                }
                umma::fence_smem_to_async();
#ifdef XW_TC_PROF
                pf_[9] += clock64() - tim_;
#endif
                umma::mbar_arrive(mDPh + hh_);
                published = true;
                if (k == 0) {
                    if (a.delta0_out) {                          // cotangent of the first pre-activation, per point
                        const long long pp = tix * 128 + j;
                        if (pp < npts) {
                            f4* dst = reinterpret_cast<f4*>(a.delta0_out + pp * 52);
#pragma unroll
                            for (int c = 0; c < 13; ++c)      // streaming stores: written once, read by another kernel -- keep L2 for the scratch
                                __stcs(reinterpret_cast<float4*>(dst + c), make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]));
                        }
                    }
                    break;
                }
                const f4 mv = ld_scratch(sb + (size_t)((k - 1) * 14 + 13) * 128, pol);
                const uint32_t m0 = __float_as_uint(mv.x), m1 = __float_as_uint(mv.y);
                XW_PF(3, mbar_wait_or_trap(mR, pR))
                umma::fence_after();
                XW_PF(10, umma::tmem_ld56(lane_addr + tc.DR, h);
                umma::fence_before())
                XW_PF(11, apply_sign_masks_2x28(h, m0, m1))
#pragma unroll
                for (int o = HV; o < KP; ++o) h[o] = 0.f;
            }
        }
        if (j == 0) XW_PF_END(1)
        if (j == 32) XW_PF_END(5)
    } else {
        // ======================================================================== P: weight gradients
        uint32_t pP = 0, pDP = 0, pFDp = 0;
        bool pending = false;
        int it = 0;
        XW_PF_DECL
        const int flush_tiles = a.flush_tiles > 0 ? a.flush_tiles : 1;
        zero_acc(lane_addr + tc.WH, lane_addr + tc.WI, kin);
        umma::tmem_wait_st();
        umma::fence_before();
        umma::group_sync(3);
        for (long long tix = blockIdx.x; tix < ntiles; tix += gridDim.x, ++it) {
            const long long p = tix * 128 + j;
            const bool valid = p < npts;
            const long long n = valid ? p / L : 0;
            const int l = (int)(p - n * L);
            const float* xr = a.p.x + n * a.p.x_sn + (long long)l * a.p.x_sl;
            const float tval = valid ? a.p.t[n * a.p.t_sn + (long long)l * a.p.t_sl] : 0.f;
            const f4* sb = scr + (size_t)(it & 1) * nvs * 14 * 128;
            XW_PF(0, mbar_wait_or_trap(mFD, pFDp))             // F's scratch writes of this tile are visible from here on
            umma::mbar_arrive(mPC);                            // (F may not finish the NEXT tile before P has seen this one)
#pragma unroll 1
            for (int k = nv; k >= 0; --k) {
                // r_{k-1} is fetched while the previous P-op still runs; only the image stores wait for it
                f4 rv4[13];
                if (k > 0) {
#pragma unroll
                    for (int c = 0; c < 13; ++c) rv4[c] = ld_scratch(sb + (size_t)((k - 1) * 14 + c) * 128, pol);
                }
                if (pending) { XW_PF(1, mbar_wait_or_trap(mPh + hh_, pP)) pending = false; }     // this half's images are free again
#ifdef XW_TC_PROF
                const long long tim_ = clock64();
#endif
                if (k > 0) {
#pragma unroll
                    for (int c = 0; c < 13; ++c) {
                        const float rv[4] = {rv4[c].x, rv4[c].y, rv4[c].z, rv4[c].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int o = 4 * c + e;
                            if (o < HV) {
                                const float hi = umma::tf32_hi(rv[e]);
                                rT[ts_off(o, j)] = hi;
                                rT[ts_off(KP + o, j)] = rv[e] - hi;
                            }
                        }
                    }
                    rT[ts_off(BIASC, j)] = 1.f; rT[ts_off(KP + BIASC, j)] = 0.f;
                    if (kin > BIASC + 1) {                    // (rows a wide input layer of the previous tile has written)
#pragma unroll
                        for (int o = BIASC + 1; o < KP; ++o) { rT[ts_off(o, j)] = 0.f; rT[ts_off(KP + o, j)] = 0.f; }
                    }
                } else {
#pragma unroll 1
                    for (int c = 0; c < kin; ++c) {
                        float v = 0.f;
                        if (valid) v = c == 0 ? tval : (c <= a.d ? __ldcs(xr + c - 1) : (c == C ? 1.f : 0.f));
                        const float hi = umma::tf32_hi(v);
                        rT[ts_off(c, j)] = hi;
                        rT[ts_off(KP + c, j)] = v - hi;
                    }
                }
                umma::fence_smem_to_async();
#ifdef XW_TC_PROF
                pf_[8] += clock64() - tim_;
#endif
                XW_PF(2, mbar_wait_or_trap(mDPh + hh_, pDP))   // R has written this half's delta_k image
                XW_PF(3, asm volatile("bar.sync %0, 64;" ::"r"(4 + hh_) : "memory"))       // the 64 threads of this half
                if ((j & 63) == 0) {
                    umma::fence_after();
                    const int ho = hh_ * THALF2;
                    XW_PF(6, if (k > 0) issue_pop_half_stacked(tbase + tc.WH, dT + ho, rT + ho, idesc_p);
                             else issue_pop_half_in(tbase + tc.WI, dT + ho, rT + ho, idesc_pin);
                             umma::commit(mPh + hh_))
                }
                pending = true;
            }
            // The accumulators are flushed into the CTA's fp32 gradient image every `flush_tiles` tiles (and after the CTA's last
            // tile): the flush is on the critical path of the R <-> P hand-off (about 1.6 layer times per flush), while
            // a tensor-core accumulator may well sum a few tiles' worth of terms (accuracy measured: tools/tc_prof.py acc)
            if ((it + 1) % flush_tiles != 0 && tix + gridDim.x < ntiles) continue;
            XW_PF(4, mbar_wait_or_trap(mPh + hh_, pP))
            pending = false;
#ifdef XW_TC_PROF
            const long long tfl_ = clock64();
#endif
            umma::group_sync(3);                                  // both halves' last P-ops are complete
            umma::fence_after();
            {
                // accumulator lane m = image row m: lanes 0..55 hold the delta_hi rows, lanes 56..111 the delta_lo rows of the
                // 56 output units; columns [0, 56) the r_hi block, [56, 112) the r_lo block.  Each thread sums its two
                // column blocks; the hi-row owners add into the CTA's gradient image first, the lo-row owners second.
                const bool row_hi = j < KP, row_lo = j >= KP && j < 2 * KP;
                float* grow = gimg + (row_hi ? j : (row_lo ? j - KP : 0)) * GS;
                float acc[KP];
#pragma unroll
                for (int c8 = 0; c8 < KP; c8 += 8) {              // (8 columns at a time: two full rows would not fit the registers)
                    float a8[8], b8[8];
                    umma::tmem_ld8(lane_addr + tc.WH + c8, a8);
                    umma::tmem_ld8(lane_addr + tc.WH + KP + c8, b8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[c8 + e] = a8[e] + b8[e];
                }
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const bool mine = pass == 0 ? row_hi : row_lo;
                    if (mine && nv > 0) {
#pragma unroll
                        for (int i = 0; i < KP; ++i) grow[i] += acc[i];
                    }
#pragma unroll 1
                    for (int c8 = 0; c8 < kin; c8 += 8) {
                        float a8[8];
                        umma::tmem_ld8(lane_addr + tc.WI + c8, a8);
                        if (mine) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) grow[KP + c8 + e] += a8[e];
                        }
                    }
                    umma::group_sync(3);
                }
                // every P-op accumulates (two issuers, no defined first MMA): clear the accumulators for the next tile
                zero_acc(lane_addr + tc.WH, lane_addr + tc.WI, kin);
            }
            umma::tmem_wait_st();
            umma::fence_before();
            umma::group_sync(3);
            umma::fence_before();
#ifdef XW_TC_PROF
            pf_[5] += clock64() - tfl_;
#endif
        }
        if (j == 0) XW_PF_END(2)
        if (j == 64) XW_PF_END(3)
        if (j == 32) XW_PF_END(6)
        if (j == 96) XW_PF_END(7)
    }
    umma::fence_before();
    __syncthreads();
    // ------------------------------------------------------------------------ write the CTA's partial
    float* out = a.gpart + (size_t)blockIdx.x * g.size;
    const int Hvr = a.Hvr;
    for (int e = tid; e < Hvr * C; e += blockDim.x) { const int o = e / C, c = e - o * C; out[g.Wi + e] = gimg[o * GS + KP + c]; }
    for (int e = tid; e < Hvr * Hvr; e += blockDim.x) { const int o = e / Hvr, i = e - o * Hvr; out[g.Wh + e] = gimg[o * GS + i]; }
    for (int o = tid; o < Hvr; o += blockDim.x) {
        out[g.bi + o] = gimg[o * GS + KP + C];
        out[g.bh + o] = gimg[o * GS + BIASC];
        out[g.Wz + o] = zacc[o];
    }
    if (tid == 0) out[g.bz] = zacc[KP];
    __syncthreads();
    if (warp == 0) umma::tmem_free(tbase, 512);
}

// =============================================================================================
// time-row 0 of every path: grad_x v by reverse mode on the tensor cores -> the a grad(phi).du and b.du phi
// terms of the weak form (src/loss.py:66-69) and the grad_x phi cache.  Replaces k_vnet_points<HV, 2>.
// Three warpgroups, each with its own stream of 128-path tiles (thread = path = TMEM lane): forward layers
// (relu masks parked in shared memory), delta chain back to the input layer, then one more MMA
// grad_(t,x) v = delta_0 . Wi  ([128 x 56] x [56 x kin]) and the per-path dot products in the epilogue.
// =============================================================================================
__global__ void __launch_bounds__(384) k_vnet_tc_row0(VnetFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int kin = kin_of(a.d), nv = a.nv, nvs = nv > 0 ? nv : 1, d = a.d, C = d + 1;
    WImages w;
    w.wh_hi = reinterpret_cast<float*>(smem_raw);
    w.wh_lo = w.wh_hi + KP * NP;
    w.wht_hi = w.wh_lo + KP * NP;
    w.wht_lo = w.wht_hi + KP * NP;
    w.wi_hi = w.wht_lo + KP * NP;
    w.wi_lo = w.wi_hi + kin * NP;
    w.wz = w.wi_lo + kin * NP;
    float* wit_hi = w.wz + 64;                                   // [KP x kin] B[n = c][k = o] = Wi[o][c]
    float* wit_lo = wit_hi + KP * kin;
    uint2* smask = reinterpret_cast<uint2*>(wit_lo + KP * kin);   // [3][nvs][128]
    double* red = reinterpret_cast<double*>(smask + 3 * nvs * 128);
    uint64_t* mbars = reinterpret_cast<uint64_t*>(red + 4 * 32);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, wq = warp & 3, row = tid & 127;
    stage_images(w, a.theta, d, a.Hvr, kin);
    for (int i = tid; i < KP * kin; i += blockDim.x) { wit_hi[i] = 0.f; wit_lo[i] = 0.f; }
    __syncthreads();
    {
        const VLayout g(d, a.Hvr);
        for (int e = tid; e < a.Hvr * C; e += blockDim.x) {
            const int o = e / C, c = e - o * C;
            const int off = (o >> 2) * (kin * 4) + (c >> 3) * 32 + (c & 7) * 4 + (o & 3);
            put_split(wit_hi, wit_lo, off, a.theta[g.Wi + e]);
        }
    }
    if (tid < 4) umma::mbar_init(mbars + tid, 1);
    umma::fence_smem_to_async();
    if (warp == 0) umma::tmem_alloc(slot, 512);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tbase = *slot + (uint32_t)(wg * (NP + 2 * KP));
    const uint32_t colD = 0, colA = NP;
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * wq) << 16);
    uint64_t* mbar = mbars + wg;
    uint2* mk = smask + (size_t)wg * nvs * 128 + row;
    const bool issuer = row == 0;
    const uint32_t idesc = umma::idesc_tf32(128, NP), idesc_g = umma::idesc_tf32(128, kin);
    uint32_t parity = 0;
    const long long ntiles = ((long long)a.n + 127) / 128;
    double accs[4] = {0.0, 0.0, 0.0, 0.0};
    auto run_layer = [&](const float* b_hi, const float* b_lo) {   // A rows are stored: one hidden-width contraction
        umma::tmem_wait_st();
        umma::fence_before();
        umma::group_sync(1 + wg);
        if (issuer) {
            umma::fence_after();
            issue_3xtf32<KP / 8>(tbase + colD, tbase + colA, tbase + colA + KP, b_hi, b_lo, 0, idesc);
            umma::commit(mbar);
        }
        mbar_wait_or_trap(mbar, parity);
        umma::fence_after();
    };
    for (long long tix = (long long)blockIdx.x * 3 + wg; tix < ntiles; tix += (long long)gridDim.x * 3) {
        const long long n = tix * 128 + row;
        const bool valid = n < a.n;
        const long long nn = valid ? n : 0;
        const float* xp = a.p.x + nn * a.p.x_sn;
        const float tval = valid ? a.p.t[nn * a.p.t_sn] : 0.f;
        // ---- input row (t, x, 1) ----------------------------------------------------------------
#pragma unroll 1
        for (int c8 = 0; c8 < kin; c8 += 8) {
            float hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int idx = c8 + e;
                float v = 0.f;
                if (valid) v = idx == 0 ? tval : (idx <= d ? xp[idx - 1] : (idx == C ? 1.f : 0.f));
                hi[e] = umma::tf32_hi(v);
                lo[e] = v - hi[e];
            }
            umma::tmem_st8(lane_addr + colA + c8, hi);
            umma::tmem_st8(lane_addr + colA + KP + c8, lo);
        }
        umma::tmem_wait_st();
        umma::fence_before();
        umma::group_sync(1 + wg);
        if (issuer) {
            umma::fence_after();
            issue_3xtf32<0>(tbase + colD, tbase + colA, tbase + colA + KP, w.wi_hi, w.wi_lo, kin / 8, idesc);
            umma::commit(mbar);
        }
        float h[KP];
        mbar_wait_or_trap(mbar, parity);
        umma::fence_after();
        umma::tmem_ld56(lane_addr + colD, h);
        // ---- forward layers, masks kept ------------------------------------------------------------
#pragma unroll 1
        for (int layer = 0; layer < nv; ++layer) {
            uint32_t m0, m1;
            sign_masks(h, m0, m1);
            mk[(size_t)layer * 128] = make_uint2(m0, m1);
#pragma unroll
            for (int o = 0; o < HV; ++o) h[o] = fmaxf(h[o], 0.f);
            h[BIASC] = 1.f;
#pragma unroll
            for (int o = BIASC + 1; o < KP; ++o) h[o] = 0.f;
            store_a_row_chunked(lane_addr + colA, lane_addr + colA + KP, h);
            run_layer(w.wh_hi, w.wh_lo);
            umma::tmem_ld56(lane_addr + colD, h);
        }
        // ---- output layer: v and the cotangent of h_nv for d v ---------------------------------------
        float v = w.wz[KP];
#pragma unroll
        for (int o = 0; o < HV; ++o) {
            const float tau = tanh_fast(h[o]);
            v = fmaf(w.wz[o], tau, v);
            h[o] = w.wz[o] * (1.f - tau * tau);
        }
#pragma unroll
        for (int o = HV; o < KP; ++o) h[o] = 0.f;
        // ---- delta chain back to the input layer ---------------------------------------------------
#pragma unroll 1
        for (int k = nv; k >= 1; --k) {
            store_a_row_chunked(lane_addr + colA, lane_addr + colA + KP, h);
            run_layer(w.wht_hi, w.wht_lo);
            umma::tmem_ld56(lane_addr + colD, h);
            const uint2 m = mk[(size_t)(k - 1) * 128];
            apply_sign_masks(h, m.x, m.y);
#pragma unroll
            for (int o = HV; o < KP; ++o) h[o] = 0.f;
        }
        // ---- grad_(t,x) v = delta_0 . Wi ----------------------------------------------------------
        store_a_row_chunked(lane_addr + colA, lane_addr + colA + KP, h);
        umma::tmem_wait_st();
        umma::fence_before();
        umma::group_sync(1 + wg);
        if (issuer) {
            umma::fence_after();
            const uint64_t dh = umma::smem_desc(wit_hi, kin * 16, 128), dl = umma::smem_desc(wit_lo, kin * 16, 128);
            const uint64_t kStep = (uint64_t)((2 * kin * 16) >> 4);
#pragma unroll 1
            for (int t3 = 0; t3 < 3; ++t3)
#pragma unroll 1
                for (int ks = 0; ks < KP / 8; ++ks)
                    umma::mma_tf32_ts(tbase + colD, tbase + colA + (t3 == 0 ? KP : 0) + 8 * ks, (t3 == 1 ? dl : dh) + kStep * ks, idesc_g,
                                      (t3 | ks) ? 1u : 0u);
            umma::commit(mbar);
        }
        mbar_wait_or_trap(mbar, parity);
        umma::fence_after();
        // ---- per-path terms --------------------------------------------------------------------------
        const DomW W = domain_w(a.dom_kind, a.dp0, a.dp1, a.dp2, tval, xp, d);
        const float phi = v * W.w;
        const float* dun = a.du + nn * d;
        float s31 = 0.f;
#pragma unroll 1
        for (int c8 = 0; c8 < kin; c8 += 8) {
            float g8[8];
            umma::tmem_ld8(lane_addr + colD + c8, g8);
            if (valid) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int i = c8 + e - 1;               // column 0 is d/dt, columns 1..d the spatial gradient
                    if (i >= 0 && i < d) {
                        const float dphi_i = fmaf(W.w, g8[e], v * domain_dw_x(W, i, xp));
                        if (a.gcache) a.gcache[nn * d + i] = dphi_i;
                        float q;
                        if (a.ca) {
                            q = 0.f;
                            for (int jj = 0; jj < d; ++jj) q = fmaf(a.ca[nn * a.ca_sn + i * d + jj], dun[jj], q);
                        } else {
                            q = dun[i];
                        }
                        s31 = fmaf(dphi_i, q, s31);
                    }
                }
            }
        }
        if (valid) {
            if (a.cb) {
                float bq = 0.f;
                for (int jj = 0; jj < d; ++jj) bq = fmaf(a.cb[nn * a.cb_sn + jj], dun[jj], bq);
                s31 = fmaf(phi, bq, s31);
            }
            accs[2] += (double)s31;
        }
        umma::fence_before();
    }
    const int idx[4] = {0, 1, 2, 3};
    block_sum_to_global<4>(accs, red, a.sums, idx);
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free(*slot, 512);
}

}  // namespace tc
}  // namespace xw
#endif

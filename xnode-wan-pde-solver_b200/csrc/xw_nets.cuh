// xw_nets.cuh -- per-thread device code of the two networks:
//   * XNODE primal net u_theta: lift 1->H->H->H, explicit fixed-grid RK integration of the
//     field MLP, projection H->1            (reference: src/model.py:87-112,133-141,153-156)
//   * test-function net v_phi                (reference: src/model.py:37-47)
// plus their hand-derived reverse sweeps and the warp-level outer-product
// accumulator used for parameter gradients.
//
// Mapping: ONE THREAD walks ONE PATH (XNODE) or ONE POINT (v net).  All state lives in registers
// (fully unrolled compile-time sizes); weights are staged once per CTA in shared memory in both
// orientations (in-major for forward, out-major for reverse) and read with broadcast 128-bit loads.
#pragma once
#include "xw_platform.cuh"

namespace xw {

// ---------------------------------------------------------------------------------------------
// flat global layouts (reference named_parameters() order, PyTorch [out][in])
// ---------------------------------------------------------------------------------------------
struct ULayout {
    int d, H, hh;
    int W0, b0, W1, b1, W2, b2, Wa, ba, Ws, bs, Wf, bf, Wo, bo, size, lda;
    XW_HD ULayout() {}
    XW_HD ULayout(int d_, int H_, int hh_) : d(d_), H(H_), hh(hh_) {
        lda = d + 1 + H;
        int o = 0;
        W0 = o; o += H;      b0 = o; o += H;
        W1 = o; o += H * H;  b1 = o; o += H;
        W2 = o; o += H * H;  b2 = o; o += H;
        Wa = o; o += hh * lda; ba = o; o += hh;
        Ws = o; o += hh * hh;  bs = o; o += hh;
        Wf = o; o += H * hh;   bf = o; o += H;
        Wo = o; o += H;        bo = o; o += 1;
        size = o;
    }
};
struct VLayout {
    int d, Hv, C;
    int Wi, bi, Wh, bh, Wz, bz, size;
    XW_HD VLayout() {}
    XW_HD VLayout(int d_, int Hv_) : d(d_), Hv(Hv_), C(d_ + 1) {
        int o = 0;
        Wi = o; o += Hv * C;  bi = o; o += Hv;
        Wh = o; o += Hv * Hv; bh = o; o += Hv;
        Wz = o; o += Hv;      bz = o; o += 1;
        size = o;
    }
};

// ---------------------------------------------------------------------------------------------
// shared-memory images of the weights.  H/HH/HV are the COMPILED capacities; real sizes may be
// smaller (zero padded: padded units stay exactly 0 through relu/tanh and contribute nothing).
// ---------------------------------------------------------------------------------------------
template <int H, int HH>
struct USmem {
    static constexpr int HP = pad4(H), HHP = pad4(HH);
    // in-major (forward): row i holds the weights of input i towards every output
    static constexpr int W0 = 0;                    // [HP]
    static constexpr int B0 = W0 + HP;              // [HP]
    static constexpr int W1T = B0 + HP;             // [H][HP]
    static constexpr int B1 = W1T + H * HP;
    static constexpr int W2T = B1 + HP;             // [H][HP]
    static constexpr int B2 = W2T + H * HP;
    static constexpr int WYT = B2 + HP;             // [H][HHP]   Wa[:, d+1+i]
    static constexpr int WT = WYT + H * HHP;        // [HHP]      Wa[:, d]
    static constexpr int BA = WT + HHP;             // [HHP]
    static constexpr int WST = BA + HHP;            // [HH][HHP]
    static constexpr int BS = WST + HH * HHP;
    static constexpr int WFT = BS + HHP;            // [HH][HP]
    static constexpr int BF = WFT + HH * HP;
    static constexpr int WO = BF + HP;              // [HP]
    static constexpr int BO = WO + HP;              // [4]
    // out-major (reverse)
    static constexpr int W1 = BO + 4;               // [H][HP]
    static constexpr int W2 = W1 + H * HP;          // [H][HP]
    static constexpr int WY = W2 + H * HP;          // [HH][HP]
    static constexpr int WS = WY + HH * HP;         // [HH][HHP]
    static constexpr int WF = WS + HH * HHP;        // [H][HHP]
    static constexpr int WXT = WF + H * HHP;        // [d][HHP]   Wa[:, j] (x part), j-major
    static constexpr int fixed_size = WXT;
    static constexpr int size(int d) { return WXT + d * HHP; }
};

template <int H, int HH>
XW_DEV void stage_theta_u(float* s, const float* XW_RESTRICT th, int d, int Hr, int HHr) {
    using S = USmem<H, HH>;
    const ULayout g(d, Hr, HHr);
    const int total = S::size(d);
    for (int i = XW_TID; i < total; i += XW_BDIM) s[i] = 0.f;
    XW_SYNCTHREADS();
    for (int i = XW_TID; i < Hr; i += XW_BDIM) {
        s[S::W0 + i] = th[g.W0 + i]; s[S::B0 + i] = th[g.b0 + i];
        s[S::B1 + i] = th[g.b1 + i]; s[S::B2 + i] = th[g.b2 + i];
        s[S::BF + i] = th[g.bf + i]; s[S::WO + i] = th[g.Wo + i];
    }
    for (int i = XW_TID; i < HHr; i += XW_BDIM) {
        s[S::BA + i] = th[g.ba + i]; s[S::BS + i] = th[g.bs + i]; s[S::WT + i] = th[g.Wa + i * g.lda + d];
    }
    if (XW_TID == 0) s[S::BO] = th[g.bo];
    for (int e = XW_TID; e < Hr * Hr; e += XW_BDIM) {
        int o = e / Hr, i = e % Hr;
        float w1 = th[g.W1 + e], w2 = th[g.W2 + e];
        s[S::W1 + o * S::HP + i] = w1;  s[S::W1T + i * S::HP + o] = w1;
        s[S::W2 + o * S::HP + i] = w2;  s[S::W2T + i * S::HP + o] = w2;
    }
    for (int e = XW_TID; e < HHr * Hr; e += XW_BDIM) {       // Wa y-part [o<hh][i<H]
        int o = e / Hr, i = e % Hr;
        float w = th[g.Wa + o * g.lda + d + 1 + i];
        s[S::WY + o * S::HP + i] = w;  s[S::WYT + i * S::HHP + o] = w;
    }
    for (int e = XW_TID; e < HHr * HHr; e += XW_BDIM) {
        int o = e / HHr, i = e % HHr;
        float w = th[g.Ws + e];
        s[S::WS + o * S::HHP + i] = w;  s[S::WST + i * S::HHP + o] = w;
    }
    for (int e = XW_TID; e < Hr * HHr; e += XW_BDIM) {       // Wf [o<H][i<hh]
        int o = e / HHr, i = e % HHr;
        float w = th[g.Wf + e];
        s[S::WF + o * S::HHP + i] = w;  s[S::WFT + i * S::HP + o] = w;
    }
    for (int e = XW_TID; e < HHr * d; e += XW_BDIM) {        // Wa x-part [o<hh][j<d]
        int o = e / d, j = e % d;
        s[S::WXT + j * S::HHP + o] = th[g.Wa + o * g.lda + j];
    }
    XW_SYNCTHREADS();
}

template <int HV>
struct VSmem {
    static constexpr int HVP = pad4(HV);
    static constexpr int BI = 0;                    // [HVP]
    static constexpr int WHT = BI + HVP;            // [HV][HVP] in-major
    static constexpr int BH = WHT + HV * HVP;       // [HVP]
    static constexpr int WH = BH + HVP;             // [HV][HVP] out-major
    static constexpr int WZ = WH + HV * HVP;        // [HVP]
    static constexpr int BZ = WZ + HVP;             // [4]
    static constexpr int WIT = BZ + 4;              // [C][HVP]  Wi[:, c], c-major
    static constexpr int size(int C) { return WIT + C * HVP; }
};

template <int HV>
XW_DEV void stage_theta_v(float* s, const float* XW_RESTRICT th, int d, int Hvr) {
    using S = VSmem<HV>;
    const VLayout g(d, Hvr);
    const int total = S::size(g.C);
    for (int i = XW_TID; i < total; i += XW_BDIM) s[i] = 0.f;
    XW_SYNCTHREADS();
    for (int i = XW_TID; i < Hvr; i += XW_BDIM) {
        s[S::BI + i] = th[g.bi + i]; s[S::BH + i] = th[g.bh + i]; s[S::WZ + i] = th[g.Wz + i];
    }
    if (XW_TID == 0) s[S::BZ] = th[g.bz];
    for (int e = XW_TID; e < Hvr * Hvr; e += XW_BDIM) {
        int o = e / Hvr, i = e % Hvr;
        float w = th[g.Wh + e];
        s[S::WH + o * S::HVP + i] = w;  s[S::WHT + i * S::HVP + o] = w;
    }
    for (int e = XW_TID; e < Hvr * g.C; e += XW_BDIM) {
        int o = e / g.C, c = e % g.C;
        s[S::WIT + c * S::HVP + o] = th[g.Wi + e];
    }
    XW_SYNCTHREADS();
}

// where the XNODE weight image lives (shared memory).  (A constant-bank variant was measured in round 1 and dropped:
// ptxas turns the constant reads into LDC + register operands and hoists them -- 255 registers, spills, no gain.)
struct WSmem {
    static constexpr bool kStage = true;
    const float* p;
    XW_DEV static WSmem make(const float* smem) { return WSmem{smem}; }
    XW_DEV const float* at(int off) const { return p + off; }
};

// ---------------------------------------------------------------------------------------------
// explicit Runge-Kutta tableaus of the reference's fixed-grid solvers (torchdiffeq 0.1.1:
// euler, midpoint, rk4 = 3/8 rule); see oracle/shims/torchdiffeq/__init__.py
// ---------------------------------------------------------------------------------------------
template <int SOLVER> struct Tableau;
template <> struct Tableau<0> {
    static constexpr int S = 1;
    XW_DEV static float c(int) { return 0.f; }
    XW_DEV static float a(int, int) { return 0.f; }
    XW_DEV static float b(int) { return 1.f; }
};
template <> struct Tableau<1> {
    static constexpr int S = 2;
    XW_DEV static float c(int s) { return s == 1 ? 0.5f : 0.f; }
    XW_DEV static float a(int s, int r) { return (s == 1 && r == 0) ? 0.5f : 0.f; }
    XW_DEV static float b(int s) { return s == 1 ? 1.f : 0.f; }
};
template <> struct Tableau<2> {
    static constexpr int S = 4;
    XW_DEV static float c(int s) { return s == 1 ? (1.f / 3.f) : s == 2 ? (2.f / 3.f) : s == 3 ? 1.f : 0.f; }
    XW_DEV static float a(int s, int r) {
        if (s == 1) return r == 0 ? (1.f / 3.f) : 0.f;
        if (s == 2) return r == 0 ? (-1.f / 3.f) : r == 1 ? 1.f : 0.f;
        if (s == 3) return r == 0 ? 1.f : r == 1 ? -1.f : r == 2 ? 1.f : 0.f;
        return 0.f;
    }
    XW_DEV static float b(int s) { return (s == 0 || s == 3) ? 0.125f : 0.375f; }
};

// ---------------------------------------------------------------------------------------------
// XNODE pieces
// ---------------------------------------------------------------------------------------------
template <int H, int HH, class W>
XW_DEV void lift_fwd(const W& s, float s0, float (&z1)[H], float (&z2)[H], float (&y)[H]) {
    using S = USmem<H, HH>;
    float w0[H], b0[H];
    load_row<H>(s.at(S::W0), w0);
    load_row<H>(s.at(S::B0), b0);
#pragma unroll
    for (int o = 0; o < H; ++o) z1[o] = fmaxf(fmaf(w0[o], s0, b0[o]), 0.f);
    load_row<H>(s.at(S::B1), z2);
    matvec_acc<H, H, S::HP>(s.at(S::W1T), z1, z2);
#pragma unroll
    for (int o = 0; o < H; ++o) z2[o] = fmaxf(z2[o], 0.f);
    load_row<H>(s.at(S::B2), y);
    matvec_acc<H, H, S::HP>(s.at(S::W2T), z2, y);
}

// recorders for the internals of one field evaluation
template <int HH>
struct RecNone {
    XW_DEV void relu_in(int, const float (&)[HH]) {}
    XW_DEV void tanh_out(const float (&)[HH]) {}
};
template <int HH>
struct RecBits {                       // relu masks as a bit stack + tanh outputs in registers
    BitStack128 m;
    float tau[HH];
    XW_DEV void relu_in(int, const float (&r)[HH]) {
        unsigned bits = 0;
#pragma unroll
        for (int i = 0; i < HH; ++i) bits |= (r[i] > 0.f ? 1u : 0u) << i;
        m.template push<HH>(bits);
    }
    XW_DEV void tanh_out(const float (&t)[HH]) {
#pragma unroll
        for (int i = 0; i < HH; ++i) tau[i] = t[i];
    }
};
template <int HH>
struct RecSmem {                       // post-relu activations of every shared layer in shared memory
    float* base;                       // [(nsh+1)][HH][blockDim] slice of this stage, + tid
    int stride;                        // blockDim
    int nsh;
    XW_DEV void relu_in(int j, const float (&r)[HH]) {
#pragma unroll
        for (int i = 0; i < HH; ++i) base[(j * HH + i) * stride] = r[i];
    }
    XW_DEV void tanh_out(const float (&t)[HH]) {   // slot nsh holds tanh(a_nsh)
#pragma unroll
        for (int i = 0; i < HH; ++i) base[(nsh * HH + i) * stride] = t[i];
    }
};

// F(t, y) = net(cat(x, t, y)); ax = Wa[:, :d] x + ba (path constant, hoisted)
template <int H, int HH, class Rec, class W>
XW_DEV void field_fwd(const W& s, const float (&ax)[HH], float t, const float (&y)[H], int nsh,
                      float (&out)[H], float (&tau)[HH], Rec& rec) {
    using S = USmem<H, HH>;
    float a[HH], wt[HH];
    load_row<HH>(s.at(S::WT), wt);
#pragma unroll
    for (int o = 0; o < HH; ++o) a[o] = fmaf(wt[o], t, ax[o]);
    matvec_acc<H, HH, S::HHP>(s.at(S::WYT), y, a);
    for (int j = 0; j < nsh; ++j) {
        float r[HH], b[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) r[i] = fmaxf(a[i], 0.f);
        rec.relu_in(j, r);
        load_row<HH>(s.at(S::BS), b);
        matvec_acc<HH, HH, S::HHP>(s.at(S::WST), r, b);
#pragma unroll
        for (int i = 0; i < HH; ++i) a[i] = b[i];
    }
#pragma unroll
    for (int i = 0; i < HH; ++i) tau[i] = tanh_fast(a[i]);
    rec.tanh_out(tau);
    load_row<H>(s.at(S::BF), out);
    matvec_acc<HH, H, S::HP>(s.at(S::WFT), tau, out);
}

// reverse of one field evaluation, input-VJP only (relu masks from a bit stack).
// gout: cotangent of F.  gy += dF/dy^T gout ; a0 += cotangent of the first pre-activation.
template <int H, int HH, class W>
XW_DEV void field_rev_bits(const W& s, RecBits<HH>& rec, int nsh, const float (&gout)[H],
                           float (&gy)[H], float (&a0)[HH]) {
    using S = USmem<H, HH>;
    float dl[HH];
#pragma unroll
    for (int i = 0; i < HH; ++i) dl[i] = 0.f;
    matvec_acc<H, HH, S::HHP>(s.at(S::WF), gout, dl);          // Wf^T gout
#pragma unroll
    for (int i = 0; i < HH; ++i) dl[i] *= (1.f - rec.tau[i] * rec.tau[i]);
    for (int j = nsh; j > 0; --j) {
        float dn[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) dn[i] = 0.f;
        matvec_acc<HH, HH, S::HHP>(s.at(S::WS), dl, dn);       // Ws^T delta
        unsigned bits = rec.m.template pop<HH>();
#pragma unroll
        for (int i = 0; i < HH; ++i) dl[i] = ((bits >> i) & 1u) ? dn[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < HH; ++i) a0[i] += dl[i];
    matvec_acc<HH, H, S::HP>(s.at(S::WY), dl, gy);             // Wy^T delta0
}

// ---------------------------------------------------------------------------------------------
// warp-level outer-product accumulation:  G[o][i] += sum_{lanes} dl_lane[o] * r_lane[i]
// Each lane stages its two vectors lane-contiguously in a per-warp buffer; every lane then owns a
// BO x BI block of G, contracts it over the 32 lanes with 128-bit loads and adds it into the
// WARP-PRIVATE gradient image (plain read-modify-write, no atomics).
// Dst(o, i) -> float* (or nullptr to discard) maps an element to its slot in the gradient image.
// ---------------------------------------------------------------------------------------------
constexpr int kStgLd = 36;   // staging row stride (floats): 16B aligned, rows 4 banks apart

template <int O, int I, int BO, int NBO, int BI, int NBI, class Dst>
XW_DEV void warp_outer(const float (&dl)[O], const float (&r)[I], float* stg_d, float* stg_r, Dst dst) {
    static_assert(NBO * NBI <= 32, "one block per lane");
    static_assert(NBI * BI >= I, "columns covered in one pass");
    const int lane = XW_TID & 31;
#pragma unroll
    for (int o = 0; o < O; ++o) stg_d[o * kStgLd + lane] = dl[o];
#pragma unroll
    for (int i = 0; i < I; ++i) stg_r[i * kStgLd + lane] = r[i];
    XW_SYNCWARP();
    const int bo = lane / NBI, bi = lane % NBI;
    if (lane < NBO * NBI) {
#pragma unroll
        for (int o0 = 0; o0 < O; o0 += NBO * BO) {
            float acc[BO][BI];
#pragma unroll
            for (int a = 0; a < BO; ++a)
#pragma unroll
                for (int b = 0; b < BI; ++b) acc[a][b] = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                f4 dv[BO], rv[BI];
#pragma unroll
                for (int a = 0; a < BO; ++a) dv[a] = ld4(stg_d + (o0 + bo + NBO * a) * kStgLd + 4 * q);
#pragma unroll
                for (int b = 0; b < BI; ++b) rv[b] = ld4(stg_r + (bi + NBI * b) * kStgLd + 4 * q);
#pragma unroll
                for (int a = 0; a < BO; ++a)
#pragma unroll
                    for (int b = 0; b < BI; ++b) {
                        float t = acc[a][b];
                        t = fmaf(dv[a].x, rv[b].x, t); t = fmaf(dv[a].y, rv[b].y, t);
                        t = fmaf(dv[a].z, rv[b].z, t); t = fmaf(dv[a].w, rv[b].w, t);
                        acc[a][b] = t;
                    }
            }
#pragma unroll
            for (int a = 0; a < BO; ++a)
#pragma unroll
                for (int b = 0; b < BI; ++b) {
                    const int o = o0 + bo + NBO * a, i = bi + NBI * b;
                    if (o < O && i < I) {
                        float* p = dst(o, i);
                        if (p) *p += acc[a][b];
                    }
                }
        }
    }
    XW_SYNCWARP();
}


#ifndef XW_EMU
// ---------------------------------------------------------------------------------------------
// the same contraction on the warp-level tensor-core path (mma.sync m16n8k8, TF32 with 3xTF32 error
// compensation: fp32-level result): M = o, N = i, K = the 32 lanes.  The fragments come straight out of
// the lane-contiguous staging rows (row stride 36 floats: the 8 x 4 threads of a fragment load hit 32
// different banks), so one 16 x 8 x 32 block costs 24 scalar shared-memory loads instead of the
// (BO + BI) x 8 LDS.128 of the FFMA version -- the XNODE backward is LSU bound, not FMA bound.
// Rows / columns beyond O / I read stale staging rows; they only reach outputs that are discarded.
// ---------------------------------------------------------------------------------------------
XW_DEV void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
XW_DEV void split_tf32(float v, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
template <int O, int I, class Dst>
XW_DEV void warp_outer_mma(const float (&dl)[O], const float (&r)[I], float* stg_d, float* stg_r, Dst dst) {
    constexpr int MT = (O + 15) / 16, NT = (I + 7) / 8;
    const int lane = XW_TID & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int o = 0; o < O; ++o) stg_d[o * kStgLd + lane] = dl[o];
#pragma unroll
    for (int i = 0; i < I; ++i) stg_r[i * kStgLd + lane] = r[i];
    XW_SYNCWARP();
    float c[MT][NT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) c[m][n][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        unsigned ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const float* pa = stg_d + (16 * m + g) * kStgLd + 8 * ks + t;
            split_tf32(pa[0], ah[m][0], al[m][0]);
            split_tf32(pa[8 * kStgLd], ah[m][1], al[m][1]);
            split_tf32(pa[4], ah[m][2], al[m][2]);
            split_tf32(pa[8 * kStgLd + 4], ah[m][3], al[m][3]);
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const float* pb = stg_r + (8 * n + g) * kStgLd + 8 * ks + t;
            split_tf32(pb[0], bh[n][0], bl[n][0]);
            split_tf32(pb[4], bh[n][1], bl[n][1]);
        }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                mma_tf32_16x8x8(c[m][n], al[m], bh[n]);       // small terms first
                mma_tf32_16x8x8(c[m][n], ah[m], bl[n]);
                mma_tf32_16x8x8(c[m][n], ah[m], bh[n]);
            }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int o = 16 * m + g + (e >> 1) * 8, i = 8 * n + 2 * t + (e & 1);
                if (o < O && i < I) {
                    float* p = dst(o, i);
                    if (p) *p += c[m][n][e];
                }
            }
    XW_SYNCWARP();
}
#endif

// same for a RUNTIME number of columns (e.g. the d spatial inputs): colval(i) returns this lane's
// i-th column value; columns are processed in chunks of 8*BI (r-side staging rows), rows in passes of 4*BO.
template <int O, int BO, int BI, class ColVal, class Dst>
XW_DEV void warp_outer_dyn(const float (&dl)[O], int ncols, ColVal colval, float* stg_d, float* stg_r, Dst dst) {
    constexpr int NBO = 4, NBI = 8;
    const int lane = XW_TID & 31;
#pragma unroll
    for (int o = 0; o < O; ++o) stg_d[o * kStgLd + lane] = dl[o];
    const int bo = lane / NBI, bi = lane % NBI;
    for (int c0 = 0; c0 < ncols; c0 += BI * NBI) {
        XW_SYNCWARP();
        for (int i = 0; i < BI * NBI; ++i) stg_r[i * kStgLd + lane] = (c0 + i < ncols) ? colval(c0 + i) : 0.f;
        XW_SYNCWARP();
#pragma unroll
        for (int o0 = 0; o0 < O; o0 += NBO * BO) {
            float acc[BO][BI];
#pragma unroll
            for (int a = 0; a < BO; ++a)
#pragma unroll
                for (int b = 0; b < BI; ++b) acc[a][b] = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                f4 dv[BO], rv[BI];
#pragma unroll
                for (int a = 0; a < BO; ++a) {
                    const int o = o0 + bo + NBO * a;
                    dv[a] = (o < O) ? ld4(stg_d + o * kStgLd + 4 * q) : f4{0.f, 0.f, 0.f, 0.f};
                }
#pragma unroll
                for (int b = 0; b < BI; ++b) rv[b] = ld4(stg_r + (bi + NBI * b) * kStgLd + 4 * q);
#pragma unroll
                for (int a = 0; a < BO; ++a)
#pragma unroll
                    for (int b = 0; b < BI; ++b) {
                        float t = acc[a][b];
                        t = fmaf(dv[a].x, rv[b].x, t); t = fmaf(dv[a].y, rv[b].y, t);
                        t = fmaf(dv[a].z, rv[b].z, t); t = fmaf(dv[a].w, rv[b].w, t);
                        acc[a][b] = t;
                    }
            }
#pragma unroll
            for (int a = 0; a < BO; ++a)
#pragma unroll
                for (int b = 0; b < BI; ++b) {
                    const int o = o0 + bo + NBO * a, i = c0 + bi + NBI * b;
                    if (o < O && i < ncols) {
                        float* p = dst(o, i);
                        if (p) *p += acc[a][b];
                    }
                }
        }
    }
    XW_SYNCWARP();
}

// ---------------------------------------------------------------------------------------------
// v net, one thread = one point
// ---------------------------------------------------------------------------------------------
constexpr int kMaxNv = 16;

// forward; masks[j] = relu mask of h_j (j = 0..nv-1) as bits; returns v, leaves tanh in tau
template <int HV, class Store>
XW_DEV float vnet_fwd(const float* s, float t, const float* XW_RESTRICT x, int d, int nv,
                      unsigned long long* masks, float (&tau)[HV], Store store) {
    using S = VSmem<HV>;
    float a[HV];
    load_row<HV>(s + S::BI, a);
    {
        float w[HV];
        load_row<HV>(s + S::WIT, w);
#pragma unroll
        for (int o = 0; o < HV; ++o) a[o] = fmaf(w[o], t, a[o]);
    }
    for (int j = 0; j < d; ++j) {
        float w[HV];
        load_row<HV>(s + S::WIT + (j + 1) * S::HVP, w);
        const float xj = x[j];
#pragma unroll
        for (int o = 0; o < HV; ++o) a[o] = fmaf(w[o], xj, a[o]);
    }
    for (int k = 0; k < nv; ++k) {
        float r[HV], b[HV];
        unsigned long long m = 0ull;
#pragma unroll
        for (int i = 0; i < HV; ++i) {
            r[i] = fmaxf(a[i], 0.f);
            m |= (unsigned long long)(r[i] > 0.f ? 1u : 0u) << i;
        }
        if (masks) masks[k] = m;
        store(k, r);
        load_row<HV>(s + S::BH, b);
        matvec_acc<HV, HV, S::HVP>(s + S::WHT, r, b);
#pragma unroll
        for (int i = 0; i < HV; ++i) a[i] = b[i];
    }
    float v = s[S::BZ];
    float wz[HV];
    load_row<HV>(s + S::WZ, wz);
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        tau[i] = tanh_fast(a[i]);
        v = fmaf(wz[i], tau[i], v);
    }
    return v;
}

struct StoreNone {
    template <int HV> XW_DEV void operator()(int, const float (&)[HV]) const {}
};

// reverse from cotangent g of v down to the cotangent of h_0 (input-VJP only, masks as bits)
template <int HV>
XW_DEV void vnet_rev_bits(const float* s, int nv, const unsigned long long* masks, const float (&tau)[HV],
                          float g, float (&dl)[HV]) {
    using S = VSmem<HV>;
    float wz[HV];
    load_row<HV>(s + S::WZ, wz);
#pragma unroll
    for (int i = 0; i < HV; ++i) dl[i] = g * wz[i] * (1.f - tau[i] * tau[i]);
    for (int k = nv; k > 0; --k) {
        float dn[HV];
#pragma unroll
        for (int i = 0; i < HV; ++i) dn[i] = 0.f;
        matvec_acc<HV, HV, S::HVP>(s + S::WH, dl, dn);
        const unsigned long long m = masks[k - 1];
#pragma unroll
        for (int i = 0; i < HV; ++i) dl[i] = ((m >> i) & 1ull) ? dn[i] : 0.f;
    }
}

// d v / d input_c = sum_o Wi[o][c] dl0[o]
template <int HV>
XW_DEV float vnet_input_grad(const float* s, int c, const float (&dl0)[HV]) {
    using S = VSmem<HV>;
    float w[HV];
    load_row<HV>(s + S::WIT + c * S::HVP, w);
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int o = 0; o + 1 < HV; o += 2) { g0 = fmaf(w[o], dl0[o], g0); g1 = fmaf(w[o + 1], dl0[o + 1], g1); }
    if (HV & 1) g0 = fmaf(w[HV - 1], dl0[HV - 1], g0);
    return g0 + g1;
}

// ---------------------------------------------------------------------------------------------
// domain weight w and its derivatives at one point (src/dataset.py:278-282, :216-218, :119-125)
// ---------------------------------------------------------------------------------------------
struct DomW {
    float w, dw_t;
    int arg;        // cube: coordinate carrying dw/dx (others 0)
    float dw_arg;   // cube: its value; sphere domains: -1/|x| (dw/dx_j = dw_arg * x_j)
};
XW_DEV DomW domain_w(int kind, float p0, float p1, float p2, float t, const float* XW_RESTRICT x, int d) {
    DomW r;
    if (kind == 0) {
        const float bot = p0, top = p1;
        float mt = 3.4e38f, mb = 3.4e38f;
        int it = 0, ib = 0;
        for (int j = 0; j < d; ++j) {
            const float xj = x[j];
            const float dt_ = fabsf(top - xj), db_ = fabsf(bot - xj);
            if (dt_ < mt) { mt = dt_; it = j; }
            if (db_ < mb) { mb = db_; ib = j; }
        }
        const bool use_top = mt <= mb;
        r.w = use_top ? mt : mb;
        r.arg = use_top ? it : ib;
        const float diff = use_top ? (top - x[r.arg]) : (bot - x[r.arg]);
        r.dw_arg = diff > 0.f ? -1.f : (diff < 0.f ? 1.f : 0.f);
        r.dw_t = 0.f;
        return r;
    }
    float n2 = 0.f;
    for (int j = 0; j < d; ++j) n2 = fmaf(x[j], x[j], n2);
    const float nrm = sqrtf(n2);
    r.arg = -1;
    r.dw_arg = -1.f / nrm;
    if (kind == 1) {
        r.w = p0 * (1.f - t) - nrm;
        r.dw_t = -p0;
    } else {
        const float span = p2 - p1;
        const bool first = t <= span * 0.5f;
        r.w = first ? (p0 * (span - t) - nrm) : (p0 * t - nrm);
        r.dw_t = first ? -p0 : p0;
    }
    return r;
}
XW_DEV float domain_dw_x(const DomW& r, int j, const float* XW_RESTRICT x) {
    return r.arg >= 0 ? (j == r.arg ? r.dw_arg : 0.f) : r.dw_arg * x[j];
}

}  // namespace xw

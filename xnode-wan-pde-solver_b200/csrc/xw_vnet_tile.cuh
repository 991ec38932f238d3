// xw_vnet_tile.cuh -- CTA-tiled test-function net (generation 2 of the v-net kernels).
//
// The hidden layers of v_phi are a [points x Hv] x [Hv x Hv] contraction (reference
// src/model.py:37-47: the SAME Hv x Hv matrix nv times).  Instead of one thread per point, a CTA of
// NG warps owns a tile of 128 "pair rows" in shared memory; a pair row interleaves two activation
// vectors (forward kernel: value and d/dt tangent of one point; backward kernel: two points):
//
//      T[row][k][c]      row stride kRS = 2*Hv + 4 floats  (16-byte rows, conflict-free LDS.128)
//
// Warp g (4 warps per CTA = one per SM sub-partition, so the barrier-coupled warps advance at the
// same rate) computes the TN=13 hidden units [13g, 13g+13) for every row; lane l owns rows l, l+32,
// l+64, l+96 (8 columns).  Per pair of input units it issues 4 LDS.128 of activations + 8 broadcast
// LDS.128 of weights for 208 FFMA (17:1); accumulators stay in registers for the whole layer and the
// layer output overwrites the tile in place between two __syncthreads.  Weights: W^T in shared
// memory grouped per warp ([k][g][16]); unit groups are padded to 14 tile positions so that every
// group starts on a 16-byte boundary (STS.128 epilogue) without padding the contraction.
#pragma once
// included at the end of xw_kernels.cuh (needs PointsView, block_sum_to_global)

namespace xw {

template <int HV, int QR_ = 3, int NH_ = 1>
struct VTile {
    static constexpr int NG = 4;                         // warps per CTA = output groups (one per SMSP)
    static constexpr int TN = (HV + NG - 1) / NG;        // hidden units per group (13)
    static constexpr int GP = TN + (TN & 1);             // group stride in tile positions (14: 16-byte aligned groups)
    static constexpr int TNP = pad4(TN);                 // padded group width in the weight image (16)
    static constexpr int WLD = NG * TNP;                 // weight row stride (64)
    static constexpr int QR = QR_;                       // pair rows per lane
    static constexpr int NH = NH_;                       // row blocks per CTA (each handled by its own NG warps)
    static constexpr int WROWS = 32 * QR;                // pair rows per warp
    static constexpr int ROWS = WROWS * NH;              // pair rows per tile
    static constexpr int RS0 = 2 * NG * GP + 4;
    static constexpr int RS = RS0 + ((RS0 / 4) % 2 == 0 ? 4 : 0);    // 116 for HV=50
    static constexpr int THREADS = NG * NH * 32;
    static_assert(TN * NG >= HV, "groups cover the hidden units");
    static_assert((RS / 4) % 2 == 1, "row stride must be an odd multiple of 16 bytes (bank-conflict free)");
    static constexpr int valid_in_group(int g) { return HV - TN * g >= TN ? TN : (HV - TN * g > 0 ? HV - TN * g : 0); }
    // shared memory image (floats); hidden unit o lives in column (o / TN) * TNP + o % TN of a weight row
    static constexpr int WT = 0;                         // [HV][WLD]  hidden, row = input unit k
    static constexpr int BH = WT + HV * WLD;             // [WLD]
    static constexpr int WZ = BH + WLD;                  // [WLD]
    static constexpr int BI = WZ + WLD;                  // [WLD]
    static constexpr int MISC = BI + WLD;                // [4]  bz
    static constexpr int WIT = MISC + 4;                 // [C][WLD]  input layer, row = input channel c
    static constexpr int wit_size(int C) { return C * WLD; }
    static constexpr int xin_ld(int C) { return C | 1; }     // odd stride: conflict-free scalar loads
};

template <int HV, int QR>
XW_DEV void stage_theta_v_tile(float* s, const float* XW_RESTRICT th, int d, int Hvr) {
    using VT = VTile<HV, QR>;
    const VLayout g(d, Hvr);
    const int total = VT::WIT + VT::wit_size(g.C);
    for (int i = XW_TID; i < total; i += XW_BDIM) s[i] = 0.f;
    XW_SYNCTHREADS();
    for (int o = XW_TID; o < Hvr; o += XW_BDIM) {
        const int pos = (o / VT::TN) * VT::TNP + (o % VT::TN);
        s[VT::BH + pos] = th[g.bh + o];
        s[VT::WZ + pos] = th[g.Wz + o];
        s[VT::BI + pos] = th[g.bi + o];
    }
    if (XW_TID == 0) s[VT::MISC] = th[g.bz];
    for (int e = XW_TID; e < Hvr * Hvr; e += XW_BDIM) {
        const int o = e / Hvr, k = e % Hvr;
        s[VT::WT + k * VT::WLD + (o / VT::TN) * VT::TNP + (o % VT::TN)] = th[g.Wh + e];
    }
    for (int e = XW_TID; e < Hvr * g.C; e += XW_BDIM) {
        const int o = e / g.C, c = e % g.C;
        s[VT::WIT + c * VT::WLD + (o / VT::TN) * VT::TNP + (o % VT::TN)] = th[g.Wi + e];
    }
    XW_SYNCTHREADS();
}

// accumulate one input unit (its (c0, c1) pair per row) into the register tile: 13 packed FFMA2 per row
template <int HV, int QR>
XW_DEV void tile_fma_unit(const float* wk, const fpair (&a)[QR], fpair (&acc)[QR][VTile<HV, QR>::TN]) {
    using VT = VTile<HV, QR>;
    float w[VT::TN];
    load_row<VT::TN>(wk, w);
#pragma unroll
    for (int o = 0; o < VT::TN; ++o) {
        const fpair ww = pack2(w[o], w[o]);
#pragma unroll
        for (int q = 0; q < VT::QR; ++q) acc[q][o] = fma2(a[q], ww, acc[q][o]);
    }
}

// contraction over the NU input units of one group (tile positions tg.., weight rows wg..): fully
// unrolled straight-line code so that the activation / weight loads of later units are scheduled
// under the FFMA2 stream of earlier ones (no exposed LDS latency at loop boundaries)
template <int HV, int QR, int NU>
XW_DEV void tile_group_fma(const float* tg, const float* wg, fpair (&acc)[QR][VTile<HV, QR>::TN]) {
    using VT = VTile<HV, QR>;
#pragma unroll
    for (int jp = 0; jp < NU / 2; ++jp) {
        fpair a[QR], b[QR];
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            const f4 v = ld4(tg + q * 32 * VT::RS + 4 * jp);
            a[q] = pack2(v.x, v.y); b[q] = pack2(v.z, v.w);
        }
        tile_fma_unit<HV, QR>(wg + (2 * jp) * VT::WLD, a, acc);
        tile_fma_unit<HV, QR>(wg + (2 * jp + 1) * VT::WLD, b, acc);
    }
    if (NU & 1) {
        fpair a[QR];
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            const f2 v = ld2(tg + q * 32 * VT::RS + 2 * (NU - 1));
            a[q] = pack2(v.x, v.y);
        }
        tile_fma_unit<HV, QR>(wg + (NU - 1) * VT::WLD, a, acc);
    }
}

// acc[q][o] += sum_k W[k-row][o] * T[row_q][k][:]  over all HV input units; `wimg` is the weight image
// (row = input unit, columns = this warp's outputs), `trow` the lane's first pair row
template <int HV, int QR>
XW_DEV void tile_contract(const float* wimg, const float* trow, fpair (&acc)[QR][VTile<HV, QR>::TN]) {
    using VT = VTile<HV, QR>;
    constexpr int NFULL = HV / VT::TN;                 // groups with all TN units
    constexpr int NLAST = HV - NFULL * VT::TN;         // units of the trailing partial group
#pragma unroll 1
    for (int gk = 0; gk < NFULL; ++gk)
        tile_group_fma<HV, QR, VT::TN>(trow + 2 * VT::GP * gk, wimg + gk * VT::TN * VT::WLD, acc);
    if (NLAST > 0)
        tile_group_fma<HV, QR, (NLAST > 0 ? NLAST : 1)>(trow + 2 * VT::GP * NFULL, wimg + NFULL * VT::TN * VT::WLD, acc);
}

// one hidden layer on the pair-row tile:  acc[q][o] = (bias[o], BIAS1 ? bias[o] : 0) + sum_k W[o][k] * T[row_q][k][:]
// BIAS1: add the bias to the second vector of the pair too (false for tangents)
template <int HV, int QR, bool BIAS1>
XW_DEV void tile_hidden_layer(const float* sw, const float* tile, int grp, int lane,
                              fpair (&acc)[QR][VTile<HV, QR>::TN]) {
    using VT = VTile<HV, QR>;
    XW_FENCE();
    {
        float b[VT::TN];
        load_row<VT::TN>(sw + VT::BH + grp * VT::TNP, b);
#pragma unroll
        for (int o = 0; o < VT::TN; ++o) {
            const fpair bb = pack2(b[o], BIAS1 ? b[o] : 0.f);
#pragma unroll
            for (int q = 0; q < QR; ++q) acc[q][o] = bb;
        }
    }
    tile_contract<HV, QR>(sw + VT::WT + grp * VT::TNP, tile + lane * VT::RS, acc);
}

// store relu(value) and the masked tangent of a layer's pre-activations back into the tile
template <int HV, int QR>
XW_DEV void tile_store_relu_tangent(float* tile, int grp, int lane, const fpair (&acc)[QR][VTile<HV, QR>::TN]) {
    using VT = VTile<HV, QR>;
#pragma unroll
    for (int q = 0; q < VT::QR; ++q) {
        float* dst = tile + (lane + 32 * q) * VT::RS + 2 * VT::GP * grp;
#pragma unroll
        for (int o = 0; o + 1 < VT::TN; o += 2) {
            float v0, t0, v1, t1;
            unpack2(acc[q][o], v0, t0);
            unpack2(acc[q][o + 1], v1, t1);
            f4 v;
            v.x = fmaxf(v0, 0.f); v.y = v0 > 0.f ? t0 : 0.f;
            v.z = fmaxf(v1, 0.f); v.w = v1 > 0.f ? t1 : 0.f;
            st4(dst + 2 * o, v);
        }
        if (VT::TN & 1) {
            constexpr int o = VT::TN - 1;
            float v0, t0;
            unpack2(acc[q][o], v0, t0);
            f2 v;
            v.x = fmaxf(v0, 0.f); v.y = v0 > 0.f ? t0 : 0.f;
            *reinterpret_cast<f2*>(dst + 2 * o) = v;
        }
    }
}

// =============================================================================================
// interior forward over all points: v, dv/dt (forward-mode tangent), weak-form integrands, seeds
// =============================================================================================
struct VtileFwdArgs {
    int d, Hvr, nv, n, L;
    const float* theta; PointsView p;
    int dom_kind; float dp0, dp1, dp2;
    float c0, c1;
    const float* Aval; const float* Ader;   // optional per-point A(u) = c(X,u) u and dA/du (general c: host-evaluated)
    const float* u; const float* h; const float* f;
    double* sums; float* cot_u; float* cot_v;
    float* vcache;           // optional [n*L] x (v, dv/dt, w, dw/dt): lets later sub-steps on the same sample
                             // and the same theta_v skip the v net entirely (k_weak_combine)
    const float* wbuf;       // optional [n*L] domain weight and
    const float* dwtbuf;     // optional [n*L] its time derivative (tensor-core forward on the virtual net, d > 54)
};

// the weak-form integrands of one point from (v, dv/dt, w, dw/dt) and (u, f, h): src/loss.py:64-73
// without the time-row-0 gradient term; accumulates into acc[0..3], writes the cotangent seeds
// the zeroth-order term A(u) = c(X, u) u and its derivative: the affine structure c = c0 + c1 u in closed form, or -- for
// a coefficient that depends on X or is not affine in u (src/training.py:30, src/loss.py:70 accept any callable) -- the
// per-point values the host evaluated with the user's callable at the u of xw_xnode_eval (xw_coef.A_val / A_der)
XW_DEV void weak_A(float c0, float c1, const float* Aval, const float* Ader, long long p, float u, float& A, float& Ap) {
    if (Aval) { A = Aval[p]; Ap = Ader[p]; return; }
    const float cu_ = fmaf(c1, u, c0);
    A = cu_ * u;
    Ap = fmaf(c1, u, cu_);
}

XW_DEV void weak_point_terms(float v, float dv_t, float w, float dw_t, float u, float fv, float hn, int l, int L,
                             float A, float Ap, double (&acc)[4], float& cu, float& cv) {
    const float phi = v * w;
    const float dphi0 = fmaf(w, dv_t, v * dw_t);
    float s1 = 0.f;
    cu = Ap * phi;
    cv = w * (A + fv);
    if (l == L - 1) { s1 = fmaf(u, v, s1); cu = fmaf((float)L, v, cu); cv = fmaf((float)L, u, cv); }
    if (l == 0) { s1 = fmaf(-hn, v, s1); cv = fmaf(-(float)L, hn, cv); }
    acc[0] += (double)s1;
    acc[1] += (double)(u * dphi0);
    acc[2] += (double)((A + fv) * phi);
    acc[3] += (double)(v * v);
}

template <int HV, int QR, int NH>
XW_GLOBAL void
#ifndef XW_EMU
__launch_bounds__(VTile<HV, QR, NH>::THREADS, (NH > 1 ? 2 : 1))
#endif
k_vnet_tile_fwd(VtileFwdArgs a) {
    using VT = VTile<HV, QR, NH>;
    XW_DYN_SMEM(smem_raw);
    const int C = a.d + 1, XLD = VT::xin_ld(C);
    float* sw = reinterpret_cast<float*>(smem_raw);
    float* tile = sw + pad4(VT::WIT + VT::wit_size(C));
    float* xin = tile + VT::ROWS * VT::RS;                        // [ROWS][XLD]  (t, x_0..x_{d-1})
    float* redv = xin + VT::ROWS * XLD;                           // [NG][ROWS][2]
    double* red = reinterpret_cast<double*>(redv + VT::NG * VT::ROWS * 2);
    stage_theta_v_tile<HV, QR>(sw, a.theta, a.d, a.Hvr);
    const int lane = XW_TID & 31, grp = (XW_TID >> 5) % VT::NG, rbase = ((XW_TID >> 5) / VT::NG) * VT::WROWS;
    float* tile_w = tile + rbase * VT::RS;                      // this warp's block of pair rows
    const long long npts = (long long)a.n * a.L;
    const long long ntiles = (npts + VT::ROWS - 1) / VT::ROWS;
    const int L = a.L;
    int* rown = reinterpret_cast<int*>(red + 4 * 32);            // [ROWS] path index of each row (-1: past the end)
    int* rowl = rown + VT::ROWS;                                  // [ROWS] time index
    double accs[4] = {0.0, 0.0, 0.0, 0.0};
#if !defined(XW_EMU) && defined(XW_PHASE_SKEW)
    // co-resident CTAs run identical phases (FMA-heavy layer, then store/barrier): skew the second
    // resident set by about half a layer so that one CTA's epilogue overlaps the other's FMA phase
    if (XW_BID >= (XW_GDIM >> 1)) {
        const long long t0 = clock64();
        while (clock64() - t0 < XW_PHASE_SKEW) {}
    }
#endif
    float wz[VT::TN], wi0[VT::TN];
    load_row<VT::TN>(sw + VT::WZ + grp * VT::TNP, wz);
    load_row<VT::TN>(sw + VT::WIT + grp * VT::TNP, wi0);          // dh_0/dt = Wi[:, 0]
    for (long long tix = XW_BID; tix < ntiles; tix += XW_GDIM) {
        const long long p0 = tix * VT::ROWS;
        // ---- stage the tile's inputs --------------------------------------------------------
        if (XW_TID < VT::ROWS) {
            const long long p = p0 + XW_TID;
            const long long n = p / L;
            rown[XW_TID] = p < npts ? (int)n : -1;
            rowl[XW_TID] = (int)(p - n * L);
        }
        XW_SYNCTHREADS();
        for (int e = XW_TID; e < VT::ROWS * C; e += XW_BDIM) {
            const unsigned r = (unsigned)e / (unsigned)C, c = (unsigned)e - r * (unsigned)C;
            const int n = rown[r], l = rowl[r];
            float val = 0.f;
            if (n >= 0)
                val = c == 0 ? a.p.t[(long long)n * a.p.t_sn + l * a.p.t_sl]
                             : a.p.x[(long long)n * a.p.x_sn + l * a.p.x_sl + (c - 1)];
            xin[r * XLD + c] = val;
        }
        XW_SYNCTHREADS();
        // ---- input layer: h_0 = Wi (t, x) + bi ; tangent dh_0/dt = Wi[:, 0] -------------------
        fpair acc[VT::QR][VT::TN];
        {
            float h0[VT::QR][VT::TN];
            float b[VT::TN];
            load_row<VT::TN>(sw + VT::BI + grp * VT::TNP, b);
#pragma unroll
            for (int q = 0; q < VT::QR; ++q)
#pragma unroll
                for (int o = 0; o < VT::TN; ++o) h0[q][o] = b[o];
            const float* wrow = sw + VT::WIT + grp * VT::TNP;
            for (int c = 0; c < C; ++c) {
                XW_FENCE();
                float w[VT::TN];
                load_row<VT::TN>(wrow + c * VT::WLD, w);
#pragma unroll
                for (int q = 0; q < VT::QR; ++q) {
                    const float xv = xin[(rbase + lane + 32 * q) * XLD + c];
#pragma unroll
                    for (int o = 0; o < VT::TN; ++o) h0[q][o] = fmaf(xv, w[o], h0[q][o]);
                }
            }
#pragma unroll
            for (int q = 0; q < VT::QR; ++q)
#pragma unroll
                for (int o = 0; o < VT::TN; ++o) acc[q][o] = pack2(h0[q][o], wi0[o]);
        }
        // ---- hidden layers ---------------------------------------------------------------------
        for (int layer = 0; layer < a.nv; ++layer) {
            tile_store_relu_tangent<HV, QR>(tile_w, grp, lane, acc);
            XW_SYNCTHREADS();
            tile_hidden_layer<HV, QR, false>(sw, tile_w, grp, lane, acc);
            XW_SYNCTHREADS();
        }
        // ---- tanh + output layer: partial dot products of this warp's outputs -------------------
#pragma unroll
        for (int q = 0; q < VT::QR; ++q) {
            float pv = 0.f, pt = 0.f;
#pragma unroll
            for (int o = 0; o < VT::TN; ++o) {
                float hv, ht;
                unpack2(acc[q][o], hv, ht);
                const float y = tanh_fast(hv);
                pv = fmaf(wz[o], y, pv);
                pt = fmaf(wz[o] * (1.f - y * y), ht, pt);
            }
            f2 out; out.x = pv; out.y = pt;
            *reinterpret_cast<f2*>(redv + (grp * VT::ROWS + rbase + lane + 32 * q) * 2) = out;
        }
        XW_SYNCTHREADS();
        // ---- per-point weak-form terms (one thread per row) ------------------------------------
        if (XW_TID < VT::ROWS) {
            const int r = XW_TID;
            const long long p = p0 + r;
            if (rown[r] >= 0) {
                float v = sw[VT::MISC], dv_t = 0.f;
#pragma unroll
                for (int g2 = 0; g2 < VT::NG; ++g2) {
                    const f2 pr = *reinterpret_cast<const f2*>(redv + (g2 * VT::ROWS + r) * 2);
                    v += pr.x; dv_t += pr.y;
                }
                const int n = rown[r], l = rowl[r];
                const float* xr = xin + r * XLD;
                const DomW W = domain_w(a.dom_kind, a.dp0, a.dp1, a.dp2, xr[0], xr + 1, a.d);
                float cu, cv;
                float Au, Ap;
                weak_A(a.c0, a.c1, a.Aval, a.Ader, p, a.u[p], Au, Ap);
                weak_point_terms(v, dv_t, W.w, W.dw_t, a.u[p], a.f[p], l == 0 ? a.h[n] : 0.f, l, L, Au, Ap, accs, cu, cv);
                a.cot_u[p] = cu;
                a.cot_v[p] = cv;
                if (a.vcache) { f4 cch; cch.x = v; cch.y = dv_t; cch.z = W.w; cch.w = W.dw_t; st4(a.vcache + 4 * p, cch); }
            }
        }
        XW_SYNCTHREADS();
    }
    const int idx[4] = {0, 1, 2, 3};
    block_sum_to_global<4>(accs, red, a.sums, idx);
}

// =============================================================================================
// v net backward, CTA-tiled:  parameter gradients of sum_p G[p] v[p],  G = k0*cot_v + k1*v + k2*w
//
// A pair row holds TWO POINTS.  Per tile of 2*ROWS points:
//   forward : input layer + nv hidden layers (tile engine); the post-relu activations r_0..r_{nv-1}
//             are parked in a per-CTA scratch (global memory, sized to stay L2 resident);
//   reverse : for k = nv..1   delta_k -> tile D ; r_{k-1} -> tile R (cp.async, overlapped with the
//             previous layer's contraction) ; P-op  dWh += D^T R (K-split over the 4 warps, each lane
//             owns a 13 x 7 block of dWh in REGISTERS for the whole kernel) ; R-op  delta_{k-1} =
//             relu'(r_{k-1}) . Wh^T delta_k (tile engine with the out-major weight image);
//   input   : dWi += delta_0^T (t, x, 1).
// =============================================================================================
struct VtileBwdArgs {
    int d, Hvr, nv, n, L;
    const float* theta; PointsView p;
    int dom_kind; float dp0, dp1, dp2;
    const float* cot; const double* coefs;
    float* scratch;          // [gridDim][nv][ROWS*RS] floats
    float* gpart;            // [gridDim][P]
    // tensor-core backward on a VIRTUAL input (d > 54, xw_capi.cu): domain weight per point from a buffer (the virtual
    // coordinates are not the spatial ones) and the cotangent of the first pre-activation dumped per point
    const float* wbuf;       // optional [n*L]
    float* delta0_out;       // optional [n*L][52]
    int tm_packed;           // k_vnet_tc_bwd3: force the fully packed tensor-memory layout (test hook, XW_TC_TMEM_PACKED=1)
    int flush_tiles;         // k_vnet_tc_bwd3: tiles per flush of the weight-gradient accumulators (0 = 1)
};

template <int HV, int QR>
struct VTileB {
    using VT = VTile<HV, QR>;
    static constexpr int NPT = 2 * VT::ROWS;                     // points per tile
    static constexpr int WR = 0;                                 // [HV][WLD]  out-major image: row o, columns i grouped
    static constexpr int IB = 7;                                 // dWh block per lane: TN (13) x IB (7)
    static_assert(VT::TN == 13 && VT::NG == 4, "lane blocking below assumes 4 groups of 13");
    static_assert(NPT <= VT::THREADS, "one thread per point in the cotangent step");
};

template <int HV, int QR>
XW_GLOBAL void k_vnet_tile_bwd(VtileBwdArgs a) {
    using VT = VTile<HV, QR>;
    using VB = VTileB<HV, QR>;
    constexpr int TN = VT::TN, NPT = VB::NPT, IB = VB::IB;
    XW_DYN_SMEM(smem_raw);
    const int C = a.d + 1, XLD = VT::xin_ld(C);
    const VLayout g(a.d, a.Hvr);
    float* sw = reinterpret_cast<float*>(smem_raw);                         // forward weight image
    float* swr = sw + pad4(VT::WIT + VT::wit_size(C));                      // [HV][WLD] out-major hidden weights
    float* tileD = swr + HV * VT::WLD;
    float* tileR = tileD + VT::ROWS * VT::RS;
    float* xin = tileR + VT::ROWS * VT::RS;                                 // [NPT][XLD]
    float* vred = xin + pad4(NPT * XLD);                                    // [NG][NPT]
    float* gpt = vred + VT::NG * NPT;                                       // [NPT]  G per point
    float* gwi = gpt + NPT;                                                 // [HV][C+1] dWi | dbi image
    int* rown = reinterpret_cast<int*>(gwi + pad4(HV * (C + 1)));           // [NPT]
    int* rowl = rown + NPT;
    stage_theta_v_tile<HV, QR>(sw, a.theta, a.d, a.Hvr);
    for (int i = XW_TID; i < HV * VT::WLD; i += XW_BDIM) swr[i] = 0.f;
    for (int i = XW_TID; i < HV * (C + 1); i += XW_BDIM) gwi[i] = 0.f;
    XW_SYNCTHREADS();
    for (int e = XW_TID; e < a.Hvr * a.Hvr; e += XW_BDIM) {
        const int o = e / a.Hvr, i = e % a.Hvr;
        swr[o * VT::WLD + (i / TN) * VT::TNP + (i % TN)] = a.theta[g.Wh + e];
    }
    XW_SYNCTHREADS();
    const int lane = XW_TID & 31, grp = XW_TID >> 5;
    const float k0 = (float)a.coefs[0], k1 = (float)a.coefs[1], k2 = (float)a.coefs[2];
    const long long npts = (long long)a.n * a.L;
    const long long ntiles = (npts + NPT - 1) / NPT;
    const int L = a.L, nv = a.nv;
    float* scr = a.scratch + (size_t)XW_BID * (nv > 0 ? nv : 1) * VT::ROWS * VT::RS;
    float wz[TN];
    load_row<TN>(sw + VT::WZ + grp * VT::TNP, wz);
    // persistent gradient accumulators
    float pacc[TN][IB];                 // dWh block: o in group (lane/8), i-slots 7*(lane%8)..+7
#pragma unroll
    for (int o = 0; o < TN; ++o)
#pragma unroll
        for (int j = 0; j < IB; ++j) pacc[o][j] = 0.f;
    float gwz[TN];
#pragma unroll
    for (int o = 0; o < TN; ++o) gwz[o] = 0.f;
    float gbz = 0.f;
    const int p_ob = lane >> 3, p_ib = lane & 7;
    // i-slot -> tile position: blocks (g, h) = (ib/2, ib%2) cover units 13g + 7h + j; slot (ib=7, j=4) is the bias column
    const int p_rpos = 2 * (VT::GP * (p_ib >> 1) + 7 * (p_ib & 1));       // float offset of the lane's first R unit

    auto issue_r_copy = [&](int layer) {          // async copy of r_layer from the scratch into tileR
        const float* src = scr + (size_t)layer * VT::ROWS * VT::RS;
        for (int i = XW_TID * 4; i < VT::ROWS * VT::RS; i += XW_BDIM * 4) XW_CP_ASYNC16(tileR + i, src + i);
    };

    for (long long tix = XW_BID; tix < ntiles; tix += XW_GDIM) {
        const long long p0 = tix * NPT;
        if (XW_TID < NPT) {
            const long long p = p0 + XW_TID;
            const long long n = p / L;
            rown[XW_TID] = p < npts ? (int)n : -1;
            rowl[XW_TID] = (int)(p - n * L);
        }
        XW_SYNCTHREADS();
        for (int e = XW_TID; e < NPT * C; e += XW_BDIM) {
            const unsigned r = (unsigned)e / (unsigned)C, c = (unsigned)e - r * (unsigned)C;
            const int n = rown[r], l = rowl[r];
            float val = 0.f;
            if (n >= 0)
                val = c == 0 ? a.p.t[(long long)n * a.p.t_sn + l * a.p.t_sl]
                             : a.p.x[(long long)n * a.p.x_sn + l * a.p.x_sl + (c - 1)];
            xin[r * XLD + c] = val;
        }
        XW_SYNCTHREADS();
        // ---------------------------------------------------------------- forward (recompute)
        fpair acc[QR][TN];
        {
            float h0[QR][TN][2];
            float b[TN];
            load_row<TN>(sw + VT::BI + grp * VT::TNP, b);
#pragma unroll
            for (int q = 0; q < QR; ++q)
#pragma unroll
                for (int o = 0; o < TN; ++o) { h0[q][o][0] = b[o]; h0[q][o][1] = b[o]; }
            const float* wrow = sw + VT::WIT + grp * VT::TNP;
            for (int c = 0; c < C; ++c) {
                XW_FENCE();
                float w[TN];
                load_row<TN>(wrow + c * VT::WLD, w);
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    const float x0 = xin[(2 * (lane + 32 * q)) * XLD + c], x1 = xin[(2 * (lane + 32 * q) + 1) * XLD + c];
#pragma unroll
                    for (int o = 0; o < TN; ++o) { h0[q][o][0] = fmaf(x0, w[o], h0[q][o][0]); h0[q][o][1] = fmaf(x1, w[o], h0[q][o][1]); }
                }
            }
#pragma unroll
            for (int q = 0; q < QR; ++q)
#pragma unroll
                for (int o = 0; o < TN; ++o) acc[q][o] = pack2(h0[q][o][0], h0[q][o][1]);
        }
        for (int layer = 0; layer < nv; ++layer) {
            // r_layer = relu(h_layer): into the working tile and into the scratch (same layout)
            float* sdst = scr + (size_t)layer * VT::ROWS * VT::RS;
#pragma unroll
            for (int q = 0; q < QR; ++q) {
                const int off = (lane + 32 * q) * VT::RS + 2 * VT::GP * grp;
#pragma unroll
                for (int o = 0; o + 1 < TN; o += 2) {
                    float v0, t0, v1, t1;
                    unpack2(acc[q][o], v0, t0);
                    unpack2(acc[q][o + 1], v1, t1);
                    f4 v;
                    v.x = fmaxf(v0, 0.f); v.y = fmaxf(t0, 0.f); v.z = fmaxf(v1, 0.f); v.w = fmaxf(t1, 0.f);
                    st4(tileD + off + 2 * o, v);
                    st4(sdst + off + 2 * o, v);
                }
                if (TN & 1) {
                    float v0, t0;
                    unpack2(acc[q][TN - 1], v0, t0);
                    f2 v;
                    v.x = fmaxf(v0, 0.f); v.y = fmaxf(t0, 0.f);
                    *reinterpret_cast<f2*>(tileD + off + 2 * (TN - 1)) = v;
                    *reinterpret_cast<f2*>(sdst + off + 2 * (TN - 1)) = v;
                }
            }
            XW_SYNCTHREADS();
            tile_hidden_layer<HV, QR, true>(sw, tileD, grp, lane, acc);
            XW_SYNCTHREADS();
        }
        // ---------------------------------------------------------------- output layer, cotangent G
        float tau[QR][TN][2];
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            float pv0 = 0.f, pv1 = 0.f;
#pragma unroll
            for (int o = 0; o < TN; ++o) {
                float h0_, h1_;
                unpack2(acc[q][o], h0_, h1_);
                tau[q][o][0] = tanh_fast(h0_);
                tau[q][o][1] = tanh_fast(h1_);
                pv0 = fmaf(wz[o], tau[q][o][0], pv0);
                pv1 = fmaf(wz[o], tau[q][o][1], pv1);
            }
            f2 out; out.x = pv0; out.y = pv1;
            *reinterpret_cast<f2*>(vred + grp * NPT + 2 * (lane + 32 * q)) = out;
        }
        XW_SYNCTHREADS();
        if (XW_TID < NPT) {
            const int r = XW_TID;
            float G = 0.f;
            if (rown[r] >= 0) {
                float v = sw[VT::MISC];
#pragma unroll
                for (int g2 = 0; g2 < VT::NG; ++g2) v += vred[g2 * NPT + r];
                const float* xr = xin + r * XLD;
                const DomW W = domain_w(a.dom_kind, a.dp0, a.dp1, a.dp2, xr[0], xr + 1, a.d);
                G = fmaf(k0, a.cot[p0 + r], fmaf(k1, v, k2 * W.w));
            }
            gpt[r] = G;
            gbz += G;
        }
        if (nv > 0) issue_r_copy(nv - 1);          // tileR is free: prefetch r_{nv-1}
        XW_SYNCTHREADS();
        // delta_nv = G * Wz * (1 - tau^2);  dWz += G * tau
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            const f2 G2 = *reinterpret_cast<const f2*>(gpt + 2 * (lane + 32 * q));
#pragma unroll
            for (int o = 0; o < TN; ++o) {
                const float t0 = tau[q][o][0], t1 = tau[q][o][1];
                gwz[o] = fmaf(G2.x, t0, fmaf(G2.y, t1, gwz[o]));
                acc[q][o] = pack2(G2.x * wz[o] * (1.f - t0 * t0), G2.y * wz[o] * (1.f - t1 * t1));
            }
        }
        // ---------------------------------------------------------------- reverse sweep
        for (int k = nv; k >= 0; --k) {
            // delta_k -> tile D
#pragma unroll
            for (int q = 0; q < QR; ++q) {
                const int off = (lane + 32 * q) * VT::RS + 2 * VT::GP * grp;
#pragma unroll
                for (int o = 0; o + 1 < TN; o += 2) {
                    float v0, t0, v1, t1;
                    unpack2(acc[q][o], v0, t0);
                    unpack2(acc[q][o + 1], v1, t1);
                    f4 v; v.x = v0; v.y = t0; v.z = v1; v.w = t1;
                    st4(tileD + off + 2 * o, v);
                }
                if (TN & 1) {
                    float v0, t0;
                    unpack2(acc[q][TN - 1], v0, t0);
                    f2 v; v.x = v0; v.y = t0;
                    *reinterpret_cast<f2*>(tileD + off + 2 * (TN - 1)) = v;
                }
            }
            if (k == 0) break;
            XW_CP_ASYNC_WAIT_ALL();
            XW_SYNCTHREADS();                      // D complete, r_{k-1} landed in R
            // relu masks of r_{k-1} for this thread's outputs of the R-op
            unsigned long long mask = 0ull;
#pragma unroll
            for (int q = 0; q < QR; ++q) {
                const float* rr = tileR + (lane + 32 * q) * VT::RS + 2 * VT::GP * grp;
#pragma unroll
                for (int o = 0; o < TN; ++o) {
                    const f2 v = *reinterpret_cast<const f2*>(rr + 2 * o);
                    mask |= (unsigned long long)(v.x > 0.f ? 1u : 0u) << (q * 2 * TN + 2 * o);
                    mask |= (unsigned long long)(v.y > 0.f ? 1u : 0u) << (q * 2 * TN + 2 * o + 1);
                }
            }
            // P-op: dWh[o][i] += sum_cols D[col][o] R[col][i]  (this warp's quarter of the rows)
            {
                const int r0 = grp * (VT::ROWS / VT::NG);
                const float* dbase = tileD + 2 * VT::GP * p_ob;
                const float* rbase = tileR + p_rpos;
#pragma unroll 2
                for (int rr = 0; rr < VT::ROWS / VT::NG; ++rr) {
                    const float* dr = dbase + (r0 + rr) * VT::RS;
                    const float* rw = rbase + (r0 + rr) * VT::RS;
                    float dv[TN][2];
#pragma unroll
                    for (int o = 0; o + 1 < TN; o += 2) {
                        const f4 v = ld4(dr + 2 * o);
                        dv[o][0] = v.x; dv[o][1] = v.y; dv[o + 1][0] = v.z; dv[o + 1][1] = v.w;
                    }
                    if (TN & 1) {
                        const f2 v = ld2(dr + 2 * (TN - 1));
                        dv[TN - 1][0] = v.x; dv[TN - 1][1] = v.y;
                    }
#pragma unroll
                    for (int j = 0; j < IB; ++j) {
                        f2 rv = ld2(rw + 2 * j);
                        if (j == 4 && p_ib == 7) { rv.x = 1.f; rv.y = 1.f; }      // bias column
#pragma unroll
                        for (int o = 0; o < TN; ++o) pacc[o][j] = fmaf(dv[o][0], rv.x, fmaf(dv[o][1], rv.y, pacc[o][j]));
                    }
                }
            }
            XW_SYNCTHREADS();                      // everyone is done with tile R
            if (k >= 2) issue_r_copy(k - 2);
            // R-op: delta_{k-1} = relu'(r_{k-1}) . Wh^T delta_k
            {
                XW_FENCE();
#pragma unroll
                for (int q = 0; q < QR; ++q)
#pragma unroll
                    for (int o = 0; o < TN; ++o) acc[q][o] = pack2(0.f, 0.f);
                tile_contract<HV, QR>(swr + grp * VT::TNP, tileD + lane * VT::RS, acc);
#pragma unroll
                for (int q = 0; q < QR; ++q)
#pragma unroll
                    for (int o = 0; o < TN; ++o) {
                        float d0, d1;
                        unpack2(acc[q][o], d0, d1);
                        const bool m0 = (mask >> (q * 2 * TN + 2 * o)) & 1ull, m1 = (mask >> (q * 2 * TN + 2 * o + 1)) & 1ull;
                        acc[q][o] = pack2(m0 ? d0 : 0.f, m1 ? d1 : 0.f);
                    }
            }
            XW_SYNCTHREADS();                      // everyone is done reading tile D
        }
        XW_SYNCTHREADS();                          // delta_0 complete in tile D
        // ---------------------------------------------------------------- input layer: dWi, dbi
        // warp grp owns the hidden units of group grp; lanes own input channels (C = bias column)
        for (int c0 = 0; c0 <= C; c0 += 32) {
            const int c = c0 + lane;
            float ga[TN];
#pragma unroll
            for (int o = 0; o < TN; ++o) ga[o] = 0.f;
            for (int r = 0; r < VT::ROWS; ++r) {
                const float* dr = tileD + r * VT::RS + 2 * VT::GP * grp;
                float x0 = 0.f, x1 = 0.f;
                if (c < C) { x0 = xin[(2 * r) * XLD + c]; x1 = xin[(2 * r + 1) * XLD + c]; }
                else if (c == C) { x0 = 1.f; x1 = 1.f; }
#pragma unroll
                for (int o = 0; o + 1 < TN; o += 2) {
                    const f4 v = ld4(dr + 2 * o);
                    ga[o] = fmaf(v.x, x0, fmaf(v.y, x1, ga[o]));
                    ga[o + 1] = fmaf(v.z, x0, fmaf(v.w, x1, ga[o + 1]));
                }
                if (TN & 1) {
                    const f2 v = ld2(dr + 2 * (TN - 1));
                    ga[TN - 1] = fmaf(v.x, x0, fmaf(v.y, x1, ga[TN - 1]));
                }
            }
            if (c <= C) {
#pragma unroll
                for (int o = 0; o < TN; ++o) {
                    const int ou = grp * TN + o;
                    if (ou < HV) gwi[ou * (C + 1) + c] += ga[o];
                }
            }
        }
        XW_SYNCTHREADS();
    }
    // ------------------------------------------------------------------------ write the CTA's partial
    float* img = tileD;                            // [P] image, reuse the tile memory
    for (int i = XW_TID; i < g.size; i += XW_BDIM) img[i] = 0.f;
    XW_SYNCTHREADS();
    {
        const int ob = p_ob;
#pragma unroll
        for (int o = 0; o < TN; ++o) {
            const int ou = ob * TN + o;
#pragma unroll
            for (int j = 0; j < IB; ++j) {
                const int slot = 7 * (p_ib & 1) + j;                 // unit index inside the group (0..13)
                const int iu = (p_ib >> 1) * TN + slot;
                if (ou < a.Hvr) {
                    if (p_ib == 7 && j == 4) XW_ATOMIC_ADD_F(img + g.bh + ou, pacc[o][j]);
                    else if (slot < TN && iu < a.Hvr && !(p_ib == 7 && j > 4)) XW_ATOMIC_ADD_F(img + g.Wh + ou * a.Hvr + iu, pacc[o][j]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < TN; ++o) {
            const float s = warp_sum(gwz[o]);
            const int ou = grp * TN + o;
            if (lane == 0 && ou < a.Hvr) XW_ATOMIC_ADD_F(img + g.Wz + ou, s);
        }
        const float sb = warp_sum(gbz);
        if (lane == 0) XW_ATOMIC_ADD_F(img + g.bz, sb);
    }
    XW_SYNCTHREADS();
    for (int e = XW_TID; e < a.Hvr * (C + 1); e += XW_BDIM) {
        const int o = e / (C + 1), c = e % (C + 1);
        if (c < C) img[g.Wi + o * C + c] = gwi[e];
        else img[g.bi + o] = gwi[e];
    }
    XW_SYNCTHREADS();
    for (int i = XW_TID; i < g.size; i += XW_BDIM) a.gpart[(size_t)XW_BID * g.size + i] = img[i];
}

}  // namespace xw

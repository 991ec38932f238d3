// xw_xnode3.cuh -- XNODE backward, generation 3: THREE warp roles per SM sub-partition.
//
// Same arithmetic as xw_xnode2.cuh (reduced state z = Wy y, shared layer in registers, FP32 FFMA2 gradient tasks); what
// changes is who does what.  The 2-role kernel (k_xnode2_bwd) ran ONE compute warp per sub-partition that re-evaluated a
// stage, reversed it and only then re-evaluated the next one, with the gradient warp idle 70 % of the time: one serial
// instruction stream at IPC 0.3, FMA pipe 42-48 % busy (profiles/r02b_*).  Here the three independent streams run side
// by side on every sub-partition (warp w, w+4, w+8 of a 384-thread CTA):
//   F (forward)  : re-evaluates the stages of step l from the stored state z_l -- it needs nothing from the reverse
//                  sweep -- and records relu outputs / tanh into an r-tile;
//   R (reverse)  : takes the r-tile, runs the reverse sweep of that stage and writes the per-layer cotangents into a
//                  delta-tile; carries the adjoint (zbar, qbar) from step to step;
//   G (gradient) : lane = (path group, task slot); 55 FFMA2 per task into register-resident 10 x 11 accumulators.
// Hand-offs are shared-memory mbarriers (two r-tiles and two delta-tiles per triple, phase parity per use).  Registers:
// the CTA is launched at 168 per thread (384 threads); G gives back to 144 and R takes 192 with setmaxnreg (the sum must stay 3 x 168).
// The forward sweep (state history) is NOT part of this kernel: the interior history comes from the interior forward
// kernel, the boundary history from a k_xnode2_fwd<.,0> launch in front of it.
#pragma once
#include "xw_xnode2.cuh"

namespace xw {
namespace x3 {

using x2::HH;
using x2::HHP;
using x2::R;
using x2::kRow;
using x2::kSlot;
using x2::kSlots;
using x2::kTile;
using x2::kZQ;
constexpr int kTriples = 4;
constexpr int kThreads = 96 * kTriples;
constexpr int kPartA = x2::kPartA;

// shared-memory image (floats)
struct Sm {
    static constexpr int RED = 0;                       // reduced image x2::R
    static constexpr int WST = RED + R::size;           // [HH][HHP]  Ws in-major (staging for the register copy)
    static constexpr int WSO = WST + HH * HHP;          // [HH][HHP]  Ws out-major: row o = Ws[o][:] (register copy of the R role)
    static constexpr int BS = WSO + HH * HHP;           // [HHP]
    static constexpr int WT = BS + HHP;                 // [HHP]  Wa[:, d]
    static constexpr int BA = WT + HHP;                 // [HHP]
    static constexpr int BO = BA + HHP;                 // [4]
    static constexpr int DV = BO + 4;                   // [kTriples][32][HHP]  per-lane dv (10) | de (1) accumulators of R
    static constexpr int GIMG = DV + kTriples * 32 * HHP;   // [kPartA]
    static constexpr int WXT = GIMG + kPartA;           // [d][HHP]   Wa[:, j] (x part), j-major
    static constexpr int fixed = WXT;
};

struct Args {
    int d, Hr, HHr, nsh, L, n;
    const float* theta; const float* x; long long x_sn; const float* times;
    const float* cot;            // MODE 0: cot_u[n*L]   MODE 1: g[n*L]
    const double* coefs;         // MODE 0: device k0,k1,k2
    const float* hloss;          // MODE 0: func_h values of loss.init
    const float* zq;             // [L][zq_row<SOLVER>][n] reduced state history incl. stage inputs (required)
    double gscale;               // MODE 1
    float* perpath;              // [x2::kPerPath][n]
    float* partA;                // [gridDim][kPartA]
    double* sums;
};

// weights straight from the flat parameter vector (reference named_parameters() order, [out][in])
XW_DEV void stage(float* sm, const float* XW_RESTRICT th, int d, int Hr, int HHr) {
    const ULayout g(d, Hr, HHr);
    const int total = Sm::fixed + d * HHP;
    for (int i = XW_TID; i < total; i += XW_BDIM) sm[i] = 0.f;
    XW_SYNCTHREADS();
    float* sr = sm + Sm::RED;
    for (int e = XW_TID; e < HHr * HHr; e += XW_BDIM) {
        const int o = e / HHr, i = e % HHr;
        float m = 0.f;
        for (int k = 0; k < Hr; ++k) m = fmaf(th[g.Wa + o * g.lda + d + 1 + k], th[g.Wf + k * HHr + i], m);
        sr[R::MT + i * HHP + o] = m;
        sr[R::MO + o * HHP + i] = m;
        sm[Sm::WST + i * HHP + o] = th[g.Ws + o * HHr + i];
        sm[Sm::WSO + o * HHP + i] = th[g.Ws + o * HHr + i];
    }
    for (int o = XW_TID; o < HHr; o += XW_BDIM) {
        float c = 0.f, v = 0.f;
        for (int k = 0; k < Hr; ++k) {
            c = fmaf(th[g.Wa + o * g.lda + d + 1 + k], th[g.bf + k], c);
            v = fmaf(th[g.Wo + k], th[g.Wf + k * HHr + o], v);
        }
        sr[R::CV + o] = c;
        sr[R::VV + o] = v;
        sm[Sm::BS + o] = th[g.bs + o];
        sm[Sm::WT + o] = th[g.Wa + o * g.lda + d];
        sm[Sm::BA + o] = th[g.ba + o];
    }
    if (XW_TID == 0) {
        float e = 0.f;
        for (int k = 0; k < Hr; ++k) e = fmaf(th[g.Wo + k], th[g.bf + k], e);
        sr[R::EE] = e;
        sm[Sm::BO] = th[g.bo];
    }
    for (int e = XW_TID; e < HHr * d; e += XW_BDIM) {
        const int o = e / d, j = e % d;
        sm[Sm::WXT + j * HHP + o] = th[g.Wa + o * g.lda + j];
    }
    XW_SYNCTHREADS();
}

// The R role runs reverse passes only, so it keeps the shared layer in the TRANSPOSED pairing
// wT[o][ip] = (Ws[o][2ip], Ws[o][2ip+1]):  dn[i] = sum_o Ws[o][i] dl[o]  becomes, per o, five FFMA2 with dl[o] as the
// broadcast operand, and the packed accumulators ARE (dn[2ip], dn[2ip+1]) -- no final lo+hi adds; the relu mask is one
// packed multiply per pair and the next layer reads its broadcast scalars straight from the register halves.
struct CoreRegsT {
    fpair wT[HH][HH / 2];
};
XW_DEV void load_core_t(const float* wso, CoreRegsT& cr) {
#pragma unroll
    for (int o = 0; o < HH; ++o)
#pragma unroll
        for (int ip = 0; ip < HH / 2; ++ip) {
            const f2 v = ld2(wso + o * HHP + 2 * ip);
            cr.wT[o][ip] = pack2(v.x, v.y);
        }
}
XW_DEV void st_pairs10(float* p, const fpair (&d)[HH / 2]) {
    float v[HH];
#pragma unroll
    for (int ip = 0; ip < HH / 2; ++ip) unpack2(d[ip], v[2 * ip], v[2 * ip + 1]);
    x2::st_vec10(p, v, 0.f);
}
// d = cotangent of a_nsh (packed pairs) -> cotangent of a_0; delta_j goes to the delta-tile slot j-1, the mask of layer
// j-1 comes from its recorded output in the r-tile (loaded before the FFMA2 block)
XW_DEV void core_rev_t(const CoreRegsT& cr, fpair (&d)[HH / 2], int nsh, const float* rrow, float* drow) {
#pragma unroll 1
    for (int j = nsh; j > 0; --j) {
        st_pairs10(drow + kSlot * (j - 1), d);
        float rm[HH];
        x2::ld_vec10(rrow + kSlot * (j - 1), rm);
        fpair acc[HH / 2];
#pragma unroll
        for (int ip = 0; ip < HH / 2; ++ip) acc[ip] = pack2(0.f, 0.f);
#pragma unroll
        for (int op = 0; op < HH / 2; ++op) {
            float d0, d1;
            unpack2(d[op], d0, d1);
            const fpair b0 = pack2(d0, d0), b1 = pack2(d1, d1);
#pragma unroll
            for (int ip = 0; ip < HH / 2; ++ip) acc[ip] = fma2(cr.wT[2 * op][ip], b0, acc[ip]);
#pragma unroll
            for (int ip = 0; ip < HH / 2; ++ip) acc[ip] = fma2(cr.wT[2 * op + 1][ip], b1, acc[ip]);
        }
#pragma unroll
        for (int ip = 0; ip < HH / 2; ++ip) {
            const fpair m = pack2(rm[2 * ip] > 0.f ? 1.f : 0.f, rm[2 * ip + 1] > 0.f ? 1.f : 0.f);
            d[ip] = fma2(acc[ip], m, pack2(0.f, 0.f));
        }
    }
}

template <int SOLVER, int MODE>
XW_GLOBAL void XW_LAUNCH_BOUNDS(kThreads, 1) k_xnode3_bwd(Args a) {
    using T = Tableau<SOLVER>;
    XW_DYN_SMEM(smem_raw);
    float* sm = reinterpret_cast<float*>(smem_raw);
    float* st = sm + pad4(Sm::fixed + a.d * HHP);
    float* tiles = st + pad4(a.L) + 4;                                  // [kTriples][4][kTile]: r0, r1, d0, d1
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(tiles + (size_t)kTriples * 4 * kTile);   // [kTriples][8]
    double* red = reinterpret_cast<double*>(mbar + kTriples * 8);
    stage(sm, a.theta, a.d, a.Hr, a.HHr);
    for (int i = XW_TID; i < a.L; i += XW_BDIM) st[i] = a.times[i];
    if (XW_TID < kTriples * 8) XW_MBAR_INIT(mbar + XW_TID, 32);
    XW_SYNCTHREADS();

    const float* sr = sm + Sm::RED;
    const int warp = XW_TID >> 5, lane = XW_TID & 31;
    const int tri = warp & (kTriples - 1), role = warp / kTriples;      // 0 F, 1 R, 2 G
    float* rt0 = tiles + tri * (4 * kTile);                             // r-tile b = rt0 + b kTile; delta-tile b = rt0 + (2 + b) kTile
    unsigned long long* mb = mbar + tri * 8;
    unsigned long long *full_r = mb, *empty_r = mb + 2, *full_d = mb + 4, *empty_d = mb + 6;
    const int L = a.L, nsh = a.nsh;
    constexpr int ROW = x2::zq_row<SOLVER>();
    const int cpaths = 32 * kTriples;
    const int nchunks = (a.n + cpaths - 1) / cpaths;
    double bd_acc = 0.0;

    if (role == 0) {
        // ================================================================================== F: forward re-evaluation
        x2::CoreRegs cr;
        x2::load_core_from(sm + Sm::WST, sm + Sm::BS, cr);
        unsigned e = 0;
        for (int c = XW_BID; c < nchunks; c += XW_GDIM) {
            const long long nraw = (long long)c * cpaths + tri * 32 + lane;
            const long long n = nraw < a.n ? nraw : (long long)a.n - 1;
            const float* xp = a.x + n * a.x_sn;
            float ax[HH];
            load_row<HH>(sm + Sm::BA, ax);
            for (int j = 0; j < a.d; ++j) {
                float w[HH];
                load_row<HH>(sm + Sm::WXT + j * HHP, w);
                const float xj = xp[j];
#pragma unroll
                for (int o = 0; o < HH; ++o) ax[o] = fmaf(w[o], xj, ax[o]);
            }
            const float* hb = a.zq + n;
            for (int l = L - 2; l >= 0; --l) {
                const float t0 = st[l], dt = st[l + 1] - st[l];
                // stage inputs straight from the history the forward kernel kept (z_l and the z_in of stages 1..S-1):
                // no unrecorded evaluation here
                float zin[T::S][HH];
#pragma unroll
                for (int i = 0; i < HH; ++i) zin[0][i] = hb[((long long)l * ROW + i) * a.n];
#pragma unroll
                for (int s = 1; s < T::S; ++s)
#pragma unroll
                    for (int i = 0; i < HH; ++i) zin[s][i] = hb[((long long)l * ROW + kZQ + (s - 1) * HH + i) * a.n];
                // recorded evaluations in the order the reverse sweep consumes them: last stage first
#pragma unroll
                for (int s = T::S - 1; s >= 0; --s) {
                    const unsigned b = e & 1u, u = e >> 1;
                    if (u >= 1) XW_MBAR_WAIT(empty_r + b, (u - 1) & 1u);           // G is done with the previous use of this tile
                    x2::TileRec rec;
                    rec.rrow = rt0 + (b ? kTile : 0) + lane * kRow;
                    rec.drow = nullptr;
                    float av[HH], tau[HH];
                    x2::stage_input_from(sm + Sm::WT, ax, fmaf(T::c(s), dt, t0), zin[s], av);
                    x2::core_fwd(cr, av, nsh, tau, rec);
                    XW_MBAR_ARRIVE(full_r + b);
                    ++e;
                }
            }
        }
    } else if (role == 1) {
        // ================================================================================== R: reverse sweep
        XW_SETMAXNREG_INC(192);     // 168 (F) + 192 (R) + 144 (G) = 3 x 168: the pool is what the CTA was launched with
        CoreRegsT cr;
        load_core_t(sm + Sm::WSO, cr);
        const float bo = sm[Sm::BO];
        float k0 = 0.f, k1 = 0.f, k2 = 0.f;
        if (MODE == 0) { k0 = (float)a.coefs[0]; k1 = (float)a.coefs[1]; k2 = (float)a.coefs[2]; }
        const float gsc2 = (float)(2.0 * a.gscale);
        float* dvrow = sm + Sm::DV + (tri * 32 + lane) * HHP;           // this lane's dv[10] | de accumulators
        unsigned e = 0;
        for (int c = XW_BID; c < nchunks; c += XW_GDIM) {
            const long long nraw = (long long)c * cpaths + tri * 32 + lane;
            const bool active = nraw < a.n;
            const long long n = active ? nraw : (long long)a.n - 1;
            const float* hb = a.zq + n;
            auto cot_at = [&](int l, float u) -> float {
                if (!active) return 0.f;
                if (MODE == 0) {
                    float G = fmaf(k0, a.cot[n * L + l], k2);
                    if (l == 0) G = fmaf(k1, u - a.hloss[n], G);
                    return G;
                } else {
                    const float r = u - a.cot[n * L + l];
                    bd_acc += (double)(r * r);
                    return gsc2 * r;
                }
            };
            float zb[HH], a0[HH], a0t[HH];
#pragma unroll
            for (int i = 0; i < HH; ++i) { zb[i] = 0.f; a0[i] = 0.f; a0t[i] = 0.f; }
            float qb = cot_at(L - 1, hb[((long long)(L - 1) * ROW + HH) * a.n] + bo);
            for (int l = L - 2; l >= 0; --l) {
                const float t0 = st[l], dt = st[l + 1] - st[l];
                const float ql = hb[((long long)l * ROW + HH) * a.n];
                float kbar[T::S][HH], zacc[HH];
#pragma unroll
                for (int s = 0; s < T::S; ++s)
#pragma unroll
                    for (int i = 0; i < HH; ++i) kbar[s][i] = (T::b(s) * dt) * zb[i];
#pragma unroll
                for (int i = 0; i < HH; ++i) zacc[i] = zb[i];
#pragma unroll
                for (int s = T::S - 1; s >= 0; --s) {
                    const unsigned b = e & 1u, u = e >> 1;
                    const float ts = fmaf(T::c(s), dt, t0);
                    x2::TileRec rec;
                    rec.rrow = rt0 + (b ? kTile : 0) + lane * kRow;
                    rec.drow = rt0 + (2 + b) * kTile + lane * kRow;
                    XW_MBAR_WAIT(full_r + b, u & 1u);                               // F has recorded this evaluation
                    float tau[HH];
                    x2::ld_vec10(rec.rrow + kSlot * (kSlots - 1), tau);
                    const float pbar = (T::b(s) * dt) * qb;
                    if (T::b(s) != 0.f) {                                           // dv += pbar tau ; de += pbar
                        float dv[HH];
                        x2::ld_vec10(dvrow, dv);
#pragma unroll
                        for (int i = 0; i < HH; ++i) dv[i] = fmaf(pbar, tau[i], dv[i]);
                        x2::st_vec10(dvrow, dv, dvrow[HH] + pbar);
                    }
                    if (u >= 1) XW_MBAR_WAIT(empty_d + b, (u - 1) & 1u);           // G is done with the previous use of this tile
                    x2::st_vec10(rec.drow + kSlot * (kSlots - 1), kbar[s], 0.f);
                    float dl[HH];
                    {
                        // cotangent of tanh's input: (M^T kbar + v pbar) (1 - tau^2), then the shared-layer stack in reverse
                        float v[HH];
                        load_row<HH>(sr + R::VV, v);
#pragma unroll
                        for (int i = 0; i < HH; ++i) dl[i] = v[i] * pbar;
                        matvec_acc<HH, HH, HHP>(sr + R::MO, kbar[s], dl);
                        fpair dp[HH / 2];
#pragma unroll
                        for (int ip = 0; ip < HH / 2; ++ip)
                            dp[ip] = pack2(dl[2 * ip] * fmaf(-tau[2 * ip], tau[2 * ip], 1.f),
                                           dl[2 * ip + 1] * fmaf(-tau[2 * ip + 1], tau[2 * ip + 1], 1.f));
                        core_rev_t(cr, dp, nsh, rec.rrow, rec.drow);
#pragma unroll
                        for (int ip = 0; ip < HH / 2; ++ip) unpack2(dp[ip], dl[2 * ip], dl[2 * ip + 1]);
                    }
                    XW_MBAR_ARRIVE(full_d + b);
                    ++e;
#pragma unroll
                    for (int i = 0; i < HH; ++i) { zacc[i] += dl[i]; a0[i] += dl[i]; a0t[i] = fmaf(ts, dl[i], a0t[i]); }
#pragma unroll
                    for (int r = 0; r < s; ++r) {
                        const float cc = T::a(s, r);
                        if (cc != 0.f) {
                            const float cd = cc * dt;
#pragma unroll
                            for (int i = 0; i < HH; ++i) kbar[r][i] = fmaf(cd, dl[i], kbar[r][i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < HH; ++i) zb[i] = zacc[i];
                qb += cot_at(l, ql + bo);
            }
            if (active) {
#pragma unroll
                for (int i = 0; i < HH; ++i) {
                    a.perpath[(long long)i * a.n + n] = a0[i];
                    a.perpath[(long long)(HH + i) * a.n + n] = a0t[i];
                    a.perpath[(long long)(2 * HH + i) * a.n + n] = zb[i];
                }
                a.perpath[(long long)(3 * HH) * a.n + n] = qb;
            }
        }
        // dv | de: per-lane rows -> warp sum -> CTA image
        {
            float dv[HH];
            x2::ld_vec10(dvrow, dv);
            const float de = dvrow[HH];
#pragma unroll
            for (int i = 0; i < HH; ++i) {
                const float s = warp_sum(dv[i]);
                if (lane == 0) XW_ATOMIC_ADD_F(sm + Sm::GIMG + 2 * HH * (HH + 1) + i, s);
            }
            const float s = warp_sum(de);
            if (lane == 0) XW_ATOMIC_ADD_F(sm + Sm::GIMG + 2 * HH * (HH + 1) + HH, s);
        }
    } else {
        // ================================================================================== G: gradient tasks
        XW_SETMAXNREG_DEC(144);
        const int slot = lane & 7, pg = (lane >> 3) * 8;
        const bool valid = slot == kSlots - 1 || slot < nsh;
        fpair acc[HH / 2][HH + 1];
#pragma unroll
        for (int op = 0; op < HH / 2; ++op)
#pragma unroll
            for (int i = 0; i <= HH; ++i) acc[op][i] = pack2(0.f, 0.f);
        int my_chunks = 0;
        for (int c = XW_BID; c < nchunks; c += XW_GDIM) ++my_chunks;
        const unsigned total = (unsigned)my_chunks * (unsigned)((L - 1) * T::S);
        for (unsigned e = 0; e < total; ++e) {
            const unsigned b = e & 1u, u = e >> 1;
            XW_MBAR_WAIT(full_d + b, u & 1u);                                       // R (and, before it, F) have written this evaluation
            const float* rbase = rt0 + (b ? kTile : 0) + pg * kRow + slot * kSlot;
            const float* dbase = rt0 + (2 + b) * kTile + pg * kRow + slot * kSlot;
#pragma unroll 2
            for (int rho = 0; rho < 8; ++rho) {
                const f4 d0 = ld4(dbase + rho * kRow), d1 = ld4(dbase + rho * kRow + 4);
                const f2 d2 = ld2(dbase + rho * kRow + 8);
                const f4 r0 = ld4(rbase + rho * kRow), r1 = ld4(rbase + rho * kRow + 4), r2 = ld4(rbase + rho * kRow + 8);
                if (valid) {
                    const fpair dp[HH / 2] = {pack2(d0.x, d0.y), pack2(d0.z, d0.w), pack2(d1.x, d1.y), pack2(d1.z, d1.w),
                                              pack2(d2.x, d2.y)};
                    const float rv[HH + 1] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z};
#pragma unroll
                    for (int i = 0; i <= HH; ++i) {
                        const fpair rr = pack2(rv[i], rv[i]);
#pragma unroll
                        for (int op = 0; op < HH / 2; ++op) acc[op][i] = fma2(dp[op], rr, acc[op][i]);
                    }
                }
            }
            XW_MBAR_ARRIVE(empty_r + b);
            XW_MBAR_ARRIVE(empty_d + b);
        }
        if (valid) {
            float* img = sm + Sm::GIMG + (slot == kSlots - 1 ? HH * (HH + 1) : 0);
#pragma unroll
            for (int op = 0; op < HH / 2; ++op)
#pragma unroll
                for (int i = 0; i <= HH; ++i) {
                    float lo, hi;
                    unpack2(acc[op][i], lo, hi);
                    XW_ATOMIC_ADD_F(img + (2 * op) * (HH + 1) + i, lo);
                    XW_ATOMIC_ADD_F(img + (2 * op + 1) * (HH + 1) + i, hi);
                }
        }
    }
    XW_SYNCTHREADS();
    for (int e = XW_TID; e < kPartA; e += XW_BDIM) a.partA[(size_t)XW_BID * kPartA + e] = sm[Sm::GIMG + e];
    if (MODE == 1) {
        double v[1] = {bd_acc};
        const int idx[1] = {5};
        block_sum_to_global<1>(v, red, a.sums, idx);
    }
}

}  // namespace x3
}  // namespace xw

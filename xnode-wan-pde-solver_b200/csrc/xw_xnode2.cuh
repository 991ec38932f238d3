// xw_xnode2.cuh -- generation 2 of the XNODE kernels (primal net u_theta; reference src/model.py:87-112,
// 133-141,153-156 + the fixed-grid RK schemes of torchdiffeq 0.1.1, call site src/model.py:103-106).
//
// Three ideas:
//  (1) REDUCED STATE.  The field MLP sees the ODE state y (H = 20) only through Wy*y, and the output only
//      through Wo*y.  With  z = Wy y (hh = 10),  q = Wo y,  M = Wy Wf,  c = Wy bf,  v = Wo Wf,  e = Wo bf:
//          a_0 = ax + wt t + z_in ;  a_j = Ws relu(a_{j-1}) + bs ;  tau = tanh(a_nsh)
//          kappa = M tau + c  (= Wy F) ;  p = v.tau + e  (= Wo F)
//          z' = z + dt sum_s b_s kappa_s ;  q' = q + dt sum_s b_s p_s ;  u_l = q_l + bo
//      which is the same arithmetic in exact numbers (fp32 re-association only), at 810 instead of 1100 MACs
//      per field evaluation and half the live state.  Parameter gradients are accumulated for the reduced
//      quantities (dM, dc, dv, de, dWs, dbs) and mapped back to (Wy, Wf, bf, Wo) once per launch
//      (k_xnode2_finish): dWf = Wy^T dM + Wo^T dv, dWy = dM Wf^T + dc bf^T + sum_paths zbar_0 y_0^T, ...
//  (2) The shared hh x hh layer (7 of the 9 layers of an evaluation) lives in REGISTERS as 50 packed pairs:
//      no shared-memory traffic for 700 of the 810 MACs (generation 1 was LSU bound: one broadcast LDS.128
//      per 4 FMAs of a single path).
//  (3) WARP-SPECIALISED BACKWARD.  Compute warps (one path per lane) re-evaluate a stage, run its reverse
//      sweep and stream the per-layer pairs (delta_j, relu(a_{j-1})) through shared-memory tiles to gradient
//      warps, whose lanes each own ONE 10 x 11 outer-product accumulator block in registers for the whole
//      kernel (55 packed FFMA2 per task, no staging, no mma.sync hi/lo splits, no read-modify-write).
//      Hand-offs are named barriers between warp c and warp c+4 (same SM sub-partition).
#pragma once
#include "xw_kernels.cuh"

#ifdef XW_EMU
#define XW_LAUNCH_BOUNDS(t, b)
#else
#define XW_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#endif

namespace xw {
namespace x2 {

constexpr int H = 20, HH = 10, HHP = 12;
using S = USmem<H, HH>;

// reduced weight image (floats, relative to its base)
struct R {
    static constexpr int MT = 0;                 // [HH][HHP]  MT[i][o] = M[o][i]   (kappa = M tau)
    static constexpr int MO = MT + HH * HHP;     // [HH][HHP]  MO[o][i] = M[o][i]   (taubar = M^T kappabar)
    static constexpr int CV = MO + HH * HHP;     // [HHP] c
    static constexpr int VV = CV + HHP;          // [HHP] v
    static constexpr int EE = VV + HHP;          // [4]   e
    static constexpr int size = EE + 4;
};

// gradient-task tiles: one row per path, 8 slots of 12 floats (+4 floats: rows 4 banks apart)
constexpr int kSlot = 12, kSlots = 8, kRow = kSlots * kSlot + 4, kTile = 32 * kRow;
constexpr int kCW = 4;                           // compute warps per CTA (= gradient warps)
constexpr int kBwdThreads = 64 * kCW;
constexpr int kPartA = 232;                      // [S: 10 x 11][M: 10 x 11][dv 10][de 1] (+1 pad)
constexpr int kPerPath = 31;                     // a0[10], a0t[10], zbar0[10], qbar0
constexpr int kZQ = 11;                          // history row: z[10], q, then the inputs z_in of stages 1..S-1 of the step
template <int SOLVER> constexpr int zq_row() { return kZQ + HH * (Tableau<SOLVER>::S - 1); }   // floats per grid point

XW_DEV void stage_reduced(float* sr, const float* su) {
    for (int i = XW_TID; i < R::size; i += XW_BDIM) sr[i] = 0.f;
    XW_SYNCTHREADS();
    for (int e = XW_TID; e < HH * HH; e += XW_BDIM) {
        const int o = e / HH, i = e % HH;
        float m = 0.f;
        for (int k = 0; k < H; ++k) m = fmaf(su[S::WY + o * S::HP + k], su[S::WF + k * S::HHP + i], m);
        sr[R::MT + i * HHP + o] = m;
        sr[R::MO + o * HHP + i] = m;
    }
    for (int o = XW_TID; o < HH; o += XW_BDIM) {
        float c = 0.f, v = 0.f;
        for (int k = 0; k < H; ++k) {
            c = fmaf(su[S::WY + o * S::HP + k], su[S::BF + k], c);
            v = fmaf(su[S::WO + k], su[S::WF + k * S::HHP + o], v);
        }
        sr[R::CV + o] = c;
        sr[R::VV + o] = v;
    }
    if (XW_TID == 0) {
        float e = 0.f;
        for (int k = 0; k < H; ++k) e = fmaf(su[S::WO + k], su[S::BF + k], e);
        sr[R::EE] = e;
    }
    XW_SYNCTHREADS();
}

// the shared layer in registers: w[i][jp] = (Ws[2jp][i], Ws[2jp+1][i]), b[jp] = (bs[2jp], bs[2jp+1])
struct CoreRegs {
    fpair w[HH][HH / 2];
    fpair b[HH / 2];
};
// wst: [HH][HHP] in-major image of Ws (wst[i][o] = Ws[o][i]); bs: [HHP]
XW_DEV void load_core_from(const float* wst, const float* bs, CoreRegs& cr) {
#pragma unroll
    for (int i = 0; i < HH; ++i)
#pragma unroll
        for (int jp = 0; jp < HH / 2; ++jp) {
            const f2 v = ld2(wst + i * HHP + 2 * jp);
            cr.w[i][jp] = pack2(v.x, v.y);
        }
#pragma unroll
    for (int jp = 0; jp < HH / 2; ++jp) {
        const f2 v = ld2(bs + 2 * jp);
        cr.b[jp] = pack2(v.x, v.y);
    }
}
XW_DEV void load_core(const float* su, CoreRegs& cr) { load_core_from(su + S::WST, su + S::BS, cr); }

XW_DEV void st_vec10(float* p, const float (&v)[HH], float v10) {
    st4(p, f4{v[0], v[1], v[2], v[3]});
    st4(p + 4, f4{v[4], v[5], v[6], v[7]});
    st4(p + 8, f4{v[8], v[9], v10, 0.f});
}
XW_DEV void ld_vec10(const float* p, float (&v)[HH]) {
    const f4 a = ld4(p), b = ld4(p + 4);
    const f2 c = ld2(p + 8);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y;
}

// recorders of the internals of one evaluation of the shared-layer stack
struct NoRec {
    XW_DEV void layer(int, const float (&)[HH]) {}
    XW_DEV void tanh_out(const float (&)[HH]) {}
    XW_DEV void delta(int, const float (&)[HH]) {}
    XW_DEV void mask_load(int, float (&)[HH]) {}
    XW_DEV void mask(const float (&)[HH], const float (&dn)[HH], float (&dl)[HH]) {
#pragma unroll
        for (int i = 0; i < HH; ++i) dl[i] = dn[i];
    }
};
struct BitsRec {                       // relu masks (bit set = active) as a bit stack + tanh outputs
    BitStack128 m;
    float tau[HH];
    XW_DEV void layer(int, const float (&r)[HH]) {
        unsigned mm = 0u;
#pragma unroll
        for (int i = 0; i < HH; ++i) mm = __funnelshift_l(0u - __float_as_uint(r[i]), mm, 1);   // r >= 0: bit = (r > 0)
        m.template push<HH>(mm & ((1u << HH) - 1u));
    }
    XW_DEV void tanh_out(const float (&t)[HH]) {
#pragma unroll
        for (int i = 0; i < HH; ++i) tau[i] = t[i];
    }
    XW_DEV void delta(int, const float (&)[HH]) {}
    XW_DEV void mask_load(int, float (&)[HH]) {}
    XW_DEV void mask(const float (&)[HH], const float (&dn)[HH], float (&dl)[HH]) {
        const unsigned mm = m.template pop<HH>();
#pragma unroll
        for (int i = 0; i < HH; ++i) dl[i] = ((mm >> (HH - 1 - i)) & 1u) ? dn[i] : 0.f;
    }
};
struct TileRec {                       // gradient-task tiles: r_j -> slot j, tanh -> slot 7, delta_j -> slot j-1
    float* rrow;
    float* drow;
    XW_DEV void layer(int j, const float (&r)[HH]) { st_vec10(rrow + kSlot * j, r, 1.f); }
    XW_DEV void tanh_out(const float (&t)[HH]) { st_vec10(rrow + kSlot * (kSlots - 1), t, 1.f); }
    XW_DEV void delta(int j, const float (&dl)[HH]) { st_vec10(drow + kSlot * (j - 1), dl, 0.f); }
    // the relu mask of layer j-1 comes from its recorded output; loaded BEFORE the layer's FFMA2 block, used after it
    XW_DEV void mask_load(int jm1, float (&r)[HH]) { ld_vec10(rrow + kSlot * jm1, r); }
    XW_DEV void mask(const float (&r)[HH], const float (&dn)[HH], float (&dl)[HH]) {
#pragma unroll
        for (int i = 0; i < HH; ++i) dl[i] = r[i] > 0.f ? dn[i] : 0.f;
    }
};

// a (first pre-activation) -> tau = tanh(a_nsh); a is destroyed
template <class Rec>
XW_DEV void core_fwd(const CoreRegs& cr, float (&a)[HH], int nsh, float (&tau)[HH], Rec& rec) {
#pragma unroll 1
    for (int j = 0; j < nsh; ++j) {
        float r[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) r[i] = fmaxf(a[i], 0.f);
        rec.layer(j, r);
        fpair acc[HH / 2];
#pragma unroll
        for (int jp = 0; jp < HH / 2; ++jp) acc[jp] = cr.b[jp];
#pragma unroll
        for (int i = 0; i < HH; ++i) {
            const fpair rr = pack2(r[i], r[i]);
#pragma unroll
            for (int jp = 0; jp < HH / 2; ++jp) acc[jp] = fma2(cr.w[i][jp], rr, acc[jp]);
        }
#pragma unroll
        for (int jp = 0; jp < HH / 2; ++jp) unpack2(acc[jp], a[2 * jp], a[2 * jp + 1]);
    }
#pragma unroll
    for (int i = 0; i < HH; ++i) tau[i] = tanh_fast(a[i]);
    rec.tanh_out(tau);
}

// dl = cotangent of a_nsh  ->  dl = cotangent of a_0 (delta^0); the recorder sees every delta_j (j = nsh..1)
template <class Rec>
XW_DEV void core_rev(const CoreRegs& cr, float (&dl)[HH], int nsh, Rec& rec) {
#pragma unroll 1
    for (int j = nsh; j > 0; --j) {
        rec.delta(j, dl);
        float rm[HH];
        rec.mask_load(j - 1, rm);
        fpair acc[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) acc[i] = pack2(0.f, 0.f);
#pragma unroll
        for (int jp = 0; jp < HH / 2; ++jp) {
            const fpair d2 = pack2(dl[2 * jp], dl[2 * jp + 1]);
#pragma unroll
            for (int i = 0; i < HH; ++i) acc[i] = fma2(cr.w[i][jp], d2, acc[i]);
        }
        float dn[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) {
            float lo, hi;
            unpack2(acc[i], lo, hi);
            dn[i] = lo + hi;
        }
        rec.mask(rm, dn, dl);
    }
}

// kappa = M tau + c ;  p = v.tau + e
XW_DEV void kappa_of(const float* sr, const float (&tau)[HH], float (&kap)[HH]) {
    load_row<HH>(sr + R::CV, kap);
    matvec_acc<HH, HH, HHP>(sr + R::MT, tau, kap);
}
XW_DEV float p_of(const float* sr, const float (&tau)[HH]) {
    float v[HH];
    load_row<HH>(sr + R::VV, v);
    float p0 = sr[R::EE], p1 = 0.f;
#pragma unroll
    for (int i = 0; i < HH; i += 2) { p0 = fmaf(v[i], tau[i], p0); p1 = fmaf(v[i + 1], tau[i + 1], p1); }
    return p0 + p1;
}
// first pre-activation of a stage: a = ax + wt t + zin
XW_DEV void stage_input_from(const float* wtp, const float (&ax)[HH], float t, const float (&zin)[HH], float (&a)[HH]) {
    float wt[HH];
    load_row<HH>(wtp, wt);
#pragma unroll
    for (int i = 0; i < HH; ++i) a[i] = fmaf(wt[i], t, ax[i]) + zin[i];
}
XW_DEV void stage_input(const float* su, const float (&ax)[HH], float t, const float (&zin)[HH], float (&a)[HH]) {
    stage_input_from(su + S::WT, ax, t, zin, a);
}

// one explicit RK step of the reduced state (z, q), recording stage internals in rec[s]
template <int SOLVER, class Rec>
XW_DEV void rk_step_red(const CoreRegs& cr, const float* su, const float* sr, const float (&ax)[HH], float t0, float dt,
                        int nsh, float (&z)[HH], float& q, Rec (&rec)[Tableau<SOLVER>::S], float* zin_out = nullptr,
                        long long zin_stride = 0) {
    using T = Tableau<SOLVER>;
    float kap[T::S][HH];
    float pq = 0.f;
#pragma unroll
    for (int s = 0; s < T::S; ++s) {
        float zin[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) zin[i] = z[i];
#pragma unroll
        for (int r = 0; r < s; ++r) {
            const float c = T::a(s, r);
            if (c != 0.f) {
                const float cd = c * dt;
#pragma unroll
                for (int i = 0; i < HH; ++i) zin[i] = fmaf(cd, kap[r][i], zin[i]);
            }
        }
        if (s > 0 && zin_out) {                 // stage inputs kept for the backward (it then needs no unrecorded pass)
#pragma unroll
            for (int i = 0; i < HH; ++i) zin_out[(long long)((s - 1) * HH + i) * zin_stride] = zin[i];
        }
        float a[HH], tau[HH];
        stage_input(su, ax, fmaf(T::c(s), dt, t0), zin, a);
        core_fwd(cr, a, nsh, tau, rec[s]);
        kappa_of(sr, tau, kap[s]);
        if (T::b(s) != 0.f) pq = fmaf(T::b(s), p_of(sr, tau), pq);
    }
#pragma unroll
    for (int s = 0; s < T::S; ++s) {
        const float c = T::b(s);
        if (c != 0.f) {
            const float cd = c * dt;
#pragma unroll
            for (int i = 0; i < HH; ++i) z[i] = fmaf(cd, kap[s][i], z[i]);
        }
    }
    q = fmaf(dt, pq, q);
}

// lift + reduction of the initial state: z0 = Wy y0, q0 = Wo y0 (no bo)
XW_DEV void lift_reduced(const float* su, float s0, float (&z)[HH], float& q) {
    const WSmem sw = WSmem::make(su);
    float z1[H], z2[H], y0[H];
    lift_fwd<H, HH>(sw, s0, z1, z2, y0);
#pragma unroll
    for (int i = 0; i < HH; ++i) z[i] = 0.f;
    matvec_acc<H, HH, S::HHP>(su + S::WYT, y0, z);
    float wo[H];
    load_row<H>(su + S::WO, wo);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < H; i += 2) { q0 = fmaf(wo[i], y0[i], q0); q1 = fmaf(wo[i + 1], y0[i + 1], q1); }
    q = q0 + q1;
}

// reverse of one stage in the reduced variables.  kbar = cotangent of kappa_s, pbar = cotangent of p_s, tau = the
// stage's tanh outputs.  Returns delta^0 (cotangent of the stage's first pre-activation = of z_in) in dl.
template <class Rec>
XW_DEV void stage_rev(const CoreRegs& cr, const float* sr, const float (&kbar)[HH], float pbar, const float (&tau)[HH],
                      int nsh, float (&dl)[HH], Rec& rec) {
    float v[HH];
    load_row<HH>(sr + R::VV, v);
#pragma unroll
    for (int i = 0; i < HH; ++i) dl[i] = v[i] * pbar;
    matvec_acc<HH, HH, HHP>(sr + R::MO, kbar, dl);
#pragma unroll
    for (int i = 0; i < HH; ++i) dl[i] *= fmaf(-tau[i], tau[i], 1.f);
    core_rev(cr, dl, nsh, rec);
}

// =============================================================================================
// forward.  MODE 0: u only.  MODE 1: interior forward (u, du = grad_x sum_l u, init sum, optional (z, q) history)
// =============================================================================================
struct FwdArgs {
    int d, Hr, HHr, nsh, L, n;
    const float* theta; const float* x; long long x_sn; const float* times; const float* s0;
    float* u_out;
    const float* grad_h; float* du_out; float* rec; double* sums; const float* hloss;
    float* zq;                   // optional: [L][kZQ][n] reduced state history kept for the backward kernels
};
constexpr int kFwdThreads = 256;     // (384 threads at 168 registers measured 10 % slower: 4.2 vs 3.8 ms)
constexpr int kRecWords2 = 4 + HH;

template <int SOLVER, int MODE>
XW_GLOBAL void XW_LAUNCH_BOUNDS(kFwdThreads, 1) k_xnode2_fwd(FwdArgs a) {
    using T = Tableau<SOLVER>;
    XW_DYN_SMEM(smem_raw);
    float* su = reinterpret_cast<float*>(smem_raw);
    float* sr = su + pad4(S::size(a.d));
    float* st = sr + R::size;
    double* red = reinterpret_cast<double*>(st + pad4(a.L) + 4);
    stage_theta_u<H, HH>(su, a.theta, a.d, a.Hr, a.HHr);
    stage_reduced(sr, su);
    for (int i = XW_TID; i < a.L; i += XW_BDIM) st[i] = a.times[i];
    XW_SYNCTHREADS();
    CoreRegs cr;
    load_core(su, cr);
    const WSmem sw = WSmem::make(su);

    const long long nthr = (long long)XW_GDIM * XW_BDIM;
    const long long gtid = (long long)XW_BID * XW_BDIM + XW_TID;
    const int L = a.L, nsh = a.nsh;
    constexpr int ROW = zq_row<SOLVER>();
    const float bo = su[S::BO];
    double init_acc = 0.0;
    for (long long n = gtid; n < a.n; n += nthr) {
        const float* xp = a.x + n * a.x_sn;
        float ax[HH];
        hoist_ax<H, HH>(sw, xp, a.d, ax);
        const float s0 = a.s0[n];
        float z[HH], q;
        lift_reduced(su, s0, z, q);
        if (a.u_out) a.u_out[n * L] = q + bo;
        if (MODE == 1) { const float hd = (q + bo) - a.hloss[n]; init_acc += (double)(hd * hd); }
        if (a.zq) {
#pragma unroll
            for (int i = 0; i < HH; ++i) a.zq[(long long)i * a.n + n] = z[i];
            a.zq[(long long)HH * a.n + n] = q;
        }
        for (int l = 0; l + 1 < L; ++l) {
            const float t0 = st[l], dt = st[l + 1] - st[l];
            if (MODE == 1) {
                BitsRec rec[T::S];
#pragma unroll
                for (int s = 0; s < T::S; ++s) rec[s].m.clear();
                rk_step_red<SOLVER>(cr, su, sr, ax, t0, dt, nsh, z, q, rec,
                                    a.zq ? a.zq + ((long long)l * ROW + kZQ) * a.n + n : nullptr, a.n);
#pragma unroll
                for (int s = 0; s < T::S; ++s) {
                    float* hp = a.rec + ((long long)(l * T::S + s) * kRecWords2) * nthr + gtid;
                    hp[0] = __uint_as_float((unsigned)rec[s].m.lo);
                    hp[nthr] = __uint_as_float((unsigned)(rec[s].m.lo >> 32));
                    hp[2 * nthr] = __uint_as_float((unsigned)rec[s].m.hi);
                    hp[3 * nthr] = __uint_as_float((unsigned)(rec[s].m.hi >> 32));
#pragma unroll
                    for (int i = 0; i < HH; ++i) hp[(4 + i) * nthr] = rec[s].tau[i];
                }
            } else {
                NoRec rec[T::S];
                rk_step_red<SOLVER>(cr, su, sr, ax, t0, dt, nsh, z, q, rec,
                                    a.zq ? a.zq + ((long long)l * ROW + kZQ) * a.n + n : nullptr, a.n);
            }
            if (a.zq) {
#pragma unroll
                for (int i = 0; i < HH; ++i) a.zq[((long long)(l + 1) * ROW + i) * a.n + n] = z[i];
                a.zq[((long long)(l + 1) * ROW + HH) * a.n + n] = q;
            }
            if (a.u_out) a.u_out[n * L + l + 1] = q + bo;
        }
        if (MODE == 1) {
            // reverse sweep with cotangent 1 on every u_l: zb = d sum_l u_l / d z_l, qb = number of later outputs
            float zb[HH], a0[HH];
#pragma unroll
            for (int i = 0; i < HH; ++i) { zb[i] = 0.f; a0[i] = 0.f; }
            float qb = 1.f;
            for (int l = L - 2; l >= 0; --l) {
                const float dt = st[l + 1] - st[l];
                BitsRec rec[T::S];
#pragma unroll
                for (int s = 0; s < T::S; ++s) {
                    const float* hp = a.rec + ((long long)(l * T::S + s) * kRecWords2) * nthr + gtid;
                    rec[s].m.lo = (unsigned long long)__float_as_uint(hp[0]) | ((unsigned long long)__float_as_uint(hp[nthr]) << 32);
                    rec[s].m.hi = (unsigned long long)__float_as_uint(hp[2 * nthr]) | ((unsigned long long)__float_as_uint(hp[3 * nthr]) << 32);
#pragma unroll
                    for (int i = 0; i < HH; ++i) rec[s].tau[i] = hp[(4 + i) * nthr];
                }
                float kbar[T::S][HH], zacc[HH];
#pragma unroll
                for (int s = 0; s < T::S; ++s)
#pragma unroll
                    for (int i = 0; i < HH; ++i) kbar[s][i] = (T::b(s) * dt) * zb[i];
#pragma unroll
                for (int i = 0; i < HH; ++i) zacc[i] = zb[i];
#pragma unroll
                for (int s = T::S - 1; s >= 0; --s) {
                    float dl[HH];
                    stage_rev(cr, sr, kbar[s], (T::b(s) * dt) * qb, rec[s].tau, nsh, dl, rec[s]);
#pragma unroll
                    for (int i = 0; i < HH; ++i) { zacc[i] += dl[i]; a0[i] += dl[i]; }
#pragma unroll
                    for (int r = 0; r < s; ++r) {
                        const float c = T::a(s, r);
                        if (c != 0.f) {
                            const float cd = c * dt;
#pragma unroll
                            for (int i = 0; i < HH; ++i) kbar[r][i] = fmaf(cd, dl[i], kbar[r][i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < HH; ++i) zb[i] = zacc[i];
                qb += 1.f;
            }
            // cotangent of y0, then the lift reverse down to s0 (s0 = h(x) depends on x through grad_h)
            float lam[H];
            {
                float wo[H];
                load_row<H>(su + S::WO, wo);
#pragma unroll
                for (int i = 0; i < H; ++i) lam[i] = wo[i] * qb;
                matvec_acc<HH, H, S::HP>(su + S::WY, zb, lam);
            }
            float z1[H], z2[H], y0[H];
            lift_fwd<H, HH>(sw, s0, z1, z2, y0);
            float dz2[H], dz1[H];
#pragma unroll
            for (int i = 0; i < H; ++i) { dz2[i] = 0.f; dz1[i] = 0.f; }
            matvec_acc<H, H, S::HP>(su + S::W2, lam, dz2);
#pragma unroll
            for (int i = 0; i < H; ++i) dz2[i] = z2[i] > 0.f ? dz2[i] : 0.f;
            matvec_acc<H, H, S::HP>(su + S::W1, dz2, dz1);
            float gs = 0.f;
#pragma unroll
            for (int i = 0; i < H; ++i) gs = fmaf(z1[i] > 0.f ? dz1[i] : 0.f, su[S::W0 + i], gs);
            for (int j = 0; j < a.d; ++j) {
                float w[HH];
                load_row<HH>(su + S::WXT + j * S::HHP, w);
                float g = gs * a.grad_h[n * a.d + j];
#pragma unroll
                for (int o = 0; o < HH; ++o) g = fmaf(w[o], a0[o], g);
                a.du_out[n * a.d + j] = g;
            }
        }
    }
    if (MODE == 1) {
        double v[1] = {init_acc};
        const int idx[1] = {4};
        block_sum_to_global<1>(v, red, a.sums, idx);
    }
}

// =============================================================================================
// backward: parameter gradients of sum_l G[n,l] u[n,l] in the reduced variables (+ per-path quantities for
// k_xnode2_lift).   MODE 0 interior: G = k0*cot_u + k2 + [l=0] k1 (u0 - h)     (G_u)
//                   MODE 1 boundary: G = 2*gscale*(u - g), sums[BDRY] += sum (u-g)^2   (src/loss.py:83-85)
// =============================================================================================
struct BwdArgs {
    int d, Hr, HHr, nsh, L, n;
    const float* theta; const float* x; long long x_sn; const float* times; const float* s0;
    const float* cot;            // MODE 0: cot_u[n*L]   MODE 1: g[n*L]
    const double* coefs;         // MODE 0: device k0,k1,k2
    const float* hloss;          // MODE 0: func_h values of loss.init
    const float* zq;             // optional: [L][kZQ][n] reduced state history written by the forward kernel
    double gscale;               // MODE 1
    float* hist;                 // [L][zq_row][gridDim*32*kCW] scratch, used when zq == nullptr
    float* perpath;              // [kPerPath][n]
    float* partA;                // [gridDim][kPartA]
    double* sums;
};

template <int SOLVER, int MODE>
XW_GLOBAL void XW_LAUNCH_BOUNDS(kBwdThreads, 1) k_xnode2_bwd(BwdArgs a) {
    using T = Tableau<SOLVER>;
    XW_DYN_SMEM(smem_raw);
    float* su = reinterpret_cast<float*>(smem_raw);
    float* sr = su + pad4(S::size(a.d));
    float* st = sr + R::size;
    float* tiles = st + pad4(a.L) + 4;                     // [kCW][3][kTile]: r-tile 0, r-tile 1, delta-tile
    float* gimg = tiles + (size_t)kCW * 3 * kTile;         // [kPartA]
    double* red = reinterpret_cast<double*>(gimg + kPartA);
    stage_theta_u<H, HH>(su, a.theta, a.d, a.Hr, a.HHr);
    stage_reduced(sr, su);
    for (int i = XW_TID; i < a.L; i += XW_BDIM) st[i] = a.times[i];
    for (int i = XW_TID; i < kPartA; i += XW_BDIM) gimg[i] = 0.f;
    XW_SYNCTHREADS();

    const int warp = XW_TID >> 5, lane = XW_TID & 31;
    const int pair = warp & (kCW - 1);
    const bool is_compute = warp < kCW;
    float* rt0 = tiles + pair * (3 * kTile);               // r-tile 0; r-tile 1 = rt0 + kTile; delta tile = rt0 + 2 kTile
    float* dtile = rt0 + 2 * kTile;
    const int bar_full = 1 + 2 * pair, bar_empty = 2 + 2 * pair;
    const int L = a.L, nsh = a.nsh;
    constexpr int ROW = zq_row<SOLVER>();
    const int cpaths = 32 * kCW;
    const int nchunks = (a.n + cpaths - 1) / cpaths;
    int my_chunks = 0;
    for (int c = XW_BID; c < nchunks; c += XW_GDIM) ++my_chunks;
    double bd_acc = 0.0;

    if (is_compute) {
        CoreRegs cr;
        load_core(su, cr);
        const WSmem sw = WSmem::make(su);
        const float bo = su[S::BO];
        float k0 = 0.f, k1 = 0.f, k2 = 0.f;
        if (MODE == 0) { k0 = (float)a.coefs[0]; k1 = (float)a.coefs[1]; k2 = (float)a.coefs[2]; }
        const float gsc2 = (float)(2.0 * a.gscale);
        const long long nthr_c = (long long)XW_GDIM * cpaths;
        const long long ctid = (long long)XW_BID * cpaths + pair * 32 + lane;
        float dv[HH], de = 0.f;
#pragma unroll
        for (int i = 0; i < HH; ++i) dv[i] = 0.f;
        unsigned ecount = 0;
        for (int c = XW_BID; c < nchunks; c += XW_GDIM) {
            const long long nraw = (long long)c * cpaths + pair * 32 + lane;
            const bool active = nraw < a.n;
            const long long n = active ? nraw : (long long)a.n - 1;
            const float* xp = a.x + n * a.x_sn;
            float ax[HH];
            hoist_ax<H, HH>(sw, xp, a.d, ax);
            const bool have_hist = a.zq != nullptr;
            const float* hbase = have_hist ? a.zq + n : a.hist + ctid;
            const long long hstr = have_hist ? (long long)a.n : nthr_c;
            if (!have_hist) {              // integrate forward, keeping the reduced state of every grid point
                float z[HH], q;
                lift_reduced(su, a.s0[n], z, q);
                float* hw = a.hist + ctid;
#pragma unroll
                for (int i = 0; i < HH; ++i) hw[(long long)i * hstr] = z[i];
                hw[(long long)HH * hstr] = q;
                for (int l = 0; l + 1 < L; ++l) {
                    NoRec rec[T::S];
                    rk_step_red<SOLVER>(cr, su, sr, ax, st[l], st[l + 1] - st[l], nsh, z, q, rec);
#pragma unroll
                    for (int i = 0; i < HH; ++i) hw[((long long)(l + 1) * ROW + i) * hstr] = z[i];
                    hw[((long long)(l + 1) * ROW + HH) * hstr] = q;
                }
            }
            // cotangent of u[n, l]
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
            float qb = cot_at(L - 1, hbase[((long long)(L - 1) * ROW + HH) * hstr] + bo);
            for (int l = L - 2; l >= 0; --l) {
                const float t0 = st[l], dt = st[l + 1] - st[l];
                float zl[HH];
#pragma unroll
                for (int i = 0; i < HH; ++i) zl[i] = hbase[((long long)l * ROW + i) * hstr];
                const float ql = hbase[((long long)l * ROW + HH) * hstr];
                // stage inputs zin[s] (forward through the stages; only the LAST stage records its internals: earlier
                // stages are re-evaluated right before their own reverse -> 2S-1 evaluations of the layer stack)
                float zin[T::S][HH];
                TileRec trec;
                trec.rrow = rt0 + ((ecount & 1u) ? kTile : 0) + lane * kRow;
                trec.drow = dtile + lane * kRow;
                float tau_last[HH];
                {
                    float kap[T::S][HH];
#pragma unroll
                    for (int s = 0; s < T::S; ++s) {
#pragma unroll
                        for (int i = 0; i < HH; ++i) zin[s][i] = zl[i];
#pragma unroll
                        for (int r = 0; r < s; ++r) {
                            const float c = T::a(s, r);
                            if (c != 0.f) {
                                const float cd = c * dt;
#pragma unroll
                                for (int i = 0; i < HH; ++i) zin[s][i] = fmaf(cd, kap[r][i], zin[s][i]);
                            }
                        }
                        float av[HH];
                        stage_input(su, ax, fmaf(T::c(s), dt, t0), zin[s], av);
                        if (s + 1 < T::S) {
                            NoRec none;
                            float tau[HH];
                            core_fwd(cr, av, nsh, tau, none);
                            kappa_of(sr, tau, kap[s]);
                        } else {
                            core_fwd(cr, av, nsh, tau_last, trec);
                        }
                    }
                }
                float kbar[T::S][HH], zacc[HH];
#pragma unroll
                for (int s = 0; s < T::S; ++s)
#pragma unroll
                    for (int i = 0; i < HH; ++i) kbar[s][i] = (T::b(s) * dt) * zb[i];
#pragma unroll
                for (int i = 0; i < HH; ++i) zacc[i] = zb[i];
#pragma unroll
                for (int s = T::S - 1; s >= 0; --s) {
                    const float ts = fmaf(T::c(s), dt, t0);
                    float tau[HH];
                    if (s + 1 < T::S) {      // re-evaluate this stage, recording its internals in the other r-tile
                        trec.rrow = rt0 + ((ecount & 1u) ? kTile : 0) + lane * kRow;
                        float av[HH];
                        stage_input(su, ax, ts, zin[s], av);
                        core_fwd(cr, av, nsh, tau, trec);
                    } else {
#pragma unroll
                        for (int i = 0; i < HH; ++i) tau[i] = tau_last[i];
                    }
                    const float pbar = (T::b(s) * dt) * qb;
                    if (T::b(s) != 0.f) {
#pragma unroll
                        for (int i = 0; i < HH; ++i) dv[i] = fmaf(pbar, tau[i], dv[i]);
                        de += pbar;
                    }
                    XW_BAR_SYNC(bar_empty, 64);                       // the gradient warp is done with the delta tile
                    st_vec10(trec.drow + kSlot * (kSlots - 1), kbar[s], 0.f);
                    float dl[HH];
                    stage_rev(cr, sr, kbar[s], pbar, tau, nsh, dl, trec);
                    XW_BAR_ARRIVE(bar_full, 64);                      // tasks of this evaluation are complete
                    ++ecount;
#pragma unroll
                    for (int i = 0; i < HH; ++i) { zacc[i] += dl[i]; a0[i] += dl[i]; a0t[i] = fmaf(ts, dl[i], a0t[i]); }
#pragma unroll
                    for (int r = 0; r < s; ++r) {
                        const float c = T::a(s, r);
                        if (c != 0.f) {
                            const float cd = c * dt;
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
        // dv, de: per-lane accumulators -> warp sum -> CTA image
#pragma unroll
        for (int i = 0; i < HH; ++i) {
            const float s = warp_sum(dv[i]);
            if (lane == 0) XW_ATOMIC_ADD_F(gimg + 2 * HH * (HH + 1) + i, s);
        }
        {
            const float s = warp_sum(de);
            if (lane == 0) XW_ATOMIC_ADD_F(gimg + 2 * HH * (HH + 1) + HH, s);
        }
    } else {
        // gradient warp: lane = (path group of 8, slot).  Slots 0..nsh-1: dWs|dbs += delta_{j+1} (x) (r_j | 1);
        // slot 7: dM|dc += kappabar (x) (tau | 1).  Eight rounds per evaluation, one path of the group per round.
        const int slot = lane & 7, pg = (lane >> 3) * 8;
        const bool valid = slot == kSlots - 1 || slot < nsh;
        fpair acc[HH / 2][HH + 1];
#pragma unroll
        for (int op = 0; op < HH / 2; ++op)
#pragma unroll
            for (int i = 0; i <= HH; ++i) acc[op][i] = pack2(0.f, 0.f);
        const unsigned total = (unsigned)my_chunks * (unsigned)((L - 1) * T::S);
        if (total > 0) XW_BAR_ARRIVE(bar_empty, 64);
        for (unsigned e = 0; e < total; ++e) {
            XW_BAR_SYNC(bar_full, 64);
            const float* rbase = rt0 + ((e & 1u) ? kTile : 0) + pg * kRow + slot * kSlot;
            const float* dbase = dtile + pg * kRow + slot * kSlot;
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
            if (e + 1 < total) XW_BAR_ARRIVE(bar_empty, 64);
        }
        if (valid) {
            float* img = gimg + (slot == kSlots - 1 ? HH * (HH + 1) : 0);
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
    for (int e = XW_TID; e < kPartA; e += XW_BDIM) a.partA[(size_t)XW_BID * kPartA + e] = gimg[e];
    if (MODE == 1) {
        double v[1] = {bd_acc};
        const int idx[1] = {5};
        block_sum_to_global<1>(v, red, a.sums, idx);
    }
}

// =============================================================================================
// per-path-once part of the backward: initial-state, lift and x-column gradients from the per-path
// cotangents (a0, a0t, zbar0, qbar0) left by k_xnode2_bwd.  One thread per path.
// =============================================================================================
struct LiftArgs {
    int d, Hr, HHr, n;
    const float* theta; const float* x; long long x_sn; const float* s0; const float* perpath;
    float* partB;                // [gridDim][theta_u size]
};

XW_GLOBAL void k_xnode2_lift(LiftArgs a) {
    XW_DYN_SMEM(smem_raw);
    const int nwarps = XW_BDIM >> 5, warp = XW_TID >> 5, lane = XW_TID & 31;
    const ULayout g(a.d, a.Hr, a.HHr);
    const int Pp = pad4(g.size);
    float* su = reinterpret_cast<float*>(smem_raw);
    const WSmem sw = WSmem::make(su);
    float* sstg = su + pad4(S::size(a.d));                               // [nwarps][kStgRowsU][kStgLd]
    float* sgrad = sstg + (size_t)nwarps * kStgRowsU * kStgLd;           // [nwarps][Pp]
    stage_theta_u<H, HH>(su, a.theta, a.d, a.Hr, a.HHr);
    for (int i = XW_TID; i < nwarps * Pp; i += XW_BDIM) sgrad[i] = 0.f;
    XW_SYNCTHREADS();
    float* stg = sstg + (size_t)warp * kStgRowsU * kStgLd;
    float* gw = sgrad + (size_t)warp * Pp;
    const long long nthr = (long long)XW_GDIM * XW_BDIM;
    const long long gtid = (long long)XW_BID * XW_BDIM + XW_TID;
    float gwo[H], gbo = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) gwo[i] = 0.f;
    const long long iters = (a.n + nthr - 1) / nthr;
    for (long long it = 0; it < iters; ++it) {
        const long long nraw = it * nthr + gtid;
        const bool active = nraw < a.n;
        const long long n = active ? nraw : (long long)a.n - 1;
        const float* xp = a.x + n * a.x_sn;
        float a0[HH], a0t[HH], zb[HH], qb;
#pragma unroll
        for (int i = 0; i < HH; ++i) {
            a0[i] = active ? a.perpath[(long long)i * a.n + n] : 0.f;
            a0t[i] = active ? a.perpath[(long long)(HH + i) * a.n + n] : 0.f;
            zb[i] = active ? a.perpath[(long long)(2 * HH + i) * a.n + n] : 0.f;
        }
        qb = active ? a.perpath[(long long)(3 * HH) * a.n + n] : 0.f;
        const float s0 = a.s0[n];
        float z1[H], z2[H], y0[H];
        lift_fwd<H, HH>(sw, s0, z1, z2, y0);
        // first field layer: dWa[:, y] += zbar0 (x) y0 ; dWa[:, t] += a0t ; dba += a0 ; dWa[:, x] += a0 (x) x
        {
            float y1[H + 2];
#pragma unroll
            for (int i = 0; i < H; ++i) y1[i] = y0[i];
            y1[H] = 0.f; y1[H + 1] = 0.f;
            outer_auto<HH, H + 2, 32>(zb, y1, stg, [&](int o, int i) -> float* {
                if (o >= a.HHr || i >= a.Hr) return nullptr;
                return gw + g.Wa + o * g.lda + g.d + 1 + i;
            });
            float one[2] = {1.f, 0.f};
            outer_auto<HH, 2, 32>(a0t, one, stg, [&](int o, int i) -> float* {
                return (o < a.HHr && i == 0) ? gw + g.Wa + o * g.lda + g.d : nullptr;
            });
            outer_auto<HH, 2, 32>(a0, one, stg, [&](int o, int i) -> float* {
                return (o < a.HHr && i == 0) ? gw + g.ba + o : nullptr;
            });
        }
        warp_outer_dyn<HH, 3, 2>(a0, a.d, [&](int j) -> float { return xp[j]; }, stg, stg + 32 * kStgLd,
                              [&](int o, int j) -> float* { return o < a.HHr ? gw + g.Wa + o * g.lda + j : nullptr; });
        // final linear: dWo += qbar0 y0 ; dbo += qbar0
#pragma unroll
        for (int i = 0; i < H; ++i) gwo[i] = fmaf(qb, y0[i], gwo[i]);
        gbo += qb;
        // cotangent of y0 and the lift reverse
        float lam[H];
        {
            float wo[H];
            load_row<H>(su + S::WO, wo);
#pragma unroll
            for (int i = 0; i < H; ++i) lam[i] = wo[i] * qb;
            matvec_acc<HH, H, S::HP>(su + S::WY, zb, lam);
        }
        {
            float r1[H + 1];
#pragma unroll
            for (int i = 0; i < H; ++i) r1[i] = z2[i];
            r1[H] = 1.f;
            outer_auto<H, H + 1, 32>(lam, r1, stg, [&](int o, int i) -> float* {
                if (o >= a.Hr) return nullptr;
                if (i == H) return gw + g.b2 + o;
                return i < a.Hr ? gw + g.W2 + o * a.Hr + i : nullptr;
            });
        }
        float dz2[H], dz1[H];
#pragma unroll
        for (int i = 0; i < H; ++i) { dz2[i] = 0.f; dz1[i] = 0.f; }
        matvec_acc<H, H, S::HP>(su + S::W2, lam, dz2);
#pragma unroll
        for (int i = 0; i < H; ++i) dz2[i] = z2[i] > 0.f ? dz2[i] : 0.f;
        {
            float r1[H + 1];
#pragma unroll
            for (int i = 0; i < H; ++i) r1[i] = z1[i];
            r1[H] = 1.f;
            outer_auto<H, H + 1, 32>(dz2, r1, stg, [&](int o, int i) -> float* {
                if (o >= a.Hr) return nullptr;
                if (i == H) return gw + g.b1 + o;
                return i < a.Hr ? gw + g.W1 + o * a.Hr + i : nullptr;
            });
        }
        matvec_acc<H, H, S::HP>(su + S::W1, dz2, dz1);
#pragma unroll
        for (int i = 0; i < H; ++i) dz1[i] = z1[i] > 0.f ? dz1[i] : 0.f;
        {
            float r2[2] = {s0, 1.f};
            outer_auto<H, 2, 32>(dz1, r2, stg, [&](int o, int i) -> float* {
                if (o >= a.Hr) return nullptr;
                return i == 0 ? gw + g.W0 + o : gw + g.b0 + o;
            });
        }
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float sgi = warp_sum(gwo[i]);
        if (lane == 0 && i < a.Hr) gw[g.Wo + i] += sgi;
    }
    {
        const float sgb = warp_sum(gbo);
        if (lane == 0) gw[g.bo] += sgb;
    }
    XW_SYNCTHREADS();
    for (int e = XW_TID; e < g.size; e += XW_BDIM) {
        float sgr = 0.f;
        for (int w = 0; w < nwarps; ++w) sgr += sgrad[(size_t)w * Pp + e];
        a.partB[(size_t)XW_BID * g.size + e] = sgr;
    }
}

// =============================================================================================
// finish: sum the per-CTA partials (fp64) and map the reduced gradients back to the reference parameters:
//   dWs, dbs direct;  dWf = Wy^T dM + Wo^T dv;  dbf = Wy^T dc + Wo^T de;
//   dWy += dM Wf^T + dc bf^T;  dWo += dv Wf^T + de bf^T
// grad_out[e] = (accumulate ? grad_out[e] : 0) + ...
// =============================================================================================
XW_GLOBAL void k_xnode2_finish(const float* partA, int nA, const float* partB, int nB, const float* theta, int d,
                               int Hr, int HHr, float* out, int accumulate) {
    XW_DYN_SMEM(smem_raw);
    double* ra = reinterpret_cast<double*>(smem_raw);    // [kPartA]
    for (int e = XW_TID; e < kPartA; e += XW_BDIM) {
        double s = 0.0;
        for (int b = 0; b < nA; ++b) s += (double)partA[(size_t)b * kPartA + e];
        ra[e] = s;
    }
    XW_SYNCTHREADS();
    const ULayout g(d, Hr, HHr);
    const double* dS = ra;                               // [o][11]
    const double* dM = ra + HH * (HH + 1);               // [o][11] (col 10 = dc)
    const double* dvv = ra + 2 * HH * (HH + 1);          // [10], then de
    const double de = dvv[HH];
    const long long e = (long long)XW_BID * XW_BDIM + XW_TID;
    if (e >= g.size) return;
    double s = 0.0;
    for (int b = 0; b < nB; ++b) s += (double)partB[(size_t)b * g.size + e];
    auto Wy = [&](int o, int k) -> double { return (double)theta[g.Wa + o * g.lda + d + 1 + k]; };
    auto Wf = [&](int k, int i) -> double { return (double)theta[g.Wf + k * HHr + i]; };
    const int ei = (int)e;
    if (ei >= g.Ws && ei < g.bs) {
        const int o = (ei - g.Ws) / HHr, i = (ei - g.Ws) % HHr;
        s += dS[o * (HH + 1) + i];
    } else if (ei >= g.bs && ei < g.Wf) {
        s += dS[(ei - g.bs) * (HH + 1) + HH];
    } else if (ei >= g.Wf && ei < g.bf) {
        const int k = (ei - g.Wf) / HHr, i = (ei - g.Wf) % HHr;
        double t = (double)theta[g.Wo + k] * dvv[i];
        for (int o = 0; o < HHr; ++o) t += Wy(o, k) * dM[o * (HH + 1) + i];
        s += t;
    } else if (ei >= g.bf && ei < g.Wo) {
        const int k = ei - g.bf;
        double t = (double)theta[g.Wo + k] * de;
        for (int o = 0; o < HHr; ++o) t += Wy(o, k) * dM[o * (HH + 1) + HH];
        s += t;
    } else if (ei >= g.Wa && ei < g.ba) {
        const int o = (ei - g.Wa) / g.lda, col = (ei - g.Wa) % g.lda;
        if (col >= d + 1) {
            const int k = col - d - 1;
            double t = dM[o * (HH + 1) + HH] * (double)theta[g.bf + k];
            for (int i = 0; i < HHr; ++i) t += dM[o * (HH + 1) + i] * Wf(k, i);
            s += t;
        }
    } else if (ei >= g.Wo && ei < g.bo) {
        const int k = ei - g.Wo;
        double t = de * (double)theta[g.bf + k];
        for (int i = 0; i < HHr; ++i) t += dvv[i] * Wf(k, i);
        s += t;
    }
    out[ei] = (float)(s + (accumulate ? (double)out[ei] : 0.0));
}

}  // namespace x2
}  // namespace xw

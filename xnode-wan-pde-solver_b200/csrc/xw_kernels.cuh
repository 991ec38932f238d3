// xw_kernels.cuh -- the __global__ kernels of the hot path (generation 1: one thread per path /
// per point; see DESIGN.md for the roofline of each).  Compiles for sm_100a with nvcc and, for the
// CPU-side logic tests only, with g++ -DXW_EMU.
#pragma once
#include "xw_nets.cuh"

#ifdef XW_EMU
#define XW_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#else
#define XW_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace xw {

struct PointsView {
    const float* t; long long t_sn, t_sl;
    const float* x; long long x_sn, x_sl;
};

// block-level sum of NV doubles per thread -> atomicAdd into out[idx[k]]
template <int NV>
XW_DEV void block_sum_to_global(double (&v)[NV], double* red /* smem [NV][32] */, double* out, const int (&idx)[NV]) {
    const int lane = XW_TID & 31, warp = XW_TID >> 5, nw = XW_BDIM >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = warp_sum_d(v[k]);
        if (lane == 0) red[k * 32 + warp] = s;
    }
    XW_SYNCTHREADS();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = lane < nw ? red[k * 32 + lane] : 0.0;
            s = warp_sum_d(s);
            if (lane == 0 && out) XW_ATOMIC_ADD_D(out + idx[k], s);
        }
    }
    XW_SYNCTHREADS();
}

// =============================================================================================
// XNODE forward (+ the "ones" reverse sweep giving du = grad_x sum_l u)
// MODE 0: u only (forward evaluation)          MODE 1: interior forward (u, du, init sum)
// =============================================================================================
struct XnodeFwdArgs {
    int d, Hr, HHr, nsh, L, n;
    const float* theta; const float* x; long long x_sn; const float* times; const float* s0;
    float* u_out;
    const float* grad_h; float* du_out; float* yhist; double* sums;
    const float* hloss;          // func_h values used by loss.init (== s0 when the batch starts at T0)
    float* ypath;                // MODE 1, optional: [L][H][n] state history kept for the interior backward
};

template <int H, int HH, class W>
XW_DEV void hoist_ax(const W& sw, const float* XW_RESTRICT xp, int d, float (&ax)[HH]) {
    using S = USmem<H, HH>;
    load_row<HH>(sw.at(S::BA), ax);
    for (int j = 0; j < d; ++j) {
        float w[HH];
        load_row<HH>(sw.at(S::WXT + j * S::HHP), w);
        const float xj = xp[j];
#pragma unroll
        for (int o = 0; o < HH; ++o) ax[o] = fmaf(w[o], xj, ax[o]);
    }
}

template <int H, int HH, class W>
XW_DEV float project_u(const W& sw, const float (&y)[H]) {
    using S = USmem<H, HH>;
    float wo[H];
    load_row<H>(sw.at(S::WO), wo);
    float u0 = *sw.at(S::BO), u1 = 0.f;
#pragma unroll
    for (int i = 0; i + 1 < H; i += 2) { u0 = fmaf(wo[i], y[i], u0); u1 = fmaf(wo[i + 1], y[i + 1], u1); }
    if (H & 1) u0 = fmaf(wo[H - 1], y[H - 1], u0);
    return u0 + u1;
}

// one explicit RK step y <- y + dt * sum_s b_s k_s, recording stage internals in rec[s] and
// (optionally) the stage inputs in yin[s]
template <int H, int HH, int SOLVER, class Rec, bool KEEP_YIN, class W>
XW_DEV void rk_step(const W& sw, const float (&ax)[HH], float t0, float dt, int nsh, float (&y)[H],
                    Rec (&rec)[Tableau<SOLVER>::S], float (*yin_keep)[H]) {
    using T = Tableau<SOLVER>;
    float k[T::S][H];
#pragma unroll
    for (int s = 0; s < T::S; ++s) {
        float yin[H];
#pragma unroll
        for (int i = 0; i < H; ++i) yin[i] = y[i];
#pragma unroll
        for (int r = 0; r < s; ++r) {
            const float c = T::a(s, r);
            if (c != 0.f) {
                const float cd = c * dt;
#pragma unroll
                for (int i = 0; i < H; ++i) yin[i] = fmaf(cd, k[r][i], yin[i]);
            }
        }
        if (KEEP_YIN) {
#pragma unroll
            for (int i = 0; i < H; ++i) yin_keep[s][i] = yin[i];
        }
        float tau[HH];
        field_fwd<H, HH>(sw, ax, fmaf(T::c(s), dt, t0), yin, nsh, k[s], tau, rec[s]);
    }
#pragma unroll
    for (int s = 0; s < T::S; ++s) {
        const float c = T::b(s);
        if (c != 0.f) {
            const float cd = c * dt;
#pragma unroll
            for (int i = 0; i < H; ++i) y[i] = fmaf(cd, k[s][i], y[i]);
        }
    }
}

// words kept per field evaluation for the reverse "ones" sweep: 128 mask bits + HH tanh outputs
template <int HH> constexpr int kRecWords = 4 + HH;

template <int H, int HH, int SOLVER, int MODE, class WS = WSmem>
XW_GLOBAL void k_xnode_fwd(XnodeFwdArgs a) {
    using S = USmem<H, HH>;
    using T = Tableau<SOLVER>;
    XW_DYN_SMEM(smem_raw);
    float* swp = reinterpret_cast<float*>(smem_raw);
    float* st = swp + pad4(S::size(a.d));
    double* red = reinterpret_cast<double*>(st + pad4(a.L) + 4);
    const WS sw = WS::make(swp);
    if (WS::kStage) stage_theta_u<H, HH>(swp, a.theta, a.d, a.Hr, a.HHr);
    for (int i = XW_TID; i < a.L; i += XW_BDIM) st[i] = a.times[i];
    XW_SYNCTHREADS();

    const long long nthr = (long long)XW_GDIM * XW_BDIM;
    const long long gtid = (long long)XW_BID * XW_BDIM + XW_TID;
    const int L = a.L, nsh = a.nsh;
    double init_acc = 0.0;
    for (long long n = gtid; n < a.n; n += nthr) {
        const float* xp = a.x + n * a.x_sn;
        float ax[HH];
        hoist_ax<H, HH>(sw, xp, a.d, ax);
        const float s0 = a.s0[n];
        float y[H];
        {
            float z1[H], z2[H];
            lift_fwd<H, HH>(sw, s0, z1, z2, y);
        }
        float u = project_u<H, HH>(sw, y);
        if (a.u_out) a.u_out[n * L] = u;
        if (MODE == 1) {
            { const float hd = u - a.hloss[n]; init_acc += (double)(hd * hd); }
            if (a.ypath) {
#pragma unroll
                for (int i = 0; i < H; ++i) a.ypath[(long long)i * a.n + n] = y[i];
            }
        }
        for (int l = 0; l + 1 < L; ++l) {
            const float t0 = st[l], dt = st[l + 1] - st[l];
            if (MODE == 1) {
                // keep what the reverse "ones" sweep needs of every field evaluation (relu masks as
                // bits + tanh outputs: 4 + HH words per stage) so that it does not recompute the stages
                RecBits<HH> rec[T::S];
#pragma unroll
                for (int s = 0; s < T::S; ++s) rec[s].m.clear();
                rk_step<H, HH, SOLVER, RecBits<HH>, false>(sw, ax, t0, dt, nsh, y, rec, nullptr);
#pragma unroll
                for (int s = 0; s < T::S; ++s) {
                    float* hp = a.yhist + ((long long)(l * T::S + s) * kRecWords<HH>) * nthr + gtid;
                    hp[0] = __uint_as_float((unsigned)rec[s].m.lo);
                    hp[nthr] = __uint_as_float((unsigned)(rec[s].m.lo >> 32));
                    hp[2 * nthr] = __uint_as_float((unsigned)rec[s].m.hi);
                    hp[3 * nthr] = __uint_as_float((unsigned)(rec[s].m.hi >> 32));
#pragma unroll
                    for (int i = 0; i < HH; ++i) hp[(4 + i) * nthr] = rec[s].tau[i];
                }
            } else {
                RecNone<HH> rec[T::S];
                rk_step<H, HH, SOLVER, RecNone<HH>, false>(sw, ax, t0, dt, nsh, y, rec, nullptr);
            }
            u = project_u<H, HH>(sw, y);
            if (a.u_out) a.u_out[n * L + l + 1] = u;
            if (MODE == 1 && a.ypath) {
#pragma unroll
                for (int i = 0; i < H; ++i) a.ypath[((long long)(l + 1) * H + i) * a.n + n] = y[i];
            }
        }
        if (MODE == 1) {
            // reverse sweep with cotangent 1 on every u[l]: lam = d sum_l u_l / d y_l
            float lam[H], a0[HH];
            load_row<H>(sw.at(S::WO), lam);
#pragma unroll
            for (int i = 0; i < HH; ++i) a0[i] = 0.f;
            for (int l = L - 2; l >= 0; --l) {
                const float t0 = st[l], dt = st[l + 1] - st[l];
                RecBits<HH> rec[T::S];
#pragma unroll
                for (int s = 0; s < T::S; ++s) {
                    const float* hp = a.yhist + ((long long)(l * T::S + s) * kRecWords<HH>) * nthr + gtid;
                    rec[s].m.lo = (unsigned long long)__float_as_uint(hp[0]) | ((unsigned long long)__float_as_uint(hp[nthr]) << 32);
                    rec[s].m.hi = (unsigned long long)__float_as_uint(hp[2 * nthr]) | ((unsigned long long)__float_as_uint(hp[3 * nthr]) << 32);
#pragma unroll
                    for (int i = 0; i < HH; ++i) rec[s].tau[i] = hp[(4 + i) * nthr];
                }
                float kbar[T::S][H], ybar[H];
#pragma unroll
                for (int s = 0; s < T::S; ++s)
#pragma unroll
                    for (int i = 0; i < H; ++i) kbar[s][i] = (T::b(s) * dt) * lam[i];
#pragma unroll
                for (int i = 0; i < H; ++i) ybar[i] = lam[i];
#pragma unroll
                for (int s = T::S - 1; s >= 0; --s) {
                    float gin[H];
#pragma unroll
                    for (int i = 0; i < H; ++i) gin[i] = 0.f;
                    field_rev_bits<H, HH>(sw, rec[s], nsh, kbar[s], gin, a0);
#pragma unroll
                    for (int i = 0; i < H; ++i) ybar[i] += gin[i];
#pragma unroll
                    for (int r = 0; r < s; ++r) {
                        const float c = T::a(s, r);
                        if (c != 0.f) {
                            const float cd = c * dt;
#pragma unroll
                            for (int i = 0; i < H; ++i) kbar[r][i] = fmaf(cd, gin[i], kbar[r][i]);
                        }
                    }
                }
                float wo[H];
                load_row<H>(sw.at(S::WO), wo);
#pragma unroll
                for (int i = 0; i < H; ++i) lam[i] = ybar[i] + wo[i];
            }
            // lift reverse: d/ds0 (s0 = h(x) depends on x through grad_h)
            float z1[H], z2[H], y0[H];
            lift_fwd<H, HH>(sw, s0, z1, z2, y0);
            float dz2[H], dz1[H];
#pragma unroll
            for (int i = 0; i < H; ++i) { dz2[i] = 0.f; dz1[i] = 0.f; }
            matvec_acc<H, H, S::HP>(sw.at(S::W2), lam, dz2);
#pragma unroll
            for (int i = 0; i < H; ++i) dz2[i] = z2[i] > 0.f ? dz2[i] : 0.f;
            matvec_acc<H, H, S::HP>(sw.at(S::W1), dz2, dz1);
            float gs = 0.f;
#pragma unroll
            for (int i = 0; i < H; ++i) gs = fmaf(z1[i] > 0.f ? dz1[i] : 0.f, (*sw.at(S::W0 + i)), gs);
            for (int j = 0; j < a.d; ++j) {
                float w[HH];
                load_row<HH>(sw.at(S::WXT + j * S::HHP), w);
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
// v net, one thread per point.  MODE 0: v only (forward evaluation).  MODE 2: time-row 0 of every path: the
// a grad(phi).du + b.du phi term of src/loss.py:66-69 (used when the tensor-core row-0 kernel does not cover d)
// =============================================================================================
struct VnetFwdArgs {
    int d, Hvr, nv, n, L;
    const float* theta; PointsView p;
    int dom_kind; float dp0, dp1, dp2;
    float c0, c1; const float* ca; const float* cb;
    long long ca_sn, cb_sn;  // elements between consecutive paths' a / b (0: one constant matrix / vector)
    const float* u; const float* du; const float* h; const float* f;
    double* sums; float* cot_u; float* cot_v; float* v_out;
    float* gcache;           // MODE 2, optional: [n*d] grad_x phi at time-row 0 (for k_weak_combine)
};

template <int HV, int MODE>
XW_GLOBAL void k_vnet_points(VnetFwdArgs a) {
    using S = VSmem<HV>;
    XW_DYN_SMEM(smem_raw);
    float* sv = reinterpret_cast<float*>(smem_raw);
    double* red = reinterpret_cast<double*>(sv + pad4(S::size(a.d + 1)) + 4);
    stage_theta_v<HV>(sv, a.theta, a.d, a.Hvr);
    const long long nthr = (long long)XW_GDIM * XW_BDIM;
    const long long gtid = (long long)XW_BID * XW_BDIM + XW_TID;
    // MODE 2 visits time-row 0 of every path only (the grad_x phi . du term of src/loss.py:66-69)
    const long long npts = MODE == 2 ? (long long)a.n : (long long)a.n * a.L;
    const int L = a.L, d = a.d;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long it = gtid; it < npts; it += nthr) {
        const long long p = MODE == 2 ? it * L : it;
        const long long n = p / L;
        const int l = (int)(p - n * L);
        const float t = a.p.t[n * a.p.t_sn + l * a.p.t_sl];
        const float* xp = a.p.x + n * a.p.x_sn + l * a.p.x_sl;
        unsigned long long masks[kMaxNv];
        float tau[HV];
        const float v = vnet_fwd<HV>(sv, t, xp, d, a.nv, masks, tau, StoreNone());
        if (MODE == 0) { a.v_out[p] = v; continue; }
        float dl[HV];
        vnet_rev_bits<HV>(sv, a.nv, masks, tau, 1.f, dl);
        const DomW W = domain_w(a.dom_kind, a.dp0, a.dp1, a.dp2, t, xp, d);
        const float phi = v * W.w;
        if (MODE == 2) {
            const float* dun = a.du + n * d;
            float s31 = 0.f;
            for (int i = 0; i < d; ++i) {
                const float dphi_i = fmaf(W.w, vnet_input_grad<HV>(sv, 1 + i, dl), v * domain_dw_x(W, i, xp));
                if (a.gcache) a.gcache[n * d + i] = dphi_i;
                float q;
                if (a.ca) {
                    q = 0.f;
                    for (int j = 0; j < d; ++j) q = fmaf(a.ca[n * a.ca_sn + i * d + j], dun[j], q);
                } else {
                    q = dun[i];
                }
                s31 = fmaf(dphi_i, q, s31);
            }
            if (a.cb) {
                float bq = 0.f;
                for (int j = 0; j < d; ++j) bq = fmaf(a.cb[n * a.cb_sn + j], dun[j], bq);
                s31 = fmaf(phi, bq, s31);
            }
            acc[2] += (double)s31;
            continue;
        }
        static_assert(MODE == 0 || MODE == 2, "generation-1 interior pass (MODE 1) was removed");
    }
    if (MODE != 0) {
        const int idx[4] = {0, 1, 2, 3};
        block_sum_to_global<4>(acc, red, a.sums, idx);
    }
}

// =============================================================================================
// XNODE backward: parameter gradients of sum_l G[n,l] u[n,l]
// MODE 0 interior: G = k0*cot_u + k2 + [l=0] k1 (u0 - h)      (G_u)
// MODE 1 boundary: G = 2*gscale*(u - g), and sums[BDRY] += sum (u-g)^2   (src/loss.py:83-85)
// =============================================================================================
struct XnodeBwdArgs {
    int d, Hr, HHr, nsh, L, n;
    const float* theta; const float* x; long long x_sn; const float* times; const float* s0;
    const float* cot;            // MODE 0: cot_u[n*L]   MODE 1: g[n*L]
    const double* coefs;         // MODE 0: device k0,k1,k2
    const float* hloss;          // MODE 0: func_h values of loss.init
    const float* ypath;          // MODE 0, optional: [L][H][n] state history written by the forward kernel
    double gscale;               // MODE 1
    float* yhist; float* gpart; double* sums;
};

template <int O, int I>
struct OuterShape {
    static constexpr int NBI = I <= 2 ? I : (I <= 24 ? 6 : 8);
    static constexpr int BI = (I + NBI - 1) / NBI;
    static constexpr int NBO = 32 / NBI;
    static constexpr int BOfull = (O + NBO - 1) / NBO;
    static constexpr int BOcap = (49 / BI) < 1 ? 1 : (49 / BI);
    static constexpr int BO = BOfull < BOcap ? BOfull : BOcap;
    static constexpr int rows_d = ((O + NBO * BO - 1) / (NBO * BO)) * (NBO * BO);
    static constexpr int rows_r = NBI * BI;
};

// staging buffer of one warp: d-side rows [0, ROFF), r-side rows [ROFF, ...)
template <int O, int I, int ROFF, class Dst>
XW_DEV void outer_auto(const float (&dl)[O], const float (&r)[I], float* stg, Dst dst) {
    using Sh = OuterShape<O, I>;
    static_assert(Sh::rows_d <= ROFF, "d-side staging rows");
#ifndef XW_EMU
    // XNODE staging (32 + 24 rows): tensor-core fragments fit when O <= 32 and I <= 24
    if constexpr (ROFF == 32 && O <= 32 && I <= 24 && O >= 8) {
        warp_outer_mma<O, I>(dl, r, stg, stg + ROFF * kStgLd, dst);
        return;
    }
#endif
    warp_outer<O, I, Sh::BO, Sh::NBO, Sh::BI, Sh::NBI>(dl, r, stg, stg + ROFF * kStgLd, dst);
}
constexpr int kStgRowsU = 32 + 24;     // XNODE: d-side <= 32 rows, r-side <= 24 rows (warp_outer_dyn chunks of 16)
constexpr int kStgRowsV = 64 + 64;     // per-point v-net kernel

// reverse of one field evaluation with parameter gradients
template <int H, int HH>
XW_DEV void field_rev_grads(const WSmem& sw, const float* acts, int stride, int nsh, float tstage,
                            const float (&yin)[H], const float (&gout)[H], float (&gy)[H], float (&a0)[HH],
                            float* stg, float* gw, const ULayout& g) {
    using S = USmem<H, HH>;
    const int Hr = g.H, HHr = g.hh;
    float tau1[HH + 1];
#pragma unroll
    for (int i = 0; i < HH; ++i) tau1[i] = acts[(nsh * HH + i) * stride];
    tau1[HH] = 1.f;
    outer_auto<H, HH + 1, 32>(gout, tau1, stg, [&](int o, int i) -> float* {
        if (o >= Hr) return nullptr;
        if (i == HH) return gw + g.bf + o;
        return i < HHr ? gw + g.Wf + o * HHr + i : nullptr;
    });
    float dl[HH];
#pragma unroll
    for (int i = 0; i < HH; ++i) dl[i] = 0.f;
    matvec_acc<H, HH, S::HHP>(sw.at(S::WF), gout, dl);
#pragma unroll
    for (int i = 0; i < HH; ++i) dl[i] *= (1.f - tau1[i] * tau1[i]);
    for (int j = nsh; j > 0; --j) {
        float r1[HH + 1];
#pragma unroll
        for (int i = 0; i < HH; ++i) r1[i] = acts[((j - 1) * HH + i) * stride];
        r1[HH] = 1.f;
        outer_auto<HH, HH + 1, 32>(dl, r1, stg, [&](int o, int i) -> float* {
            if (o >= HHr) return nullptr;
            if (i == HH) return gw + g.bs + o;
            return i < HHr ? gw + g.Ws + o * HHr + i : nullptr;
        });
        float dn[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) dn[i] = 0.f;
        matvec_acc<HH, HH, S::HHP>(sw.at(S::WS), dl, dn);
#pragma unroll
        for (int i = 0; i < HH; ++i) dl[i] = r1[i] > 0.f ? dn[i] : 0.f;
    }
    float yt1[H + 2];
#pragma unroll
    for (int i = 0; i < H; ++i) yt1[i] = yin[i];
    yt1[H] = tstage;
    yt1[H + 1] = 1.f;
    outer_auto<HH, H + 2, 32>(dl, yt1, stg, [&](int o, int i) -> float* {
        if (o >= HHr) return nullptr;
        if (i == H + 1) return gw + g.ba + o;
        if (i == H) return gw + g.Wa + o * g.lda + g.d;
        return i < Hr ? gw + g.Wa + o * g.lda + g.d + 1 + i : nullptr;
    });
#pragma unroll
    for (int i = 0; i < HH; ++i) a0[i] += dl[i];
    matvec_acc<HH, H, S::HP>(sw.at(S::WY), dl, gy);
}

template <int H, int HH, int SOLVER, int MODE>
XW_GLOBAL void k_xnode_bwd(XnodeBwdArgs a) {
    using S = USmem<H, HH>;
    using T = Tableau<SOLVER>;
    XW_DYN_SMEM(smem_raw);
    const int nwarps = XW_BDIM >> 5, warp = XW_TID >> 5, lane = XW_TID & 31;
    const ULayout g(a.d, a.Hr, a.HHr);
    const int Pp = pad4(g.size);
    float* swp = reinterpret_cast<float*>(smem_raw);
    const WSmem sw = WSmem::make(swp);
    float* st = swp + pad4(S::size(a.d));
    float* sacts = st + pad4(a.L) + 4;                                  // [(nsh+1)*HH][BDIM]  ONE stage at a time
    float* sstg = sacts + (size_t)(a.nsh + 1) * HH * XW_BDIM;    // [nwarps][kStgRowsU][kStgLd]
    float* sgrad = sstg + (size_t)nwarps * kStgRowsU * kStgLd;                // [nwarps][Pp]
    double* red = reinterpret_cast<double*>(sgrad + (size_t)nwarps * Pp);
    stage_theta_u<H, HH>(swp, a.theta, a.d, a.Hr, a.HHr);
    for (int i = XW_TID; i < a.L; i += XW_BDIM) st[i] = a.times[i];
    for (int i = XW_TID; i < nwarps * Pp; i += XW_BDIM) sgrad[i] = 0.f;
    XW_SYNCTHREADS();
    float* stg = sstg + (size_t)warp * kStgRowsU * kStgLd;
    float* gw = sgrad + (size_t)warp * Pp;

    const long long nthr = (long long)XW_GDIM * XW_BDIM;
    const long long gtid = (long long)XW_BID * XW_BDIM + XW_TID;
    const int L = a.L, nsh = a.nsh;
    float k0 = 0.f, k1 = 0.f, k2 = 0.f;
    if (MODE == 0) { k0 = (float)a.coefs[0]; k1 = (float)a.coefs[1]; k2 = (float)a.coefs[2]; }
    const float gsc2 = (float)(2.0 * a.gscale);
    float gwo[H], gbo = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) gwo[i] = 0.f;
    double bd_acc = 0.0;
    const long long iters = (a.n + nthr - 1) / nthr;
    for (long long it = 0; it < iters; ++it) {
        const long long nraw = it * nthr + gtid;
        const bool active = nraw < a.n;
        const long long n = active ? nraw : (long long)a.n - 1;
        const float* xp = a.x + n * a.x_sn;
        float ax[HH];
        hoist_ax<H, HH>(sw, xp, a.d, ax);
        const float s0 = a.s0[n];
        float y[H];
        const bool have_hist = MODE == 0 && a.ypath != nullptr;
        if (have_hist) {               // the forward kernel kept the state history: no forward sweep here
#pragma unroll
            for (int i = 0; i < H; ++i) y[i] = a.ypath[((long long)(L - 1) * H + i) * a.n + n];
        } else {
            {
                float z1[H], z2[H];
                lift_fwd<H, HH>(sw, s0, z1, z2, y);
            }
            for (int l = 0; l + 1 < L; ++l) {
#pragma unroll
                for (int i = 0; i < H; ++i) a.yhist[((long long)l * H + i) * nthr + gtid] = y[i];
                RecNone<HH> rec[T::S];
                rk_step<H, HH, SOLVER, RecNone<HH>, false>(sw, ax, st[l], st[l + 1] - st[l], nsh, y, rec, nullptr);
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
        float lam[H], a0[HH];
#pragma unroll
        for (int i = 0; i < HH; ++i) a0[i] = 0.f;
        {
            const float G = cot_at(L - 1, project_u<H, HH>(sw, y));
            float wo[H];
            load_row<H>(sw.at(S::WO), wo);
#pragma unroll
            for (int i = 0; i < H; ++i) { lam[i] = wo[i] * G; gwo[i] = fmaf(G, y[i], gwo[i]); }
            gbo += G;
        }
        for (int l = L - 2; l >= 0; --l) {
            const float t0 = st[l], dt = st[l + 1] - st[l];
            float yl[H], ycur[H];
#pragma unroll
            for (int i = 0; i < H; ++i) {
                yl[i] = have_hist ? a.ypath[((long long)l * H + i) * a.n + n] : a.yhist[((long long)l * H + i) * nthr + gtid];
                ycur[i] = yl[i];
            }
            // stage inputs yin[s] (forward through the stages; only the LAST stage records its
            // internals: the shared-memory activation buffer holds one stage at a time, earlier
            // stages are re-evaluated right before their own reverse -> 2S-1 field evaluations)
            RecSmem<HH> rec;
            rec.base = sacts + XW_TID;
            rec.stride = XW_BDIM;
            rec.nsh = nsh;
            float yin[T::S][H];
            {
                float kst[T::S][H];
#pragma unroll
                for (int s = 0; s < T::S; ++s) {
#pragma unroll
                    for (int i = 0; i < H; ++i) yin[s][i] = ycur[i];
#pragma unroll
                    for (int r = 0; r < s; ++r) {
                        const float c = T::a(s, r);
                        if (c != 0.f) {
                            const float cd = c * dt;
#pragma unroll
                            for (int i = 0; i < H; ++i) yin[s][i] = fmaf(cd, kst[r][i], yin[s][i]);
                        }
                    }
                    float tau[HH];
                    if (s + 1 < T::S) {
                        RecNone<HH> none;
                        field_fwd<H, HH>(sw, ax, fmaf(T::c(s), dt, t0), yin[s], nsh, kst[s], tau, none);
                    } else {
                        field_fwd<H, HH>(sw, ax, fmaf(T::c(s), dt, t0), yin[s], nsh, kst[s], tau, rec);
                    }
                }
            }
            float kbar[T::S][H], ybar[H];
#pragma unroll
            for (int s = 0; s < T::S; ++s)
#pragma unroll
                for (int i = 0; i < H; ++i) kbar[s][i] = (T::b(s) * dt) * lam[i];
#pragma unroll
            for (int i = 0; i < H; ++i) ybar[i] = lam[i];
#pragma unroll
            for (int s = T::S - 1; s >= 0; --s) {
                if (s + 1 < T::S) {      // re-evaluate this stage, recording its internals
                    float kdummy[H], tau[HH];
                    field_fwd<H, HH>(sw, ax, fmaf(T::c(s), dt, t0), yin[s], nsh, kdummy, tau, rec);
                }
                float gin[H];
#pragma unroll
                for (int i = 0; i < H; ++i) gin[i] = 0.f;
                field_rev_grads<H, HH>(sw, rec.base, XW_BDIM, nsh, fmaf(T::c(s), dt, t0), yin[s], kbar[s], gin, a0,
                                       stg, gw, g);
#pragma unroll
                for (int i = 0; i < H; ++i) ybar[i] += gin[i];
#pragma unroll
                for (int r = 0; r < s; ++r) {
                    const float c = T::a(s, r);
                    if (c != 0.f) {
                        const float cd = c * dt;
#pragma unroll
                        for (int i = 0; i < H; ++i) kbar[r][i] = fmaf(cd, gin[i], kbar[r][i]);
                    }
                }
            }
            const float G = cot_at(l, project_u<H, HH>(sw, yl));
            float wo[H];
            load_row<H>(sw.at(S::WO), wo);
#pragma unroll
            for (int i = 0; i < H; ++i) { lam[i] = fmaf(wo[i], G, ybar[i]); gwo[i] = fmaf(G, yl[i], gwo[i]); }
            gbo += G;
        }
        // x part of the first field layer: dWa[:, j] += a0 (x) x_j
        warp_outer_dyn<HH, 3, 2>(a0, a.d, [&](int j) -> float { return xp[j]; }, stg, stg + 32 * kStgLd,
                              [&](int o, int j) -> float* { return o < a.HHr ? gw + g.Wa + o * g.lda + j : nullptr; });
        // lift reverse
        float z1[H], z2[H], y0[H];
        lift_fwd<H, HH>(sw, s0, z1, z2, y0);
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
        matvec_acc<H, H, S::HP>(sw.at(S::W2), lam, dz2);
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
        matvec_acc<H, H, S::HP>(sw.at(S::W1), dz2, dz1);
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
    // final_linear grads: per-lane accumulators -> warp sum -> warp image
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
        a.gpart[(size_t)XW_BID * g.size + e] = sgr;
    }
    if (MODE == 1) {
        double v[1] = {bd_acc};
        const int idx[1] = {5};
        block_sum_to_global<1>(v, red, a.sums, idx);
    }
}

// grad_out[e] = (accumulate ? grad_out[e] : 0) + sum_b gpart[b][e]
XW_GLOBAL void k_reduce_partials(const float* gpart, int nblocks, int P, float* out, int accumulate) {
    const long long gtid = (long long)XW_BID * XW_BDIM + XW_TID;
    if (gtid >= P) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += (double)gpart[(size_t)b * P + gtid];
    out[gtid] = (float)(s + (accumulate ? (double)out[gtid] : 0.0));
}

}  // namespace xw

#include "xw_vnet_tile.cuh"

namespace xw {

// =============================================================================================
// Adam step on one flat parameter vector (reference: torch.optim.Adam with default betas / eps, no weight decay, no
// amsgrad, src/training.py:103-104,138,162).  ONE CTA (the nets have 1.5-7.7 k parameters): reads the step counter,
// updates fp64 parameters and moments from the kernels' flat fp32 gradient, refreshes the fp32 copy the kernels read,
// then bumps the counter -- one launch instead of torch's ~15 foreach launches + 14 gradient casts + the re-packing.
// =============================================================================================
XW_GLOBAL void k_adam_step(double* p, const float* g, double* m, double* v, long long* step, float* p32, int n, double lr,
                           double b1, double b2, double eps) {
    const long long t = step[0] + 1;
    const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    const double step_size = lr / bc1, bc2s = sqrt(bc2);
    for (int i = XW_TID; i < n; i += XW_BDIM) {
        const double gi = (double)g[i];
        const double mi = m[i] + (gi - m[i]) * (1.0 - b1);          // exp_avg.lerp_(grad, 1 - beta1)
        const double vi = v[i] * b2 + (1.0 - b2) * gi * gi;
        const double pi = p[i] - step_size * (mi / (sqrt(vi) / bc2s + eps));
        m[i] = mi; v[i] = vi; p[i] = pi;
        if (p32) p32[i] = (float)pi;
    }
    XW_SYNCTHREADS();
    if (XW_TID == 0) step[0] = t;
}

// =============================================================================================
// The scalars of one sub-step from the (all-reduced) Monte-Carlo sums, on the device in fp64, one launch instead of ~22
// one-element torch kernels between the forward and the backward kernels (at the shipped N = 4000 those were a tenth of
// a graph-replayed sub-step).  Reference: I, S (src/loss.py:64-76), init / bdry / int (:78-90), loss u / v (:92-96);
// k[] = the cotangent coefficients the backward entries take (include/xnode_wan_b200.h).
//   out[0] loss   out[1] I   out[2] S   out[3] init   out[4] bdry   out[5..7] k0, k1, k2
// =============================================================================================
XW_GLOBAL void k_loss_scalars(const double* sums, int phase, double V, double n, double L, double nb, double Lb, double alpha,
                              double side, double* out) {
    if (XW_TID != 0 || XW_BID != 0) return;
    const double I = (V / n) * sums[0] - (V / (n * L)) * (sums[1] - sums[2]);
    const double S = V * sums[3] / (n * L);
    const double init = sums[4] / n;
    const double bdry = nb > 0.0 ? sums[5] / (nb * Lb) : 0.0;
    const double integ = log(I * I) - log(S);
    out[1] = I; out[2] = S; out[3] = init; out[4] = bdry;
    if (phase == 0) {
        out[0] = integ + alpha * (init + bdry);
        out[5] = (2.0 / I) * (V / (n * L)); out[6] = 2.0 * alpha / n; out[7] = side;
    } else {
        out[0] = -integ;
        out[5] = -(2.0 / I) * (V / (n * L)); out[6] = 2.0 / sums[3]; out[7] = side;
    }
}

// =============================================================================================
// weak-form sums from the CACHED test-function values: when the sample and theta_v are unchanged
// (2nd u-step and the v-step of one outer iteration, src/training.py:125-162) only u, du change, so
// the v net need not be evaluated again.  One thread per point; time-row-0 threads add the
// a grad(phi).du + b.du phi term from the cached grad_x phi.
// =============================================================================================
struct CombineArgs {
    int d, n, L;
    float c0, c1; const float* ca; const float* cb;
    long long ca_sn, cb_sn;
    const float* Aval; const float* Ader;
    const float* vcache; const float* gcache;
    const float* u; const float* du; const float* h; const float* f;
    double* sums; float* cot_u; float* cot_v;
};

XW_GLOBAL void k_weak_combine(CombineArgs a) {
    XW_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);
    const long long nthr = (long long)XW_GDIM * XW_BDIM;
    const long long npts = (long long)a.n * a.L;
    const int L = a.L, d = a.d;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long p = (long long)XW_BID * XW_BDIM + XW_TID; p < npts; p += nthr) {
        const long long n = p / L;
        const int l = (int)(p - n * L);
        const f4 c = ld4(a.vcache + 4 * p);
        float cu, cv;
        float Au, Ap;
        weak_A(a.c0, a.c1, a.Aval, a.Ader, p, a.u[p], Au, Ap);
        weak_point_terms(c.x, c.y, c.z, c.w, a.u[p], a.f[p], l == 0 ? a.h[n] : 0.f, l, L, Au, Ap, acc, cu, cv);
        a.cot_u[p] = cu;
        a.cot_v[p] = cv;
        if (l == 0) {
            const float* dun = a.du + n * d;
            const float* gp = a.gcache + n * d;
            float s31 = 0.f;
            for (int i = 0; i < d; ++i) {
                float q;
                if (a.ca) {
                    q = 0.f;
                    for (int j = 0; j < d; ++j) q = fmaf(a.ca[n * a.ca_sn + i * d + j], dun[j], q);
                } else {
                    q = dun[i];
                }
                s31 = fmaf(gp[i], q, s31);
            }
            if (a.cb) {
                float bq = 0.f;
                for (int j = 0; j < d; ++j) bq = fmaf(a.cb[n * a.cb_sn + j], dun[j], bq);
                s31 = fmaf(c.x * c.z, bq, s31);
            }
            acc[2] += (double)s31;
        }
    }
    const int idx[4] = {0, 1, 2, 3};
    block_sum_to_global<4>(acc, red, a.sums, idx);
}

}  // namespace xw

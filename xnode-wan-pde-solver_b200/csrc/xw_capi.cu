// xw_capi.cu -- C ABI (include/xnode_wan_b200.h) over the kernels in xw_kernels.cuh.
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a -> libxnode_wan_b200.so.
// The same file compiles with g++ -DXW_EMU -x c++ into the CPU logic-test library
// (tests/host_emu); that build is test infrastructure and is never shipped or loaded by the
// product package.
#include "../../include/xnode_wan_b200.h"
#include "xw_kernels.cuh"
#include "xw_xnode2.cuh"
#include "xw_xnode3.cuh"
#include "xw_umma.cuh"
#include "xw_vnet_tc.cuh"
#ifndef XW_EMU
#include "xw_vnet_virtual.cuh"
#endif

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#ifdef XW_EMU
struct Dev { int sms = 2; size_t smem_optin = 227 * 1024; int ordinal = 0; };
const Dev* device() { static Dev d; return &d; }
#define XW_LAUNCH(kern, grid, block, smem, stream, ...)                              \
    do { emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); }); } while (0)
#define XW_CHECK_LAUNCH(name) 0
#define XW_SET_SMEM(kern, bytes) 0
#else
// properties of the CURRENT device (per-ordinal cache: a process may drive several GPUs)
struct Dev { int sms; size_t smem_optin; int ordinal; };
constexpr int kMaxDevices = 64;
const Dev* device() {
    static Dev devs[kMaxDevices];
    static bool have[kMaxDevices] = {};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev]) {
        int sms = 0, optin = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
            return nullptr;
        devs[dev] = Dev{sms, (size_t)optin, dev};
        have[dev] = true;
    }
    return &devs[dev];
}
#define XW_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
int check_launch(const char* name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("%s: launch failed: %s", name, cudaGetErrorString(e));
    return 0;
}
#define XW_CHECK_LAUNCH(name) check_launch(name)
template <class K>
int set_smem(K kern, size_t bytes, const char* name) {
    if (bytes <= 48 * 1024) return 0;
    // remember the largest opt-in per kernel: no CUDA API call on the steady-state path (keeps the
    // launch sequence capturable into a CUDA graph)
    // (cudaFuncSetAttribute is a per-device setting: the memo is keyed by device ordinal as well)
    static std::map<std::pair<int, const void*>, size_t> done;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("%s: cudaGetDevice failed", name);
    std::lock_guard<std::mutex> lk(mu);
    size_t& have = done[std::make_pair(dev, (const void*)kern)];
    if (have >= bytes) return 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    if (e != cudaSuccess) return fail("%s: cannot opt in to %zu bytes of shared memory: %s", name, bytes, cudaGetErrorString(e));
    return 0;
}
#define XW_SET_SMEM(kern, bytes) set_smem(kern, bytes, #kern)
#endif

// compiled capacities (exact for configs/cube_pde.yaml: H=20, hh=10, Hv=50); smaller nets are
// zero-padded into them, larger ones are rejected (no silent fallback).
constexpr int kH = 20, kHH = 10, kHV = 50;
#ifndef XW_TILE_QR
#define XW_TILE_QR 4
#endif
#ifndef XW_TILE_NH
#define XW_TILE_NH 1
#endif
constexpr int kBlkFwd = 128;

int check_dims(const xw_dims* m) {
    if (!m) return fail("dims is NULL");
    if (m->d < 1) return fail("dim must be >= 1 (got %d)", m->d);
    if (m->H < 1 || m->H > kH) return fail("u_hidden_dim %d unsupported (compiled capacity %d)", m->H, kH);
    if (m->hh < 1 || m->hh > kHH) return fail("u_hidden_hidden_dim %d unsupported (compiled capacity %d)", m->hh, kHH);
    if (m->nu < 1) return fail("u_layers %d unsupported (need >= 1)", m->nu);
    if ((m->nu - 1) * kHH > 128) return fail("u_layers %d unsupported (relu-mask stack holds %d layers)", m->nu, 128 / kHH + 1);
    if (m->Hv < 1 || m->Hv > kHV) return fail("v_hidden_dim %d unsupported (compiled capacity %d)", m->Hv, kHV);
    if (m->nv < 0 || m->nv > xw::kMaxNv) return fail("v_layers %d unsupported (max %d)", m->nv, xw::kMaxNv);
    if (m->solver < 0 || m->solver > 2) return fail("solver %d unsupported (0 euler, 1 midpoint, 2 rk4)", m->solver);
    return 0;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
int grid_for(long long items, int block, int ctas_per_sm);
int ctas_per_sm_for(size_t smem, int cap);
// per-thread scratch of the interior forward kernel: (mask bits + tanh) of every field evaluation
size_t fwd_hist_floats(const xw_dims* m, int L) {
    const int S = m->solver == 0 ? 1 : m->solver == 1 ? 2 : 4;
    return (size_t)std::max(L, 1) * S * xw::kRecWords<kHH>;
}

int stages_of(int solver) { return solver == 0 ? 1 : solver == 1 ? 2 : 4; }
int bwd_block(int solver) { (void)solver; return 128; }

size_t smem_xnode_fwd(int d, int L) {
    using S = xw::USmem<kH, kHH>;
    return (size_t)(xw::pad4(S::size(d)) + xw::pad4(L) + 4) * 4 + 32 * 8;
}
size_t smem_xnode_bwd(const xw_dims* m, int L, int block) {
    using S = xw::USmem<kH, kHH>;
    const int nw = block / 32;
    const xw::ULayout g(m->d, m->H, m->hh);
    size_t f = (size_t)xw::pad4(S::size(m->d)) + xw::pad4(L) + 4;
    f += (size_t)m->nu * kHH * block;                              // (nsh+1) = nu, one stage at a time
    f += (size_t)nw * xw::kStgRowsU * xw::kStgLd;
    f += (size_t)nw * xw::pad4(g.size);
    return f * 4 + 32 * 8;
}
size_t smem_vnet_fwd(int d) {
    using S = xw::VSmem<kHV>;
    return (size_t)(xw::pad4(S::size(d + 1)) + 4) * 4 + 4 * 32 * 8;
}
constexpr int kQR = XW_TILE_QR, kNH = XW_TILE_NH;
size_t smem_vtile_fwd(int d) {
    using VT = xw::VTile<kHV, kQR, kNH>;
    const int C = d + 1;
    size_t f = (size_t)xw::pad4(VT::WIT + VT::wit_size(C)) + (size_t)VT::ROWS * VT::RS + (size_t)VT::ROWS * VT::xin_ld(C) +
               (size_t)VT::NG * VT::ROWS * 2;
    return f * 4 + 4 * 32 * 8 + 2 * VT::ROWS * 4;
}
constexpr int kQRB = 2;      // pair rows per lane in the tiled backward (2 x 64 rows = 128 points per tile)
size_t smem_vtile_bwd(int d) {
    using VT = xw::VTile<kHV, kQRB>;
    const int C = d + 1, NPT = 2 * VT::ROWS;
    size_t f = (size_t)xw::pad4(VT::WIT + VT::wit_size(C)) + (size_t)kHV * VT::WLD + 2 * (size_t)VT::ROWS * VT::RS +
               (size_t)xw::pad4(NPT * VT::xin_ld(C)) + (size_t)VT::NG * NPT + NPT + (size_t)xw::pad4(kHV * (C + 1)) + 2 * NPT;
    return f * 4 + 64;
}
struct VtileBwdPlan { int grid; size_t smem, scratch_bytes, part_bytes; };
int plan_vtile_bwd(const xw_dims* m, int n, int L, VtileBwdPlan* p) {
    using VT = xw::VTile<kHV, kQRB>;
    p->smem = smem_vtile_bwd(m->d);
    if (p->smem > device()->smem_optin) return fail("tiled v-net backward needs %zu B shared memory (> %zu): dim too large", p->smem, device()->smem_optin);
    const long long ntiles = ((long long)n * L + 2 * VT::ROWS - 1) / (2 * VT::ROWS);
    p->grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)device()->sms * ctas_per_sm_for(p->smem, 2)));
    p->scratch_bytes = align_up((size_t)p->grid * std::max(m->nv, 1) * VT::ROWS * VT::RS * 4, 256);
    p->part_bytes = align_up((size_t)p->grid * xw::VLayout(m->d, m->Hv).size * 4, 256);
    return 0;
}
#ifndef XW_EMU
// tensor-core backward: one 128-point tile per CTA iteration, one CTA per SM (tensor memory: 512 columns)
// test hook: XW_TC_TMEM_PACKED=1 runs k_vnet_tc_bwd3 on the fully packed tensor-memory layout for every input width
// (the layout kin = 56 needs anyway), so that the parity tests of the narrow configs cover it too
// tiles per flush of k_vnet_tc_bwd3's weight-gradient accumulators (tensor memory -> the CTA's fp32 image).  The tensor core
// truncates when it accumulates, so the error of dWh grows linearly with the number of tiles summed in one accumulator
// (measured against fp64 at 2^20 paths, tools/tc_prof.py acc: 4.4e-6 / 6.8e-6 / 1.2e-5 / 2.1e-5 rel-L2 at 1 / 2 / 4 / 8
// tiles, 2.7e-3 without flushing); 4 keeps it at the level of the other tensors (dWi 8.6e-6) and takes the flush off the
// critical path (26.8 -> 25.2 ms).  XW_TC_FLUSH overrides (tests).
int tc_flush_tiles() { static const int v = []() { const char* e = getenv("XW_TC_FLUSH"); const int k = e ? atoi(e) : 4; return k > 0 ? k : 4; }(); return v; }
// k_vnet_tc_fwd: MMAs of a layer issued by three warps instead of one thread (XW_TC_SPLIT=1; 18.9 -> 17.2 ms per evaluated
// interior forward at 2^20 paths).  The order in which the three 3xTF32 terms reach the accumulator then varies from run to
// run: v agrees to fp32 rounding (1e-7) at every point, not bit for bit -- and dv/dt, which is discontinuous where a
// pre-activation crosses 0, may land on the other side of a relu kink at the few points that sit within rounding of one.
// The DEFAULT is the single-issuer kernel: bit-reproducible runs (and tests).  Read per call, so that one process can
// measure both (bench.py `variants`).
int tc_split_issue() { const char* e = getenv("XW_TC_SPLIT"); return e ? atoi(e) : 0; }
// the same for the F-op / R-op of k_vnet_tc_bwd3: OFF (XW_TC_SPLIT_BWD=1 selects it).  It gains 3 % (25.2 -> 24.5 ms) and is as
// accurate as the single-issuer kernel against fp64 (tools/tc_prof.py acc), but it lands on the other side of a relu kink
// more often: the parameter gradient of this net is DISCONTINUOUS where a pre-activation crosses 0, so any two fp32
// evaluation orders disagree with the fp64 reference by O(weight of one point) at the few pre-activations that sit within
// rounding of 0 (seed sweep at d = 20 and d = 100, 6 seeds each: single issuer 1-2 configurations off by 2e-4..6e-4 in dWi,
// three issuers 3 off by 9e-4..1e-2; the others 4e-6).  The deterministic kernel keeps test_mid_size_against_oracle
// reproducible run after run; the run-to-run varying one would make it flaky.
int tc_split_bwd() { const char* e = getenv("XW_TC_SPLIT_BWD"); return e ? atoi(e) : 0; }
int tc_tmem_packed() { static const int v = []() { const char* e = getenv("XW_TC_TMEM_PACKED"); return e && e[0] == '1' ? 1 : 0; }(); return v; }
bool vtc_bwd_ok(const xw_dims* m) { return xw::tc::kin_of(m->d) <= xw::tc::KP; }
int plan_vtc_bwd(const xw_dims* m, int n, int L, VtileBwdPlan* p) {
    const int kin = xw::tc::kin_of(m->d);
    p->smem = (size_t)(4 * xw::tc::KP * xw::tc::NP + 2 * kin * xw::tc::NP + 64 + 2 * xw::tc::TIMG2 + xw::tc::KP * (xw::tc::KP + kin + 1) + 512 + 64 + 256) * 4 + 128;
    if (p->smem > device()->smem_optin) return fail("tensor-core v-net backward needs %zu B shared memory (> %zu)", p->smem, device()->smem_optin);
    const long long ntiles = ((long long)n * L + 127) / 128;
    p->grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)device()->sms));
    p->scratch_bytes = align_up((size_t)p->grid * 2 * std::max(m->nv, 1) * 14 * 128 * 16, 256);   // double-buffered (pipelined variant)
    p->part_bytes = align_up((size_t)p->grid * xw::VLayout(m->d, m->Hv).size * 4, 256);
    return 0;
}
#endif
#ifndef XW_EMU
// wide inputs (d > 54): the tensor-core backward runs on the VIRTUAL net of input width Hv (xw_vnet_virtual.cuh)
struct VvPlan { xw_dims mv; VtileBwdPlan pl; int grid_prep, grid_dwx, Pv; size_t y_b, th_b, gv_b, w_b, d0_b, px_b, total; };
bool vv_wanted(const xw_dims* m) {
    const char* e = getenv("XW_VNET_WIDE");
    return !vtc_bwd_ok(m) && m->Hv + 2 <= xw::tc::KP && !(e && strcmp(e, "tile") == 0);
}
int plan_vv(const xw_dims* m, int n, int L, VvPlan* p) {
    p->mv = *m;
    p->mv.d = m->Hv;
    if (plan_vtc_bwd(&p->mv, n, L, &p->pl)) return 1;
    p->Pv = xw::VLayout(m->Hv, m->Hv).size;
    p->grid_prep = (int)std::max<long long>(1, std::min<long long>(((long long)n + 127) / 128, (long long)device()->sms * 4));
    p->grid_dwx = (int)std::max<long long>(1, std::min<long long>(((long long)n + xw::vv::kDwxPaths - 1) / xw::vv::kDwxPaths, (long long)device()->sms * 2));
    p->y_b = align_up((size_t)n * m->Hv * 4, 256);
    p->th_b = align_up((size_t)p->Pv * 4, 256);
    p->gv_b = p->th_b;
    p->w_b = align_up((size_t)n * L * 4, 256);
    p->d0_b = align_up((size_t)n * L * 52 * 4, 256);
    p->px_b = align_up((size_t)p->grid_dwx * m->Hv * m->d * 4, 256);
    p->total = p->pl.scratch_bytes + p->pl.part_bytes + p->y_b + p->th_b + p->gv_b + p->w_b + p->d0_b + p->px_b;
    return 0;
}
#endif
// XW_VNET_IMPL: "tc" (default on the device build: tcgen05 3xTF32 kernels), "tile" (FP32 FFMA2 tile engine)
bool use_tc_kernels() {
#ifdef XW_EMU
    return false;
#else
    const char* e = getenv("XW_VNET_IMPL");
    return !e || strcmp(e, "tc") == 0;
#endif
}

int grid_for(long long items, int block, int ctas_per_sm) {
    const Dev* dv = device();
    long long want = (items + block - 1) / block;
    long long cap = (long long)dv->sms * ctas_per_sm;
    return (int)std::max<long long>(1, std::min(want, cap));
}
int ctas_per_sm_for(size_t smem, int cap) {
    const Dev* dv = device();
    int k = (int)((dv->smem_optin + 1024) / (smem + 1024));
    return std::max(1, std::min(k, cap));
}

template <int MODE>
int launch_xnode_fwd(const xw_dims* m, const xw::XnodeFwdArgs& a, int grid, size_t smem, void* stream) {
    using namespace xw;
#define XW_CASE(SOLV)                                                                       \
    case SOLV: {                                                                            \
        if (XW_SET_SMEM((k_xnode_fwd<kH, kHH, SOLV, MODE, WSmem>), smem)) return 1;         \
        XW_LAUNCH((k_xnode_fwd<kH, kHH, SOLV, MODE, WSmem>), grid, kBlkFwd, smem, stream, a); \
        break;                                                                              \
    }
    switch (m->solver) { XW_CASE(0) XW_CASE(1) XW_CASE(2) }
#undef XW_CASE
    return XW_CHECK_LAUNCH("k_xnode_fwd");
}

template <int MODE>
int launch_xnode_bwd(const xw_dims* m, const xw::XnodeBwdArgs& a, int grid, int block, size_t smem, void* stream) {
    using namespace xw;
#define XW_CASE(SOLV)                                                                       \
    case SOLV: {                                                                            \
        if (XW_SET_SMEM((k_xnode_bwd<kH, kHH, SOLV, MODE>), smem)) return 1;                \
        XW_LAUNCH((k_xnode_bwd<kH, kHH, SOLV, MODE>), grid, block, smem, stream, a);        \
        break;                                                                              \
    }
    switch (m->solver) { XW_CASE(0) XW_CASE(1) XW_CASE(2) }
#undef XW_CASE
    return XW_CHECK_LAUNCH("k_xnode_bwd");
}

xw::PointsView view_of(const xw_points* p) {
    xw::PointsView v;
    v.t = p->t; v.t_sn = p->t_sn; v.t_sl = p->t_sl;
    v.x = p->x; v.x_sn = p->x_sn; v.x_sl = p->x_sl;
    return v;
}

int reduce_partials(const float* gpart, int nblocks, int P, float* out, int accumulate, void* stream) {
    XW_LAUNCH(xw::k_reduce_partials, (P + 127) / 128, 128, 0, stream, gpart, nblocks, P, out, accumulate);
    return XW_CHECK_LAUNCH("k_reduce_partials");
}

struct XnodeBwdPlan { int block, grid; size_t smem, hist_bytes, part_bytes; };
int plan_xnode_bwd(const xw_dims* m, int n, int L, XnodeBwdPlan* p) {
    const Dev* dv = device();
    p->block = bwd_block(m->solver);
    p->smem = smem_xnode_bwd(m, L, p->block);
    if (p->smem > dv->smem_optin)
        return fail("xnode backward needs %zu B shared memory per CTA (> %zu): u_layers/N_t/dim too large", p->smem, dv->smem_optin);
    p->grid = grid_for(n, p->block, ctas_per_sm_for(p->smem, 8));
    p->hist_bytes = align_up((size_t)L * kH * p->grid * p->block * 4, 256);
    p->part_bytes = align_up((size_t)p->grid * xw::ULayout(m->d, m->H, m->hh).size * 4, 256);
    return 0;
}


// ---------------------------------------------------------------------------------------------
// generation-2 XNODE engine (xw_xnode2.cuh): reduced state, shared layer in registers, warp-specialised
// backward.  Covers u_layers <= 8 (7 shared-layer uses + the tanh slot = the 8 task slots of a tile row);
// deeper field nets run on the generation-1 kernels.  XW_XNODE_IMPL=v1 forces generation 1 (A/B runs).
// ---------------------------------------------------------------------------------------------
bool use_x2(const xw_dims* m) {
    if (m->nu - 1 > xw::x2::kSlots - 1) return false;
    const char* e = getenv("XW_XNODE_IMPL");
    return !(e && strcmp(e, "v1") == 0);
}
int g_last_xnode_impl = 0;       // 1 / 2 / 3: which generation the last XNODE launch used
int g_last_xnode_bwd = 0;        // same for the last BACKWARD launch (a forward launch follows it in every step)
int g_last_vnet_fwd = 0, g_last_vnet_bwd = 0;   // 1 points, 2 FP32 tile engine, 3 tcgen05 (last v-net forward / backward launch)

size_t x2_smem_fwd(int d, int L) {
    using S = xw::USmem<kH, kHH>;
    return (size_t)(xw::pad4(S::size(d)) + xw::x2::R::size + xw::pad4(L) + 4) * 4 + 32 * 8;
}
int x2_grid_fwd(int n) { return grid_for(n, xw::x2::kFwdThreads, 1); }
size_t x2_rec_bytes(const xw_dims* m, int n, int L) {
    return align_up((size_t)std::max(L, 1) * stages_of(m->solver) * xw::x2::kRecWords2 * x2_grid_fwd(n) * xw::x2::kFwdThreads * 4, 256);
}
struct X2BwdPlan { int grid, lgrid, lblock; size_t smem, lsmem, hist_b, pp_b, pa_b, pb_b; bool three; };
// XW_XNODE_BWD=v2 forces the 2-role backward kernel (k_xnode2_bwd) instead of the 3-role one (k_xnode3_bwd): A/B runs
bool use_three_roles() {
    const char* e = getenv("XW_XNODE_BWD");
    return !(e && strcmp(e, "v2") == 0);
}
int x2_plan_bwd(const xw_dims* m, int n, int L, X2BwdPlan* p) {
    using S = xw::USmem<kH, kHH>;
    const Dev* dv = device();
    p->smem = (size_t)(xw::pad4(S::size(m->d)) + xw::x2::R::size + xw::pad4(L) + 4 + xw::x2::kCW * 3 * xw::x2::kTile + xw::x2::kPartA) * 4 + 32 * 8;
    const size_t smem3 = (size_t)(xw::pad4(xw::x3::Sm::fixed + m->d * xw::x2::HHP) + xw::pad4(L) + 4 + xw::x3::kTriples * 4 * xw::x2::kTile) * 4 +
                         xw::x3::kTriples * 8 * 8 + 32 * 8;
    p->three = use_three_roles() && smem3 <= dv->smem_optin;
    if (p->three) p->smem = smem3;
    if (p->smem > dv->smem_optin)
        return fail("xnode backward needs %zu B shared memory per CTA (> %zu): N_t/dim too large", p->smem, dv->smem_optin);
    const int cpaths = 32 * xw::x2::kCW;
    p->grid = (int)std::max<long long>(1, std::min<long long>(((long long)n + cpaths - 1) / cpaths, dv->sms));
    p->lblock = 128;
    const int nw = p->lblock / 32;
    const int P = xw::ULayout(m->d, m->H, m->hh).size;
    p->lsmem = (size_t)(xw::pad4(S::size(m->d)) + nw * xw::kStgRowsU * xw::kStgLd + nw * xw::pad4(P)) * 4;
    if (p->lsmem > dv->smem_optin)
        return fail("xnode lift backward needs %zu B shared memory per CTA (> %zu): dim too large", p->lsmem, dv->smem_optin);
    p->lgrid = grid_for(n, p->lblock, ctas_per_sm_for(p->lsmem, 4));
    p->hist_b = align_up((size_t)L * (xw::x2::kZQ + xw::x2::HH * (stages_of(m->solver) - 1)) * (p->three ? (size_t)n : (size_t)p->grid * cpaths) * 4, 256);
    p->pp_b = align_up((size_t)xw::x2::kPerPath * n * 4, 256);
    p->pa_b = align_up((size_t)p->grid * xw::x2::kPartA * 4, 256);
    p->pb_b = align_up((size_t)p->lgrid * P * 4, 256);
    return 0;
}
size_t x2_bwd_bytes(const X2BwdPlan& p) { return p.hist_b + p.pp_b + p.pa_b + p.pb_b; }

template <int MODE>
int x2_launch_fwd(const xw_dims* m, const xw::x2::FwdArgs& a, void* stream) {
    using namespace xw::x2;
    const size_t smem = x2_smem_fwd(a.d, a.L);
    const int grid = x2_grid_fwd(a.n);
#define XW_CASE(SOLV)                                                                       \
    case SOLV: {                                                                            \
        if (XW_SET_SMEM((k_xnode2_fwd<SOLV, MODE>), smem)) return 1;                        \
        XW_LAUNCH((k_xnode2_fwd<SOLV, MODE>), grid, kFwdThreads, smem, stream, a);          \
        break;                                                                              \
    }
    switch (m->solver) { XW_CASE(0) XW_CASE(1) XW_CASE(2) }
#undef XW_CASE
    g_last_xnode_impl = 2;
    return XW_CHECK_LAUNCH("k_xnode2_fwd");
}

// backward kernel + per-path kernel + finish.  grad_u == NULL: sums only (no lift / finish).
template <int MODE>
int x2_run_bwd(const xw_dims* m, xw::x2::BwdArgs a, const X2BwdPlan& p, void* workspace, float* grad_u, int accumulate,
               void* stream) {
    using namespace xw::x2;
    char* ws = (char*)workspace;
    a.hist = (float*)ws;
    a.perpath = (float*)(ws + p.hist_b);
    a.partA = (float*)(ws + p.hist_b + p.pp_b);
    float* partB = (float*)(ws + p.hist_b + p.pp_b + p.pa_b);
    if (p.three) {
        // 3-role kernel: the reduced state history is an input; without one (boundary batch, or an interior call that was
        // not handed the forward kernel's history) a forward-only launch writes it into the workspace first
        if (!a.zq) {
            FwdArgs f{};
            f.d = a.d; f.Hr = a.Hr; f.HHr = a.HHr; f.nsh = a.nsh; f.L = a.L; f.n = a.n;
            f.theta = a.theta; f.x = a.x; f.x_sn = a.x_sn; f.times = a.times; f.s0 = a.s0; f.zq = a.hist;
            if (x2_launch_fwd<0>(m, f, stream)) return 1;
            a.zq = a.hist;
        }
        xw::x3::Args q{};
        q.d = a.d; q.Hr = a.Hr; q.HHr = a.HHr; q.nsh = a.nsh; q.L = a.L; q.n = a.n;
        q.theta = a.theta; q.x = a.x; q.x_sn = a.x_sn; q.times = a.times; q.cot = a.cot; q.coefs = a.coefs; q.hloss = a.hloss;
        q.zq = a.zq; q.gscale = a.gscale; q.perpath = a.perpath; q.partA = a.partA; q.sums = a.sums;
#define XW_CASE(SOLV)                                                                                   \
    case SOLV: {                                                                                        \
        if (XW_SET_SMEM((xw::x3::k_xnode3_bwd<SOLV, MODE>), p.smem)) return 1;                          \
        XW_LAUNCH((xw::x3::k_xnode3_bwd<SOLV, MODE>), p.grid, xw::x3::kThreads, p.smem, stream, q);    \
        break;                                                                                          \
    }
        switch (m->solver) { XW_CASE(0) XW_CASE(1) XW_CASE(2) }
#undef XW_CASE
        g_last_xnode_impl = 3; g_last_xnode_bwd = 3;
        if (XW_CHECK_LAUNCH("k_xnode3_bwd")) return 1;
    } else {
#define XW_CASE(SOLV)                                                                       \
    case SOLV: {                                                                            \
        if (XW_SET_SMEM((k_xnode2_bwd<SOLV, MODE>), p.smem)) return 1;                      \
        XW_LAUNCH((k_xnode2_bwd<SOLV, MODE>), p.grid, kBwdThreads, p.smem, stream, a);      \
        break;                                                                              \
    }
        switch (m->solver) { XW_CASE(0) XW_CASE(1) XW_CASE(2) }
#undef XW_CASE
        g_last_xnode_impl = 2; g_last_xnode_bwd = 2;
        if (XW_CHECK_LAUNCH("k_xnode2_bwd")) return 1;
    }
    if (!grad_u) return 0;
    LiftArgs l{};
    l.d = a.d; l.Hr = a.Hr; l.HHr = a.HHr; l.n = a.n; l.theta = a.theta; l.x = a.x; l.x_sn = a.x_sn; l.s0 = a.s0;
    l.perpath = a.perpath; l.partB = partB;
    if (XW_SET_SMEM(k_xnode2_lift, p.lsmem)) return 1;
    XW_LAUNCH(k_xnode2_lift, p.lgrid, p.lblock, p.lsmem, stream, l);
    if (XW_CHECK_LAUNCH("k_xnode2_lift")) return 1;
    const int P = xw::ULayout(m->d, m->H, m->hh).size;
    XW_LAUNCH(k_xnode2_finish, (P + 255) / 256, 256, kPartA * 8, stream, a.partA, p.grid, partB, p.lgrid, a.theta, a.d, a.Hr,
              a.HHr, grad_u, accumulate);
    return XW_CHECK_LAUNCH("k_xnode2_finish");
}

}  // namespace

extern "C" {

int xw_abi_version(void) { return XW_ABI_VERSION; }
#ifdef XW_EMU
// only the CPU emulation build (tests/host_emu) has this symbol: it tells the Python layer that THIS library takes host
// pointers.  The product library does not export it, so the product path keeps refusing CPU tensors.
int xw_emu_marker(void) { return 1; }
#endif
const char* xw_last_error(void) { return g_err; }

int xw_theta_u_size(const xw_dims* m) { return m ? xw::ULayout(m->d, m->H, m->hh).size : -1; }
int xw_theta_v_size(const xw_dims* m) { return m ? xw::VLayout(m->d, m->Hv).size : -1; }
size_t xw_yhist_floats(const xw_dims* m, int n, int L) {      // generation 2: (z[10], q) per grid point; generation 1: y[H]
    if (!m) return 0;
    return (size_t)L * (use_x2(m) ? xw::x2::kZQ + xw::x2::HH * (stages_of(m->solver) - 1) : kH) * n;
}
int xw_last_xnode_impl(void) { return g_last_xnode_impl | (g_last_xnode_bwd << 4); }
int xw_last_vnet_impl(void) { return g_last_vnet_fwd | (g_last_vnet_bwd << 4); }
size_t xw_vcache_floats(const xw_dims* m, int n, int L) { return m ? (size_t)4 * n * L + (size_t)n * m->d : 0; }

size_t xw_workspace_bytes(const xw_dims* m, int n, int L) {
    if (check_dims(m) || !device() || n < 1 || L < 1) return 0;
    // interior forward: du[n*d] + u[n*L] + yhist
    int gf = grid_for(n, kBlkFwd, 8);
    size_t fwd = align_up((size_t)n * m->d * 4, 256) + align_up((size_t)n * L * 4, 256) +
                 align_up(fwd_hist_floats(m, L) * gf * kBlkFwd * 4, 256);
    size_t fwd_virtual = 0;
#ifndef XW_EMU
    if (vv_wanted(m))      // tensor-core forward on the virtual net: y[n][Hv], theta', (w, dw/dt) per point
        fwd_virtual = align_up((size_t)n * m->Hv * 4, 256) + align_up((size_t)xw::VLayout(m->Hv, m->Hv).size * 4, 256) + 2 * align_up((size_t)n * L * 4, 256);
    fwd += fwd_virtual;
#endif
    XnodeBwdPlan pb;
    if (plan_xnode_bwd(m, n, L, &pb)) return 0;
    size_t bwd_u = pb.hist_bytes + pb.part_bytes;
    if (use_x2(m)) {
        fwd = std::max(fwd, align_up((size_t)n * m->d * 4, 256) + align_up((size_t)n * L * 4, 256) + x2_rec_bytes(m, n, L) + fwd_virtual);
        X2BwdPlan p2;
        if (x2_plan_bwd(m, n, L, &p2)) return 0;
        bwd_u = std::max(bwd_u, x2_bwd_bytes(p2));
    }
    size_t bwd_v = 0;
    VtileBwdPlan pv;
    if (plan_vtile_bwd(m, n, L, &pv)) return 0;
    bwd_v = std::max(bwd_v, pv.scratch_bytes + pv.part_bytes);
#ifndef XW_EMU
    if (vtc_bwd_ok(m)) {
        if (plan_vtc_bwd(m, n, L, &pv)) return 0;
        bwd_v = std::max(bwd_v, pv.scratch_bytes + pv.part_bytes);
    } else if (vv_wanted(m)) {
        VvPlan vp;
        if (plan_vv(m, n, L, &vp)) return 0;
        bwd_v = std::max(bwd_v, vp.total);
    }
#endif
    return std::max(fwd, std::max(bwd_u, bwd_v)) + 1024;
}

int xw_adam_step(double* params, const float* grad, double* exp_avg, double* exp_avg_sq, long long* step, float* params_f32,
                 int n, double lr, double beta1, double beta2, double eps, void* stream) {
    if (!params || !grad || !exp_avg || !exp_avg_sq || !step) return fail("NULL pointer argument");
    if (n < 1) return fail("empty parameter vector");
    if (!device()) return fail("no CUDA device");
    XW_LAUNCH(xw::k_adam_step, 1, 1024, 0, stream, params, grad, exp_avg, exp_avg_sq, step, params_f32, n, lr, beta1, beta2, eps);
    return XW_CHECK_LAUNCH("k_adam_step");
}

int xw_loss_scalars(const double* sums, int phase, double V, double n_glob, double L, double nb_glob, double Lb,
                    double alpha, double side, double* out, void* stream) {
    if (!sums || !out) return fail("NULL pointer argument");
    if (phase != 0 && phase != 1) return fail("phase must be 0 (u) or 1 (v), got %d", phase);
    if (!(n_glob >= 1.0) || !(L >= 1.0) || nb_glob < 0.0 || (nb_glob > 0.0 && !(Lb >= 1.0))) return fail("bad sample counts");
    if (!device()) return fail("no CUDA device");
    XW_LAUNCH(xw::k_loss_scalars, 1, 32, 0, stream, sums, phase, V, n_glob, L, nb_glob, Lb, alpha, side, out);
    return XW_CHECK_LAUNCH("k_loss_scalars");
}

int xw_xnode_eval(const xw_dims* m, const float* theta_u, const float* x, long long x_sn, const float* times,
                  int L, const float* s0, int n, float* u_out, void* stream) {
    if (check_dims(m)) return 1;
    if (!device()) return fail("no CUDA device");
    if (n < 1 || L < 1) return fail("empty batch (n=%d, L=%d)", n, L);
    if (!theta_u || !x || !times || !s0 || !u_out) return fail("NULL pointer argument");
    if (use_x2(m)) {
        xw::x2::FwdArgs q{};
        q.d = m->d; q.Hr = m->H; q.HHr = m->hh; q.nsh = m->nu - 1; q.L = L; q.n = n;
        q.theta = theta_u; q.x = x; q.x_sn = x_sn; q.times = times; q.s0 = s0; q.u_out = u_out;
        return x2_launch_fwd<0>(m, q, stream);
    }
    xw::XnodeFwdArgs a{};
    a.d = m->d; a.Hr = m->H; a.HHr = m->hh; a.nsh = m->nu - 1; a.L = L; a.n = n;
    a.theta = theta_u; a.x = x; a.x_sn = x_sn; a.times = times; a.s0 = s0; a.u_out = u_out;
    g_last_xnode_impl = 1;
    return launch_xnode_fwd<0>(m, a, grid_for(n, kBlkFwd, 8), smem_xnode_fwd(m->d, L), stream);
}

int xw_vnet_eval(const xw_dims* m, const float* theta_v, const xw_points* pts, int n, int L, float* v_out, void* stream) {
    if (check_dims(m)) return 1;
    if (!device()) return fail("no CUDA device");
    if (n < 1 || L < 1) return fail("empty batch (n=%d, L=%d)", n, L);
    if (!theta_v || !pts || !pts->t || !pts->x || !v_out) return fail("NULL pointer argument");
    xw::VnetFwdArgs a{};
    a.d = m->d; a.Hvr = m->Hv; a.nv = m->nv; a.n = n; a.L = L; a.theta = theta_v; a.p = view_of(pts);
    a.v_out = v_out;
    const size_t smem = smem_vnet_fwd(m->d);
    if (XW_SET_SMEM((xw::k_vnet_points<kHV, 0>), smem)) return 1;
    XW_LAUNCH((xw::k_vnet_points<kHV, 0>), grid_for((long long)n * L, 128, 8), 128, smem, stream, a);
    return XW_CHECK_LAUNCH("k_vnet_points<eval>");
}

int xw_interior_forward(const xw_dims* m, const xw_domain* dom, const xw_coef* coef, const float* theta_u,
                        const float* theta_v, const float* x, long long x_sn, const float* times, int L,
                        const xw_points* xv, const float* h, const float* grad_h, const float* f, int n,
                        double* sums, float* cot_u, float* cot_v, float* u_out, void* workspace,
                        size_t workspace_bytes, void* stream, const float* s0, float* vcache, int vcache_mode,
                        float* y_hist, size_t vcache_floats, size_t y_hist_floats) {
    if (check_dims(m)) return 1;
    if (!device()) return fail("no CUDA device");
    if (n < 1 || L < 1) return fail("empty batch (n=%d, L=%d)", n, L);
    if (!dom || !coef || !theta_u || !theta_v || !x || !times || !xv || !xv->t || !xv->x || !h || !grad_h || !f ||
        !sums || !cot_u || !cot_v || !workspace)
        return fail("NULL pointer argument");
    if (dom->kind < 0 || dom->kind > 2) return fail("unknown domain kind %d", dom->kind);
    if ((coef->A_val == nullptr) != (coef->A_der == nullptr)) return fail("xw_coef: A_val and A_der go together");
    if (coef->a_sn < 0 || coef->b_sn < 0) return fail("xw_coef: negative path stride");
    if (vcache_mode < 0 || vcache_mode > 2 || (vcache_mode != 0 && !vcache)) return fail("bad vcache arguments");
    if (vcache_mode != 0 && vcache_floats < xw_vcache_floats(m, n, L))
        return fail("test-function cache too small: %zu floats < %zu for n=%d, L=%d", vcache_floats, xw_vcache_floats(m, n, L), n, L);
    if (y_hist && y_hist_floats < xw_yhist_floats(m, n, L))
        return fail("state-history buffer too small: %zu floats < %zu for n=%d, L=%d", y_hist_floats, xw_yhist_floats(m, n, L), n, L);
    float* gcache = vcache ? vcache + (size_t)4 * n * L : nullptr;
    const bool x2 = use_x2(m);
    const int gf = grid_for(n, kBlkFwd, 8);
    const size_t du_b = align_up((size_t)n * m->d * 4, 256), u_b = align_up((size_t)n * L * 4, 256);
    const size_t hist_b = x2 ? x2_rec_bytes(m, n, L) : align_up(fwd_hist_floats(m, L) * gf * kBlkFwd * 4, 256);
    if (workspace_bytes < du_b + u_b + hist_b) return fail("workspace too small: %zu < %zu", workspace_bytes, du_b + u_b + hist_b);
    char* ws = (char*)workspace;
    float* du = (float*)ws;
    float* ubuf = u_out ? u_out : (float*)(ws + du_b);
    float* yhist = (float*)(ws + du_b + u_b);

    if (x2) {
        xw::x2::FwdArgs q{};
        q.d = m->d; q.Hr = m->H; q.HHr = m->hh; q.nsh = m->nu - 1; q.L = L; q.n = n;
        q.theta = theta_u; q.x = x; q.x_sn = x_sn; q.times = times; q.s0 = s0 ? s0 : h; q.u_out = ubuf;
        q.grad_h = grad_h; q.du_out = du; q.rec = yhist; q.sums = sums; q.hloss = h; q.zq = y_hist;
        if (x2_launch_fwd<1>(m, q, stream)) return 1;
    } else {
        xw::XnodeFwdArgs a{};
        a.d = m->d; a.Hr = m->H; a.HHr = m->hh; a.nsh = m->nu - 1; a.L = L; a.n = n;
        a.theta = theta_u; a.x = x; a.x_sn = x_sn; a.times = times; a.s0 = s0 ? s0 : h; a.u_out = ubuf;
        a.grad_h = grad_h; a.du_out = du; a.yhist = yhist; a.sums = sums; a.hloss = h; a.ypath = y_hist;
        g_last_xnode_impl = 1;
        if (launch_xnode_fwd<1>(m, a, gf, smem_xnode_fwd(m->d, L), stream)) return 1;
    }

    xw::VnetFwdArgs b{};
    b.d = m->d; b.Hvr = m->Hv; b.nv = m->nv; b.n = n; b.L = L; b.theta = theta_v; b.p = view_of(xv);
    b.dom_kind = dom->kind; b.dp0 = dom->p0; b.dp1 = dom->p1; b.dp2 = dom->p2;
    b.c0 = coef->c0; b.c1 = coef->c1; b.ca = coef->a; b.cb = coef->b; b.ca_sn = coef->a_sn; b.cb_sn = coef->b_sn;
    b.u = ubuf; b.du = du; b.h = h; b.f = f; b.sums = sums; b.cot_u = cot_u; b.cot_v = cot_v; b.v_out = nullptr;
    b.gcache = vcache_mode == 1 ? gcache : nullptr;
    if (vcache_mode == 2) {          // sample and theta_v unchanged: reuse the cached test-function values
        xw::CombineArgs q{};
        q.d = m->d; q.n = n; q.L = L; q.c0 = coef->c0; q.c1 = coef->c1; q.ca = coef->a; q.cb = coef->b;
        q.ca_sn = coef->a_sn; q.cb_sn = coef->b_sn; q.Aval = coef->A_val; q.Ader = coef->A_der;
        q.vcache = vcache; q.gcache = gcache; q.u = ubuf; q.du = du; q.h = h; q.f = f;
        q.sums = sums; q.cot_u = cot_u; q.cot_v = cot_v;
        XW_LAUNCH(xw::k_weak_combine, grid_for((long long)n * L, 256, 8), 256, 4 * 32 * 8, stream, q);
        return XW_CHECK_LAUNCH("k_weak_combine");
    }
    const size_t smem = smem_vnet_fwd(m->d);
    // generation 2: time-row-0 gradient term (one thread per path) + CTA-tiled pass over all points
    bool row0_done = false;
#ifndef XW_EMU
    if (use_tc_kernels() && vtc_bwd_ok(m) && m->nv <= xw::kMaxNv) {
        // time-row 0 on the tensor cores (reverse mode for grad_x v), 128 paths per tile, three tile streams per CTA
        const int kin = xw::tc::kin_of(m->d), nvs = std::max(m->nv, 1);
        size_t sm0 = (size_t)(4 * xw::tc::KP * xw::tc::NP + 2 * kin * xw::tc::NP + 64 + 2 * xw::tc::KP * kin) * 4 +
                     (size_t)3 * nvs * 128 * 8 + 4 * 32 * 8 + 128;
        sm0 = std::max(sm0, (size_t)(device()->smem_optin / 2 + 1024));        // (one CTA per SM: it owns the tensor memory)
        if (sm0 <= device()->smem_optin) {
            if (XW_SET_SMEM(xw::tc::k_vnet_tc_row0, sm0)) return 1;
            const long long nt0 = ((long long)n + 127) / 128;
            const int g0 = (int)std::max<long long>(1, std::min<long long>((nt0 + 2) / 3, (long long)device()->sms));
            xw::tc::k_vnet_tc_row0<<<g0, 384, sm0, (cudaStream_t)stream>>>(b);
            if (XW_CHECK_LAUNCH("k_vnet_tc_row0")) return 1;
            row0_done = true;
        }
    }
#endif
    if (!row0_done) {
        if (XW_SET_SMEM((xw::k_vnet_points<kHV, 2>), smem)) return 1;
        XW_LAUNCH((xw::k_vnet_points<kHV, 2>), grid_for(n, 128, 8), 128, smem, stream, b);
        if (XW_CHECK_LAUNCH("k_vnet_points<row0>")) return 1;
    }
    xw::VtileFwdArgs t{};
    t.d = m->d; t.Hvr = m->Hv; t.nv = m->nv; t.n = n; t.L = L; t.theta = theta_v; t.p = view_of(xv);
    t.dom_kind = dom->kind; t.dp0 = dom->p0; t.dp1 = dom->p1; t.dp2 = dom->p2;
    t.c0 = coef->c0; t.c1 = coef->c1; t.Aval = coef->A_val; t.Ader = coef->A_der; t.u = ubuf; t.h = h; t.f = f; t.sums = sums; t.cot_u = cot_u; t.cot_v = cot_v;
    t.vcache = vcache_mode == 1 ? vcache : nullptr;
#ifndef XW_EMU
    xw_dims mvirt = *m;
    if (use_tc_kernels() && vv_wanted(m)) {
        // d > 54: the tiled pass runs on the VIRTUAL net of input width Hv (y_n = Wx x_n per path, xw_vnet_virtual.cuh):
        // the real input width would leave room for ONE tile stream in tensor memory instead of three
        char* wv = ws + du_b + u_b + hist_b;
        if (workspace_bytes < du_b + u_b + hist_b + align_up((size_t)n * m->Hv * 4, 256) + align_up((size_t)xw::VLayout(m->Hv, m->Hv).size * 4, 256) + 2 * align_up((size_t)n * L * 4, 256))
            return fail("workspace too small for the wide-input forward");
        float* y = (float*)wv;                   wv += align_up((size_t)n * m->Hv * 4, 256);
        float* thv = (float*)wv;                 wv += align_up((size_t)xw::VLayout(m->Hv, m->Hv).size * 4, 256);
        float* wbuf = (float*)wv;                wv += align_up((size_t)n * L * 4, 256);
        float* dwtbuf = (float*)wv;
        xw::vv::PrepArgs pa{};
        pa.d = m->d; pa.Hvr = m->Hv; pa.n = n; pa.L = L; pa.theta = theta_v; pa.p = view_of(xv);
        pa.dom_kind = dom->kind; pa.dp0 = dom->p0; pa.dp1 = dom->p1; pa.dp2 = dom->p2;
        pa.y = y; pa.wbuf = wbuf; pa.dwtbuf = dwtbuf; pa.theta_virtual = thv;
        const size_t sm_prep = (size_t)m->d * xw::vv::HVP * 4;
        if (sm_prep > device()->smem_optin) return fail("dim %d too large for the wide-input test-function forward", m->d);
        if (XW_SET_SMEM(xw::vv::k_vv_prep, sm_prep)) return 1;
        const int gp = (int)std::max<long long>(1, std::min<long long>(((long long)n + 127) / 128, (long long)device()->sms * 4));
        xw::vv::k_vv_prep<<<gp, 128, sm_prep, (cudaStream_t)stream>>>(pa);
        if (XW_CHECK_LAUNCH("k_vv_prep")) return 1;
        mvirt.d = m->Hv;
        t.d = mvirt.d; t.theta = thv; t.wbuf = wbuf; t.dwtbuf = dwtbuf;
        t.p.x = y; t.p.x_sn = m->Hv; t.p.x_sl = 0;
        m = &mvirt;
    }
    if (use_tc_kernels() && xw::tc::kin_of(m->d) <= 224) {
        // generation 3: hidden-layer contractions on the tensor cores (tcgen05, 3xTF32), 64 points per tile
        const int kin = xw::tc::kin_of(m->d), KA = std::max(kin, xw::tc::KP);
        // one CTA per SM owns the 512 tensor-memory columns; each warpgroup needs 56 + 2 KA of them
        const int ng = std::max(1, std::min(3, 512 / (xw::tc::NP + 2 * KA)));
        size_t sm = (size_t)(2 * xw::tc::KP * xw::tc::NP + 2 * kin * xw::tc::NP + 64) * 4 + 4 * 32 * 8 + 128;
        sm = std::max(sm, (size_t)(device()->smem_optin / 2 + 1024));          // (keeps a second CTA off the SM)
        const long long nt = ((long long)n * L + 63) / 64;
        const int g3 = (int)std::max<long long>(1, std::min<long long>((nt + ng - 1) / ng, (long long)device()->sms));
        if (tc_split_issue()) {
            if (XW_SET_SMEM(xw::tc::k_vnet_tc_fwd<true>, sm)) return 1;
            xw::tc::k_vnet_tc_fwd<true><<<g3, 128 * ng, sm, (cudaStream_t)stream>>>(t, ng);
        } else {
            if (XW_SET_SMEM(xw::tc::k_vnet_tc_fwd<false>, sm)) return 1;
            xw::tc::k_vnet_tc_fwd<false><<<g3, 128 * ng, sm, (cudaStream_t)stream>>>(t, ng);
        }
        g_last_vnet_fwd = 3;
        return XW_CHECK_LAUNCH("k_vnet_tc_fwd");
    }
#endif
    using VT = xw::VTile<kHV, kQR, kNH>;
    const size_t tsmem = smem_vtile_fwd(m->d);
    if (tsmem > device()->smem_optin) return fail("tiled v-net forward needs %zu B shared memory (> %zu): dim too large", tsmem, device()->smem_optin);
    if (XW_SET_SMEM((xw::k_vnet_tile_fwd<kHV, kQR, kNH>), tsmem)) return 1;
    const long long ntiles = ((long long)n * L + VT::ROWS - 1) / VT::ROWS;
    const int tgrid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)device()->sms * ctas_per_sm_for(tsmem, 6)));
    XW_LAUNCH((xw::k_vnet_tile_fwd<kHV, kQR, kNH>), tgrid, VT::THREADS, tsmem, stream, t);
    g_last_vnet_fwd = 2;
    return XW_CHECK_LAUNCH("k_vnet_tile_fwd");
}

int xw_boundary_u(const xw_dims* m, const float* theta_u, const float* xb, long long xb_sn, const float* times_b,
                  int Lb, const float* s0b, const float* g, int nb, double gscale, double* sums, float* grad_u,
                  int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    if (check_dims(m)) return 1;
    if (!device()) return fail("no CUDA device");
    if (nb < 1 || Lb < 1) return fail("empty batch (n=%d, L=%d)", nb, Lb);
    if (!theta_u || !xb || !times_b || !s0b || !g || !sums || !workspace) return fail("NULL pointer argument");
    if (use_x2(m)) {
        X2BwdPlan p2;
        if (x2_plan_bwd(m, nb, Lb, &p2)) return 1;
        if (workspace_bytes < x2_bwd_bytes(p2)) return fail("workspace too small: %zu < %zu", workspace_bytes, x2_bwd_bytes(p2));
        xw::x2::BwdArgs q{};
        q.d = m->d; q.Hr = m->H; q.HHr = m->hh; q.nsh = m->nu - 1; q.L = Lb; q.n = nb;
        q.theta = theta_u; q.x = xb; q.x_sn = xb_sn; q.times = times_b; q.s0 = s0b; q.cot = g; q.gscale = gscale; q.sums = sums;
        return x2_run_bwd<1>(m, q, p2, workspace, grad_u, accumulate, stream);
    }
    g_last_xnode_impl = 1; g_last_xnode_bwd = 1;
    XnodeBwdPlan p;
    if (plan_xnode_bwd(m, nb, Lb, &p)) return 1;
    if (workspace_bytes < p.hist_bytes + p.part_bytes) return fail("workspace too small: %zu < %zu", workspace_bytes, p.hist_bytes + p.part_bytes);
    xw::XnodeBwdArgs a{};
    a.d = m->d; a.Hr = m->H; a.HHr = m->hh; a.nsh = m->nu - 1; a.L = Lb; a.n = nb;
    a.theta = theta_u; a.x = xb; a.x_sn = xb_sn; a.times = times_b; a.s0 = s0b; a.cot = g; a.coefs = nullptr;
    a.gscale = gscale; a.yhist = (float*)workspace; a.gpart = (float*)((char*)workspace + p.hist_bytes); a.sums = sums;
    if (launch_xnode_bwd<1>(m, a, p.grid, p.block, p.smem, stream)) return 1;
    if (grad_u) return reduce_partials(a.gpart, p.grid, xw_theta_u_size(m), grad_u, accumulate, stream);
    return 0;
}

int xw_interior_backward_u(const xw_dims* m, const float* theta_u, const float* x, long long x_sn, const float* times,
                           int L, const float* h, const float* cot_u, int n, const double* coefs_dev, float* grad_u,
                           int accumulate, void* workspace, size_t workspace_bytes, void* stream, const float* s0,
                           const float* y_hist) {
    if (check_dims(m)) return 1;
    if (!device()) return fail("no CUDA device");
    if (n < 1 || L < 1) return fail("empty batch (n=%d, L=%d)", n, L);
    if (!theta_u || !x || !times || !h || !cot_u || !coefs_dev || !grad_u || !workspace) return fail("NULL pointer argument");
    if (use_x2(m)) {
        X2BwdPlan p2;
        if (x2_plan_bwd(m, n, L, &p2)) return 1;
        if (workspace_bytes < x2_bwd_bytes(p2)) return fail("workspace too small: %zu < %zu", workspace_bytes, x2_bwd_bytes(p2));
        xw::x2::BwdArgs q{};
        q.d = m->d; q.Hr = m->H; q.HHr = m->hh; q.nsh = m->nu - 1; q.L = L; q.n = n;
        q.theta = theta_u; q.x = x; q.x_sn = x_sn; q.times = times; q.s0 = s0 ? s0 : h; q.hloss = h; q.cot = cot_u;
        q.coefs = coefs_dev; q.zq = y_hist;
        return x2_run_bwd<0>(m, q, p2, workspace, grad_u, accumulate, stream);
    }
    g_last_xnode_impl = 1; g_last_xnode_bwd = 1;
    XnodeBwdPlan p;
    if (plan_xnode_bwd(m, n, L, &p)) return 1;
    if (workspace_bytes < p.hist_bytes + p.part_bytes) return fail("workspace too small: %zu < %zu", workspace_bytes, p.hist_bytes + p.part_bytes);
    xw::XnodeBwdArgs a{};
    a.d = m->d; a.Hr = m->H; a.HHr = m->hh; a.nsh = m->nu - 1; a.L = L; a.n = n;
    a.theta = theta_u; a.x = x; a.x_sn = x_sn; a.times = times; a.s0 = s0 ? s0 : h; a.hloss = h; a.cot = cot_u; a.coefs = coefs_dev; a.ypath = y_hist;
    a.gscale = 0.0; a.yhist = (float*)workspace; a.gpart = (float*)((char*)workspace + p.hist_bytes); a.sums = nullptr;
    if (launch_xnode_bwd<0>(m, a, p.grid, p.block, p.smem, stream)) return 1;
    return reduce_partials(a.gpart, p.grid, xw_theta_u_size(m), grad_u, accumulate, stream);
}

int xw_interior_backward_v(const xw_dims* m, const xw_domain* dom, const float* theta_v, const xw_points* xv,
                           const float* cot_v, int n, int L, const double* coefs_dev, float* grad_v, int accumulate,
                           void* workspace, size_t workspace_bytes, void* stream) {
    if (check_dims(m)) return 1;
    if (!device()) return fail("no CUDA device");
    if (n < 1 || L < 1) return fail("empty batch (n=%d, L=%d)", n, L);
    if (!dom || !theta_v || !xv || !xv->t || !xv->x || !cot_v || !coefs_dev || !grad_v || !workspace)
        return fail("NULL pointer argument");
#ifndef XW_EMU
    if (use_tc_kernels() && vtc_bwd_ok(m)) {
        VtileBwdPlan pl;
        if (plan_vtc_bwd(m, n, L, &pl)) return 1;
        if (workspace_bytes < pl.scratch_bytes + pl.part_bytes) return fail("workspace too small: %zu < %zu", workspace_bytes, pl.scratch_bytes + pl.part_bytes);
        xw::VtileBwdArgs t{};
        t.d = m->d; t.Hvr = m->Hv; t.nv = m->nv; t.n = n; t.L = L; t.theta = theta_v; t.p = view_of(xv);
        t.dom_kind = dom->kind; t.dp0 = dom->p0; t.dp1 = dom->p1; t.dp2 = dom->p2;
        t.cot = cot_v; t.coefs = coefs_dev; t.scratch = (float*)workspace; t.gpart = (float*)((char*)workspace + pl.scratch_bytes);
        t.tm_packed = tc_tmem_packed(); t.flush_tiles = tc_flush_tiles();
        if (tc_split_bwd()) {
            if (XW_SET_SMEM(xw::tc::k_vnet_tc_bwd3<true>, pl.smem)) return 1;
            xw::tc::k_vnet_tc_bwd3<true><<<pl.grid, 512, pl.smem, (cudaStream_t)stream>>>(t);
        } else {
            if (XW_SET_SMEM(xw::tc::k_vnet_tc_bwd3<false>, pl.smem)) return 1;
            xw::tc::k_vnet_tc_bwd3<false><<<pl.grid, 512, pl.smem, (cudaStream_t)stream>>>(t);
        }
        if (XW_CHECK_LAUNCH("k_vnet_tc_bwd3")) return 1;
        g_last_vnet_bwd = 3;
        return reduce_partials(t.gpart, pl.grid, xw_theta_v_size(m), grad_v, accumulate, stream);
    }
    if (use_tc_kernels() && vv_wanted(m)) {
        // d > 54: tensor-core backward on the virtual net of input width Hv (y_n = Wx x_n per path), see xw_vnet_virtual.cuh
        VvPlan vp;
        if (plan_vv(m, n, L, &vp)) return 1;
        if (workspace_bytes < vp.total) return fail("workspace too small: %zu < %zu", workspace_bytes, vp.total);
        char* ws = (char*)workspace;
        float* scratch = (float*)ws;                       ws += vp.pl.scratch_bytes;
        float* gpart = (float*)ws;                         ws += vp.pl.part_bytes;
        float* y = (float*)ws;                             ws += vp.y_b;
        float* thv = (float*)ws;                           ws += vp.th_b;
        float* gradv = (float*)ws;                         ws += vp.gv_b;
        float* wbuf = (float*)ws;                          ws += vp.w_b;
        float* d0 = (float*)ws;                            ws += vp.d0_b;
        float* px = (float*)ws;
        xw::vv::PrepArgs pa{};
        pa.d = m->d; pa.Hvr = m->Hv; pa.n = n; pa.L = L; pa.theta = theta_v; pa.p = view_of(xv);
        pa.dom_kind = dom->kind; pa.dp0 = dom->p0; pa.dp1 = dom->p1; pa.dp2 = dom->p2;
        pa.y = y; pa.wbuf = wbuf; pa.theta_virtual = thv;
        const size_t sm_prep = (size_t)m->d * xw::vv::HVP * 4;
        if (sm_prep > device()->smem_optin) return fail("dim %d too large for the wide-input test-function backward", m->d);
        if (XW_SET_SMEM(xw::vv::k_vv_prep, sm_prep)) return 1;
        xw::vv::k_vv_prep<<<vp.grid_prep, 128, sm_prep, (cudaStream_t)stream>>>(pa);
        if (XW_CHECK_LAUNCH("k_vv_prep")) return 1;
        xw::VtileBwdArgs t{};
        t.d = vp.mv.d; t.Hvr = m->Hv; t.nv = m->nv; t.n = n; t.L = L; t.theta = thv;
        t.p.t = xv->t; t.p.t_sn = xv->t_sn; t.p.t_sl = xv->t_sl; t.p.x = y; t.p.x_sn = m->Hv; t.p.x_sl = 0;
        t.dom_kind = dom->kind; t.dp0 = dom->p0; t.dp1 = dom->p1; t.dp2 = dom->p2;
        t.cot = cot_v; t.coefs = coefs_dev; t.scratch = scratch; t.gpart = gpart; t.wbuf = wbuf; t.delta0_out = d0;
        t.tm_packed = tc_tmem_packed(); t.flush_tiles = tc_flush_tiles();
        if (tc_split_bwd()) {
            if (XW_SET_SMEM(xw::tc::k_vnet_tc_bwd3<true>, vp.pl.smem)) return 1;
            xw::tc::k_vnet_tc_bwd3<true><<<vp.pl.grid, 512, vp.pl.smem, (cudaStream_t)stream>>>(t);
        } else {
            if (XW_SET_SMEM(xw::tc::k_vnet_tc_bwd3<false>, vp.pl.smem)) return 1;
            xw::tc::k_vnet_tc_bwd3<false><<<vp.pl.grid, 512, vp.pl.smem, (cudaStream_t)stream>>>(t);
        }
        if (XW_CHECK_LAUNCH("k_vnet_tc_bwd3")) return 1;
        if (reduce_partials(gpart, vp.pl.grid, vp.Pv, gradv, 0, stream)) return 1;
        xw::vv::DwxArgs da{};
        da.d = m->d; da.Hvr = m->Hv; da.n = n; da.L = L; da.delta0 = d0; da.x = xv->x; da.x_sn = xv->x_sn; da.part = px;
        const size_t sm_dwx = (size_t)(xw::vv::kDwxPaths * xw::vv::HVP + xw::vv::kDwxPaths * m->d) * 4;
        if (m->d > 32 * xw::vv::kDwxJ) return fail("dim %d too large for the wide-input test-function backward (max %d)", m->d, 32 * xw::vv::kDwxJ);
        xw::vv::k_vv_dwx<<<vp.grid_dwx, 256, sm_dwx, (cudaStream_t)stream>>>(da);
        if (XW_CHECK_LAUNCH("k_vv_dwx")) return 1;
        const int P = xw_theta_v_size(m);
        xw::vv::k_vv_finish<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gradv, px, vp.grid_dwx, m->d, m->Hv, grad_v, accumulate);
        g_last_vnet_bwd = 4;
        return XW_CHECK_LAUNCH("k_vv_finish");
    }
#endif
    {
        VtileBwdPlan pl;
        if (plan_vtile_bwd(m, n, L, &pl)) return 1;
        if (workspace_bytes < pl.scratch_bytes + pl.part_bytes) return fail("workspace too small: %zu < %zu", workspace_bytes, pl.scratch_bytes + pl.part_bytes);
        xw::VtileBwdArgs t{};
        t.d = m->d; t.Hvr = m->Hv; t.nv = m->nv; t.n = n; t.L = L; t.theta = theta_v; t.p = view_of(xv);
        t.dom_kind = dom->kind; t.dp0 = dom->p0; t.dp1 = dom->p1; t.dp2 = dom->p2;
        t.cot = cot_v; t.coefs = coefs_dev; t.scratch = (float*)workspace; t.gpart = (float*)((char*)workspace + pl.scratch_bytes);
        using VT = xw::VTile<kHV, kQRB>;
        if (XW_SET_SMEM((xw::k_vnet_tile_bwd<kHV, kQRB>), pl.smem)) return 1;
        XW_LAUNCH((xw::k_vnet_tile_bwd<kHV, kQRB>), pl.grid, VT::THREADS, pl.smem, stream, t);
        if (XW_CHECK_LAUNCH("k_vnet_tile_bwd")) return 1;
        g_last_vnet_bwd = 2;
        return reduce_partials(t.gpart, pl.grid, xw_theta_v_size(m), grad_v, accumulate, stream);
    }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// FP32 FMA micro-benchmark (roofline denominator)
// ---------------------------------------------------------------------------------------------
#ifndef XW_EMU
namespace {
template <int VARIANT>
__global__ void __launch_bounds__(256) k_fma_probe(int iters, float* sink, float seed) {
    // 16 independent accumulator chains per thread
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
    const float b = 1.0000001f, c = 1e-7f;
    if (VARIANT == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        }
    } else {
        // packed FFMA2 (fma.rn.f32x2, sm_100+)
        unsigned long long p[8], pb, pc;
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}
// legacy warp-level tensor-core path (mma.sync m16n8k8 TF32): 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) k_mma_probe_tf32(int iters, float* sink) {
    float c[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
    unsigned a0 = 0x3f800000u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3f000000u + threadIdx.x, b1 = b0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
    if (s == 123.456f) sink[0] = s;
}
}  // namespace
#endif

extern "C" int xw_fma_probe(int variant, int iters, double* flops_host, void* stream) {
#ifdef XW_EMU
    (void)variant; (void)iters; (void)flops_host; (void)stream;
    return fail("xw_fma_probe needs a CUDA device");
#else
    const Dev* dv = device();
    if (!dv) return fail("no CUDA device");
    static float* sink = nullptr;
    if (!sink && cudaMalloc(&sink, 256) != cudaSuccess) return fail("cudaMalloc failed");
    const int grid = dv->sms * 8, block = 256;
    if (variant == 2) {
        k_mma_probe_tf32<<<grid, block, 0, (cudaStream_t)stream>>>(iters, sink);
        if (flops_host) *flops_host = 2.0 * 16 * 8 * 8 * 8 * (double)iters * (double)grid * (block / 32);
        return check_launch("k_mma_probe_tf32");
    }
    if (variant == 0) k_fma_probe<0><<<grid, block, 0, (cudaStream_t)stream>>>(iters, sink, 1.f);
    else k_fma_probe<1><<<grid, block, 0, (cudaStream_t)stream>>>(iters, sink, 1.f);
    if (flops_host) *flops_host = 2.0 * 16 * 8 * (double)iters * (double)grid * block;
    return check_launch("k_fma_probe");
#endif
}

// tcgen05 probe: D[128 x N] = A[128 x K] * B[N x K]^T on the 5th-gen tensor cores (kind::tf32),
// terms = 1 (plain TF32) or 3 (3xTF32 error-compensated).  err_dev[0] != 0: a bounded wait expired.
extern "C" int xw_umma_probe(const float* A, const float* B, float* D, int K, int N, int terms, int* err_dev, void* stream) {
#ifdef XW_EMU
    (void)A; (void)B; (void)D; (void)K; (void)N; (void)terms; (void)err_dev; (void)stream;
    return fail("xw_umma_probe needs a CUDA device");
#else
    if (terms < 0) {   // transposed (weight-gradient shaped) product: A = P[128 x K], B = Q[128 x N], D[128 x N] = P^T Q (rows >= K are zero)
        terms = -terms;
        if (K % 4 || K < 4 || K > 128 || N % 8 || N < 8 || N > 64 || (terms != 1 && terms != 3)) return fail("bad probe shape");
        const size_t smem_t = (size_t)(2 * 32 * 128 * 4 + 2 * N * 128) * 4 + 64;
        if (XW_SET_SMEM(xw::umma::k_umma_probe_t, smem_t)) return 1;
        xw::umma::k_umma_probe_t<<<1, 128, smem_t, (cudaStream_t)stream>>>(A, B, D, K, N, terms, err_dev, getenv("XW_UMMA_VARIANT") ? atoi(getenv("XW_UMMA_VARIANT")) : 0);
        return check_launch("k_umma_probe_t");
    }
    const int ts = terms >= 10;                        /* 11 / 13: A operand from tensor memory */
    if (ts) terms -= 10;
    if (K % 8 || K < 8 || K > 64 || N % 8 || N < 8 || N > 64 || (terms != 1 && terms != 3)) return fail("bad probe shape");
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * 4 + 64;
    if (XW_SET_SMEM(xw::umma::k_umma_probe, smem)) return 1;
    xw::umma::k_umma_probe<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, terms, err_dev, ts, getenv("XW_UMMA_M64") ? 1 : 0);
    return check_launch("k_umma_probe");
#endif
}

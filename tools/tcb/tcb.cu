// tcb.cu -- development harness: k_vnet_tc_bwd3 alone in a small library (compiles in well under a minute; the product
// library is one translation unit of ~3 minutes).  Same kernel source as the product (xw_vnet_tc.cuh); exports a launch
// entry with xw_interior_backward_v's meaning minus the final reduction, plus the XW_TC_PROF counters.
#include "../../include/xnode_wan_b200.h"
#include "../../xnode-wan-pde-solver_b200/csrc/xw_kernels.cuh"
#include "../../xnode-wan-pde-solver_b200/csrc/xw_umma.cuh"
#include "../../xnode-wan-pde-solver_b200/csrc/xw_vnet_tc.cuh"
#include <cstdio>
#include <cstdlib>
#ifndef XW_TC_BWD_SPLIT
#define XW_TC_BWD_SPLIT 0
#endif

extern "C" size_t tcb_workspace_bytes(int d, int Hv, int nv, int sms) {
    return (size_t)sms * 2 * (nv > 0 ? nv : 1) * 14 * 128 * 16 + (size_t)sms * xw::VLayout(d, Hv).size * 4 + 512;
}
extern "C" int tcb_run(int d, int Hv, int nv, int n, int L, const float* theta, const float* t, long long t_sn, long long t_sl,
                       const float* x, long long x_sn, long long x_sl, const float* cot, const double* coefs, void* ws,
                       int packed, int flush_tiles, float* grad_out, void* stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int kin = xw::tc::kin_of(d);
    const size_t smem = (size_t)(4 * xw::tc::KP * xw::tc::NP + 2 * kin * xw::tc::NP + 64 + 2 * xw::tc::TIMG2 + xw::tc::KP * (xw::tc::KP + kin + 1) + 512 + 64 + 256) * 4 + 128;
    const long long ntiles = ((long long)n * L + 127) / 128;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    xw::VtileBwdArgs a{};
    a.d = d; a.Hvr = Hv; a.nv = nv; a.n = n; a.L = L; a.theta = theta;
    a.p.t = t; a.p.t_sn = t_sn; a.p.t_sl = t_sl; a.p.x = x; a.p.x_sn = x_sn; a.p.x_sl = x_sl;
    a.dom_kind = 0; a.dp0 = -1.f; a.dp1 = 1.f; a.dp2 = 0.f;
    a.cot = cot; a.coefs = coefs; a.scratch = (float*)ws;
    a.gpart = (float*)((char*)ws + (size_t)grid * 2 * (nv > 0 ? nv : 1) * 14 * 128 * 16);
    a.tm_packed = packed; a.flush_tiles = flush_tiles;
    if (cudaFuncSetAttribute(xw::tc::k_vnet_tc_bwd3<XW_TC_BWD_SPLIT != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    xw::tc::k_vnet_tc_bwd3<XW_TC_BWD_SPLIT != 0><<<grid, 512, smem, (cudaStream_t)stream>>>(a);
    if (grad_out) {
        const int P = xw::VLayout(d, Hv).size;
        xw::k_reduce_partials<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a.gpart, grid, P, grad_out, 0);
    }
    return cudaGetLastError() != cudaSuccess;
}
#ifdef XW_TC_PROF
extern "C" int tcb_prof_read(unsigned long long* out_host) {
    unsigned long long z[96] = {0};
    if (cudaMemcpyFromSymbol(out_host, xw::tc::g_tc_prof, sizeof(z)) != cudaSuccess) return 1;
    return cudaMemcpyToSymbol(xw::tc::g_tc_prof, z, sizeof(z)) != cudaSuccess;
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// tcgen05.mma issue-rate microbenchmark: `count` MMAs of one shape from one thread, K = 8 (one tf32 k-step) each,
// round-robin over `nacc` accumulators (nacc = 1: every MMA depends on the previous one's accumulator).
// mode 0: A from tensor memory (TS), mode 1: A from shared memory (SS).  Prints cycles per MMA.
// ---------------------------------------------------------------------------------------------------------------
namespace {
template <int MODE, int NACC>
__global__ void __launch_bounds__(128) k_mma_bench(int M, int N, int batches, int a_rows, long long* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* a_img = reinterpret_cast<float*>(smem_raw);                 // [2 chunks][a_rows][4] K-major, 16 KB reserved
    float* b_img = a_img + 4096;                                        // [2 chunks][N][4]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b_img + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += 128) a_img[i] = 0.f;
    if (tid == 0) xw::umma::mbar_init(mbar, 1);
    xw::umma::fence_smem_to_async();
    if (warp == 0) xw::umma::tmem_alloc(slot, 512);
    xw::umma::fence_before();
    __syncthreads();
    xw::umma::fence_after();
    const uint32_t tbase = *slot;
    if (tid == 0) {
        const uint32_t idesc = xw::umma::idesc_tf32(M, N);
        const uint64_t da = xw::umma::smem_desc(a_img, a_rows * 16, 128), db = xw::umma::smem_desc(b_img, N * 16, 128);
        const uint32_t a_tm = tbase + 448;                              // A operand columns [448, 456)
        uint32_t dacc[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) dacc[q] = tbase + (uint32_t)(q * N);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int b = 0; b < batches; ++b) {
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    if (MODE == 0) xw::umma::mma_tf32_ts(dacc[i % NACC], a_tm, db, idesc, 1u);
                    else xw::umma::mma_tf32(dacc[i % NACC], da, db, idesc, 1u);
                }
            }
            const long long t1 = clock64();
            xw::umma::commit(mbar);
            xw::umma::mbar_wait(mbar, rep & 1);
            const long long t2 = clock64();
            out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0;
        }
    }
    xw::umma::fence_before();
    __syncthreads();
    if (warp == 0) xw::umma::tmem_free(tbase, 512);
}
}
extern "C" int tcb_mma_bench(int mode, int M, int N, int nacc, int count, int a_rows, long long* out_dev, void* stream) {
    const size_t smem = 8192 * 4 + 64;
    const int batches = count / 24;
#define TCB_CASE(MO, NA) if (mode == MO && nacc == NA) k_mma_bench<MO, NA><<<1, 128, smem, (cudaStream_t)stream>>>(M, N, batches, a_rows, out_dev);
    TCB_CASE(0, 1) TCB_CASE(0, 2) TCB_CASE(0, 3) TCB_CASE(0, 4) TCB_CASE(1, 1) TCB_CASE(1, 2) TCB_CASE(1, 3) TCB_CASE(1, 4)
    return cudaGetLastError() != cudaSuccess;
}

// several issuing warps at once (lane 0 of warps 0..nissue-1), each with its own accumulator and mbarrier: is the ~46-cycle
// minimum per MMA a property of the tensor pipe or of one issuing thread?
namespace {
__global__ void __launch_bounds__(128) k_mma_bench_multi(int M, int N, int nissue, int batches, long long* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* b_img = reinterpret_cast<float*>(smem_raw);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b_img + 8192);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8192; i += 128) b_img[i] = 0.f;
    if (tid == 0) for (int q = 0; q < 4; ++q) xw::umma::mbar_init(mbar + q, 1);
    xw::umma::fence_smem_to_async();
    if (warp == 0) xw::umma::tmem_alloc(slot, 512);
    xw::umma::fence_before();
    __syncthreads();
    xw::umma::fence_after();
    const uint32_t tbase = *slot;
    const long long t0 = clock64();
    if (lane == 0 && warp < nissue) {
        const uint32_t idesc = xw::umma::idesc_tf32(M, N);
        const uint64_t db = xw::umma::smem_desc(b_img, N * 16, 128);
        const uint32_t a_tm = tbase + 448, d = tbase + (uint32_t)(warp * N);
#pragma unroll 1
        for (int b = 0; b < batches; ++b) {
#pragma unroll
            for (int i = 0; i < 24; ++i) xw::umma::mma_tf32_ts(d, a_tm, db, idesc, 1u);
        }
        xw::umma::commit(mbar + warp);
        xw::umma::mbar_wait(mbar + warp, 0);
        out[warp] = clock64() - t0;
    }
    xw::umma::fence_before();
    __syncthreads();
    if (warp == 0) xw::umma::tmem_free(tbase, 512);
}
}
extern "C" int tcb_mma_bench_multi(int M, int N, int nissue, int count, long long* out_dev, void* stream) {
    k_mma_bench_multi<<<1, 128, 8192 * 4 + 128, (cudaStream_t)stream>>>(M, N, nissue, count / 24, out_dev);
    return cudaGetLastError() != cudaSuccess;
}

// xw_umma.cuh -- tcgen05 (5th-gen tensor core) building blocks and a self-checking probe.
//
// Groundwork for a tensor-core variant of the Hv x Hv contractions of the test-function net with
// 3xTF32 error compensation (a = a_hi + a_lo, a*b ~ a_hi*b_hi + a_lo*b_hi + a_hi*b_lo: fp32-level
// accuracy from kind::tf32 MMAs).  Operands live in shared memory in the canonical K-major,
// no-swizzle UMMA layout; the accumulator lives in TMEM and comes back with tcgen05.ld (one TMEM lane
// = one row = one thread).  Every wait is bounded: a wrong descriptor can give wrong numbers or an
// error flag, never a hang.
#pragma once
#ifndef XW_EMU
#include <cuda_runtime.h>
#include <stdint.h>

namespace xw {
namespace umma {

// canonical K-major / SWIZZLE_NONE operand tile: 16-byte chunks of 4 fp32 along K, 8-row core
// matrices of 128 contiguous bytes; chunk c, row r:  c * (ROWS*16) + (r/8)*128 + (r%8)*16
__device__ __forceinline__ int chunk_off_floats(int rows, int c, int r) { return c * rows * 4 + (r >> 3) * 32 + (r & 7) * 4; }

__device__ __forceinline__ uint64_t smem_desc(const void* p, int lbo_bytes, int sbo_bytes) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);                 // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // leading (K-direction) byte offset
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride (M/N-direction) byte offset
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    return d;                                            // layout type [61,64) = 0: no swizzle
}
// instruction descriptor: fp32 accumulate, tf32 x tf32; a_mn / b_mn = 1: that operand is MN-major
// (the 16-byte chunk runs along M / N and the 8 rows of a core matrix are 8 consecutive k)
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// 8 consecutive columns of this thread's TMEM lane  <-  registers
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}
// barrier among the 128 threads of one warpgroup (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
// bounded wait: returns false if the phase did not complete within `spins` probes
__device__ __forceinline__ bool mbar_wait(uint64_t* mbar, uint32_t parity, int spins = 1 << 22) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
    for (int i = 0; i < spins; ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int ncols) {      // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, int ncols) {       // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 56 consecutive columns of this thread's TMEM lane -> registers (one wait for the three loads)
__device__ __forceinline__ void tmem_ld56(uint32_t taddr, float (&v)[56]) {
    uint32_t r[56];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%56];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47}, [%57];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%48, %49, %50, %51, %52, %53, %54, %55}, [%58];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55])
                 : "r"(taddr), "r"(taddr + 32), "r"(taddr + 48) : "memory");
#pragma unroll
    for (int i = 0; i < 56; ++i) v[i] = __uint_as_float(r[i]);
}
// registers -> 56 consecutive columns of this thread's TMEM lane (complete after tmem_wait_st)
__device__ __forceinline__ void tmem_st56(uint32_t taddr, const uint32_t (&r)[56]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%56], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x16.b32 [%57], {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47};\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x8.b32 [%58], {%48, %49, %50, %51, %52, %53, %54, %55};"
                 :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]), "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]), "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]),
                 "r"(taddr), "r"(taddr + 32), "r"(taddr + 48) : "memory");
}

// 28 consecutive columns (half a 56-wide row) of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld28(uint32_t taddr, float (&v)[28]) {
    uint32_t r[28];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%28];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%29];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%24, %25, %26, %27}, [%30];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27])
                 : "r"(taddr), "r"(taddr + 16), "r"(taddr + 24) : "memory");
#pragma unroll
    for (int i = 0; i < 28; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st28(uint32_t taddr, const uint32_t (&r)[28]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%28], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x8.b32 [%29], {%16, %17, %18, %19, %20, %21, %22, %23};\n\t"
                 "tcgen05.st.sync.aligned.32x32b.x4.b32 [%30], {%24, %25, %26, %27};"
                 :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
                 "r"(taddr), "r"(taddr + 16), "r"(taddr + 24) : "memory");
}

// 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

// D[128 x N] = A[128 x K] * B[N x K]^T  (row-major fp32 in global memory), N multiple of 16 <= 64, K multiple of 8
// terms = 1: plain TF32, terms = 3: 3xTF32.  err[0] != 0 if a bounded wait expired.
__global__ void __launch_bounds__(128) k_umma_probe(const float* A, const float* B, float* D, int K, int N, int terms, int* err, int ts, int ts_m64) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* a_hi = reinterpret_cast<float*>(smem_raw);
    float* a_lo = a_hi + 128 * K;
    float* b_hi = a_lo + 128 * K;
    float* b_lo = b_hi + N * K;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + N * K);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c = 0; c < K / 4; ++c) {                    // thread = row of A
        float4 hi, lo;
        const float4 v = *reinterpret_cast<const float4*>(A + (size_t)tid * K + 4 * c);
        hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        *reinterpret_cast<float4*>(a_hi + chunk_off_floats(128, c, tid)) = hi;
        *reinterpret_cast<float4*>(a_lo + chunk_off_floats(128, c, tid)) = lo;
        if (tid < N) {
            const float4 w = *reinterpret_cast<const float4*>(B + (size_t)tid * K + 4 * c);
            hi.x = tf32_hi(w.x); hi.y = tf32_hi(w.y); hi.z = tf32_hi(w.z); hi.w = tf32_hi(w.w);
            lo.x = w.x - hi.x; lo.y = w.y - hi.y; lo.z = w.z - hi.z; lo.w = w.w - hi.w;
            *reinterpret_cast<float4*>(b_hi + chunk_off_floats(N, c, tid)) = hi;
            *reinterpret_cast<float4*>(b_lo + chunk_off_floats(N, c, tid)) = lo;
        }
    }
    if (tid == 0) mbar_init(mbar, 1);
    fence_smem_to_async();
    if (warp == 0) tmem_alloc(slot, 256);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = *slot;
    if (ts) {                                            // A (hi at columns 64.., lo at 128..) into tensor memory
        const uint32_t lane_base = tbase + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < K / 8; ++c) {
            float hi[8], lo[8];
            for (int e = 0; e < 8; ++e) { const float v = A[(size_t)tid * K + 8 * c + e]; hi[e] = tf32_hi(v); lo[e] = v - hi[e]; }
            tmem_st8(lane_base + 64 + 8 * c, hi);
            tmem_st8(lane_base + 128 + 8 * c, lo);
        }
        tmem_wait_st();
        fence_before();
        __syncthreads();
        fence_after();
    }
    if (tid == 0 && ts) {
        const uint32_t idesc = idesc_tf32(128, N);
        uint32_t acc = 0;
        for (int t = 0; t < terms; ++t)
            for (int ks = 0; ks < K / 8; ++ks) {
                mma_tf32_ts(tbase, tbase + (t == 1 ? 128 : 64) + 8 * ks,
                            smem_desc((t == 2 ? b_lo : b_hi) + (size_t)(2 * ks) * N * 4, N * 16, 128), idesc, acc);
                acc = 1;
            }
        commit(mbar);
    } else if (tid == 0) {
        const uint32_t idesc = idesc_tf32(ts_m64 ? 64 : 128, N);
        const int lbo_a = 128 * 16, lbo_b = N * 16, sbo = 128;
        uint32_t acc = 0;
        for (int t = 0; t < terms; ++t) {
            const float* pa = t == 1 ? a_lo : a_hi;
            const float* pb = t == 2 ? b_lo : b_hi;
            for (int ks = 0; ks < K / 8; ++ks) {
                mma_tf32(tbase, smem_desc(pa + (size_t)(2 * ks) * 128 * 4, lbo_a, sbo),
                         smem_desc(pb + (size_t)(2 * ks) * N * 4, lbo_b, sbo), idesc, acc);
                acc = 1;
            }
        }
        commit(mbar);
    }
    const bool ok = mbar_wait(mbar, 0);
    if (!ok && tid == 0) err[0] = 1;
    fence_after();
    if (ok) {
        for (int n0 = 0; n0 < N; n0 += 16) {
            float v[16];
            tmem_ld16(tbase + ((uint32_t)(32 * warp) << 16) + n0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) if (n0 + i < N) D[(size_t)tid * N + n0 + i] = v[i];
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, 256);
}

// transposed product on the SAME operand images: out[o][i] = sum_r P[r][o] Q[r][i]  (r = 0..127), the
// shape of the weight-gradient contraction.  P, Q are stored exactly as K-major A operands
// ([chunk of 4 units][row] with 16-byte cells) and read back as MN-major operands: unit chunk
// stride = SBO, 8-row group stride = 128 bytes (one MMA = one 8-row group).
__global__ void __launch_bounds__(128) k_umma_probe_t(const float* P, const float* Q, float* out, int MO, int NI, int terms, int* err, int variant) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* p_hi = reinterpret_cast<float*>(smem_raw);          // 32 chunks x 128 rows x 4 (M = 128: zero beyond MO)
    float* p_lo = p_hi + 32 * 128 * 4;
    float* q_hi = p_lo + 32 * 128 * 4;
    float* q_lo = q_hi + NI * 128;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(q_lo + NI * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c = 0; c < 32; ++c) {
        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
        if (4 * c < MO) {
            const float4 v = *reinterpret_cast<const float4*>(P + (size_t)tid * MO + 4 * c);
            hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
            lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        }
        *reinterpret_cast<float4*>(p_hi + chunk_off_floats(128, c, tid)) = hi;
        *reinterpret_cast<float4*>(p_lo + chunk_off_floats(128, c, tid)) = lo;
    }
    for (int c = 0; c < NI / 4; ++c) {
        float4 hi, lo;
        const float4 v = *reinterpret_cast<const float4*>(Q + (size_t)tid * NI + 4 * c);
        hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        *reinterpret_cast<float4*>(q_hi + chunk_off_floats(128, c, tid)) = hi;
        *reinterpret_cast<float4*>(q_lo + chunk_off_floats(128, c, tid)) = lo;
    }
    if (tid == 0) mbar_init(mbar, 1);
    fence_smem_to_async();
    if (warp == 0) tmem_alloc(slot, 64);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = *slot;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, NI, (variant & 2) ? 0 : 1, (variant & 4) ? 0 : 1);
        const int lbo = (variant & 1) ? 128 * 16 : 128, sbo = (variant & 1) ? 128 : 128 * 16;
        uint32_t acc = 0;
        for (int t = 0; t < terms; ++t) {
            const float* pa = t == 1 ? p_lo : p_hi;
            const float* pb = t == 2 ? q_lo : q_hi;
            for (int ks = 0; ks < 16; ++ks) {              // 8 rows per MMA
                mma_tf32(tbase, smem_desc(pa + ks * 32, lbo, sbo), smem_desc(pb + ks * 32, lbo, sbo), idesc, acc);
                acc = 1;
            }
        }
        commit(mbar);
    }
    const bool ok = mbar_wait(mbar, 0);
    if (!ok && tid == 0) err[0] = 1;
    fence_after();
    if (ok) {
        for (int n0 = 0; n0 < NI; n0 += 8) {               // NI multiple of 8: read 16 columns, keep what exists
            float v[16];
            if (n0 % 16 == 0) {
                tmem_ld16(tbase + ((uint32_t)(32 * warp) << 16) + n0, v);
                for (int i = 0; i < 16; ++i) if (n0 + i < NI) out[(size_t)tid * NI + n0 + i] = v[i];
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tbase, 64);
}

}  // namespace umma
}  // namespace xw
#endif

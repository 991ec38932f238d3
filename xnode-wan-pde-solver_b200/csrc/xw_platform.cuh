// xw_platform.cuh -- the handful of CUDA builtins the kernels use, behind macros.
//
// Product build (nvcc, sm_100a): maps 1:1 to the CUDA builtins.
// Test build (g++ -DXW_EMU, tests/host_emu/): maps to a tiny std::thread-based emulator
// (tests/host_emu/cuda_emu.h) so that the indexing / staging / reduction logic of every kernel
// can be checked against the oracle on a CPU-only box.  The emulator is TEST INFRASTRUCTURE: the
// shipped library never contains it and there is no CPU fallback in the product path.
#pragma once

#ifdef XW_EMU
#include "cuda_emu.h"
#define XW_DEV inline
#define XW_HD inline
#define XW_GLOBAL inline
#define XW_RESTRICT
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define XW_DEV __device__ __forceinline__
#define XW_HD __host__ __device__ inline
#define XW_GLOBAL __global__
#define XW_RESTRICT __restrict__
#define XW_SYNCTHREADS() __syncthreads()
#define XW_SYNCWARP() __syncwarp()
// named barriers (producer / consumer hand-offs between warp roles of one CTA)
#define XW_BAR_SYNC(id, n) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory")
#define XW_BAR_ARRIVE(id, n) asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory")
#define XW_SHFL_XOR(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#define XW_SHFL_IDX(v, l) __shfl_sync(0xffffffffu, (v), (l))
// shared-memory mbarriers (phase-parity producer / consumer hand-offs; any number of them, unlike the 16 named barriers).
// arrive = release.cta, wait = acquire.cta: data written before the arrive is visible after the matching wait.
#define XW_MBAR_INIT(p, n)                                                                                             \
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"((int)(n)) : "memory")
#define XW_MBAR_ARRIVE(p) \
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(p)) : "memory")
#define XW_MBAR_WAIT(p, parity) xw_mbar_wait((p), (parity))
// per-warpgroup register re-allocation (warp-specialised kernels)
#define XW_SETMAXNREG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(n))
#define XW_SETMAXNREG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(n))
#define XW_TID ((int)threadIdx.x)
#define XW_BID ((int)blockIdx.x)
#define XW_BDIM ((int)blockDim.x)
#define XW_GDIM ((int)gridDim.x)
#define XW_ATOMIC_ADD_F(p, v) atomicAdd((p), (v))
#define XW_ATOMIC_ADD_D(p, v) atomicAdd((p), (v))
#endif

#ifndef XW_EMU
// spin on mbarrier.try_wait.parity (the instruction itself suspends the thread for a bounded time per probe)
// The wait is bounded in TIME (about 10 s of SM clock), not in probes: a protocol bug must surface as a CUDA error, never
// as a hung GPU, and a slow phase (contention, a debugger, throttled clocks) must not be mistaken for one.
__device__ __forceinline__ void xw_mbar_wait(const void* p, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    unsigned ok;
    long long t0 = 0;
    for (unsigned it = 0;; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if ((it & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000LL) __trap();
        }
    }
}
#endif

// 16-byte asynchronous global->shared copy (LDGSTS) and its completion wait
#ifdef XW_EMU
#define XW_CP_ASYNC16(dst, src) memcpy((dst), (src), 16)
#define XW_CP_ASYNC_WAIT_ALL() ((void)0)
#else
#define XW_CP_ASYNC16(dst, src)                                                                   \
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define XW_CP_ASYNC_WAIT_ALL() asm volatile("cp.async.wait_all;" ::: "memory")
#endif

// compiler-only memory barrier: keeps ptxas/nvcc from hoisting the (loop-invariant) shared-memory
// weight loads out of the layer / time-step loops, which would need thousands of registers
#define XW_FENCE() asm volatile("" ::: "memory")

#ifdef XW_EMU
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    sh &= 31u;
    return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
}
#endif

namespace xw {

constexpr int pad4(int x) { return (x + 3) & ~3; }

struct alignas(16) f4 { float x, y, z, w; };
struct alignas(8) f2 { float x, y; };

XW_DEV f4 ld4(const float* p) { return *reinterpret_cast<const f4*>(p); }
XW_DEV f2 ld2(const float* p) { return *reinterpret_cast<const f2*>(p); }
XW_DEV void st4(float* p, f4 v) { *reinterpret_cast<f4*>(p) = v; }

// load N consecutive floats (p 16-byte aligned) into registers with the widest loads
template <int N>
XW_DEV void load_row(const float* p, float (&w)[N]) {
#pragma unroll
    for (int j = 0; j + 4 <= N; j += 4) {
        f4 v = ld4(p + j);
        w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
    }
    constexpr int R = N & ~3;
    if constexpr ((N & 3) >= 2) {
        f2 v = ld2(p + R);
        w[R] = v.x; w[R + 1] = v.y;
    }
    if constexpr ((N & 3) == 1) w[R] = p[R];
    if constexpr ((N & 3) == 3) w[R + 2] = p[R + 2];
}

// out[j] += sum_i M[i*LD + j] * in[i]        (M "in-major": row i holds the OUT weights of input i)
// (matvec_acc: see below, after the packed-pair helpers)

// tanh as 1 - 2/(1 + e^{2x}) on the SFU (ex2 + rcp approximations): absolute error ~1e-7, i.e. the
// fp32 rounding level of the surrounding FMAs, at ~6 instructions instead of ~30 for tanhf()
XW_DEV float tanh_fast(float x) {
#ifdef XW_EMU
    return tanhf(x);
#else
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, 1.f + e);
#endif
}

XW_DEV float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += XW_SHFL_XOR(v, m);
    return v;
}
XW_DEV double warp_sum_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += XW_SHFL_XOR(v, m);
    return v;
}

// packed pair of floats for the Blackwell FFMA2 path (fma.rn.f32x2, sm_100+): two IEEE fmas per
// instruction, i.e. the same results as two fmaf() at half the issue slots
#ifdef XW_EMU
struct fpair { float lo, hi; };
XW_DEV fpair pack2(float lo, float hi) { return fpair{lo, hi}; }
XW_DEV void unpack2(fpair v, float& lo, float& hi) { lo = v.lo; hi = v.hi; }
XW_DEV fpair fma2(fpair a, fpair b, fpair c) { return fpair{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
#else
typedef unsigned long long fpair;
XW_DEV fpair pack2(float lo, float hi) { fpair r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
XW_DEV void unpack2(fpair v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
XW_DEV fpair fma2(fpair a, fpair b, fpair c) { fpair d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#endif


// out[j] += sum_i M[i*LD + j] * in[i]        (M "in-major": row i holds the OUT weights of input i)
// Two adjacent outputs share one packed FFMA2 (fma.rn.f32x2: the weight pair comes straight out of the 128-bit
// row load, the input is broadcast to both halves): same IEEE results as OUT fmaf() per input at half the issue
// slots -- the XNODE kernels are issue bound as much as LSU bound (profiles/r01i_xnode_bwd_stalls.txt).
template <int IN, int OUT, int LD>
XW_DEV void matvec_acc(const float* M, const float (&in)[IN], float (&out)[OUT]) {
    XW_FENCE();
    if constexpr (OUT % 2 == 0) {
        fpair acc[OUT / 2];
#pragma unroll
        for (int j = 0; j < OUT / 2; ++j) acc[j] = pack2(out[2 * j], out[2 * j + 1]);
#pragma unroll
        for (int i = 0; i < IN; ++i) {
            float w[OUT];
            load_row<OUT>(M + i * LD, w);
            const fpair xx = pack2(in[i], in[i]);
#pragma unroll
            for (int j = 0; j < OUT / 2; ++j) acc[j] = fma2(pack2(w[2 * j], w[2 * j + 1]), xx, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < OUT / 2; ++j) unpack2(acc[j], out[2 * j], out[2 * j + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < IN; ++i) {
            float w[OUT];
            load_row<OUT>(M + i * LD, w);
            const float xi = in[i];
#pragma unroll
            for (int j = 0; j < OUT; ++j) out[j] = fmaf(w[j], xi, out[j]);
        }
    }
}

// 128-bit LIFO of fixed-width bit groups (relu masks of the shared field layers)
struct BitStack128 {
    unsigned long long lo, hi;
    XW_DEV void clear() { lo = 0ull; hi = 0ull; }
    template <int B> XW_DEV void push(unsigned bits) {
        hi = (hi << B) | (lo >> (64 - B));
        lo = (lo << B) | (unsigned long long)bits;
    }
    template <int B> XW_DEV unsigned pop() {
        unsigned bits = (unsigned)(lo & ((1ull << B) - 1ull));
        lo = (lo >> B) | (hi << (64 - B));
        hi >>= B;
        return bits;
    }
};

}  // namespace xw

// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
// A minimal std::thread emulation of the CUDA execution model (one std::thread per CUDA thread,
// std::barrier for __syncthreads/__syncwarp, exchange buffers for warp shuffles) so that the
// kernels under xnode-wan-pde-solver_b200/csrc can be compiled with g++ -DXW_EMU and their
// results compared with the oracle on a CPU-only machine.  Never linked into the product library.
#pragma once
#include <barrier>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace emu {

// named barrier (PTX bar.sync / bar.arrive with an explicit thread count)
struct NamedBar {
    std::mutex m;
    std::condition_variable cv;
    int count = 0;
    unsigned gen = 0;
};

// mbarrier (PTX mbarrier.init / arrive / try_wait.parity) keyed by the address of its 8-byte slot in shared memory
struct MBar {
    int count = 0, pending = 0;
    unsigned phase = 0;
};

struct Block {
    int bdim = 0;
    NamedBar nbar[16];
    std::mutex mb_mu;
    std::condition_variable mb_cv;
    std::map<const void*, MBar> mbars;
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<std::barrier<>>> wbar;
    std::vector<uint64_t> xchg;           // [nwarps][32]
    std::vector<unsigned char> smem;
};

inline thread_local int tid = 0, bid = 0, bdim = 1, gdim = 1;
inline thread_local Block* blk = nullptr;
inline std::mutex atomic_mu;

inline void warp_sync() { blk->wbar[tid >> 5]->arrive_and_wait(); }

// bar.sync id, n (wait = true) / bar.arrive id, n (wait = false): the barrier completes when n threads have arrived
inline void named_bar(int id, int n, bool wait) {
    NamedBar& b = blk->nbar[id];
    std::unique_lock<std::mutex> lk(b.m);
    const unsigned g = b.gen;
    if (++b.count == n) {
        b.count = 0;
        ++b.gen;
        b.cv.notify_all();
    } else if (wait) {
        b.cv.wait(lk, [&] { return b.gen != g; });
    }
}

template <typename T>
inline T shfl_idx(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    uint64_t* buf = &blk->xchg[(size_t)(tid >> 5) * 32];
    buf[tid & 31] = raw;
    warp_sync();
    uint64_t got = buf[src_lane & 31];
    warp_sync();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T>
inline T shfl_xor(T v, int m) { return shfl_idx(v, (tid & 31) ^ m); }

template <typename T>
inline T atomic_add(T* p, T v) {
    std::lock_guard<std::mutex> g(atomic_mu);
    T old = *p;
    *p = old + v;
    return old;
}

inline void mbar_init(const void* p, int count) {
    std::lock_guard<std::mutex> lk(blk->mb_mu);
    MBar& m = blk->mbars[p];
    m.count = m.pending = count;
    m.phase = 0;
}
inline void mbar_arrive(const void* p) {
    std::lock_guard<std::mutex> lk(blk->mb_mu);
    MBar& m = blk->mbars[p];
    if (--m.pending == 0) {
        m.pending = m.count;
        m.phase ^= 1u;
        blk->mb_cv.notify_all();
    }
}
// waits for the completion of the phase with parity `parity` (the barrier starts in phase 0)
inline void mbar_wait(const void* p, unsigned parity) {
    std::unique_lock<std::mutex> lk(blk->mb_mu);
    MBar& m = blk->mbars[p];
    blk->mb_cv.wait(lk, [&] { return (m.phase & 1u) != (parity & 1u); });
}

// run `f()` as a kernel body over grid x block threads with `smem_bytes` of dynamic shared memory
template <typename F>
inline void launch(int grid, int block, size_t smem_bytes, F f) {
    for (int b = 0; b < grid; ++b) {
        Block B;
        B.bdim = block;
        B.bar = std::make_unique<std::barrier<>>(block);
        int nw = (block + 31) / 32;
        for (int w = 0; w < nw; ++w) {
            int cnt = std::min(32, block - 32 * w);
            B.wbar.emplace_back(std::make_unique<std::barrier<>>(cnt));
        }
        B.xchg.assign((size_t)nw * 32, 0);
        B.smem.assign(smem_bytes + 64, 0);
        std::vector<std::thread> th;
        th.reserve(block);
        for (int t = 0; t < block; ++t) {
            th.emplace_back([&, t]() {
                tid = t; bid = b; bdim = block; gdim = grid; blk = &B;
                f();
            });
        }
        for (auto& x : th) x.join();
    }
}

inline unsigned char* dyn_smem() {
    uintptr_t p = reinterpret_cast<uintptr_t>(blk->smem.data());
    p = (p + 15) & ~uintptr_t(15);
    return reinterpret_cast<unsigned char*>(p);
}

}  // namespace emu

#define XW_SYNCTHREADS() emu::blk->bar->arrive_and_wait()
#define XW_SYNCWARP() emu::warp_sync()
#define XW_BAR_SYNC(id, n) emu::named_bar((id), (n), true)
#define XW_BAR_ARRIVE(id, n) emu::named_bar((id), (n), false)
#define XW_MBAR_INIT(p, n) emu::mbar_init((p), (n))
#define XW_MBAR_ARRIVE(p) emu::mbar_arrive((p))
#define XW_MBAR_WAIT(p, parity) emu::mbar_wait((p), (parity))
#define XW_SETMAXNREG_INC(n) ((void)0)
#define XW_SETMAXNREG_DEC(n) ((void)0)
#define XW_SHFL_XOR(v, m) emu::shfl_xor((v), (m))
#define XW_SHFL_IDX(v, l) emu::shfl_idx((v), (l))
#define XW_TID (emu::tid)
#define XW_BID (emu::bid)
#define XW_BDIM (emu::bdim)
#define XW_GDIM (emu::gdim)
#define XW_ATOMIC_ADD_F(p, v) emu::atomic_add<float>((p), (v))
#define XW_ATOMIC_ADD_D(p, v) emu::atomic_add<double>((p), (v))

"""tcgen05.mma (kind::tf32, K = 8 per instruction) issue-rate microbenchmark on tools/tcb/libtcb.so: cycles per MMA for the
shapes of the test-function kernels, dependent (one accumulator) vs independent (round-robin accumulators)."""
import ctypes as C, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "tools", "tcb", "libtcb.so"))
lib.tcb_mma_bench.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p]
out = torch.zeros(6, dtype=torch.int64, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
rows = []
count = 336
for mode, M, N, a_rows in ((0, 128, 56, 128), (0, 128, 112, 128), (0, 128, 24, 128), (0, 128, 224, 128), (1, 64, 56, 56), (1, 128, 56, 112), (1, 128, 112, 112), (1, 64, 24, 56),
                           (1, 128, 24, 112), (1, 128, 224, 112)):
    for nacc in (1, 2, 3, 4):
        if nacc * N > 440:
            continue
        assert lib.tcb_mma_bench(mode, M, N, nacc, count, a_rows, out.data_ptr(), st) == 0
        torch.cuda.synchronize()
        o = out.cpu().tolist()
        rows.append({"mode": "TS" if mode == 0 else "SS", "M": M, "N": N, "nacc": nacc, "floor": max(M, 128) * N / 256,
                     "issue_cyc_per_mma": round(o[4] / count, 1), "done_cyc_per_mma": round(o[5] / count, 1)})
        print(rows[-1], flush=True)

# several issuing warps at once: total MMAs / cycles
lib.tcb_mma_bench_multi.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
for N in (24, 56, 112):
    for nissue in (1, 2, 3, 4):
        if nissue * N > 440:
            continue
        out.zero_()
        assert lib.tcb_mma_bench_multi(128, N, nissue, count, out.data_ptr(), st) == 0
        torch.cuda.synchronize()
        o = out.cpu().tolist()
        print({"multi_issuer": nissue, "N": N, "floor": 128 * N / 256, "cyc_per_mma_overall": round(max(o[:nissue]) / (count * nissue), 1)}, flush=True)

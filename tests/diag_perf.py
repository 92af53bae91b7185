"""GPU diagnostic (not a pytest): throughput of the K-major tcgen05 GEMM alone, small vs large M, both epilogues."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import npp_b200

nat = npp_b200._native
lib = nat.lib()
torch.manual_seed(0)
CHAIN = int(os.environ.get('NPP_DEBUG_CHAIN', '1'))
print('chain length per launch:', CHAIN, 'dependent' if os.environ.get('NPP_DEBUG_CHAIN_DEP') else 'independent')
for (m, n, k) in [(16384, 512, 512), (16384, 512, 1024), (131072, 512, 512)]:
    a = torch.randn(m, k, device="cuda").half()
    b = (torch.randn(n, k, device="cuda") * 0.05).half()
    o0 = torch.empty(m, n, device="cuda", dtype=torch.half)
    o1 = torch.empty(m, n, device="cuda", dtype=torch.half)
    for epi in (0, 1):
        ms = C.c_float()
        nat.check(lib.npp_debug_gemm_bench(a.data_ptr(), b.data_ptr(), o0.data_ptr(), o1.data_ptr(), m, n, k, epi, 50, C.byref(ms)))
        us = ms.value * 1e3 / 50
        print(f"m={m} n={n} k={k} epi={'snake' if epi else 'linear'}: {us:8.2f} us/launch  {2.0*m*n*k/us/1e6:8.1f} TFLOP/s", flush=True)
    # cuBLAS reference point for the same shape (library GEMM, not part of the product)
    torch.matmul(a, b.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        torch.matmul(a, b.t())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    print(f"m={m} n={n} k={k} cuBLAS fp16            : {us:8.2f} us/launch  {2.0*m*n*k/us/1e6:8.1f} TFLOP/s", flush=True)

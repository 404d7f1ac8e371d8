"""Ad-hoc: does the projection kernel's time scale with M (bandwidth/issue bound) or carry a large fixed cost?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import ops, _lib
_lib.load()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
def timeit(fn, reps=8):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for (K, N) in [(36, 108), (108, 36)]:
    W = torch.randn(K, N, device="cuda")
    for M in [128 * 148, 25600, 102400, 409600, 1638400]:
        X = torch.randn(M, K, device="cuda")
        t = timeit(lambda: ops.gemm(X, W))
        print(f"K={K} N={N} M={M:8d}: {t:8.1f} us  {4*M*(K+N)/t/1e3:8.1f} GB/s  tiles/SM {M/128/148:.1f}")

"""Ad-hoc: phase timestamps (globaltimer, ns) of CTA 0 inside the projection kernel."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glam_b200 import ops, _lib
lib = _lib.load()
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.glam_debug_tc_timestamps.argtypes = [ctypes.c_void_p]
lib.glam_debug_tc_timestamps(buf.data_ptr())
names = ["entry", "setup done", "W staged (mma)", "first X tile landed", "first accumulator ready", "first tile stored", "all done", "tmem freed"]
for (M, K, N, nt) in [(102400, 36, 108, 0), (102400, 36, 108, 1), (102400, 108, 36, 0), (4096, 108, 144, 1), (4096, 144, 108, 0)]:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") if nt else torch.randn(K, N, device="cuda")
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); ops.gemm(X, W, transpose_w=bool(nt)); e1.record(); torch.cuda.synchronize()
    t = buf.cpu().tolist()
    print(f"M={M} K={K} N={N} nt={nt}: event {e0.elapsed_time(e1)*1e3:.1f} us;", ", ".join(f"{n} +{(t[i]-t[0])/1e3:.2f}" for i, n in enumerate(names)))

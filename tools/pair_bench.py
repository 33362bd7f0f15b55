"""CTA-pair GEMM prototype vs the single-CTA tensor-core kernels on the weight-gradient shape (256 x 256 x 256, batch 512)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sac_expert_b200 import lib as L
lib = L.load()
fn = lib.saceo_test_pair_gemm
fn.restype = C.c_int
fn.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
st = torch.cuda.current_stream().cuda_stream
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for K in (256, 512):
    A = torch.randn(batch, 256, K, device="cuda"); B = torch.randn(batch, 256, K, device="cuda"); Cc = torch.empty(batch, 256, 256, device="cuda")
    def t(f, n=10):
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    us_pair = t(lambda: L.check(fn(batch, K, A.data_ptr(), B.data_ptr(), Cc.data_ptr(), st)))
    # same product with the single-CTA kernels: C = A . B^T  (transB = 1): streaming kernel (mode 1) and register-staged (variant 2)
    us_stream = t(lambda: L.check(lib.saceo_test_gemm(1, batch, 256, 256, K, 0, 1, A.data_ptr(), B.data_ptr(), Cc.data_ptr(), st)))
    us_reg = t(lambda: L.check(lib.saceo_test_gemm(1 | (2 << 8), batch, 256, 256, K, 0, 1, A.data_ptr(), B.data_ptr(), Cc.data_ptr(), st)))
    gb = 4.0 * batch * (2 * 256 * K + 256 * 256) / 1e3
    print(f"K={K}: pair {us_pair:7.1f} us ({gb / us_pair:5.0f} GB/s)   stream {us_stream:7.1f} us   register-staged {us_reg:7.1f} us")

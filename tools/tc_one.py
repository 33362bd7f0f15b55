"""One tcgen05 GEMM shape, a few launches (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sac_expert_b200 import lib as L
lib = L.load()
batch, M, N, K = 512, 256, 256, 256
ta, tb = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 0
A = torch.randn(batch, *((K, M) if ta else (M, K)), device="cuda")
B = torch.randn(batch, *((N, K) if tb else (K, N)), device="cuda")
C = torch.empty(batch, M, N, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    L.check(lib.saceo_test_gemm(1, batch, M, N, K, ta, tb, A.data_ptr(), B.data_ptr(), C.data_ptr(), st))
torch.cuda.synchronize()
print("ok")

"""Times the standalone batched GEMM surface (saceo_test_gemm) with CUDA events: both engines, the three
operand orientations of the update (forward NN, input-gradient NT, weight-gradient TN)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sac_expert_b200 import lib as L

lib = L.load()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
shapes = [(256, 256, 256), (256, 256, 35), (276, 256, 256), (256, 256, 276)]
st = torch.cuda.current_stream().cuda_stream
for mode in [int(m) for m in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['0', '1', '257'])]:
    for (M, N, K) in shapes:
        for ta, tb in ((0, 0), (0, 1), (1, 0)):
            A = torch.randn(batch, *((K, M) if ta else (M, K)), device="cuda")
            B = torch.randn(batch, *((N, K) if tb else (K, N)), device="cuda")
            C = torch.empty(batch, M, N, device="cuda")
            for _ in range(3):
                L.check(lib.saceo_test_gemm(mode, batch, M, N, K, ta, tb, A.data_ptr(), B.data_ptr(), C.data_ptr(), st))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 10
            for _ in range(n):
                L.check(lib.saceo_test_gemm(mode, batch, M, N, K, ta, tb, A.data_ptr(), B.data_ptr(), C.data_ptr(), st))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            tf = 2.0 * batch * M * N * K / (ms * 1e-3) / 1e12
            gb = 4.0 * batch * (M * K + K * N + M * N) / (ms * 1e-3) / 1e9
            print(f"mode={mode} M={M} N={N} K={K} ta={ta} tb={tb}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s  {gb:7.0f} GB/s (compulsory)")

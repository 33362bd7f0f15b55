"""Per-CTA phase timing of the tcgen05 streaming GEMM kernel (globaltimer stamps)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from sac_expert_b200 import lib as L
lib = L.load()
lib.saceo_test_set_tc_debug.argtypes = [C.c_void_p]
batch, M, N, K = 512, 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = torch.randn(batch, M, K, device="cuda"); B = torch.randn(batch, K, N, device="cuda"); Cc = torch.empty(batch, M, N, device="cuda")
nct = batch * 2
dbg = torch.zeros(nct * 8, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    L.check(lib.saceo_test_gemm(1, batch, M, N, K, 0, 0, A.data_ptr(), B.data_ptr(), Cc.data_ptr(), st))
torch.cuda.synchronize()
lib.saceo_test_set_tc_debug(dbg.data_ptr())
L.check(lib.saceo_test_gemm(1, batch, M, N, K, 0, 0, A.data_ptr(), B.data_ptr(), Cc.data_ptr(), st))
torch.cuda.synchronize()
lib.saceo_test_set_tc_debug(None)
t = dbg.cpu().numpy().reshape(nct, 8).astype(np.float64)
t0 = t[:, 0].min()
names = ["start->alloc+prologue", "main loop issue", "wait last MMA", "epilogue", "final sync"]
d = np.diff(t[:, :6], axis=1)
print("K=%d kernel span %.1f us, CTA duration mean %.2f us (min %.2f max %.2f)" % (K, (t[:, 5].max() - t0) / 1e3, (t[:, 5] - t[:, 0]).mean() / 1e3, (t[:, 5] - t[:, 0]).min() / 1e3, (t[:, 5] - t[:, 0]).max() / 1e3))
for i, n in enumerate(names):
    print("  %-24s mean %7.2f us  p10 %7.2f  p90 %7.2f" % (n, d[:, i].mean() / 1e3, np.percentile(d[:, i], 10) / 1e3, np.percentile(d[:, i], 90) / 1e3))
starts = np.sort(t[:, 0] - t0) / 1e3
print("  CTA start times (us): first 148: %.1f..%.1f ; gaps between waves ~ %s" % (starts[0], starts[147], np.round(starts[[148, 296, 444, 592, 740, 888]], 1)))

"""HBM bandwidth probes (GPU box): torch read-only / copy / fill vs our streaming kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zeroshotsemanticsegmentation_b200 import _lib
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = 1 << 30  # 4 GiB fp32
x = torch.empty(N, device="cuda"); x.normal_()
y = torch.empty_like(x)
gb = N * 4 / 1e9
print("torch sum      (read)       %.0f GB/s" % (gb / t(lambda: x.sum()) * 1e3))
print("torch copy     (read+write) %.0f GB/s" % (2 * gb / t(lambda: y.copy_(x)) * 1e3))
print("torch fill     (write)      %.0f GB/s" % (gb / t(lambda: y.zero_()) * 1e3))
print("torch mul      (r+w)        %.0f GB/s" % (2 * gb / t(lambda: torch.mul(x, 2.0, out=y)) * 1e3))
rows, C = N // 256, 256
db = torch.zeros(C, device="cuda")
print("szn_bias_grad  (read)       %.0f GB/s" % (gb / t(lambda: _lib.call("szn_bias_grad", 0, x.data_ptr(), db.data_ptr(), rows, C, C, st)) * 1e3))
rows, C = N // 64, 64
print("szn_bias_grad C=64 (read)   %.0f GB/s" % (gb / t(lambda: _lib.call("szn_bias_grad", 0, x.data_ptr(), db.data_ptr(), rows, C, C, st)) * 1e3))

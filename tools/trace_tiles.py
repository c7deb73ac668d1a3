"""Debug: per-tile clock64 stamps of CTA 0's roles for one conv launch (needs the SZN_TRACE hook in szn_umma.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
buf = torch.zeros(3 * 64 * 4, dtype=torch.int64, device="cuda")
os.environ["SZN_TRACE"] = str(buf.data_ptr())
from zeroshotsemanticsegmentation_b200 import _lib
B, H, W, cin, cout, k, pad = 8, 710, 710, 64, 64, 3, 1
if len(sys.argv) > 1:
    H = W = int(sys.argv[1]); cin = int(sys.argv[2]); cout = int(sys.argv[3])
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(B, H, W, cin, device="cuda"); y = torch.empty(B, H, W, cout, device="cuda")
wt = torch.randn(cout, 9, cin, device="cuda") * 0.01; bias = torch.zeros(cout, device="cuda")
for _ in range(2):
    buf.zero_()
    _lib.call("szn_conv_fwd", 0, x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, cin, cout, k, k, pad, 1, None, 0, 0, cout, st)
    torch.cuda.synchronize()
t = buf.cpu().view(3, 64, 4)
t0 = int(t[0, 0, 0])
print("epilogue chunk 0 detail: accf ok -> tmem ld done -> math done -> wait_read+bar -> sts+fence+bar -> chunk0 done")
for i in range(40, 46):
    r = lambda v: int(v) - int(t[2, i, 1])
    print("%4d | ld %6d | math %6d | wait+bar %6d | sts+fence+bar %6d | tma issue %6d" % (i, r(t[0, i, 1]), r(t[0, i, 2]), r(t[0, i, 3]), r(t[1, i, 3]), r(t[2, i, 2])))
print("tile | producer start | mma: start, acce ok, committed | epi: start, accf ok, chunk0 done, chunk1 done   (cycles since first producer stamp)")
for i in list(range(0, 12)) + list(range(40, 46)):
    r = lambda v: int(v) - t0 if int(v) else -1
    print("%4d | %8d | %8d %8d %8d | %8d %8d %8d %8d" % (i, r(t[0, i, 0]), r(t[1, i, 0]), r(t[1, i, 1]), r(t[1, i, 2]), r(t[2, i, 0]), r(t[2, i, 1]), r(t[2, i, 2]), r(t[2, i, 3])))

"""Per-layer timing of the tensor-core conv kernels (fwd / dgrad / wgrad) at the bench shapes, CUDA events, GPU box.
usage: python tools/bench_layers.py [tf32|bf16|fp32] [B] [reps] [layer,layer,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zeroshotsemanticsegmentation_b200 import _lib
from zeroshotsemanticsegmentation_b200.engine import TRUNK

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
dt, td = {"tf32": (0, torch.float32), "bf16": (1, torch.bfloat16), "fp32": (2, torch.bfloat16)}[prec]
npl = 2 if prec == "fp32" else 1  # split storage: [hi | lo] bf16 planes per pixel row / two weight planes
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
layers = []
h = w = 512 + 198
for row in TRUNK:
    if len(row) == 1:
        h, w = (h + 1) // 2, (w + 1) // 2
    elif row[0] != "conv1_1":
        layers.append((row[0], h, w, row[1], row[2], row[3], row[4]))
layers += [("fc6", h, w, 512, 4096, 7, 0), ("fc7", h - 6, w - 6, 4096, 4096, 1, 0), ("head", h - 6, w - 6, 4096, 320, 1, 0)]


def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print("%-8s %5s %5s %5s %5s | %22s | %22s | %22s" % ("layer", "H", "Cin", "Cout", "k", "fwd ms / TF/s", "dgrad ms / TF/s", "wgrad ms / TF/s"))
for name, H, W, cin, cout, k, pad in layers:
    if only and name not in only:
        continue
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x = torch.randn(B, H, W, npl * cin, device=dev).to(td)
    y = torch.empty(B, Ho, Wo, npl * cout, device=dev, dtype=td)
    dy = torch.randn(B, Ho, Wo, npl * cout, device=dev).to(td)
    dx = torch.empty(B, H, W, npl * cin, device=dev, dtype=td)
    wt = (torch.randn(npl * cout, k * k, cin, device=dev) * 0.01).to(td)
    wd = (torch.randn(npl * cin, k * k, cout, device=dev) * 0.01).to(td)
    bias = torch.zeros(cout, device=dev)
    dw = torch.zeros(cout, k * k * cin, device=dev)
    fl = 2.0 * B * Ho * Wo * cout * k * k * cin
    t_f = timeit(lambda: _lib.call("szn_conv_fwd", dt, x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, cin, cout, k, k, pad, 1, None, 0, 0, cout, st))
    if name == "fc6":
        dcol = torch.empty(B, Ho, Wo, npl * k * k * cin, device=dev, dtype=td)
        def dg():
            _lib.call("szn_conv_dgrad", dt, dy.data_ptr(), wd.data_ptr(), dcol.data_ptr(), B, Ho, Wo, k * k * cin, cout, 1, 1, 0, None, None, 0, cout, None, st)
            _lib.call("szn_col2im", dt, dcol.data_ptr(), dx.data_ptr(), B, H, W, cin, k, k, st)
        t_d = timeit(dg)
    else:
        t_d = timeit(lambda: _lib.call("szn_conv_dgrad", dt, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), B, H, W, cin, cout, k, k, pad, x.data_ptr(), None, 0, cout, None, st))
    t_w = timeit(lambda: _lib.call("szn_conv_wgrad", dt, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), B, H, W, cin, cout, k, k, pad, cout, st))
    tot["fwd"] += t_f; tot["dgrad"] += t_d; tot["wgrad"] += t_w
    print("%-8s %5d %5d %5d %5d | %9.3f ms %7.0f TF/s | %9.3f ms %7.0f TF/s | %9.3f ms %7.0f TF/s" %
          (name, H, cin, cout, k, t_f, fl / t_f / 1e9, t_d, fl / t_d / 1e9, t_w, fl / t_w / 1e9))
    del x, y, dy, dx, wt, wd, dw
    torch.cuda.empty_cache()
print("total ms: fwd %.2f dgrad %.2f wgrad %.2f" % (tot["fwd"], tot["dgrad"], tot["wgrad"]))

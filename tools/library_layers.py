"""Per-layer "library GPU" table: every conv layer of the trunk through stock PyTorch (cuDNN / cuBLAS,
cudnn.benchmark on, channels_last, TF32 allowed or bf16) -- forward, data gradient and weight gradient timed separately
with CUDA events -- next to the same layer through libszn's tcgen05 kernel (tools/bench_layers.py prints those).
The reference ships no GPU kernels of its own: this is what "the reference's conv on a B200" is.

    python tools/library_layers.py [tf32|bf16] [B] [reps] [layer,layer,...]

(with a layer list under `ncu --metrics gpu__time_duration.sum` the launch list shows which cuDNN kernels won the
cudnn.benchmark search: they are the ones repeated `reps` times at the end of each pass.)

Not part of the product or of any parity claim; plain torch only (no oracle import).  Prints a table and one JSON line.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from zeroshotsemanticsegmentation_b200.engine import TRUNK

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dtype = torch.bfloat16 if prec == "bf16" else torch.float32
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"

layers = []
h = w = 512 + 198
for row in TRUNK:
    if len(row) == 1:
        h, w = (h + 1) // 2, (w + 1) // 2
    elif row[0] != "conv1_1":
        layers.append((row[0], h, w, row[1], row[2], row[3], row[4]))
layers += [("fc6", h, w, 512, 4096, 7, 0), ("fc7", h - 6, w - 6, 4096, 4096, 1, 0)]
if len(sys.argv) > 4:
    layers = [l for l in layers if l[0] in sys.argv[4].split(",")]


def timeit(fn):
    for _ in range(3):  # cudnn.benchmark picks its algorithm on the first calls
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows, tot = [], {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print("%-8s %5s %5s %5s | %20s | %20s | %20s" % ("layer", "H", "Cin", "Cout", "cuDNN fwd ms / TF/s", "cuDNN dgrad", "cuDNN wgrad"))
for name, H, W, cin, cout, k, pad in layers:
    x = torch.randn(B, cin, H, W, device=dev, dtype=dtype).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, k, k, device=dev, dtype=dtype) * 0.01).contiguous(memory_format=torch.channels_last)
    y = F.conv2d(x, wt, padding=pad)
    dy = torch.randn_like(y)
    fl = 2.0 * B * y.shape[2] * y.shape[3] * cout * k * k * cin
    t_f = timeit(lambda: F.conv2d(x, wt, padding=pad))
    t_d = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, wt, None, [1, 1], [pad, pad], [1, 1], False, [0, 0], 1,
                                                             [True, False, False]))
    t_w = timeit(lambda: torch.ops.aten.convolution_backward(dy, x, wt, None, [1, 1], [pad, pad], [1, 1], False, [0, 0], 1,
                                                             [False, True, False]))
    tot["fwd"] += t_f
    tot["dgrad"] += t_d
    tot["wgrad"] += t_w
    rows.append({"layer": name, "fwd_ms": t_f, "dgrad_ms": t_d, "wgrad_ms": t_w, "gflop": fl / 1e9})
    print("%-8s %5d %5d %5d | %8.3f ms %6.0f TF/s | %8.3f ms %6.0f TF/s | %8.3f ms %6.0f TF/s" %
          (name, H, cin, cout, t_f, fl / t_f / 1e9, t_d, fl / t_d / 1e9, t_w, fl / t_w / 1e9))
    del x, wt, y, dy
    torch.cuda.empty_cache()
print("total ms: fwd %.2f dgrad %.2f wgrad %.2f" % (tot["fwd"], tot["dgrad"], tot["wgrad"]))
print(json.dumps({"tool": "library_layers", "precision": prec, "B": B, "torch": torch.__version__,
                  "cudnn": torch.backends.cudnn.version(), "layers": rows, "total_ms": tot}))

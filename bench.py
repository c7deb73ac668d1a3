#!/usr/bin/env python
"""bench.py — Mpixel/s of the SZN pixel-embedding hot path (forward + cosine loss + backward + nearest-embedding
labels) on synthetic PASCAL-Context-shaped batches, B x 3 x 512 x 512, 59 classes, 300-d embeddings.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|4] [--impl reference]

One "step" = one pass of the hot path over one batch per GPU, exactly the call sequence of the reference's training
iteration (trainer_fcn.py:83-158 without optimizer/metrics): model(x, mode='fcn') -> utils.cosine_loss -> backward ->
utils.infer_lbl.  N > 1: one process per GPU (torchrun), images sharded (weak scaling, B per GPU fixed), NCCL
all-reduce of the gradients and of the loss accumulator inside the timed region.

Prints ONE JSON line (rank 0).  `value` = device-resident inputs, CUDA events; `e2e` = same call sequence with pinned
HOST inputs copied in and loss + labels copied out every step; `roofline` = the tcgen05 implicit-GEMM conv kernel family
(timed per launch with CUDA events in a separate instrumented pass); `cpu_baseline` = the CPU oracle port of the
reference (oracle/, torch CPU fp32) on the box's host cores, one image.

--impl reference: times that CPU port alone (the reference is Python/torch and cannot travel to the GPU box; the
oracle is its line-by-line restatement, pinned to golden vectors produced by the unmodified reference).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 512
CONFIGS = {
    # BASELINE.json configs[1], [2], [4] (per-GPU batch); [3] is config 2's shape at 8 images per GPU
    1: dict(B=8, D=300, C=59, precision="tf32", name="configs[1]: PASCAL-Context-shaped 59-class SZN, 300-d, B=8/GPU, "
            "512x512, fp32 storage + TF32 tcgen05 products, fp32 accumulate"),
    2: dict(B=32, D=300, C=59, precision="bf16", name="configs[2]: SZN bf16 tensor-core path, B=32/GPU, 512x512"),
    4: dict(B=16, D=1024, C=256, precision="bf16", name="configs[4]: 256-class x 1024-d stress, B=16/GPU, 512x512"),
}
METRIC = "Mpixels/sec fwd+bwd at 512x512/59-class/300-d"
UNIT = "Mpixel/s"


def conv_flops(name, a):
    """Algorithmic FLOPs (2*MAC) of one tensor-core conv launch from its C-ABI arguments (include/szn.h)."""
    if name == "szn_conv_fwd":
        B, Hh, Ww, Cin, Cout, R, S, pad = a[5:13]
    elif name == "szn_conv_dgrad":
        B, Hh, Ww, Cin, Cout, R, S, pad = a[4:12]
    elif name == "szn_conv_wgrad":
        B, Hh, Ww, Cin, Cout, R, S, pad = a[4:12]
    else:
        return 0
    Ho, Wo = Hh + 2 * pad - R + 1, Ww + 2 * pad - S + 1
    return 2.0 * B * Ho * Wo * Cout * R * S * Cin


def trunk_flops_per_image(D):
    """algorithmic fwd+bwd FLOPs of one 512x512 image (dense upscore counted as the bilinear upsample)."""
    from zeroshotsemanticsegmentation_b200.engine import TRUNK
    h = w = H + 198
    fwd = 0.0
    first = True
    bwd = 0.0
    for row in TRUNK:
        if len(row) == 1:
            h, w = (h + 1) // 2, (w + 1) // 2
            continue
        _, cin, cout, k, _ = row
        f = 2.0 * h * w * cout * k * k * cin
        fwd += f
        bwd += f if first else 2 * f  # no dgrad into the image
        first = False
    hs, ws = h - 6, w - 6
    for cin, cout, k in ((512, 4096, 7), (4096, 4096, 1), (4096, D + 2, 1)):
        f = 2.0 * hs * ws * cout * k * k * cin
        fwd += f
        bwd += 2 * f
    return fwd, bwd


def clocks_sampler(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


def clocks_summary(path, index):
    sm, smax, reasons = [], 0.0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not f[0].isdigit() or int(f[0]) != index:
                continue
            sm.append(float(f[1]))
            smax = max(smax, float(f[2]))
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    busy = [v for v in sm if v > 0.5 * max(sm)] or sm
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_step_fn(cfg):
    """The reference's CPU path for this workload through the oracle port: 1 image per step (the reference's loss and
    inference functions only support n == 1).  upscore.weight's gradient is skipped (the reference never
    optimises it, train.py:324-327; computing it as written costs ~150 s more per image on 8 cores)."""
    import torch
    from oracle import szn_oracle as O
    D, C = cfg["D"], cfg["C"]
    torch.set_num_threads(os.cpu_count() or 1)
    params = O.init_params(D, seed=1337)
    x, lab, table = O.synth_batch(1, H, W, C, D, seed=1337)
    pr = {k: v.clone().requires_grad_("upscore" not in k) for k, v in params.items()}

    def step():
        for v in pr.values():
            v.grad = None
        f = O.forward(x, pr, "fcn")
        loss = O.cosine_loss(f, lab, O.target_embed_from_labels(lab, table))
        loss.backward()
        O.infer_lbl(f.detach(), table)
        return float(loss.detach())
    return step, torch.get_num_threads()


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, cores = cpu_oracle_step_fn(cfg)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps * H * W / 1e6 / dt
    sample = "1 image 512x512 per step (B=1: the reference cannot batch its loss), fwd+cosine loss+bwd+infer_lbl, " \
             "upscore.weight grad skipped"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["name"], "D": cfg["D"], "C": cfg["C"], "H": H, "W": W, "batch_per_step": 1},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)  # ~1 s per leg: long enough to sit at the sustained (power-capped) clock
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fused-head", action="store_true",
                    help="EXPERIMENTAL: FCN32s(fused_head=True), loss / labels / d s17 from the 17x17 score map (not the default, "
                         "not a bench line until its GPU parity tests are green)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["B"] = args.batch
    if args.impl == "reference":
        return run_reference(args, cfg)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import zeroshotsemanticsegmentation_b200 as szn
    from zeroshotsemanticsegmentation_b200 import _lib, ddp, synth
    U = szn.utils

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B, D, C = cfg["B"], cfg["D"], cfg["C"]
    model = synth.init_model_(szn.FCN32s(D, precision=cfg["precision"], fused_head=args.fused_head), seed=1337).to(dev).train()
    reducer = ddp.GradientAllReduce(model)
    # every rank draws its own images (seed + rank), like a sharded loader would
    x_h, lab_h, table = synth.synth_batch(B, H, W, C, D, seed=1337 + rank)
    x_h, lab_h = x_h.pin_memory(), lab_h.pin_memory()
    table = table.to(dev)
    x_d, lab_d = x_h.to(dev), lab_h.to(dev)

    def step(x, lab):
        model.zero_grad(set_to_none=True)
        f = model(x, mode="fcn")
        loss = U.cosine_loss(f, lab, table=table, accum_hook=reducer.accum_hook)
        loss.backward()
        lbl = U.infer_lbl_device(f if args.fused_head else f.detach(), table)  # detach() would drop the fused-head handle
        return loss, lbl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, finish=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        if finish is not None:
            finish()  # still inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    last = {}

    def resident():
        last["loss"], last["lbl"] = step(x_d, lab_d)

    # end-to-end: every step copies ITS inputs from pinned host memory and returns ITS loss + labels to the host.  Like a
    # pinned-memory DataLoader with non_blocking copies, the H2D copy of step i+1 is issued on a copy stream while step
    # i computes, and the D2H copy of step i's loss + labels runs on a second copy stream; the host reads (and NaN-checks)
    # the result of step i-1 while step i runs, so the device never waits for Python.  Every copy and every host read
    # happens inside the timed region, once per step; `drain` reads the last step's result before the clock stops.
    copy_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    slots = [dict(x=torch.empty_like(x_d), lab=torch.empty_like(lab_d), ev=torch.cuda.Event()) for _ in range(2)]
    results = [dict(lbl=torch.empty((B, H, W), dtype=torch.int64).pin_memory(),
                    loss=torch.empty((1,), dtype=torch.float32).pin_memory(), ev=torch.cuda.Event()) for _ in range(2)]
    state = {"i": 0, "primed": False, "pending": None, "host_reads": 0}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            slot["x"].copy_(x_h, non_blocking=True)
            slot["lab"].copy_(lab_h, non_blocking=True)
            slot["ev"].record(copy_stream)

    def read_pending():
        r = state["pending"]
        if r is None:
            return
        r["ev"].synchronize()  # host waits for THAT step's D2H only; the next step is already queued behind it
        last["loss_host"] = float(r["loss"][0])
        if last["loss_host"] != last["loss_host"]:
            raise SystemExit("loss is nan while training")  # trainer_fcn.py:152-153
        state["host_reads"] += 1
        state["pending"] = None

    def end_to_end():
        if not state["primed"]:
            upload(slots[0])
            state["primed"] = True
        cur = slots[state["i"] & 1]
        nxt = slots[(state["i"] + 1) & 1]
        res = results[state["i"] & 1]
        state["i"] += 1
        main = torch.cuda.current_stream()
        main.wait_event(cur["ev"])
        copy_stream.wait_stream(main)  # the slot being refilled was consumed by the previous step
        upload(nxt)                    # next step's inputs travel while this step computes
        loss, lbl = step(cur["x"], cur["lab"])
        d2h_stream.wait_stream(main)
        with torch.cuda.stream(d2h_stream):
            res["lbl"].copy_(lbl, non_blocking=True)
            res["loss"].copy_(loss.detach().reshape(1), non_blocking=True)
            res["ev"].record(d2h_stream)
        lbl.record_stream(d2h_stream)
        loss.record_stream(d2h_stream)
        read_pending()                 # result of the PREVIOUS step (its slot is the other one)
        state["pending"] = res

    def drain():
        read_pending()

    for _ in range(args.warmup):
        resident()
    torch.cuda.synchronize()
    if not torch.isfinite(last["loss"]).item():
        raise SystemExit("non-finite loss in warm-up")

    clk_path = os.path.join(tempfile.gettempdir(), "szn_clocks_%d.csv" % rank)
    sampler = clocks_sampler(clk_path) if rank == 0 else None
    n0 = _lib.launch_count()
    ms_total = timed(resident, args.steps)
    launches = _lib.launch_count() - n0
    e2e_ms = None
    if not args.no_e2e:
        for _ in range(2):
            end_to_end()
        drain()
        reads0 = state["host_reads"]
        e2e_ms = timed(end_to_end, args.steps, finish=drain)
        assert state["host_reads"] - reads0 == args.steps, "every timed step's loss must reach the host"
    if sampler is not None:
        time.sleep(0.25)
        sampler.terminate()
        sampler.wait()

    # ---- per-launch CUDA-event timing of every C-ABI call (instrumented pass, not part of `value`) ----
    recs = []

    def prof(name, a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()

        def done():
            e1.record()
            recs.append((name, a, e0, e1))
        return done

    _lib.set_profiler(prof)
    prof_steps = 2
    for _ in range(prof_steps):
        resident()
    torch.cuda.synchronize()
    _lib.set_profiler(None)
    per = {}
    for name, a, e0, e1 in recs:
        d = per.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
        d[2] += conv_flops(name, a)
    tot_ms = sum(d[1] for d in per.values())
    kernels = {k: {"launches_per_step": d[0] // prof_steps, "ms_per_step": d[1] / prof_steps,
                   "share": d[1] / tot_ms, **({"tflops": d[2] / d[1] / 1e9} if d[2] else {})}
               for k, d in sorted(per.items(), key=lambda kv: -kv[1][1])}
    umma = [per[k] for k in ("szn_conv_fwd", "szn_conv_dgrad", "szn_conv_wgrad") if k in per]
    umma_flops, umma_ms, umma_n = sum(d[2] for d in umma), sum(d[1] for d in umma), sum(d[0] for d in umma)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
    if cfg["precision"] == "tf32":
        peak, peak_note = bf16_peak / 2, peak_src + " / 2: kind::tf32 issues at half the bf16 rate"
    else:
        peak, peak_note = bf16_peak, peak_src
    achieved = umma_flops / umma_ms / 1e9 if umma_ms else 0.0
    # DRAM bytes per launch of the same kernel family, from the committed ncu capture of this command (config 1 only)
    traffic, traffic_src = None, None
    try:
        if args.config == 1 and not args.batch:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_umma_traffic.json")))
            traffic, traffic_src = tj["umma_family_dram_bytes_per_launch"], "profiles/r01_umma_traffic.json (ncu dram__bytes)"
    except Exception:
        pass

    ms_step = ms_total / args.steps
    pix = world * B * H * W / 1e6
    value = pix / (ms_step / 1e3)
    fwd, bwd = trunk_flops_per_image(D)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": cfg["precision"], "data": "synthetic",
        "config": {"workload": cfg["name"], "batch_per_gpu": B, "global_batch": B * world, "H": H, "W": W, "D": D, "C": C,
                   "loss": "cosine", "mode": "train (Dropout2d live)", "parallelism": "dp%d" % world,
                   "l2": "no explicit flush: one step streams >10 GB of activations per GPU, far above the 126 MB L2",
                   "weights": "seeded random init (no network for VGG16 weights)",
                   **({"fused_head": "EXPERIMENTAL"} if args.fused_head else {})},
        "gpu_launches": launches,
        "step_tflops": (fwd + bwd) * B * world / (ms_step / 1e3) / 1e12,
        "roofline": {"bound": "tensor", "kernel": "umma_conv_kernel<T,MODE> (tcgen05 implicit-GEMM conv fwd/dgrad/wgrad)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flops_per_launch": umma_flops / max(umma_n, 1),
                     "launches_per_step": umma_n // prof_steps,
                     "share_of_step": umma_ms / tot_ms if tot_ms else None, "peak_source": peak_note},
        "kernels": kernels,
    }
    if e2e_ms is not None:
        out["e2e"] = {"value": pix / (e2e_ms / args.steps / 1e3), "unit": UNIT,
                      "h2d_bytes_per_step": x_h.numel() * 4 + lab_h.numel() * 8,
                      "d2h_bytes_per_step": results[0]["lbl"].numel() * 8 + 4, "ms_per_step": e2e_ms / args.steps,
                      "pipeline": "H2D of step i+1 and D2H + host read of step i-1 overlap step i (copy streams); "
                                  "all copies and reads of the timed steps are inside the timed region"}
    if rank == 0:
        out["clocks"] = clocks_summary(clk_path, local)
        out["loss"] = float(last["loss"].item())
    if world > 1:
        out["allreduce_bytes_per_step"] = reducer.bytes_reduced // max(1, (args.warmup + args.steps * 2 + 2 + prof_steps))
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        cstep, cores = cpu_oracle_step_fn(cfg)
        t0 = time.perf_counter()
        cstep()
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": H * W / 1e6 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "1 image 512x512 (B=1), one cold pass of fwd+cosine loss+bwd+infer_lbl through the "
                                         "oracle port (torch CPU fp32), upscore.weight grad skipped; %.1f s" % dt}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — Mpixel/s of the SZN pixel-embedding hot path (forward + cosine loss + backward + nearest-embedding
labels) on synthetic PASCAL-Context-shaped batches, B x 3 x 512 x 512, 59 classes, 300-d embeddings.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 0|1|2|3|4] [--precision tf32|bf16|fp32] [--impl reference]

One "step" = one pass of the hot path over one batch per GPU, exactly the call sequence of the reference's training
iteration (trainer_fcn.py:83-158 without optimizer/metrics): model(x, mode='fcn') -> utils.cosine_loss -> backward ->
utils.infer_lbl.  N > 1: one process per GPU (torchrun), images sharded (weak scaling, B per GPU fixed), NCCL
all-reduce of the gradients and of the loss accumulator inside the timed region.

--config selects BASELINE.json's configs[i]:
  0  1 x 3 x 256 x 256, 21 classes, cross_entropy2d(sum) (trainer_fcn.py with pixel_embeddings off; plumbing case)
  1  B=8 per GPU, D=300, C=59, "fp32" config: fp32 storage + TF32 products (default; the headline line)
  2  B=32 per GPU, bf16
  3  zero-shot split (49 seen / 10 unseen at validation, {0, 12} held out of training -> 47 seen), B=8 per GPU, bf16:
     one step = a phase-1 iteration (trainer_fcn.Trainer.forward + backward, train.py:138-161), a phase-2 iteration
     (trainer_seenmask.Trainer.forward + backward with everything but the seen-mask head frozen, train.py:164-194) and
     an SZN inference pass (trainer_fcn.Trainer.forward_szn: mode='both' + infer_lbl_szn, utils.py:195-205)
  4  B=16 per GPU, D=1024, C=256, bf16

Prints ONE JSON line (rank 0).  `value` = device-resident inputs, CUDA events; `e2e` = same call sequence with pinned
HOST inputs copied in and loss + labels copied out every step; `roofline` = the tcgen05 implicit-GEMM conv kernel family
(timed per launch with CUDA events in a separate instrumented pass); `kernels` = every C-ABI call of a step with its
share and, for the HBM-bound ones, achieved GB/s against the measured copy bandwidth; `cpu_baseline` = the CPU oracle
port of the reference (oracle/, torch CPU fp32) on the box's host cores, one image, with the SAME weights and image as
the GPU model, so the same pass also yields `parity` (forward / loss / label agreement of this build at full size);
`grad_check` (N > 1) = sharded + all-reduced gradients against a micro-batched single-rank replica.

--impl reference: times the reference's CPU path alone: the UNMODIFIED models.py / utils.py when a reference checkout
sits in baseline/_ref (kind "reference"), else the oracle port (kind "port": the reference is a Python script
directory that cannot travel to the GPU box; the port is pinned to golden vectors produced by the unmodified reference).
"""
import argparse
import hashlib
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

CONFIGS = {
    0: dict(B=1, H=256, D=21, C=21, precision="fp32", loss="ce",  # the plumbing / parity case runs in the fp32-grade mode
            name="configs[0]: 1x3x256x256 PASCAL-VOC 21-class FCN32s forward + cross_entropy2d(sum) + backward (plumbing)"),
    1: dict(B=8, H=512, D=300, C=59, precision="tf32", loss="cos",
            name="configs[1]: PASCAL-Context-shaped 59-class SZN, 300-d, B=8/GPU, 512x512, fp32 storage + TF32 tcgen05 "
                 "products, fp32 accumulate"),
    2: dict(B=32, H=512, D=300, C=59, precision="bf16", loss="cos",
            name="configs[2]: SZN bf16 tensor-core path, B=32/GPU, 512x512"),
    3: dict(B=8, H=512, D=300, C=59, precision="bf16", loss="cos", zeroshot=True,
            name="configs[3]: zero-shot split (49 seen / 10 unseen; {0,12} held out of training), B=8/GPU, 512x512, bf16: "
                 "phase-1 iteration + phase-2 (seen-mask head) iteration + SZN inference per step"),
    4: dict(B=16, H=512, D=1024, C=256, precision="bf16", loss="cos",
            name="configs[4]: 256-class x 1024-d stress, B=16/GPU, 512x512, bf16"),
}
VAL_UNSEEN = list(range(49, 59))   # mirrors configs.py:114-126 scaled to 59 classes
TRAIN_UNSEEN = [0, 12]
METRIC = "Mpixels/sec fwd+bwd at 512x512/59-class/300-d"
UNIT = "Mpixel/s"
DTYPE_NAME = {"tf32": "tf32", "bf16": "bf16", "fp32": "bf16x3"}
ES = {"tf32": 4, "bf16": 2, "fp32": 4}  # bytes per stored trunk element


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


def hbm_bytes(name, a, es):
    """ALGORITHMIC HBM bytes of one launch of an HBM-bound kernel (every operand byte moved once), from its C-ABI
    arguments (include/szn.h); 0 for kernels that are not HBM-bound.  es = bytes per trunk element."""
    if name == "szn_conv1_1_fwd":       # (dtype, x, w, bias, y, B, H, W, pad): read the image, write 64 channels
        B, Hh, Ww, pad = a[5:9]
        return B * 3 * Hh * Ww * 4 + B * (Hh + 2 * pad - 2) * (Ww + 2 * pad - 2) * 64 * es
    if name == "szn_conv1_1_wgrad":     # only dY pixels whose 3x3 window touches the image contribute
        B, Hh, Ww, pad = a[4:8]
        return B * 3 * Hh * Ww * 4 + B * (Hh + 2) * (Ww + 2) * 64 * es
    if name == "szn_pool_fwd":          # (dtype, in, out, B, H, W, C)
        B, Hh, Ww, C = a[3:7]
        return B * Hh * Ww * C * es * 1.25
    if name == "szn_pool_bwd":          # reads y and dp, writes dy
        B, Hh, Ww, C = a[4:8]
        return B * Hh * Ww * C * es * 2.25
    if name == "szn_pool_fwd_code":     # (dtype, in, out, code, B, H, W, C): + one routing byte per pooled element
        B, Hh, Ww, C = a[4:8]
        return B * Hh * Ww * C * (es * 1.25 + 0.25)
    if name == "szn_pool_bwd_code":     # (dtype, code, dp, dy, B, H, W, C, ...): reads dp and the codes, writes dy
        B, Hh, Ww, C = a[4:8]
        return B * Hh * Ww * C * (es * 1.25 + 0.25)
    if name == "szn_upsample32_crop_fwd":   # (s, out, B, D, H, W, ...): writes the score
        B, D, Hh, Ww = a[2:6]
        return B * D * Hh * Ww * 4
    if name == "szn_upsample32_crop_bwd":   # (dtype, g, ds, B, D, H, W, ...): reads the score gradient
        B, D, Hh, Ww = a[3:7]
        return B * D * Hh * Ww * 4
    if name == "szn_embed_loss_fwd":    # (kind, score, target, te, table, table_rows, n, c, h, w, ...)
        n, c, h, w = a[6:10]
        return n * h * w * (c * 4 + 8) * (2 if a[3] else 1)
    if name == "szn_embed_loss_bwd":
        n, c, h, w = a[6:10]
        return n * h * w * (c * 4 * 2 + 8) + (n * h * w * c * 4 if a[3] else 0)
    if name == "szn_embed_argmax":      # (score, table, n, c, h, w, C, scratch, labels): reads the score, writes labels
        n, c, h, w = a[2:6]
        return n * h * w * (c * 4 + 8)
    return 0


def trunk_flops_per_image(D, H):
    """algorithmic fwd+bwd FLOPs of one H x H image (dense upscore counted as the bilinear upsample)."""
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


def so_hash():
    try:
        from zeroshotsemanticsegmentation_b200 import _lib
        return hashlib.sha256(open(_lib.LIB_PATH, "rb").read()).hexdigest()[:16]
    except Exception:
        return None


def build_id():
    """Hash of the kernel sources + flags libszn.so was built from (include/szn_build.h).  Unlike the hash of the .so it is
    the same for every rebuild of the same sources (nvcc objects are not byte-reproducible)."""
    try:
        from zeroshotsemanticsegmentation_b200 import _lib
        return _lib.build_id()
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ CPU arm
def reference_modules():
    """(models, utils) of an UNMODIFIED reference checkout under baseline/_ref, or None.  /root/reference is never read
    from here: it does not exist on the GPU box."""
    from oracle import ref_import
    root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(root, "models.py")):
        return None
    try:
        return ref_import.load_reference(root)
    except Exception:
        return None


def cpu_step_fn(cfg, params=None, x=None, lab=None, table=None, upscore_grad=False):
    """One image of the reference's CPU path for this workload (the reference's loss and inference functions only
    support n == 1): forward + loss + backward + infer_lbl.  Through the unmodified reference when
    baseline/_ref holds it, else through the oracle port.  upscore.weight's gradient is skipped unless asked for (the
    reference never optimises it, train.py:324-327; as written it costs ~150 s more per image on 8 cores).
    Returns (step, cores, kind); step() -> dict(loss, f, lbl)."""
    import torch
    from oracle import szn_oracle as O
    D, C, H = cfg["D"], cfg["C"], cfg["H"]
    torch.set_num_threads(os.cpu_count() or 1)
    if params is None:
        params = O.init_params(D, seed=1337)
    if x is None:
        x, lab, table = O.synth_batch(1, H, H, C, D, seed=1337)
    ref = reference_modules()
    if ref is not None:
        rmodels, rutils = ref
        m = rmodels.FCN32s(n_class=D)
        m.load_state_dict(params)
        m.eval()
        for n, p_ in m.named_parameters():
            p_.requires_grad_(upscore_grad or "upscore" not in n)

        def step():
            m.zero_grad()
            f = m(x, mode="fcn")
            if cfg["loss"] == "ce":
                loss = rutils.cross_entropy2d(f, lab, size_average=False)
                lbl = f.detach().max(1)[1].numpy()
            else:
                loss = rutils.cosine_loss(f, lab, O.target_embed_from_labels(lab, table))
                lbl = rutils.infer_lbl(f.detach(), table, False)
            loss.backward()
            return dict(loss=float(loss.detach()), f=f.detach(), lbl=lbl)
        return step, torch.get_num_threads(), "reference"

    pr = {k: v.clone().requires_grad_(upscore_grad or "upscore" not in k) for k, v in params.items()}

    def step():
        for v in pr.values():
            v.grad = None
        f = O.forward(x, pr, "fcn")
        if cfg["loss"] == "ce":
            loss = O.cross_entropy2d(f, lab, size_average=False)
            lbl = f.detach().max(1)[1].numpy()
        else:
            loss = O.cosine_loss(f, lab, O.target_embed_from_labels(lab, table))
            lbl = O.infer_lbl(f.detach(), table)
        loss.backward()
        return dict(loss=float(loss.detach()), f=f.detach(), lbl=lbl)
    return step, torch.get_num_threads(), "port"


def cpu_sample_text(cfg, extra=""):
    what = "cross_entropy2d(sum)" if cfg["loss"] == "ce" else "cosine loss"
    return ("1 image %dx%d per step (B=1: the reference cannot batch its loss), eval-mode fwd + %s + bwd + labels, "
            "upscore.weight grad skipped%s" % (cfg["H"], cfg["H"], what, extra))


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H = cfg["H"]
    step, cores, kind = cpu_step_fn(cfg, upscore_grad=args.cpu_as_written)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps * H * H / 1e6 / dt
    sample = cpu_sample_text(cfg) if not args.cpu_as_written else \
        "1 image per step AS WRITTEN: dense upscore ConvTranspose2d(D,D,64,32) with its weight gradient (trainer_fcn.py:157,161)"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["name"], "D": cfg["D"], "C": cfg["C"], "H": H, "W": H, "batch_per_step": 1},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)  # ~1 s per leg: long enough to sit at the sustained (power-capped) clock
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=None, choices=["tf32", "bf16", "fp32"],
                    help="override the config's arithmetic: fp32 = fp32-grade 3 x bf16 split products (the strict-parity mode)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-grad-check", action="store_true")
    ap.add_argument("--cpu-as-written", action="store_true",
                    help="--impl reference only: also compute the dense upscore.weight gradient, as the reference does")
    ap.add_argument("--no-fused-head", action="store_true",
                    help="time only the API-preserving path (loss / labels read the materialised (B,D,H,W) score); by default "
                         "`value` is the fast path FCN32s(fused_head=True) -- loss, labels and d s17 from the 17x17 score map -- and "
                         "the API-preserving path is timed beside it as `api_path`")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["B"] = args.batch
    if args.precision:
        cfg["precision"] = args.precision
        cfg["name"] += " [precision overridden: %s]" % args.precision
    if args.impl == "reference":
        return run_reference(args, cfg)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import zeroshotsemanticsegmentation_b200 as szn
    from zeroshotsemanticsegmentation_b200 import _lib, ddp, synth, trainer as T
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

    B, D, C, H = cfg["B"], cfg["D"], cfg["C"], cfg["H"]
    W = H
    prec = cfg["precision"]
    zeroshot = bool(cfg.get("zeroshot"))
    use_ce = cfg["loss"] == "ce"
    args.fused_head = not args.no_fused_head and not use_ce  # the fused head covers the embedding losses

    def build_model(fused=None):
        fused = args.fused_head if fused is None else fused
        return synth.init_model_(szn.FCN32s(D, precision=prec, fused_head=fused), seed=1337).to(dev)

    model = build_model().train()
    reducer = ddp.GradientAllReduce(model)

    def rank_batch(r):
        """Rank r's images and labels (seed + rank), like a sharded loader would hand them out.  The class table is the
        dataset's: the same on every rank (rank 0's draw)."""
        x, lab, tab = synth.synth_batch(B, H, W, C, D, seed=1337 + r)
        if r:
            tab = synth.synth_batch(B, H, W, C, D, seed=1337)[2]
        if zeroshot:  # phase-1 training images contain seen classes only (pascal_dataset.py:78-84 filters the others)
            seen = torch.tensor([c for c in range(C) if c not in VAL_UNSEEN + TRAIN_UNSEEN])
            lab = torch.where(lab >= 0, seen[lab.clamp_min(0) % len(seen)], lab)
        return x, lab, tab

    x_h, lab_h, table = rank_batch(rank)
    x_h, lab_h = x_h.pin_memory(), lab_h.pin_memory()
    table_h = table
    table = table.to(dev)
    x_d, lab_d = x_h.to(dev), lab_h.to(dev)
    seen_t, unseen_t = (t.to(dev) for t in U.split_embeddings(table_h, VAL_UNSEEN)) if zeroshot else (None, None)
    head_names = ("seenmask_score.weight", "seenmask_score.bias", "seenmask_upscore.weight")

    def set_phase(m, phase):
        """phase 1: everything trainable except the two deconvs (train.py:324-327); phase 2: only the seen-mask head
        (train.py:166-171)."""
        for n, p_ in m.named_parameters():
            p_.requires_grad_(n in head_names if phase == 2 else ("upscore" not in n and not n.startswith("seenmask")))

    def fcn_iteration(m, x, lab, hook):
        m.zero_grad(set_to_none=True)
        f = m(x, mode="fcn")
        if use_ce:
            loss = U.cross_entropy2d(f, lab, size_average=False, accum_hook=hook)
        else:
            loss = U.cosine_loss(f, lab, table=table, accum_hook=hook)
        loss.backward()
        if use_ce:
            lbl = f.detach().max(1)[1]
        else:
            lbl = U.infer_lbl_device(f if m.fused_head else f.detach(), table)  # detach() would drop the fused-head handle
        return loss, lbl

    def step(x, lab):
        if not zeroshot:
            return fcn_iteration(model, x, lab, reducer.accum_hook)
        set_phase(model, 1)
        loss, _ = fcn_iteration(model, x, lab, reducer.accum_hook)
        set_phase(model, 2)
        model.zero_grad(set_to_none=True)
        s = model(x, mode="seenmask")
        loss2 = U.cross_entropy2d(s, U.seenmask_target(lab, TRAIN_UNSEEN, C), size_average=True, accum_hook=reducer.accum_hook)
        loss2.backward()
        with torch.no_grad():
            f, s = model(x, mode="both")
            U.cosine_loss(f, lab, table=table, accum_hook=reducer.accum_hook)
            lbl = U.infer_lbl_szn_device(f, s, seen_t, unseen_t)
        return loss + loss2.detach(), lbl

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

    # ---- N > 1: gradient check before anything is timed: the sharded, all-reduced gradients of one step against a
    # single-rank replica that runs every rank's shard as a micro-batch (same per-launch shapes) and lets autograd add ----
    grad_check = None
    if world > 1 and not args.no_grad_check:
        model.eval()  # no dropout on either side
        set_phase(model, 1)
        loss_p, _ = fcn_iteration(model, x_d, lab_d, reducer.accum_hook)
        torch.cuda.synchronize()
        if rank == 0:
            ref = build_model().eval()
            ref.load_state_dict(model.state_dict())
            set_phase(ref, 1)
            shards = [(x_d, lab_d)] + [tuple(t.to(dev) for t in rank_batch(r)[:2]) for r in range(1, world)]
            accs = []
            with torch.no_grad():
                for xs, ls in shards:
                    f = ref(xs, mode="fcn")
                    (U.cross_entropy2d(f, ls, accum_hook=lambda a: accs.append(a.clone())) if use_ce else
                     U.cosine_loss(f, ls, table=table, accum_hook=lambda a: accs.append(a.clone())))
            total = sum(accs)
            loss_1 = None
            for xs, ls in shards:
                f = ref(xs, mode="fcn")
                loss_1 = (U.cross_entropy2d(f, ls, accum_hook=lambda a: a.copy_(total)) if use_ce else
                          U.cosine_loss(f, ls, table=table, accum_hook=lambda a: a.copy_(total)))
                loss_1.backward()
            worst, worst_name, n_checked = 0.0, None, 0
            g1 = dict(ref.named_parameters())
            for n, p_ in model.named_parameters():
                if p_.grad is None:
                    continue
                e = float(((p_.grad - g1[n].grad).norm() / g1[n].grad.norm().clamp_min(1e-30)).item())
                n_checked += 1
                if e > worst:
                    worst, worst_name = e, n
            grad_check = {"worst_rel": worst, "worst_param": worst_name, "params_checked": n_checked,
                          "loss_abs_diff": abs(float(loss_p.item()) - float(loss_1.item())),
                          "what": "rel-L2 of every parameter gradient: %d ranks sharded + NCCL all-reduced vs a single-rank "
                                  "replica running the %d shards as micro-batches (eval mode, same weights)" % (world, world)}
            del ref, shards, g1
            torch.cuda.empty_cache()
        model.zero_grad(set_to_none=True)
        model.train()
        barrier()

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
        main_s = torch.cuda.current_stream()
        main_s.wait_event(cur["ev"])
        copy_stream.wait_stream(main_s)  # the slot being refilled was consumed by the previous step
        upload(nxt)                      # next step's inputs travel while this step computes
        loss, lbl = step(cur["x"], cur["lab"])
        d2h_stream.wait_stream(main_s)
        with torch.cuda.stream(d2h_stream):
            res["lbl"].copy_(lbl, non_blocking=True)
            res["loss"].copy_(loss.detach().reshape(1), non_blocking=True)
            res["ev"].record(d2h_stream)
        lbl.record_stream(d2h_stream)
        loss.record_stream(d2h_stream)
        read_pending()                   # result of the PREVIOUS step (its slot is the other one)
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
    bytes0, nccl0 = reducer.bytes_reduced, reducer.launches
    n0 = _lib.launch_count()
    ms_total = timed(resident, args.steps)
    launches = _lib.launch_count() - n0
    allreduce_bytes = (reducer.bytes_reduced - bytes0) // max(1, args.steps)
    allreduce_launches = (reducer.launches - nccl0) / max(1, args.steps)
    e2e_ms = None
    if not args.no_e2e:
        for _ in range(2):
            end_to_end()
        drain()
        reads0 = state["host_reads"]
        e2e_ms = timed(end_to_end, args.steps, finish=drain)
        assert state["host_reads"] - reads0 == args.steps, "every timed step's loss must reach the host"
    api_ms = None
    if args.fused_head and not zeroshot:
        # the API-preserving path beside it (the parity reference: loss and labels read the materialised score), same
        # weights, same inputs, same barriers, half as many steps
        model.fused_head = False
        for _ in range(3):
            resident()
        api_steps = max(5, args.steps // 2)
        api_ms = timed(resident, api_steps) / api_steps
        model.fused_head = True
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

    # one untimed step first: the legs before this one (API path, CPU arm, host copies) leave other kernels' data in L2 and
    # the clocks wherever the power cap put them; a 2-step sample taken cold once showed the weight-gradient family 11 %
    # slow while ms_per_step had not moved
    resident()
    torch.cuda.synchronize()
    _lib.set_profiler(prof)
    prof_steps = 4
    for _ in range(prof_steps):
        resident()
    torch.cuda.synchronize()
    _lib.set_profiler(None)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"
    per = {}
    for name, a, e0, e1 in recs:
        d = per.setdefault(name, [0, 0.0, 0.0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
        d[2] += conv_flops(name, a)
        d[3] += hbm_bytes(name, a, ES[prec])
    tot_ms = sum(d[1] for d in per.values())
    kernels = {}
    for k, d in sorted(per.items(), key=lambda kv: -kv[1][1]):
        e = {"launches_per_step": d[0] // prof_steps, "ms_per_step": d[1] / prof_steps, "share": d[1] / tot_ms}
        if d[2]:
            e["tflops"] = d[2] / d[1] / 1e9
        if d[3]:
            e["bound"] = "hbm"
            e["algorithmic_bytes_per_step"] = d[3] / prof_steps
            e["gbs"] = d[3] / d[1] / 1e6
            e["frac"] = e["gbs"] / hbm_peak
        kernels[k] = e
    umma = [per[k] for k in ("szn_conv_fwd", "szn_conv_dgrad", "szn_conv_wgrad") if k in per]
    umma_flops, umma_ms, umma_n = sum(d[2] for d in umma), sum(d[1] for d in umma), sum(d[0] for d in umma)

    # ---- roofline denominator.  bf16: the driver's measured cuBLAS number.  TF32: measured HERE, the same way (torch.matmul
    # with TF32 allowed, 8192^3, back to back for ~1 s = sustained under the power cap), because MEASURED_PEAKS has no TF32
    # entry and "half of bf16" undersold it in round 1.  fp32-grade (3 x bf16 per product): bf16 / 3. ----
    bf16_peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
    tf32_measured = None
    if prec == "tf32" and rank == 0:
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            n = 8192
            a_ = torch.randn(n, n, device=dev)
            b_ = torch.randn(n, n, device=dev)
            for _ in range(3):
                a_ @ b_
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200
            e0.record()
            for _ in range(reps):
                a_ @ b_
            e1.record()
            torch.cuda.synchronize()
            tf32_measured = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) / 1e3) / 1e12
            del a_, b_
            torch.backends.cuda.matmul.allow_tf32 = False
        except Exception:
            tf32_measured = None
    if prec == "tf32":
        if tf32_measured:
            peak, peak_note = tf32_measured, "measured in this run: torch.matmul (cuBLAS) TF32 8192^3, 200 back-to-back (sustained)"
        else:
            peak, peak_note = bf16_peak / 2, peak_src + " / 2 (no in-run TF32 measurement on this rank)"
    elif prec == "fp32":
        peak, peak_note = bf16_peak / 3, peak_src + " / 3: every fp32-grade product is three bf16 MMAs"
    else:
        peak, peak_note = bf16_peak, peak_src
    achieved = umma_flops / umma_ms / 1e9 if umma_ms else 0.0
    # DRAM bytes per launch of the same kernel family from the committed ncu capture of this command; only trusted when it
    # was taken from a libszn.so built from THESE kernel sources (build id = source hash, or the same .so file) and this
    # configuration
    traffic, traffic_src = None, "no ncu capture for this build/config (profiles/r02_umma_traffic.json carries the build id it was taken from)"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_umma_traffic.json")))
        if tj.get("config") == args.config and tj.get("precision") == prec and not args.batch:
            if tj.get("so_sha256_16") == so_hash():
                traffic, traffic_src = tj["umma_family_dram_bytes_per_launch"], "profiles/r02_umma_traffic.json (ncu dram__bytes, same libszn.so)"
            elif tj.get("build_id") and tj.get("build_id") == build_id():
                traffic, traffic_src = (tj["umma_family_dram_bytes_per_launch"],
                                        "profiles/r02_umma_traffic.json (ncu dram__bytes, libszn.so built from the same kernel sources: build id %s)" % build_id())
            else:
                traffic, traffic_src = (tj["umma_family_dram_bytes_per_launch"],
                                        "profiles/r02_umma_traffic.json, taken from ANOTHER build of libszn.so (build id %s, this run %s): "
                                        "indicative only" % (tj.get("build_id") or tj.get("so_sha256_16"), build_id() or so_hash()))
    except Exception:
        pass

    ms_step = ms_total / args.steps
    pix = world * B * H * W / 1e6
    value = pix / (ms_step / 1e3)
    fwd, bwd = trunk_flops_per_image(D, H)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE_NAME[prec], "data": "synthetic",
        "config": {"workload": cfg["name"], "batch_per_gpu": B, "global_batch": B * world, "H": H, "W": W, "D": D, "C": C,
                   "loss": "cross_entropy2d(sum)" if use_ce else "cosine", "mode": "train (Dropout2d live)",
                   "parallelism": "dp%d" % world, "precision": prec,
                   "l2": "no explicit flush: one step streams >10 GB of activations per GPU, far above the 126 MB L2",
                   "weights": "seeded random init (no network for VGG16 weights)",
                   "head": ("fused (loss, labels and d s17 from the 17x17 score map; the score is still returned)" if args.fused_head
                            else "API-preserving (loss and labels read the materialised score)")},
        "gpu_launches": launches,
        "step_tflops": (fwd + bwd) * B * world / (ms_step / 1e3) / 1e12 if not zeroshot else None,
        "roofline": {"bound": "tensor", "kernel": "umma_conv_kernel<T,MODE,SPLIT> (tcgen05 implicit-GEMM conv fwd/dgrad/wgrad)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flops_per_launch": umma_flops / max(umma_n, 1),
                     "launches_per_step": umma_n // prof_steps,
                     "share_of_step": umma_ms / tot_ms if tot_ms else None, "peak_source": peak_note,
                     "frac_of_half_bf16_sustained": achieved / (bf16_peak / 2) if prec == "tf32" else None,
                     "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src},
        "kernels": kernels,
        "libszn_sha256_16": so_hash(), "libszn_build_id": build_id(),
    }
    if tf32_measured:
        out["tf32_tflops_measured"] = tf32_measured
    if api_ms is not None:
        out["api_path"] = {"value": pix / (api_ms / 1e3), "unit": UNIT, "ms_per_step": api_ms,
                           "what": "same step with FCN32s(fused_head=False): cosine loss, its gradient and infer_lbl each make "
                                   "their pass over the (B,D,H,W) score (utils.py:75-102,159-185 as written)"}
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
        out["allreduce_bytes_per_step"] = int(allreduce_bytes)
        out["allreduce"] = {"path": "szn_allreduce_bucket (C ABI, own NCCL communicator, one fused launch per bucket, side stream)"
                            if reducer.comm is not None else "torch.distributed.all_reduce per tensor",
                            "bucket_launches_per_step": allreduce_launches}
        if grad_check is not None:
            out["grad_check"] = grad_check
        dist.barrier()
    cpu_ok = D * D * 64 * 64 * 4 <= 4e9  # the reference's dense upscore weight (models.py:94) is 17 GB at D = 1024
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not cpu_ok:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                               "sample": "skipped: the reference's dense ConvTranspose2d(D, D, 64, 32) weight alone is %.0f GB at D = %d"
                                         % (D * D * 64 * 64 * 4 / 1e9, D)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and cpu_ok:
        # CPU arm on the SAME weights and the SAME image 0 as the GPU model, eval mode on both sides: one cold pass gives the
        # baseline time AND this build's full-size parity numbers
        model.eval()
        set_phase(model, 1)
        with torch.no_grad():
            f_gpu = model(x_d[:1].contiguous(), mode="fcn")
            if use_ce:
                loss_gpu = float(U.cross_entropy2d(f_gpu, lab_d[:1], size_average=False).item())
                lbl_gpu = f_gpu.max(1)[1].cpu().numpy()
            else:
                loss_gpu = float(U.cosine_loss(f_gpu, lab_d[:1], table=table).item())
                lbl_gpu = U.infer_lbl_device(f_gpu, table).cpu().numpy()             # the timed path (fused head if on)
                lbl_api = U.infer_lbl_device(f_gpu.detach(), table).cpu().numpy()    # from the materialised score
        f_gpu = f_gpu.cpu()
        params = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}
        torch.cuda.empty_cache()
        cstep, cores, kind = cpu_step_fn(cfg, params, x_h[:1].clone(), lab_h[:1].clone(), table_h)
        t0 = time.perf_counter()
        r = cstep()                      # the first (cold) pass also provides the parity numbers below
        dt, passes = time.perf_counter() - t0, 1
        while dt < 10.0 and passes < 8:  # a bounded sample of about 10 s of CPU work
            cstep()
            dt, passes = time.perf_counter() - t0, passes + 1
        out["cpu_baseline"] = {"value": passes * H * W / 1e6 / dt, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": cpu_sample_text(cfg, "; %d passes over the same image (the first one cold), %.1f s in "
                                                              "total, same weights and image as the GPU model" % (passes, dt))}
        import numpy as np
        out["parity"] = {
            "vs": "the CPU arm above (fp32), image 0 of the batch at full size, eval mode",
            "fwd_rel_err": float((f_gpu - r["f"]).abs().max() / r["f"].abs().max()),
            "loss_abs_err": abs(loss_gpu - r["loss"]), "loss": [loss_gpu, r["loss"]],
            "label_agreement": float((np.asarray(lbl_gpu) == np.asarray(r["lbl"])).mean()),
            **({} if use_ce else {"label_agreement_api_path": float((np.asarray(lbl_api) == np.asarray(r["lbl"])).mean())}),
            "tolerance": "north star: forward within 1e-3 (max-abs-diff / max-abs-ref); precision='fp32' is the mode held to it "
                         "strictly in tests/test_full_size_gpu.py",
        }
        if args.config == 1:
            # SURVEY 8d's other CPU row: BASELINE configs[0] (256x256, 21 classes, CE) through the same arm
            c0 = dict(CONFIGS[0])
            s0, _, _ = cpu_step_fn(c0)
            s0()
            t0 = time.perf_counter()
            s0()
            d0 = time.perf_counter() - t0
            out["cpu_baseline"]["config0"] = {"value": 256 * 256 / 1e6 / d0, "unit": UNIT,
                                              "sample": "configs[0]: 1x3x256x256, 21 classes, fwd + cross_entropy2d(sum) + bwd, warm, %.2f s" % d0}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

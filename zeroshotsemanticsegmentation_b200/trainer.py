"""The two callers of the hot path, device-resident: the FCN-phase trainer (reference ``trainer_fcn.py:19-306``) and the
seen-mask-phase trainer (``trainer_seenmask.py:18-175``), plus the batching step in front of them (``train.py:81-84``).
SURVEY §8f rows 2-4.

Same constructor arguments, attributes (``epoch``, ``iteration``, ``best_mean_iu``, ``n_class``, ``seen``, ``embeddings``,
``seen_embeddings``, ``unseen_embeddings``), method names, return contracts, CSV log files and checkpoint dictionary as
the reference, so ``train.py``'s phase sequencing (``train.py:138-199``) drives them unchanged.  What differs is where
the per-iteration work runs:

* labels, predictions and the confusion matrices stay on the GPU (``utils.infer_lbl_device``,
  ``utils.confusion_hist_device``); per iteration the host receives the loss scalar and ``n_class**2`` counters instead
  of two label maps, and validation accumulates ONE histogram instead of keeping every label map of the epoch;
* a loader may yield ``(img, lbl)`` instead of ``(img, (lbl, lbl_vec))``: the target vectors are then gathered from the
  class table inside the fused loss (no 315 MB/img ``lbl_vec`` built on the host and pushed over PCIe); an explicit
  ``lbl_vec`` is still honoured (the dataset may use another table than the trainer, ``pascal_dataset.py:92-96``);
* any batch size works (the reference's losses are only correct for n == 1): ``collate_padded`` batches variable-size
  images by padding images with 0 (the mean pixel after ``transform``) and labels with the ignore label -1.

* ``reducer=ddp.GradientAllReduce(model)`` makes a trainer data-parallel (one process per GPU over its own shard of the
  loader): gradients are summed over ranks inside backward and every loss divides by the GLOBAL valid-pixel count.

Out of scope: tensorboard images, segmentation visualisations, US/Eastern timestamps; ``tb_writer`` may be
``None`` or anything with ``add_scalar``; ``visualize`` is an optional callback.  There is no CPU path: ``cuda=False`` raises.
"""
from __future__ import annotations

import math
import os
import os.path as osp
import shutil
import time

import numpy as np
import torch

from . import utils

TRAIN_HEADERS = ["epoch", "iteration", "train/loss", "train/pxl_acc", "train/class_acc", "train/mean_iu", "train/fwavacc",
                 "elapsed_time"]
VAL_HEADERS = ["epoch", "iteration", "val/loss", "val/pxl_acc", "val/class_acc", "val/mean_iu", "val/fwavacc"]
# trainer_fcn.py:301-306: zero-shot runs stop once they have seen as many images as 50 epochs of the full set
EARLY_STOP_ITERS = {"pascal": 425000, "context": 247000}


PAD_LABEL = -2  # label of pixels added by collate_padded: ignored like -1 everywhere, but never re-mapped (see below)


def collate_padded(samples, multiple=1, size=None):
    """``collate_fn`` for ``DataLoader(batch_size > 1)`` over the reference datasets (``pascal_dataset.py:106-133``,
    ``context_dataset.py:116-138``): items are ``(img (3,h,w) fp32, lbl (h,w) int64)`` or ``(img, (lbl, lbl_vec))``;
    ``lbl_vec`` is dropped (the fused loss gathers it from the table).  Images are padded bottom/right with 0, labels
    with ``PAD_LABEL`` (-2), to the batch maximum rounded up to ``multiple`` (or to ``size=(H, W)``).  Returns ``(data
    (B,3,H,W), target (B,H,W))``.  Padded pixels are ignored by every loss and metric (``target >= 0`` masks,
    ``utils.py:36,60,85,106``) -- also in the seen-mask phase, whose target maps the dataset's ignore label -1 to class 0
    like upstream (``trainer_seenmask.py:55-56``) but leaves the padding negative; with B == 1 and no rounding the item is
    returned untouched, i.e. exactly the reference's batch."""
    imgs, lbls = [], []
    for img, tgt in samples:
        lbl = tgt[0] if isinstance(tgt, (tuple, list)) else tgt
        img, lbl = torch.as_tensor(img), torch.as_tensor(lbl)
        if img.dim() != 3 or lbl.dim() != 2 or img.shape[1:] != lbl.shape:
            raise ValueError("expected img (3,h,w) and lbl (h,w), got %s / %s" % (tuple(img.shape), tuple(lbl.shape)))
        imgs.append(img.float())
        lbls.append(lbl.long())
    if not imgs:
        raise ValueError("empty batch")
    if size is not None:
        H, W = size
    else:
        H = max(i.shape[1] for i in imgs)
        W = max(i.shape[2] for i in imgs)
        H, W = -(-H // multiple) * multiple, -(-W // multiple) * multiple
    data = torch.zeros((len(imgs), imgs[0].shape[0], H, W), dtype=torch.float32)
    target = torch.full((len(imgs), H, W), PAD_LABEL, dtype=torch.int64)
    for b, (img, lbl) in enumerate(zip(imgs, lbls)):
        h, w = lbl.shape
        if h > H or w > W:
            raise ValueError("image %dx%d does not fit the fixed batch size %dx%d" % (h, w, H, W))
        data[b, :, :h, :w] = img
        target[b, :h, :w] = lbl
    return data, target


def save_checkpoint(path, model, optimizer, epoch, iteration, best_mean_iu):
    """The reference's checkpoint dictionary (``trainer_fcn.py:281-288``), readable by ``train.py:110-116,135-136``."""
    torch.save({
        "epoch": epoch,
        "iteration": iteration,
        "arch": model.__class__.__name__,
        "optim_state_dict": optimizer.state_dict(),
        "model_state_dict": model.state_dict(),
        "best_mean_iu": best_mean_iu,
    }, path)


def load_checkpoint(path, model, optimizer=None, map_location=None):
    """Resume like ``train.py:110-116`` (``strict=False`` for old checkpoints) and ``:135-136``; returns the dictionary."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ckpt["model_state_dict"], strict=False)
    if optimizer is not None and "optim_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optim_state_dict"])
    return ckpt


def _require_cuda(device):
    if device.type != "cuda":
        raise RuntimeError("move the model to the GPU before building a trainer (train.py:121-122)")


class _Base(object):
    log_prefix = ""
    tb_prefix = "fcn"

    def _setup(self, cuda, model, optimizer, train_loader, val_loader, log_dir, dataset, max_epoch, tb_writer, n_class,
               reducer=None):
        if not cuda:
            raise RuntimeError("the B200 trainers run on CUDA only (no CPU fallback): pass cuda=True")
        self.cuda = cuda
        self.model = model
        self.optim = optimizer
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.log_dir = log_dir
        self.dataset = dataset
        self.max_epoch = max_epoch
        self.tb_writer = tb_writer
        self.epoch = 0
        self.iteration = 0
        self.best_mean_iu = 0
        if n_class is None:
            n_class = len(train_loader.dataset.class_names)  # trainer_fcn.py:42
        self.n_class = n_class
        self.timestamp_start = time.time()
        self.device = next(model.parameters()).device
        _require_cuda(self.device)
        self.verbose = True
        # data parallel (one process per GPU, ddp.GradientAllReduce attached to the model): the losses then normalise by
        # the valid-pixel count of the GLOBAL batch; gradients are all-reduced inside backward
        self.reducer = reducer
        self._accum_hook = reducer.accum_hook if reducer is not None else None
        # rank awareness: validation histograms are summed over the ranks (each rank sees its shard of the loader), and only
        # the main rank writes CSV logs, tensorboard scalars and checkpoints (all ranks would otherwise race on one log_dir)
        self.is_main = getattr(reducer, "is_main", True)
        if not self.is_main:
            self.log_dir = log_dir = None
            self.tb_writer = None
            self.verbose = False
        if log_dir:
            os.makedirs(log_dir, exist_ok=True)

    def _global_hist(self, hist):
        fn = getattr(self.reducer, "all_reduce_hist", None)
        return fn(hist) if fn is not None else hist

    # ---- logging (same files and columns as the reference) ----
    def _init_log(self, name, headers):
        if not self.log_dir:
            return
        path = osp.join(self.log_dir, name)
        if not osp.exists(path):
            with open(path, "w") as f:
                f.write(",".join(headers) + "\n")

    def _append_log(self, name, row):
        if not self.log_dir:
            return
        with open(osp.join(self.log_dir, name), "a") as f:
            f.write(",".join(map(str, row)) + "\n")

    def _scalars(self, split, values, step, names=("loss", "pxl_acc", "class_acc", "mean_iu", "fwavacc")):
        if self.tb_writer is None:
            return
        for n, v in zip(names, values):
            self.tb_writer.add_scalar("%s/%s/%s" % (self.tb_prefix, split, n), v, step)

    def _elapsed(self):
        return time.time() - self.timestamp_start

    def _say(self, msg):
        if self.verbose:
            print(msg)

    @staticmethod
    def _split_target(target):
        """Loader item -> (labels, target_embed or None): reference loaders yield (lbl, lbl_vec), labels-only loaders lbl."""
        if isinstance(target, (tuple, list)):
            return target[0], target[1]
        return target, None

    def _to_device(self, t):
        return t.to(self.device, non_blocking=True) if t is not None else None

    @staticmethod
    def _check_loss(loss):
        val = float(loss.item())
        if math.isnan(val):
            raise ValueError("loss is nan while training")  # trainer_fcn.py:107-108
        return val

    def _train_iteration(self, data, target):
        score, loss, lbl_pred, lbl_true = self._forward_device(data, target)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return score, loss, lbl_pred, lbl_true

    def train_epoch(self):
        """``trainer_fcn.py:145-179`` / ``trainer_seenmask.py:72-101``: forward, backward, optimizer step, metrics, log."""
        self.model.train()
        for batch_idx, (data, target) in enumerate(self.train_loader):
            score, loss, lbl_pred, lbl_true = self._train_iteration(data, target)
            loss_val = self._check_loss(loss)
            metrics = utils.label_accuracy_score(lbl_true, lbl_pred, self.n_class)  # device histograms
            self._say("%s Train Epoch %-5d | Iteration %-5d | Loss %5.5f" % (self.tb_prefix, self.epoch, batch_idx, loss_val))
            self._append_log(self.log_prefix + "train_log.csv",
                             [self.epoch, self.iteration, loss_val] + list(metrics) + [self._elapsed()])
            self._scalars("train", [loss_val] + list(metrics), self.iteration)
            self.iteration += 1

    def train(self):
        for epoch in range(self.max_epoch):
            self.epoch = epoch
            self.train_epoch()
            self.validate()
            if self._stop_early():
                break

    def _stop_early(self):
        return False


class Trainer(_Base):
    """FCN phase: pixel embeddings (cosine / MSE against the class table) or plain 21-way cross entropy
    (``trainer_fcn.py:19-306``).  Extra keywords: ``embed_arr`` (the (C,D) table itself instead of the pickle under
    ``datasets/<dataset>/embeddings/``), ``n_class``, ``visualize`` (callback(img, lbl_true, lbl_pred) in validation)."""

    def __init__(self, cuda, model, optimizer, train_loader, val_loader, log_dir, dataset, max_epoch, tb_writer=None,
                 pixel_embeddings=None, loss_func=None, unseen=None, val_unseen=None, label_names=None, forced_unseen=False,
                 embed_arr=None, n_class=None, visualize=None, reducer=None):
        if pixel_embeddings and embed_arr is None:
            # each embedding has norm between 0 and 1 (trainer_fcn.py:47-49); path relative to the reference root
            embed_arr = utils.load_obj("datasets/%s/embeddings/norm_embed_arr_%s" % (dataset, str(pixel_embeddings)))
        if n_class is None and embed_arr is not None and not hasattr(getattr(train_loader, "dataset", None), "class_names"):
            n_class = int(np.asarray(embed_arr).shape[0])
        self._setup(cuda, model, optimizer, train_loader, val_loader, log_dir, dataset, max_epoch, tb_writer, n_class,
                    reducer)
        self.pixel_embeddings = pixel_embeddings
        self.loss_func = loss_func
        self.unseen = list(unseen) if unseen else []  # all unseen classes (train_unseen + val_unseen)
        self.val_unseen = list(val_unseen) if val_unseen else []
        self.label_names = label_names
        self.forced_unseen = forced_unseen
        self.visualize = visualize
        self.seen = [c for c in range(self.n_class) if c not in self.unseen]  # trainer_fcn.py:44
        if loss_func not in ("cos", "mse", "cross_entropy"):
            raise ValueError("loss_func must be 'cos', 'mse' or 'cross_entropy' (trainer_fcn.py:100-105)")
        if (loss_func == "cross_entropy") == bool(pixel_embeddings):
            raise ValueError("'cos'/'mse' need pixel_embeddings, 'cross_entropy' must not have them (configs.py)")
        if self.pixel_embeddings:
            table = torch.as_tensor(np.asarray(embed_arr)).float()
            seen_t, unseen_t = utils.split_embeddings(table, self.unseen)  # trainer_fcn.py:55-64
            self.embeddings = table.to(self.device)
            self.seen_embeddings = seen_t.to(self.device)
            self.unseen_embeddings = unseen_t.to(self.device)
        self.train_log_headers = list(TRAIN_HEADERS)
        self.val_log_headers = list(VAL_HEADERS)
        if self.unseen:
            for grp in ("seen", "unseen"):
                self.val_log_headers += ["val/%s/%s" % (grp, m) for m in ("pxl_acc", "class_acc", "mean_iu", "fwavacc")]
        self.val_log_headers.append("elapsed_time")
        self._init_log("train_log.csv", self.train_log_headers)
        self._init_log("val_log.csv", self.val_log_headers)

    # ---- the hot path, in the reference's call order (trainer_fcn.py:83-120) ----
    def _loss(self, score, target, target_embed):
        if self.loss_func == "cross_entropy":
            return utils.cross_entropy2d(score, target, size_average=False, accum_hook=self._accum_hook)
        table = None if target_embed is not None else self.embeddings  # labels-only loader: gather E[label] on the device
        fn = utils.cosine_loss if self.loss_func == "cos" else utils.mse_loss
        return fn(score, target, target_embed, table=table, accum_hook=self._accum_hook)

    def _forward_device(self, data, target, szn=False):
        target, target_embed = self._split_target(target)
        data, target, target_embed = self._to_device(data), self._to_device(target), self._to_device(target_embed)
        if szn:
            score, seen_mask_score = self.model(data, mode="both")  # trainer_fcn.py:135
        else:
            score = self.model(data, mode="fcn")  # trainer_fcn.py:97
        loss = self._loss(score, target, target_embed)
        # the label functions detach internally; the score itself is passed on so that a fused-head handle survives
        if szn:
            lbl_pred = utils.infer_lbl_szn_device(score, seen_mask_score.detach(), self.seen_embeddings,
                                                  self.unseen_embeddings)
        elif self.pixel_embeddings and self.forced_unseen:
            lbl_pred = utils.infer_lbl_forced_unseen_device(score, target, self.seen_embeddings, self.unseen_embeddings,
                                                            self.unseen)
        elif self.pixel_embeddings:
            lbl_pred = utils.infer_lbl_device(score, self.embeddings)
        else:
            lbl_pred = score.detach().max(1)[1]
        return score, loss, lbl_pred, target

    def forward(self, data, target):
        """Reference contract (``trainer_fcn.py:83-120``): ``(score, loss, lbl_pred np.ndarray, lbl_true CPU tensor)``."""
        score, loss, lbl_pred, lbl_true = self._forward_device(data, target)
        self._check_loss(loss)
        return score, loss, lbl_pred.cpu().numpy(), lbl_true.cpu()

    def forward_szn(self, data, target):
        """``trainer_fcn.py:123-143``: both heads, labels stitched by the seen-mask head."""
        if not self.pixel_embeddings or self.loss_func == "cross_entropy":
            raise ValueError("forward_szn needs a pixel-embedding model (trainer_fcn.py:138-141)")
        score, loss, lbl_pred, lbl_true = self._forward_device(data, target, szn=True)
        return score, loss, lbl_pred.cpu().numpy(), lbl_true.cpu()

    def validate(self, both_fcn_and_seenmask=False):
        """``trainer_fcn.py:181-292``.  Returns ``(val_loss, metrics)``; metrics = 4-tuple, or (all, seen, unseen)."""
        self.model.eval()
        val_loss, batches = 0.0, 0
        hist = None
        with torch.no_grad():
            for batch_idx, (data, target) in enumerate(self.val_loader):
                score, loss, lbl_pred, lbl_true = self._forward_device(data, target, szn=both_fcn_and_seenmask)
                h = utils.confusion_hist_device(lbl_true, lbl_pred, self.n_class, self.val_unseen if self.unseen else None)
                hist = h if hist is None else hist + h
                loss_val = float(loss.item())
                val_loss += loss_val
                batches += 1
                self._say("Test Epoch %-5d | Iteration %-5d | Loss %5.5f" % (self.epoch, batch_idx, loss_val))
                if self.visualize is not None:
                    self.visualize(data, lbl_true, lbl_pred)
        if hist is None:
            raise ValueError("empty validation loader")
        res = utils.metrics_from_hist(self._global_hist(hist))
        val_loss /= batches  # averaged over the batches of the loader (trainer_fcn.py:248)
        if self.unseen:
            metrics, seen_metrics, unseen_metrics = res
            self._scalars("val/seen", seen_metrics, self.epoch, names=("pxl_acc", "class_acc", "mean_iu", "fwavacc"))
            self._scalars("val/unseen", unseen_metrics, self.epoch, names=("pxl_acc", "class_acc", "mean_iu", "fwavacc"))
            row = [self.epoch, self.iteration, val_loss] + list(metrics) + list(seen_metrics) + list(unseen_metrics)
        else:
            metrics = res[0]
            row = [self.epoch, self.iteration, val_loss] + list(metrics)
        self._append_log("val_log.csv", row + [self._elapsed()])
        self._scalars("val", [val_loss] + list(metrics), self.epoch)
        mean_iu = metrics[2]
        is_best = mean_iu > self.best_mean_iu
        if is_best:
            self.best_mean_iu = mean_iu
        if self.log_dir:
            save_checkpoint(osp.join(self.log_dir, "checkpoint"), self.model, self.optim, self.epoch, self.iteration,
                            self.best_mean_iu)
            if is_best:
                shutil.copy(osp.join(self.log_dir, "checkpoint"), osp.join(self.log_dir, "best"))
        return val_loss, (res if self.unseen else metrics)

    def _stop_early(self):
        limit = EARLY_STOP_ITERS.get(self.dataset)
        return limit is not None and self.epoch * len(self.train_loader) > limit


class SeenmaskTrainer(_Base):
    """Seen-mask phase (``trainer_seenmask.py:18-175``): the frozen trunk feeds the 2-way ``seenmask_score`` head, target =
    "label is a seen class" (ignore label -1 -> 0, as upstream), mean cross entropy."""
    log_prefix = "seenmask_"
    tb_prefix = "seenmask"

    def __init__(self, cuda, model, optimizer, train_loader, val_loader, log_dir, dataset, max_epoch, tb_writer=None,
                 checkpoint=None, unseen=None, n_class=None, visualize=None, reducer=None):
        self._setup(cuda, model, optimizer, train_loader, val_loader, log_dir, dataset, max_epoch, tb_writer, n_class,
                    reducer)
        self.checkpoint = checkpoint if checkpoint is not None else {}
        self.unseen = list(unseen) if unseen else []
        self.visualize = visualize
        self.train_log_headers = list(TRAIN_HEADERS)
        self.val_log_headers = list(VAL_HEADERS) + ["elapsed_time"]
        self._init_log("seenmask_train_log.csv", self.train_log_headers)
        self._init_log("seenmask_val_log.csv", self.val_log_headers)

    def _forward_device(self, data, target):
        target, _ = self._split_target(target)
        data, target = self._to_device(data), self._to_device(target)
        target = utils.seenmask_target(target, self.unseen, self.n_class)  # trainer_seenmask.py:53-58, on the device
        score = self.model(data, mode="seenmask")                           # :64
        loss = utils.cross_entropy2d(score, target, size_average=True, accum_hook=self._accum_hook)  # :65
        lbl_pred = score.detach().max(1)[1]                                  # :67
        return score, loss, lbl_pred, target

    def forward(self, data, target):
        """Reference contract (``trainer_seenmask.py:50-70``)."""
        score, loss, lbl_pred, lbl_true = self._forward_device(data, target)
        return score, loss, lbl_pred.cpu().numpy(), lbl_true.cpu()

    def validate(self):
        """``trainer_seenmask.py:103-169``; the checkpoint handed in by ``train.py:177-181`` is rewritten as ``best``."""
        self.model.eval()
        val_loss, batches = 0.0, 0
        hist = None
        with torch.no_grad():
            for batch_idx, (data, target) in enumerate(self.val_loader):
                score, loss, lbl_pred, lbl_true = self._forward_device(data, target)
                h = utils.confusion_hist_device(lbl_true, lbl_pred, self.n_class)
                hist = h if hist is None else hist + h
                loss_val = float(loss.item())
                val_loss += loss_val
                batches += 1
                self._say("Seenmask Test Epoch %-5d | Iteration %-5d | Loss %5.5f" % (self.epoch, batch_idx, loss_val))
                if self.visualize is not None:
                    self.visualize(data, lbl_true, lbl_pred)
        if hist is None:
            raise ValueError("empty validation loader")
        metrics = utils.metrics_from_hist(self._global_hist(hist))[0]
        val_loss /= batches
        self._append_log("seenmask_val_log.csv", [self.epoch, self.iteration, val_loss] + list(metrics) + [self._elapsed()])
        self._scalars("val", [val_loss] + list(metrics), self.epoch)
        if metrics[2] > self.best_mean_iu:
            self.best_mean_iu = metrics[2]
        self.checkpoint["model_state_dict"] = self.model.state_dict()  # trainer_seenmask.py:167-168
        if self.log_dir:
            torch.save(self.checkpoint, osp.join(self.log_dir, "best"))
        return val_loss, metrics


def freeze_for_seenmask(model):
    """``train.py:166-171``: freeze everything but the seen-mask head; returns its parameters for the optimizer
    (``get_parameters(model, seenmask=True)``, ``train.py:309-313``: weights and bias of the two layers)."""
    for p in model.parameters():
        p.requires_grad = False
    params = list(model.seenmask_score.parameters()) + list(model.seenmask_upscore.parameters())
    for p in params:
        p.requires_grad = True
    return params

"""Host-side mirror of the reference's `attacker.py` for the hot path: `Patcher` (attacker.py:344-498) and
`PatchAttacker` call / train_step / test_step / fit / save_weights (attacker.py:24-342).

Same class names, constructor arguments, call signatures and metric names as the reference; what differs from the
Keras surface is listed in INTEGRATION.md ("API differences").  Tensors are torch CUDA tensors (device memory +
streams are the only thing torch provides here) and every per-pixel / per-anchor operation runs in libeotpatch.so:

    images' = Patcher([boxes, images])          -> eot_apply_fwd
    second_pass + max_scores + loss             -> victim forward (framework convs) + score_max_fwd
    tape.gradient(loss, [scale, patch])         -> score_max_bwd -> victim backward -> eot_apply_bwd (+ patch_tv_grad)
    optimizer.apply_gradients + constraints     -> adam_clip_update (after an NCCL all-reduce when sharded)
"""
from __future__ import annotations

import ast
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import anchors as _anchors
from . import ops, patch_io
from .ragged import RaggedBoxes
from .sampler import TransformSampler


def resolve_first_image(first_image: Optional[int], local_batch: int) -> int:
    """Global index of this rank's first image: explicit, else rank * local batch (even data-parallel shards)."""
    if first_image is not None:
        return int(first_image)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank() * int(local_batch)
    return 0


class Patcher:
    """apply patch to persons in an image (reference: attacker.py:344)."""

    def __init__(self, patch: torch.Tensor, scale_regressor: torch.Tensor, *args, min_patch_area=4, name=None,
                 seed: int = 0, perspective: float = 0.0, **kwargs):
        self._patch = patch
        self._scale = scale_regressor
        self.min_patch_area = min_patch_area
        self.name = name
        self.geometry = ops.PatchGeometry(tolerance=0.2, noise_amp=0.01, min_patch_area=float(min_patch_area))
        self.sampler = TransformSampler(seed, perspective=perspective)
        # global index of this rank's first image (data-parallel sharding: the draws hash the GLOBAL image index).
        # None: rank * local batch when torch.distributed is initialised (even shards), else 0; set it for uneven shards.
        self.first_image: Optional[int] = None
        self._step = 0
        self._workspace = None
        self.last_context: Optional[ops.ApplyContext] = None

    def __call__(self, inputs, transforms=None, out=None):
        return self.call(inputs, transforms=transforms, out=out)

    def call(self, inputs, transforms=None, out=None):
        """called during training by the attacker for each batch (attacker.py:490-498).

        inputs = [boxes (RaggedBoxes), images [B,H,W,3]].  `transforms=(params_u8 [N,48], print_wb [B,6])`
        replaces the internally drawn transform seeds (parity tests, replay)."""
        boxes, images = inputs
        if not isinstance(boxes, RaggedBoxes):
            boxes = RaggedBoxes.from_rows(boxes, images.device)
        n = int(boxes.values.shape[0])              # box capacity; the count in use is row_splits[-1], read on the device
        if transforms is None:
            params, print_wb = self.sampler.draw(self._step, resolve_first_image(self.first_image, images.shape[0]),
                                                 boxes.row_splits, n)
            self._step += 1
        else:
            params, print_wb = transforms
        out_images, _, ctx = ops.apply_forward(self._patch, self._scale, images, boxes.values, boxes.row_splits,
                                               params, print_wb, self.geometry, out=out, workspace=self._workspace)
        self._workspace = ctx.workspace
        self.last_context = ctx
        return out_images

    def backward(self, grad_images: torch.Tensor, grad_patch: Optional[torch.Tensor] = None, accumulate=False):
        """dL/dpatch for the last call (the part of tape.gradient at attacker.py:217 that crosses this layer)."""
        if self.last_context is None:
            raise RuntimeError("Patcher.backward called before Patcher.__call__")
        return ops.apply_backward(self.last_context, grad_images, grad_patch=grad_patch, accumulate=accumulate)


class PatchAttacker:
    """attack with malicious patches (reference: attacker.py:24)."""

    def __init__(self, model, initial_patch=None, config_override=None, visualize_freq=200, *,
                 patch_size: int = 640, device=None, seed: int = 0, process_group=None, perspective: float = 0.0,
                 cuda_graphs: bool = False, always_first_pass: bool = False, box_capacity: Optional[int] = None):
        self.model = model
        self.config = model.config
        if config_override:
            self.config.override(config_override)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if initial_patch is None:
            # np.random.uniform(-1., 1., size=(640, 640, 3)), scale .4 (attacker.py:43-44)
            patch_img = np.random.default_rng(seed).uniform(-1.0, 1.0, size=(patch_size, patch_size, 3))
            scale = 0.4
        else:
            patch_img, scale = patch_io.load_weights(initial_patch)
        # the only two variables updated during training (attacker.py:50-54); constraints = clip
        self._patch = torch.tensor(np.asarray(patch_img), dtype=torch.float32, device=self.device).contiguous()
        self._scale_regressor = torch.tensor(float(scale), dtype=torch.float32, device=self.device)
        self.visualize_freq = visualize_freq
        self._patcher = Patcher(self._patch, self._scale_regressor, name="Patcher", seed=seed, perspective=perspective)
        self._trainable_variables = [self._scale_regressor, self._patch]
        self.bins = np.arange(self.config.nms_configs["score_thresh"], .805, .01, dtype="float32")
        self.process_group = process_group
        self.learning_rate = 1e-2
        self._opt_step = 0
        n = self._patch.numel() + 1
        self._flat_var = None
        self._adam_m = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._adam_v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._anchors = None
        self._packed = None
        self.metrics = {}
        # CUDA graphs of the two fixed-shape victim passes (clean forward + score; attacked forward + score +
        # objective gradient + victim backward): the ~2000 framework launches of a step become two replays.
        # The patcher itself stays outside (its box count changes from step to step).
        self.cuda_graphs = bool(cuda_graphs)
        # run the clean victim pass + NMS even when `boxes` are supplied (synthetic benchmarks that must pay for it)
        self.always_first_pass = bool(always_first_pass)
        # None: the first pass reads its box count back (exact-size launches).  An int: no host read at all -- the ragged
        # NMS output is handed to the patcher at this capacity (total boxes per batch) and the count stays on the
        # device; surplus boxes are dropped and flagged (ops.check_workspace / `boxes_dropped`).
        self.box_capacity = box_capacity
        self._graphs = None
        self.graph_launches = 0           # libeotpatch kernels launched through CUDA-graph replays (see _build_graphs)

    # -- Keras-like surface ------------------------------------------------------------------------
    def compile(self, optimizer=None, learning_rate: Optional[float] = None, run_eagerly=False):
        """attacker_train.py:38 compiles with Adam(1e-2); only the learning rate is configurable here."""
        if learning_rate is not None:
            self.learning_rate = float(learning_rate)
        elif optimizer is not None and hasattr(optimizer, "learning_rate"):
            self.learning_rate = float(optimizer.learning_rate)

    @property
    def trainable_variables(self):
        return self._trainable_variables

    def _anchor_table(self, images: torch.Tensor) -> torch.Tensor:
        hw = (int(images.shape[1]), int(images.shape[2]))
        if self._anchors is None or self._anchors[0] != hw:
            c = self.config
            tab = _anchors.anchor_table(hw, c.min_level, c.max_level, c.num_scales, tuple(c.aspect_ratios), c.anchor_scale)
            self._anchors = (hw, torch.from_numpy(tab).to(self.device))
        return self._anchors[1]

    # -- passes ------------------------------------------------------------------------------------
    def _score(self, images: torch.Tensor):
        cls_outputs, box_outputs = self.model(images, pre_mode=None, post_mode=None)
        M, argmax, ncand, ctx = ops.score_max_forward(cls_outputs, box_outputs, self._anchor_table(images),
                                                      (images.shape[1], images.shape[2]),
                                                      num_classes=self.config.num_classes)
        return cls_outputs, box_outputs, M, argmax, ncand, ctx

    def first_pass(self, images: torch.Tensor):
        """clean pass through the victim (attacker.py:91-116).  Returns (boxes RaggedBoxes, scores list | None)."""
        from . import postprocess
        with torch.no_grad():
            cls_outputs, box_outputs, _, _, _, ctx = self._score(images)
            return postprocess.person_boxes_after_nms(self.config, ctx, box_outputs, self._anchor_table(images),
                                                      images.shape[1:3], thresh=True, box_capacity=self.box_capacity)

    def second_pass(self, images: torch.Tensor):
        """pass after addition of patches (attacker.py:118-141): per-image max candidate score + context."""
        cls_outputs, box_outputs, M, argmax, ncand, ctx = self._score(images)
        return cls_outputs, M, argmax, ncand, ctx

    # -- CUDA-graph replay of the victim passes ------------------------------------------------------
    def _build_graphs(self, images: torch.Tensor):
        dev = images.device
        g = dict(shape=tuple(images.shape))
        g["clean_in"] = torch.empty_like(images)
        g["patched"] = torch.empty_like(images).requires_grad_(True)
        g["clean_in"].copy_(images)
        with torch.no_grad():
            g["patched"].copy_(images)
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)   # the capture stream differs by design
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):                                   # warm-up: cuDNN algorithm search, lazy allocations
                with torch.no_grad():
                    self._score(g["clean_in"])
                cls_outputs, M, argmax, ncand, sctx = self.second_pass(g["patched"])
                dcls, dscale, data_loss = ops.score_max_backward(sctx, self._scale_regressor)
                torch.autograd.backward(cls_outputs, dcls)
                g["patched"].grad = None
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        count = _lib.load().eot_launch_count
        n0 = count()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            with torch.no_grad():
                c_cls, c_box, _, _, _, c_ctx = self._score(g["clean_in"])
        g["g1"], g["clean_box"], g["clean_ctx"] = g1, c_box, c_ctx
        n1 = count()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            cls_outputs, M, argmax, ncand, sctx = self.second_pass(g["patched"])
            dcls, dscale, data_loss = ops.score_max_backward(sctx, self._scale_regressor)
            torch.autograd.backward(cls_outputs, dcls)
        # libeotpatch kernels recorded in each graph: a replay launches them without passing through the C ABI's counter
        g["launches"] = (int(n1 - n0), int(count() - n1))
        g.update(g2=g2, M=M, argmax=argmax, ncand=ncand, dscale=dscale, data_loss=data_loss, grad=g["patched"].grad)
        self._graphs = g

    def _call_graphed(self, images: torch.Tensor, boxes, transforms):
        if self._graphs is None or self._graphs["shape"] != tuple(images.shape):
            self._build_graphs(images)
        g = self._graphs
        if boxes is None or self.always_first_pass:
            from . import postprocess
            g["clean_in"].copy_(images)
            g["g1"].replay()
            self.graph_launches += g["launches"][0]
            det_boxes, _ = postprocess.person_boxes_after_nms(self.config, g["clean_ctx"], g["clean_box"],
                                                              self._anchor_table(images), images.shape[1:3], thresh=True,
                                                              box_capacity=self.box_capacity)
            if boxes is None:
                boxes = det_boxes
        self._patcher([boxes, images], transforms=transforms, out=g["patched"].detach())
        g["g2"].replay()
        self.graph_launches += g["launches"][1]
        grad_patch = self._patcher.backward(g["grad"], grad_patch=self._grad_view())
        self._last = dict(max_scores=g["M"], data_loss=g["data_loss"], dscale=g["dscale"], ncand=g["ncand"])
        return [g["dscale"], grad_patch]

    def call(self, images: torch.Tensor, *, training=True, boxes: Optional[RaggedBoxes] = None, transforms=None):
        """called on each batch (attacker.py:172-219).  Returns [dL/dscale, dL/dpatch] when training, else
        (max_scores, argmax_anchor).  `boxes` overrides the first pass' detections (synthetic benchmarks)."""
        if training and self.cuda_graphs:
            return self._call_graphed(images, boxes, transforms)
        det_boxes, det_scores = self.first_pass(images) if boxes is None or self.always_first_pass else (None, None)
        if not isinstance(det_scores, list):
            det_scores = None                                           # sync-free first pass: scores stay on the device
        if boxes is None:
            boxes = det_boxes
        patched = self._patcher([boxes, images], transforms=transforms)
        if not training:
            # validation (attacker.py:318-326 -> call(training=False)): metrics of the same loss, and the reference's
            # return value: the second pass' person boxes / scores after NMS (attacker.py:204, 219)
            from . import postprocess
            with torch.no_grad():
                cls_outputs, box_outputs, M, argmax, ncand, sctx = self._score(patched)
                self._record_eval_metrics(M)
                boxes_pred, scores_pred = postprocess.person_boxes_after_nms(
                    self.config, sctx, box_outputs, self._anchor_table(patched), patched.shape[1:3], thresh=False)
            self._last = dict(max_scores=M, argmax=argmax, ncand=ncand, first_pass_scores=det_scores, boxes=boxes)
            if det_scores is not None:
                asr = postprocess.calc_asr(det_scores, scores_pred)
                self.metrics.update(asr=asr, asr_to_scale=asr / max(float(self._scale_regressor), 1e-12))
            return boxes_pred, scores_pred
        patched.requires_grad_(True)
        cls_outputs, M, argmax, ncand, sctx = self.second_pass(patched)
        dcls, dscale, data_loss = ops.score_max_backward(sctx, self._scale_regressor)
        torch.autograd.backward(cls_outputs, dcls)                       # victim backward on the framework path
        grad_patch = self._patcher.backward(patched.grad, grad_patch=self._grad_view())
        self._last = dict(max_scores=M, data_loss=data_loss, dscale=dscale, ncand=ncand)
        return [dscale, grad_patch]

    def __call__(self, images, *, training=True, **kw):
        return self.call(images, training=training, **kw)

    # -- steps -------------------------------------------------------------------------------------
    def _grad_view(self) -> torch.Tensor:
        """dL/dpatch is written straight into the head of the packed all-reduce buffer [dpatch | dscale | loss | sum M | sum M^2]."""
        n = self._patch.numel()
        if self._packed is None:
            self._packed = torch.empty(n + 4, dtype=torch.float32, device=self.device)
        return self._packed[:n].view_as(self._patch)

    def _pack(self, dscale, grad_patch, M):
        n = self._patch.numel()
        buf = self._packed
        if grad_patch.data_ptr() != buf.data_ptr():
            buf[:n] = grad_patch.reshape(-1)
        ops.pack_scalars_(buf[n:], M, dscale, self._last["data_loss"])
        return buf

    def train_step(self, inputs, boxes: Optional[RaggedBoxes] = None, transforms=None, global_batch: Optional[int] = None):
        """called for each batch during training (attacker.py:307-316)."""
        dscale, grad_patch = self.call(inputs, training=True, boxes=boxes, transforms=transforms)
        M = self._last["max_scores"]
        buf = self._pack(dscale, grad_patch, M)
        B = inputs.shape[0]
        if self.process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                              and torch.distributed.get_world_size() > 1):
            # loss is a SUM over images (attacker.py:193): a sum all-reduce reproduces the single-GPU step
            torch.distributed.all_reduce(buf, group=self.process_group)
            B = global_batch if global_batch is not None else B * torch.distributed.get_world_size(self.process_group)
        n = self._patch.numel()
        g_patch = buf[:n].view_as(self._patch)
        tv = ops.tv_grad_(self._patch, g_patch, 1e-5)                   # + 1e-5 * d TV/d patch, once (attacker.py:192-193)
        self._record_metrics(None, buf, tv, B)                          # as the reference: metrics of the step's own variables,
        self._apply_gradients(buf[n:n + 1], g_patch)                    # recorded before apply_gradients (attacker.py:196-201, 315)
        return self.metrics

    def _apply_gradients(self, g_scale: torch.Tensor, g_patch: torch.Tensor):
        """optimizer.apply_gradients + variable constraints (attacker.py:315, 51-54): one fused Adam+clip per variable."""
        self._opt_step += 1
        n = self._patch.numel()
        ops.adam_clip_(self._patch.view(-1), self._adam_m[:n], self._adam_v[:n], g_patch.reshape(-1), self._opt_step,
                       lr=self.learning_rate, lo=-1.0, hi=1.0)
        ops.adam_clip_(self._scale_regressor.view(1), self._adam_m[n:], self._adam_v[n:], g_scale, self._opt_step,
                       lr=self.learning_rate, lo=0.0, hi=1.0)

    def test_step(self, inputs, boxes: Optional[RaggedBoxes] = None, transforms=None):
        """called for each batch during validation (attacker.py:318-326)."""
        self.call(inputs, training=False, boxes=boxes, transforms=transforms)
        return self.metrics

    def _record_metrics(self, M, buf, tv, B=None):
        """add_metric calls of attacker.py:196-201 (device scalars; nothing is synchronised here).  `asr` / `asr_to_scale`
        need both passes' NMS results on the host and are recorded by the validation path only."""
        n = self._patch.numel()
        m = ops.step_metrics(buf[n:], tv, self._scale_regressor, B)             # one launch; the dict holds views of it
        self.metrics = dict(loss=m[0], scale_loss=m[1], mean_max_score=m[2], std_max_score=m[3], tv_loss=m[4], scale=m[5])

    def _record_eval_metrics(self, M):
        """the same metrics for a validation batch (attacker.py:190-201 run with training=False)."""
        sc = self._scale_regressor
        scale_losses = (M - sc) ** 2
        tv = ops.tv_value(self._patch)
        self.metrics = dict(loss=(M * M + scale_losses).sum() + 1e-5 * tv, scale=sc, scale_loss=scale_losses.sum(), tv_loss=tv,
                            mean_max_score=M.mean(), std_max_score=M.std(unbiased=False))

    def fit(self, train_data, validation_data=None, epochs: int = 1, steps_per_epoch: Optional[int] = None,
            validation_steps: Optional[int] = None, save_dir: Optional[str] = None,
            save_file: str = "patch_{epoch:02d}_{val_asr_to_scale:.4f}", verbose: bool = True):
        """The loop `attacker_train.py:52-65` drives through Keras' Model.fit + ModelCheckpoint(save_weights_only=True,
        save_freq='epoch'): per epoch `steps_per_epoch` train steps, `validation_steps` validation steps, epoch means
        of the metrics (validation ones prefixed `val_`), one `save_weights` under `save_dir / save_file` (same name
        template).  Batches are [B,H,W,3] float32 CUDA tensors (or anything `torch.as_tensor` takes); Keras callback
        objects are not interpreted.  Returns the history as a list of dicts."""
        def mean_metrics(acc, count, prefix=""):
            return {prefix + k: float(v) / max(count, 1) for k, v in acc.items()}

        def run(data, steps, step_fn):
            acc, count = {}, 0
            for i, batch in enumerate(data):
                if steps is not None and i >= steps:
                    break
                x = torch.as_tensor(batch[0] if isinstance(batch, (tuple, list)) else batch, dtype=torch.float32, device=self.device)
                for k, v in step_fn(x).items():
                    acc[k] = acc.get(k, 0.0) + float(v)
                count += 1
            return acc, count

        history = []
        for epoch in range(1, epochs + 1):
            logs = mean_metrics(*run(train_data, steps_per_epoch, self.train_step))
            if validation_data is not None:
                logs.update(mean_metrics(*run(validation_data, validation_steps, self.test_step), prefix="val_"))
            logs["epoch"] = epoch
            if save_dir is not None:
                fmt = {"val_asr_to_scale": float("nan"), "val_loss": float("nan"), **logs}
                self.save_weights(os.path.join(save_dir, save_file.format(**fmt)))
            if verbose:
                print("epoch %d: " % epoch + ", ".join(f"{k}={v:.4f}" for k, v in logs.items() if k != "epoch"))
            history.append(logs)
        return history

    def save_weights(self, dirpath, **kwargs):
        """save patch and current scale to disk (attacker.py:328-341): scale.txt, patch.png, patch.tiff."""
        patch_io.save_weights(dirpath, self._patch.detach().cpu().numpy(), float(self._scale_regressor),
                              self.config.mean_rgb, self.config.stddev_rgb)

"""TensorFlow-side binding of the custom ops in eot_patch_ops.cc: drop-in `Patcher` for attacker.py.

STATUS: source only -- TensorFlow is absent from the build image and the GPU box, so this module is not
imported or tested there.  It shows the reference-side change: in
/root/reference/attacker.py replace `self._patcher = Patcher(...)` (line 59) by this class; everything else
(`call`, `train_step`, `attacker_train.py`) stays as it is.
"""
import os

import numpy as np
import tensorflow as tf

_ops = tf.load_op_library(os.path.join(os.path.dirname(__file__), "_eot_patch_ops.so"))


def _draw_params(n):
    """Transform seeds of n boxes as uint8 [n,48] EotBoxParams records (attacker.py:426-427,436,473-474)."""
    uy = tf.random.uniform([n]); ux = tf.random.uniform([n])
    delta = tf.random.uniform([n], -.3, .3)
    ang = tf.random.uniform([n], -20. * np.pi / 180., 20. * np.pi / 180.)
    zeros = tf.zeros([n])
    floats = tf.stack([uy, ux, delta, tf.cos(ang), tf.sin(ang), zeros, zeros, -tf.ones([n])], axis=1)      # [n,8]
    keys = tf.random.uniform([n, 2], 0, 2 ** 31 - 1, dtype=tf.int32)
    words = tf.concat([tf.bitcast(floats, tf.int32), keys, tf.zeros([n, 2], tf.int32)], axis=1)            # [n,12]
    return tf.reshape(tf.bitcast(words, tf.uint8), [n, 48])


@tf.custom_gradient
def _apply(patch, scale, images, boxes, row_splits, params, print_wb):
    patched, _, ws = _ops.eot_patch_apply(patch=patch, scale=scale, images=images, boxes=boxes, row_splits=row_splits,
                                          params=params, print_wb=print_wb)

    def grad(g):
        gp = _ops.eot_patch_apply_grad(patch=patch, print_wb=print_wb, grad_patched=g, workspace=ws,
                                       images_shape=tf.shape(images), num_boxes=tf.shape(boxes)[0])
        return gp, None, None, None, None, None, None       # images are not variables (attacker.py:217)
    return patched, grad


class Patcher(tf.keras.layers.Layer):
    """apply patch to persons in an image -- same constructor and call signature as attacker.Patcher."""

    def __init__(self, patch: tf.Variable, scale_regressor: tf.Variable, *args, min_patch_area=4, **kwargs):
        super().__init__(*args, trainable=False, **kwargs)
        self._patch = patch
        self._scale = scale_regressor
        self.min_patch_area = min_patch_area

    def call(self, inputs):
        boxes, images = inputs                         # boxes: tf.RaggedTensor [B,(n),4]
        b = tf.shape(images)[0]
        w = tf.random.normal((b, 3), .5, .1)
        bias = tf.random.normal((b, 3), 0., .01)
        flat = boxes.flat_values
        return _apply(self._patch, self._scale, images, flat, tf.cast(boxes.row_splits, tf.int32),
                      _draw_params(tf.shape(flat)[0]), tf.concat([w, bias], axis=1))


def first_pass_boxes(cls_outputs, box_outputs, anchors, image_hw, nms_configs):
    """Replacement for the post-victim part of `PatchAttacker.first_pass` (attacker.py:100-116,143-170): person
    candidates -> NonMaxSuppressionV5 -> clip_boxes -> (ragged boxes, ragged scores), one op on the GPU."""
    gaussian = nms_configs.method == 'gaussian'
    out = _ops.eot_first_pass_boxes(
        cls_levels=cls_outputs, box_levels=box_outputs, anchors=anchors, image_height=float(image_hw[0]),
        image_width=float(image_hw[1]), max_output_size=nms_configs.max_output_size,
        iou_threshold=1.0 if gaussian else (nms_configs.iou_thresh or .5), score_threshold=nms_configs.score_thresh or 0.,
        soft_nms_sigma=(nms_configs.sigma or .5) / 2 if gaussian else 0., score_floor=nms_configs.score_thresh or 0.)
    n = out.row_splits[-1]
    boxes = tf.RaggedTensor.from_row_splits(out.ragged_boxes[:n], tf.cast(out.row_splits, tf.int64))
    scores = tf.RaggedTensor.from_row_splits(out.ragged_scores[:n], tf.cast(out.row_splits, tf.int64))
    return boxes, scores

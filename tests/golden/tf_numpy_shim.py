"""A NumPy stand-in for the ~60 TensorFlow / TensorFlow-Addons entry points that the reference's hot path calls, so that
THE REFERENCE'S OWN CODE runs verbatim in the build container (TensorFlow is not installed):

  * `attacker.Patcher` (/root/reference/attacker.py:344-498) and `brightness_matcher.BrightnessMatcher`
    (/root/reference/brightness_matcher.py:14-73)
  * `attack_detection.Masker` (/root/reference/attack_detection.py:321-498), evaluation and training branch
  * `PatchAttacker.first_pass / second_pass / filter_valid_boxes / _postprocessing / calc_asr`
    (/root/reference/attacker.py:69-170,238-255) together with the vendored automl code they call, imported for real:
    `tf2/postprocess.py` (pre_nms, nms, clip_boxes), `tf2/anchors.py`, `utils.py`, `hparams_config.py`

What this pins and what it does not: the reference's Python -- control flow, the order and association of every
arithmetic expression (each evaluated as one float32 NumPy op, as TF eager does), casts, pads, the where / clip /
scatter sequence, the box filters, ragged masks, the loops over boxes and images, the NMS call arguments -- is
executed as written.  The LEAF kernels that live inside the TF / TFA wheels (ScaleAndTranslate,
ImageProjectiveTransformV3, rgb_to_yuv / yuv_to_rgb, reduce_mean, NonMaxSuppressionV5, sigmoid / exp) are the oracle's
restatements (oracle/tfops.py, oracle/nms.py); they stay unpinned.  Random draws are served from a queue the caller
fills with the explicit transform seeds, mapped to the requested range the way TF's random ops do
(u * (maxval - minval) + minval; mean + stddev * z).

Only tests/golden/make_golden.py imports this module (it needs /root/reference); its outputs are committed as
tests/golden/{patcher_ref,patcher_ref2,brightness_ref,masker_ref,objective_ref}.npz.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

import numpy as np

from oracle import tfops

F = np.float32


class Variable(np.ndarray):
    """tf.Variable: an ndarray that can be assigned in place (0-d integer variables index Python lists)."""

    def __new__(cls, value, trainable=True, **kw):
        return np.array(value).view(cls)

    def assign(self, v):
        self[...] = v
        return self

    def assign_add(self, v):
        self[...] = self + v
        return self

    def __index__(self):
        return int(self)


class RandomQueue:
    """Serves tf.random.* calls in call order from pre-drawn values."""

    def __init__(self):
        self.items = []

    def push(self, kind, value):
        self.items.append((kind, value))

    def pop(self, kind):
        if not self.items:
            raise RuntimeError(f"random queue empty (wanted {kind})")
        k, v = self.items.pop(0)
        if k != kind:
            raise RuntimeError(f"random queue order: wanted {kind}, next is {k}")
        return v


QUEUE = RandomQueue()


def _f(x):
    return np.asarray(x, dtype=F)


def _uniform(shape=(), minval=0.0, maxval=1.0, **kw):
    u = _f(QUEUE.pop("uniform"))
    shape = tuple(int(s) for s in np.atleast_1d(shape)) if np.ndim(shape) else tuple(shape)
    u = u.reshape(shape)
    lo, hi = _f(minval), _f(maxval)
    return (u * (hi - lo) + lo).astype(F)                     # random_uniform: rnd * (maxval - minval) + minval


def _normal(shape, mean=0.0, stddev=1.0, **kw):
    z = _f(QUEUE.pop("normal")).reshape(tuple(shape))
    return (z * _f(stddev) + _f(mean)).astype(F)              # random_normal: rnd * stddev + mean


def _random_brightness(image, max_delta):
    delta = _uniform((), -max_delta, max_delta)               # tf.image.random_brightness -> adjust_brightness(image, delta)
    return (image + delta).astype(F)


def _resize(images, size, method="bilinear", antialias=False, **kw):
    assert antialias and method == "bilinear"
    return tfops.aa_resize(_f(images), int(size[0]), int(size[1]))


def _rgb_to_yuv(x):
    return tfops.dot3(_f(x), tfops.RGB2YUV)


def _yuv_to_rgb(x):
    return tfops.dot3(_f(x), tfops.YUV2RGB)


def _rotate(images, angles, interpolation="nearest", fill_mode="constant", fill_value=0.0, **kw):
    assert interpolation == "bilinear"
    img = _f(images)
    a = F(angles)
    return tfops.projective_bilinear(img, tfops.rotation_transform(F(np.cos(a)), F(np.sin(a)), img.shape[0]), float(fill_value))


def _where(cond, x=None, y=None):
    if x is None:
        return np.argwhere(np.asarray(cond))
    return np.where(cond, x, y)


def _while_loop(cond, body, loop_vars, **kw):
    vars_ = list(loop_vars)
    while bool(cond(*vars_)):
        vars_ = list(body(*vars_))
    return vars_


def _map_rows(fn, elems, **kw):
    out = [fn(e) for e in elems]
    if len(out) and isinstance(out[0], (tuple, list)):          # map_fn with a tuple signature (Masker: image, mask)
        return tuple(np.stack([o[k] for o in out]) for k in range(len(out[0])))
    return np.stack(out) if len(out) else np.zeros((0,), F)


def _shuffle(x):
    return np.asarray(x)[np.asarray(QUEUE.pop("perm"))]       # tf.random.shuffle: permutation of the first axis


def _random_flip(axis):
    def flip(x):
        flags = np.asarray(QUEUE.pop("flip"), bool)             # one coin per image of the batch
        x = np.array(x, copy=True)
        x[flags] = np.flip(x[flags], axis=axis)
        return x
    return flip


def _cast(x, dtype):
    x = np.asarray(x)
    if np.issubdtype(dtype, np.integer) and np.issubdtype(x.dtype, np.floating):
        return np.trunc(x).astype(dtype)                      # tf.cast float -> int truncates toward zero
    return x.astype(dtype)


def _pad(x, paddings, constant_values=0):
    return np.pad(x, [tuple(int(v) for v in p) for p in np.asarray(paddings)], constant_values=constant_values).astype(x.dtype)


def _scatter_nd_update(tensor, indices, updates):
    out = np.array(tensor, copy=True)
    idx = np.asarray(indices)
    out[idx[..., 0], idx[..., 1]] = updates
    return out


def _gather_nd(params, indices):
    idx = np.asarray(indices)
    return np.asarray(params)[tuple(idx[:, k] for k in range(idx.shape[1]))] if len(idx) else np.asarray(params)[:0]


class Ragged:
    """tf.RaggedTensor with one ragged dimension: a list of per-row arrays."""

    def __init__(self, rows):
        self.rows = [np.asarray(r) for r in rows]

    def nrows(self):
        return len(self.rows)

    def __getitem__(self, key):
        if isinstance(key, tuple):                              # r[:, :, k]
            assert key[0] == slice(None) and key[1] == slice(None)
            return Ragged([r[(slice(None),) + tuple(key[2:])] for r in self.rows])
        return self.rows[int(key)]

    def _zip(self, other, fn):
        if isinstance(other, Ragged):
            return Ragged([fn(a, b) for a, b in zip(self.rows, other.rows)])
        return Ragged([fn(a, other) for a in self.rows])

    def __sub__(self, o): return self._zip(o, lambda a, b: a - b)
    def __mul__(self, o): return self._zip(o, lambda a, b: a * b)
    def __truediv__(self, o): return self._zip(o, lambda a, b: a / b)

    @property
    def flat_values(self):
        return np.concatenate(self.rows, axis=0) if self.rows else np.zeros((0,), F)

    @staticmethod
    def from_tensor(tensor, lengths):
        return Ragged([np.asarray(tensor)[i][:int(n)] for i, n in enumerate(np.asarray(lengths))])


def _rag(fn):
    def op(a, b):
        if isinstance(a, Ragged):
            return a._zip(b, fn)
        return fn(a, b)
    return op


def _boolean_mask(data, mask):
    rows = data.rows if isinstance(data, Ragged) else list(np.asarray(data))
    masks = mask.rows if isinstance(mask, Ragged) else list(np.asarray(mask))
    return Ragged([np.asarray(r)[np.asarray(m, bool)] for r, m in zip(rows, masks)])


def _reduce_max(x, axis=None):
    if isinstance(x, Ragged):                                   # empty rows reduce to the lowest float (attacker.py:190)
        assert axis == 1
        return np.array([r.max() if len(r) else np.finfo(F).min for r in x.rows], F)
    return np.max(x, axis=axis)


def _zeros_like(x):
    return Ragged([np.zeros_like(r) for r in x.rows]) if isinstance(x, Ragged) else np.zeros_like(x)


def _nms_v5(boxes, scores, max_output_size, iou_threshold, score_threshold, soft_nms_sigma, pad_to_max_output_size=False):
    from oracle import nms as onms                              # the restated leaf kernel (unpinned against TF)
    sel, sel_scores = onms.non_max_suppression_v5(boxes, scores, int(max_output_size), float(iou_threshold),
                                                  float(score_threshold), float(soft_nms_sigma))
    n = len(sel)
    if pad_to_max_output_size:
        idx = np.zeros(int(max_output_size), np.int32); idx[:n] = sel
        sc = np.zeros(int(max_output_size), F); sc[:n] = sel_scores
        return idx, sc, np.int32(n)
    return sel.astype(np.int32), sel_scores, np.int32(n)


def _sigmoid(x):
    x = _f(x)
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F)   # correctly rounded float32 sigmoid, as the oracle


class Layer:
    def __init__(self, *args, trainable=True, name=None, **kwargs):
        self.name = name

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


def build_modules():
    tf = types.ModuleType("tensorflow")
    tf.__getattr__ = lambda name: mock.MagicMock(name="tensorflow." + name)     # anything else (TensorSpec, ...) is inert
    tf.float32, tf.int32 = np.float32, np.int32
    tf.Variable = Variable
    tf.constant = lambda v, dtype=None: np.asarray(v, dtype=dtype if dtype is not None else (F if isinstance(v, float) else None))
    tf.cast = _cast
    tf.shape = lambda x: np.asarray(np.shape(x), np.int32)
    tf.unstack = lambda x, num=None, axis=0: [np.take(np.asarray(x), i, axis=axis) for i in range(np.shape(x)[axis])]
    tf.stack = lambda xs, axis=0: np.stack([np.asarray(x) for x in xs], axis=axis)
    tf.reshape = lambda x, shape: np.reshape(x, tuple(int(s) for s in shape))
    tf.where = _where
    tf.less = _rag(np.less)
    tf.gather_nd = _gather_nd
    tf.while_loop = _while_loop
    tf.map_fn = _map_rows
    tf.vectorized_map = _map_rows
    tf.meshgrid = lambda *a, indexing="xy": np.meshgrid(*a, indexing=indexing)
    tf.range = lambda a, b=None: np.arange(a, b, dtype=np.int32) if b is not None else np.arange(a, dtype=np.int32)
    tf.floor, tf.maximum, tf.minimum = np.floor, np.maximum, np.minimum
    tf.cond = lambda pred, t, f: t() if bool(pred) else f()
    tf.pad = _pad
    tf.clip_by_value = lambda x, lo, hi: np.clip(x, F(lo), F(hi)).astype(F)
    tf.tensor_scatter_nd_update = _scatter_nd_update
    tf.reduce_mean = lambda x: tfops.mean_f64(_f(x))
    tf.function = lambda fn=None, **kw: fn if fn is not None else (lambda f: f)
    tf.math = types.SimpleNamespace(floor=np.floor, ceil=np.ceil, sigmoid=_sigmoid,
                                    exp=lambda x: np.exp(_f(x).astype(np.float64)).astype(F),
                                    argmax=lambda x, axis=-1, output_type=np.int32: np.argmax(x, axis=axis).astype(output_type))
    tf.__path__ = []                                            # a package: tensorflow.compat.v1 etc. resolve to inert mocks
    tf.zeros_like = _zeros_like
    tf.Tensor = np.ndarray
    tf.size = lambda x: np.int32(np.size(x))
    tf.convert_to_tensor = lambda x, dtype=None: np.asarray(x, dtype=dtype)
    tf.equal = _rag(np.equal)
    tf.less_equal = _rag(np.less_equal)
    tf.greater_equal = _rag(np.greater_equal)
    tf.greater = _rag(np.greater)
    tf.logical_and = _rag(np.logical_and)
    tf.concat = lambda xs, axis: np.concatenate([np.asarray(x) for x in xs], axis=int(axis))
    tf.transpose = lambda x, perm: np.transpose(x, perm)
    tf.tile = lambda x, reps: np.tile(x, [int(r) for r in reps])
    tf.expand_dims = lambda x, axis: np.expand_dims(x, axis)
    tf.gather = lambda params, indices: np.asarray(params)[np.asarray(indices)]
    tf.reduce_max = _reduce_max
    tf.ragged = types.SimpleNamespace(boolean_mask=_boolean_mask)
    tf.RaggedTensor = Ragged
    tf.TensorSpec = lambda *a, **k: None
    tf.raw_ops = types.SimpleNamespace(NonMaxSuppressionV5=_nms_v5)
    tf.name_scope = lambda name: __import__("contextlib").nullcontext()
    tf.random = types.SimpleNamespace(uniform=_uniform, normal=_normal, shuffle=_shuffle)
    tf.image = types.SimpleNamespace(resize=_resize, random_brightness=_random_brightness, rgb_to_yuv=_rgb_to_yuv,
                                     yuv_to_rgb=_yuv_to_rgb, random_flip_left_right=_random_flip(2),
                                     random_flip_up_down=_random_flip(1))
    tf.keras = types.SimpleNamespace(layers=types.SimpleNamespace(Layer=Layer), Model=Layer,
                                     utils=types.SimpleNamespace(Sequence=object),
                                     backend=types.SimpleNamespace(epsilon=lambda: 1e-7))
    tfa = types.ModuleType("tensorflow_addons")
    tfa.image = types.SimpleNamespace(rotate=_rotate)
    return tf, tfa


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Any other third-party import of the reference modules (tfplot, tifffile, matplotlib, util, tf2, ...) becomes an
    inert mock: none of it is touched by Patcher / BrightnessMatcher."""
    NAMES = ("tfplot", "tifffile", "matplotlib", "util", "tf2", "requests", "automl", "visualize", "generator", "metrics",
             "custom_callbacks", "train_data_generator", "hparams_config", "utils", "seaborn")

    REAL = ("tf2", "tf2.postprocess", "tf2.anchors", "utils", "hparams_config")     # with with_automl=True

    def find_spec(self, name, path, target=None):
        if self.with_automl and name in self.REAL:
            return None
        if name.split(".")[0] in self.NAMES or name.startswith("tensorflow.") or name.split(".")[0] in ("absl", "object_detection"):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def __init__(self, with_automl=False):
        self.created = []
        self.with_automl = with_automl

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__name__, m.__spec__, m.__loader__ = [], spec.name, spec, self
        self.created.append(spec.name)
        return m

    def exec_module(self, module):
        pass


def import_reference_attacker(reference_root="/root/reference", module="attacker", with_automl=False):
    """-> the reference's `attacker` (or `attack_detection`) module, imported on top of the shim.  with_automl: the
    vendored automl modules the objective path calls (tf2/postprocess.py, tf2/anchors.py, utils.py, hparams_config.py)
    are imported for real as well (on the same shim); everything else stays an inert mock."""
    tf, tfa = build_modules()
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow_addons"] = tfa
    finder = _StubFinder(with_automl)
    sys.meta_path.insert(0, finder)
    paths = [reference_root] + ([reference_root + "/automl/efficientdet"] if with_automl else [])
    sys.path[:0] = paths
    try:
        for name in ("attacker", "brightness_matcher", "attack_detection", "tf2", "tf2.postprocess", "tf2.anchors", "utils",
                     "hparams_config", "nms_np"):
            sys.modules.pop(name, None)
        mod = importlib.import_module(module)                   # the reference module
        if with_automl:
            mod.hparams_config = importlib.import_module("hparams_config")
        return mod
    finally:
        sys.meta_path.remove(finder)
        for p in paths:
            sys.path.remove(p)
        for name in finder.created:                             # the inert stand-ins must not leak into later imports
            sys.modules.pop(name, None)

"""tf_shim — TEST INFRASTRUCTURE ONLY: a numpy-eager stand-in for the TensorFlow 1.x symbols that the reference's
four detection-head layer files touch, so that those files can be imported UNMODIFIED from /root/reference and
executed here (TensorFlow itself is not installable in this image: no network, Python 3.12).

    import tests.tf_shim as shim
    shim.install()                       # registers `tensorflow`, `keras`, `keras.backend`, `keras.layers`,
                                         # `skimage`, `skimage.transform`, `h5py` in sys.modules
    from MaskRCNN.building_blocks.proposals_tf import Proposals      # the reference's own code, verbatim

What runs where
---------------
* Every line of graph-building Python in proposals_tf.py / maskrcnn.py (roi_pooling) / data_processor.py
  (BuildDetectionTargets) / detection.py (DetectionLayer) / utils.py (norm_boxes_tf) executes as written: the index
  plumbing (`where`, `gather_nd`, `boolean_mask`, `meshgrid`, `stack`, `pad`, `unique`, `map_fn`,
  `sets.set_intersection`, `sparse_tensor_to_dense`, …), the operator arithmetic and its dtype rules are emulated
  op by op on numpy arrays, eagerly, in the dtype TensorFlow would use:
    - a Python scalar or numpy array combined with a Tensor is converted to the TENSOR's dtype first
      (`ops.convert_to_tensor(y, dtype=x.dtype)`), so `rpn_bbox * np.reshape(stddev, [1,1,4])` is an fp32 multiply
      and `(1 / 0.33) * tf.cast(n, tf.float32)` multiplies by float32(3.0303...);
    - two Tensors of different dtype raise, as in TensorFlow;
    - float32 add/sub/mul/div/sqrt are IEEE single operations (numpy ufuncs, one rounding per op, no FMA);
    - `tf.exp` / `tf.log` are evaluated in float64 and rounded to float32 (correctly rounded). TensorFlow's CPU
      kernels use Eigen's vectorised pexp/plog, which may differ from this by 1 ulp on some inputs; the north-star
      tolerances (rtol 1e-5 on boxes/deltas) cover that difference, the integer outputs do not depend on it;
    - `tf.round` is half-to-even; float -> int32 casts follow x86 `cvttss2si` (NaN/overflow -> INT_MIN);
    - `tf.minimum` / `tf.maximum` follow Eigen's `(b < a) ? b : a` / `(a < b) ? b : a` NaN behaviour.
* The three library KERNELS whose arithmetic lives inside TensorFlow's C++ (`tf.nn.top_k`,
  `tf.image.non_max_suppression`, `tf.image.crop_and_resize`) delegate to the oracle's restatements of those kernels
  (oracle.topk / oracle.nms / oracle.crop_and_resize, anchored to TensorFlow's kernel-test vectors in
  tests/test_tf_known_answers.py).  So the goldens made with this shim pin the reference's COMPOSITIONS (what the
  reference itself owns) and leave the three kernels anchored, not pinned.
* `tf.random_shuffle` has no seed in the reference (data_processor.py:587,:597).  The shim replaces it by an injected
  permutation (`set_shuffle_perms`): shuffle(x) = x[[p for p in perm if p < len(x)]] — the same explicit-permutation
  contract the CUDA path and the oracle expose.
* keras / skimage / h5py are inert stubs (the dense classifier head of maskrcnn.py is out of scope).
"""
from __future__ import annotations

import collections
import sys
import types

import numpy as np

f32 = np.float32

_shuffle_perms: list = []


def set_shuffle_perms(perms):
    """Queue the permutations consumed, in call order, by the following tf.random_shuffle calls."""
    _shuffle_perms[:] = [np.asarray(p) for p in perms]


# ----------------------------------------------------------------------------------------------- Tensor
class TensorShape(tuple):
    def as_list(self):
        return list(self)


def _x86_f2i32(x):
    x = np.asarray(x)
    if x.dtype.kind != "f":
        return x.astype(np.int32)
    ok = (x > -2147483904.0) & (x < 2147483648.0)
    with np.errstate(invalid="ignore"):
        v = np.where(ok, x, 0).astype(np.int64)
    return np.where(ok, v, -2 ** 31).astype(np.int32)


def _unwrap(x):
    return x.a if isinstance(x, Tensor) else x


def _conv(x, dtype):
    """ops.convert_to_tensor(x, dtype): Tensors must already have the dtype, anything else is converted."""
    if isinstance(x, Tensor):
        if dtype is not None and x.a.dtype != np.dtype(dtype):
            raise TypeError(f"tf_shim: Tensor dtype {x.a.dtype} where {np.dtype(dtype)} is required "
                            f"(TensorFlow raises here as well)")
        return x.a
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        x = [np.asarray(_unwrap(e)) for e in x]
    a = np.asarray(x)
    if dtype is None:
        if a.dtype == np.float64 and not isinstance(x, np.ndarray):
            return a.astype(f32)          # Python floats become float32 tensors
        if a.dtype == np.int64 and not isinstance(x, np.ndarray):
            return a.astype(np.int32)     # Python ints become int32 tensors
        return a
    return a.astype(dtype)


def _pair(x, y):
    """Operands of a binary op in the dtype TensorFlow computes in."""
    if isinstance(x, Tensor):
        return x.a, _conv(y, x.a.dtype)
    if isinstance(y, Tensor):
        return _conv(x, y.a.dtype), y.a
    a = _conv(x, None)
    return a, _conv(y, a.dtype)


class Tensor:
    __array_priority__ = 1000     # numpy hands binary ops with a Tensor on the right over to the Tensor
    __hash__ = object.__hash__

    def __init__(self, a):
        self.a = np.asarray(a)

    # -- static information the reference prints / reads
    @property
    def shape(self):
        return TensorShape(self.a.shape)

    @property
    def dtype(self):
        return self.a.dtype

    def get_shape(self):
        return TensorShape(self.a.shape)

    def set_shape(self, shape):
        want = [d for d in shape]
        if len(want) != self.a.ndim or any(w is not None and w != s for w, s in zip(want, self.a.shape)):
            raise ValueError(f"tf_shim.set_shape: {self.a.shape} is not compatible with {shape}")

    def numpy(self):
        return self.a

    def __repr__(self):
        return f"<tf_shim.Tensor shape={self.a.shape} dtype={self.a.dtype}>"

    def __len__(self):
        return self.a.shape[0]

    def __iter__(self):
        return (Tensor(self.a[i]) for i in range(self.a.shape[0]))

    def __int__(self):
        return int(self.a)

    def __index__(self):
        return int(self.a)

    def __float__(self):
        return float(self.a)

    def __bool__(self):
        raise TypeError("tf_shim: using a Tensor as a Python bool is not allowed (as in TensorFlow graph mode)")

    # -- indexing (strided_slice): slice bounds may be scalar Tensors
    def __getitem__(self, key):
        def fix(k):
            if isinstance(k, slice):
                return slice(*(None if v is None else int(v) for v in (k.start, k.stop, k.step)))
            if isinstance(k, Tensor):
                return int(k)
            return k
        key = tuple(fix(k) for k in key) if isinstance(key, tuple) else fix(key)
        return Tensor(self.a[key])

    # -- arithmetic
    def _bin(self, other, fn, swap=False):
        a, b = _pair(self, other)
        if swap:
            a, b = b, a
        with np.errstate(all="ignore"):
            return Tensor(fn(a, b))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)

    def _div(self, a, b):
        if a.dtype.kind != "f":      # tf.truediv on ints goes through float64
            a, b = a.astype(np.float64), b.astype(np.float64)
        return np.true_divide(a, b)

    def __truediv__(self, o): return self._bin(o, self._div)
    def __rtruediv__(self, o): return self._bin(o, self._div, True)
    def __neg__(self): return Tensor(np.negative(self.a))
    def __gt__(self, o): return self._bin(o, np.greater)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    def __lt__(self, o): return self._bin(o, np.less)
    def __le__(self, o): return self._bin(o, np.less_equal)


def _T(x, dtype=None):
    return Tensor(_conv(x, dtype))


def _ints(seq):
    """A Python list of ints out of a shape / multiples / paddings argument that may contain scalar Tensors."""
    if isinstance(seq, Tensor):
        return [int(v) for v in np.atleast_1d(seq.a)]
    if isinstance(seq, (list, tuple)):
        return [_ints(v) if isinstance(v, (list, tuple)) else int(v) for v in seq]
    return [int(v) for v in np.atleast_1d(np.asarray(seq))]


# ----------------------------------------------------------------------------------------------- ops
def constant(value, dtype=None, shape=None, name=None):
    return _T(value, dtype)


def convert_to_tensor(value, dtype=None, name=None):
    return _T(value, dtype)


def identity(x, name=None):
    return _T(x)


def stop_gradient(x, name=None):
    return _T(x)


def cast(x, dtype, name=None):
    a = _conv(x, None) if not isinstance(x, np.ndarray) else x
    dt = np.dtype(dtype)
    if dt == np.bool_:
        return Tensor(a != 0)
    if dt.kind == "i" and a.dtype.kind == "f":
        v = _x86_f2i32(a)
        return Tensor(v if dt == np.int32 else v.astype(dt))
    return Tensor(a.astype(dt))


def to_float(x, name=None):
    return cast(x, f32)


def shape(x, name=None, out_type=np.int32):
    return Tensor(np.array(np.shape(_unwrap(x)), dtype=out_type))


def reshape(x, shape, name=None):     # noqa: A002 (TensorFlow's own argument name)
    return Tensor(np.reshape(_conv(x, None), _ints(shape)))


def expand_dims(x, axis=None, name=None, dim=None):
    return Tensor(np.expand_dims(_conv(x, None), axis if axis is not None else dim))


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    ax = axis if axis is not None else squeeze_dims
    a = _conv(x, None)
    if ax is None:
        return Tensor(np.squeeze(a))
    return Tensor(np.squeeze(a, axis=tuple(ax) if isinstance(ax, (list, tuple)) else ax))


def transpose(x, perm=None, name=None):
    return Tensor(np.transpose(_conv(x, None), perm))


def stack(values, axis=0, name="stack"):
    return Tensor(np.stack([_conv(v, None) for v in values], axis=axis))


def concat(values, axis, name="concat"):
    arrs = [_conv(v, None) for v in values]
    dts = {a.dtype for a in arrs}
    if len(dts) != 1:
        raise TypeError(f"tf_shim.concat: mixed dtypes {dts}")
    return Tensor(np.concatenate(arrs, axis=axis))


def split(value, num_or_size_splits, axis=0, num=None, name="split"):
    return [Tensor(p) for p in np.split(_conv(value, None), num_or_size_splits, axis=axis)]


def tile(x, multiples, name=None):
    return Tensor(np.tile(_conv(x, None), _ints(multiples)))


def pad(x, paddings, mode="CONSTANT", name=None, constant_values=0):
    if mode.upper() != "CONSTANT":
        raise NotImplementedError(mode)
    a = _conv(x, None)
    return Tensor(np.pad(a, [tuple(p) for p in _ints(paddings)], mode="constant",
                         constant_values=np.asarray(constant_values).astype(a.dtype)))


def zeros(shape, dtype=f32, name=None):     # noqa: A002
    return Tensor(np.zeros(_ints(shape), dtype=dtype))


def range(start, limit=None, delta=1, dtype=None, name="range"):     # noqa: A001
    if limit is None:
        start, limit = 0, start
    return Tensor(np.arange(int(start), int(limit), int(delta), dtype=dtype or np.int32))


def meshgrid(*args, **kwargs):
    if kwargs.get("indexing", "xy") != "xy":
        raise NotImplementedError
    return [Tensor(m) for m in np.meshgrid(*[_conv(a, None) for a in args], indexing="xy")]


def gather(params, indices, validate_indices=None, name=None, axis=0):
    p, i = _conv(params, None), _conv(indices, None)
    if i.size and (i.min() < 0 or i.max() >= p.shape[axis]):
        raise IndexError("tf_shim.gather: index out of range (TensorFlow's CPU kernel raises InvalidArgument)")
    return Tensor(np.take(p, i, axis=axis))


def gather_nd(params, indices, name=None):
    p, i = _conv(params, None), _conv(indices, None)
    k = i.shape[-1]
    if i.size and ((i < 0).any() or (i >= np.array(p.shape[:k])).any()):
        raise IndexError("tf_shim.gather_nd: index out of range")
    return Tensor(p[tuple(i[..., j] for j in np.arange(k))])


def boolean_mask(tensor, mask, name="boolean_mask", axis=None):
    return Tensor(_conv(tensor, None)[_conv(mask, None).astype(bool)])


def where(condition, x=None, y=None, name=None):
    c = _conv(condition, None).astype(bool)
    if x is None and y is None:
        return Tensor(np.argwhere(c).astype(np.int64))       # row-major ascending coordinates
    a, b = _pair(x, y) if isinstance(x, Tensor) or isinstance(y, Tensor) else (np.asarray(x), np.asarray(y))
    return Tensor(np.where(c, a, b))


def is_nan(x, name=None):
    return Tensor(np.isnan(_conv(x, None)))


def equal(x, y, name=None):
    a, b = _pair(x, y)
    return Tensor(a == b)


def greater(x, y, name=None):
    a, b = _pair(x, y)
    return Tensor(a > b)


def add(x, y, name=None):
    a, b = _pair(x, y)
    return Tensor(a + b)


def multiply(x, y, name=None):
    a, b = _pair(x, y)
    return Tensor(a * b)


def divide(x, y, name=None):
    return _T(x) / y if not isinstance(x, Tensor) else x / y


def minimum(x, y, name=None):
    a, b = _pair(x, y)
    return Tensor(np.where(b < a, b, a))


def maximum(x, y, name=None):
    a, b = _pair(x, y)
    return Tensor(np.where(a < b, b, a))


def abs(x, name=None):     # noqa: A001
    return Tensor(np.abs(_conv(x, None)))


def sqrt(x, name=None):
    with np.errstate(all="ignore"):
        return Tensor(np.sqrt(_conv(x, None)))


def _via_f64(fn, x):
    a = _conv(x, None)
    if a.dtype != f32:
        raise TypeError(f"tf_shim: exp/log expect float32, got {a.dtype}")
    with np.errstate(all="ignore"):
        return Tensor(fn(a.astype(np.float64)).astype(f32))


def exp(x, name=None):
    return _via_f64(np.exp, x)


def log(x, name=None):
    return _via_f64(np.log, x)


def round(x, name=None):     # noqa: A001
    return Tensor(np.rint(_conv(x, None)))


def argmax(x, axis=None, name=None, dimension=None, output_type=np.int64):
    ax = axis if axis is not None else dimension
    return Tensor(np.argmax(_conv(x, None), axis=ax).astype(output_type))     # first maximal index


def reduce_max(x, axis=None, keepdims=False, name=None, reduction_indices=None):
    a = _conv(x, None)
    ax = axis if axis is not None else reduction_indices
    if a.shape[ax if ax is not None else 0] == 0:      # Eigen's max-reducer initial value
        shp = [s for i, s in enumerate(a.shape) if i != ax]
        return Tensor(np.full(shp, -np.inf if a.dtype.kind == "f" else np.iinfo(a.dtype).min, a.dtype))
    with np.errstate(all="ignore"):
        return Tensor(np.max(a, axis=ax, keepdims=keepdims))


def reduce_sum(x, axis=None, keepdims=False, name=None, reduction_indices=None):
    a = _conv(x, None)
    ax = axis if axis is not None else reduction_indices
    if a.dtype == f32 and ax is not None:       # strictly sequential fp32 accumulation along the axis
        m = np.moveaxis(a, ax, -1)
        acc = np.zeros(m.shape[:-1], f32)
        for j in np.arange(m.shape[-1]):
            acc = acc + m[..., j]
        return Tensor(np.expand_dims(acc, ax) if keepdims else acc)
    return Tensor(np.sum(a, axis=ax, keepdims=keepdims))


def unique(x, out_idx=np.int32, name=None):
    a = _conv(x, None)
    vals, first, inv = np.unique(a, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                  # first-occurrence order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    Unique = collections.namedtuple("Unique", ["y", "idx"])
    return Unique(Tensor(vals[order]), Tensor(rank[inv].astype(out_idx)))


def map_fn(fn, elems, dtype=None, parallel_iterations=None, back_prop=True, swap_memory=False, infer_shape=True,
           name=None):
    e = _conv(elems, None)
    outs = [_conv(fn(Tensor(e[i])), None) for i in np.arange(e.shape[0])]
    if not outs:
        return Tensor(np.zeros((0,), dtype=dtype if dtype is not None else e.dtype))
    out = np.stack(outs, axis=0)
    return Tensor(out.astype(dtype) if dtype is not None else out)


def random_shuffle(value, seed=None, name=None):
    a = _conv(value, None)
    if not _shuffle_perms:
        raise RuntimeError("tf_shim.random_shuffle: no permutation queued (tf_shim.set_shuffle_perms)")
    perm = _shuffle_perms.pop(0)
    order = [int(q) for q in perm if 0 <= int(q) < a.shape[0]]
    if len(order) != a.shape[0]:
        raise ValueError("tf_shim.random_shuffle: the injected permutation does not cover the value")
    return Tensor(a[order])


def Assert(condition, data, summarize=None, name=None):
    if not bool(np.all(_conv(condition, None))):
        raise AssertionError(f"tf.Assert failed: {name}")
    return None


class _NullCtx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


control_dependencies = _NullCtx
variable_scope = _NullCtx
name_scope = _NullCtx
AUTO_REUSE = object()


def placeholder(dtype, shape=None, name=None):
    raise RuntimeError("tf_shim is eager: pass the inputs to the layer constructors instead of feeding placeholders")


def global_variables_initializer():
    return None


class Session:
    """`with tf.Session() as sess: sess.run(fetches)` -> numpy values of already-computed (eager) tensors."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        if feed_dict:
            raise RuntimeError("tf_shim is eager: feed_dict is not supported")

        def ev(f):
            if f is None:
                return None
            if isinstance(f, Tensor):
                return f.a.copy()
            if isinstance(f, (list, tuple)):
                return type(f)(ev(v) for v in f) if not hasattr(f, "_fields") else type(f)(*(ev(v) for v in f))
            if isinstance(f, dict):
                return {k: ev(v) for k, v in f.items()}
            return f
        return ev(fetches)

    def close(self):
        pass


# ----------------------------------------------------------------------------------------------- library kernels
def _oracle():
    import oracle
    return oracle


_TopKV2 = collections.namedtuple("TopKV2", ["values", "indices"])


def _top_k(input, k=1, sorted=True, name=None):     # noqa: A002
    a = _conv(input, None)
    k = int(k)
    if a.dtype == f32:
        a2 = np.ascontiguousarray(a.reshape(-1, a.shape[-1]))
        val, idx = _oracle().topk(a2, k)
        return _TopKV2(Tensor(val.reshape(a.shape[:-1] + (k,))), Tensor(idx.reshape(a.shape[:-1] + (k,))))
    # integer keys (maskrcnn.py:171 sorts batch*100000+index): descending, ties -> lower index
    idx = np.argsort(-a.astype(np.int64), axis=-1, kind="stable")[..., :k]
    return _TopKV2(Tensor(np.take_along_axis(a, idx, axis=-1)), Tensor(idx.astype(np.int32)))


def _non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5, name=None):
    b, s = _conv(boxes, f32), _conv(scores, f32)
    if b.ndim != 2 or b.shape[1] != 4 or s.shape != (b.shape[0],):
        raise ValueError(f"tf_shim.non_max_suppression: boxes {b.shape} / scores {s.shape}")
    return Tensor(_oracle().nms(b, s, int(max_output_size), float(iou_threshold)).astype(np.int32))


def _crop_and_resize(image, boxes, box_ind, crop_size, method="bilinear", extrapolation_value=0, name=None):
    if method != "bilinear":
        raise NotImplementedError(method)
    img, bx, bi = _conv(image, f32), _conv(boxes, f32), _conv(box_ind, np.int32)
    ch, cw = _ints(crop_size)
    return Tensor(_oracle().crop_and_resize(img, bx, bi, ch, cw, float(extrapolation_value)))


class _Sparse:
    def __init__(self, rows):
        self.rows = rows


def _set_intersection(a, b, validate_indices=True):
    x, y = _conv(a, None), _conv(b, None)
    if x.dtype != y.dtype:
        raise TypeError(f"tf_shim.set_intersection: dtypes {x.dtype} / {y.dtype}")
    if x.ndim != 2 or y.ndim != 2 or x.shape[0] != y.shape[0]:
        raise ValueError("tf_shim.set_intersection: expects two [n, ?] operands")
    return _Sparse([np.intersect1d(x[r], y[r]).astype(x.dtype) for r in np.arange(x.shape[0])])   # ascending, unique


def sparse_tensor_to_dense(sp, default_value=0, validate_indices=True, name=None):
    w = max([r.size for r in sp.rows] + [0])
    dt = sp.rows[0].dtype if sp.rows else np.int64
    out = np.full((len(sp.rows), w), default_value, dtype=dt)
    for r, v in enumerate(sp.rows):
        out[r, :v.size] = v
    return Tensor(out)


# ----------------------------------------------------------------------------------------------- inert stubs
class _Inert:
    """Stands in for keras layers / tensors of the out-of-scope dense head: every call, attribute and index yields
    another inert object, so MaskRCNN.classifier_with_fpn_keras runs to completion without computing anything."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __getitem__(self, key):
        return _Inert()

    def call(self, *a, **k):
        return _Inert()


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert


def install():
    """Register the shim as `tensorflow` (+ inert keras / skimage / h5py) in sys.modules. Idempotent."""
    if getattr(sys.modules.get("tensorflow"), "__tf_shim__", False):
        return sys.modules["tensorflow"]
    tf = types.ModuleType("tensorflow")
    tf.__tf_shim__ = True
    g = globals()
    for name in ("constant convert_to_tensor identity stop_gradient cast to_float shape reshape expand_dims squeeze "
                 "transpose stack concat split tile pad zeros range meshgrid gather gather_nd boolean_mask where "
                 "is_nan equal greater add multiply divide minimum maximum abs sqrt exp log round argmax reduce_max "
                 "reduce_sum unique map_fn random_shuffle Assert control_dependencies variable_scope name_scope "
                 "AUTO_REUSE placeholder global_variables_initializer Session sparse_tensor_to_dense Tensor").split():
        setattr(tf, name, g[name])
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = np.float32, np.float64, np.int32, np.int64, np.bool_
    tf.nn = types.ModuleType("tensorflow.nn")
    tf.nn.top_k = _top_k
    tf.image = types.ModuleType("tensorflow.image")
    tf.image.non_max_suppression = _non_max_suppression
    tf.image.crop_and_resize = _crop_and_resize
    tf.sets = types.ModuleType("tensorflow.sets")
    tf.sets.set_intersection = _set_intersection
    sys.modules["tensorflow"] = tf
    for sub in ("nn", "image", "sets"):
        sys.modules[f"tensorflow.{sub}"] = getattr(tf, sub)
    for name in ("keras", "keras.backend", "keras.layers", "skimage", "skimage.transform", "h5py"):
        sys.modules[name] = _InertModule(name)
    sys.modules["keras"].backend = sys.modules["keras.backend"]
    sys.modules["keras"].layers = sys.modules["keras.layers"]
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    return tf

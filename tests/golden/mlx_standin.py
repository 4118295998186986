"""A torch-backed stand-in for the small part of Apple MLX that the reference's MLX sources call, so that the reference's
OWN model code (python/src/diffusionkit/mlx/mmdit.py, vae.py, sampler.py, config.py — imported from /root/reference,
never copied) can execute in this container.

MLX 0.17.3 itself cannot run here (Metal-only, no wheel).  What this file restates is the documented behaviour of the
~50 primitives those sources use (`mx.array` arithmetic with MLX's type promotion, reshape/transpose/split/repeat/
concatenate/stack, `mx.fast.layer_norm` / `rms_norm` / `scaled_dot_product_attention`, and the `mlx.nn` layers Linear,
Conv2d (NHWC, OHWI weights), Embedding, GroupNorm, RMSNorm, Sequential, GELU, SiLU, plus `nn.Module`'s dict-based
parameter tree).  Everything above the primitives — block wiring, adaLN slicing, RoPE tables, patchify / unpack,
modulation cache, the [text | image] concatenation order, the fc2-bias zeroing, the sampler formulas — is the reference's
code running unmodified.  So fixtures produced this way pin the oracle's WIRING against the reference source; they do
not pin MLX's kernel numerics (fp32 here).  Test infrastructure only.
"""
import math
import sys
import types

import numpy as np
import torch

_BACKING = "_dkb200_mlx_standin"


# ------------------------------------------------------------------------------------------------ dtypes
class Dtype:
    def __init__(self, name, tdtype, size):
        self.name, self.t, self.size = name, tdtype, size

    def __repr__(self):
        return f"mlx.core.{self.name}"


float32 = Dtype("float32", torch.float32, 4)
float16 = Dtype("float16", torch.float16, 2)
bfloat16 = Dtype("bfloat16", torch.bfloat16, 2)
int32 = Dtype("int32", torch.int32, 4)
int64 = Dtype("int64", torch.int64, 8)
int16 = Dtype("int16", torch.int16, 2)
uint32 = Dtype("uint32", torch.int64, 4)
bool_ = Dtype("bool", torch.bool, 1)
_ALL = [float32, float16, bfloat16, int32, int64, int16, bool_]


def _dt(t: torch.dtype) -> Dtype:
    for d in _ALL:
        if d.t == t:
            return d
    raise TypeError(t)


def _unwrap(x):
    return x.t if isinstance(x, array) else x


def _binary(a, b, fn):
    """MLX type promotion: two arrays promote like numpy regardless of their rank (no torch-style 0-dim weakness);
    python scalars are weak (an int/float never widens a floating array; a float turns an integer array into float32)"""
    if isinstance(a, array) and isinstance(b, array):
        ct = torch.promote_types(a.t.dtype, b.t.dtype)
        return array(fn(a.t.to(ct), b.t.to(ct)))
    arr, other, flip = (a, b, False) if isinstance(a, array) else (b, a, True)
    if isinstance(other, (np.ndarray, np.generic)) and not isinstance(other, (np.floating, np.integer)):
        other = array(other)
        return _binary(other, arr, fn) if flip else _binary(arr, other, fn)
    other = other.item() if isinstance(other, np.generic) else other
    t = arr.t
    if isinstance(other, float) and not t.dtype.is_floating_point:
        t = t.to(torch.float32)
    o = torch.tensor(other, dtype=t.dtype)
    return array(fn(o, t) if flip else fn(t, o))


class array:
    __array_priority__ = 1000

    def __init__(self, v, dtype=None):
        if isinstance(v, array):
            t = v.t
        elif isinstance(v, torch.Tensor):
            t = v
        elif isinstance(v, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(v))
            if t.dtype == torch.float64:
                t = t.to(torch.float32)            # mx.array(np.float64 array) -> float32
        elif isinstance(v, (bool, int)):
            t = torch.tensor(v, dtype=torch.int32 if not isinstance(v, bool) else torch.bool)
        elif isinstance(v, float):
            t = torch.tensor(v, dtype=torch.float32)
        elif isinstance(v, (list, tuple)):
            flat = np.asarray([_unwrap(x).numpy() if isinstance(x, array) else x for x in v]) if v else np.zeros((0,), np.float32)
            t = torch.from_numpy(flat)
            if t.dtype == torch.float64:
                t = t.to(torch.float32)
            elif t.dtype == torch.int64:
                t = t.to(torch.int32)
        elif isinstance(v, np.generic):
            t = torch.tensor(v.item(), dtype=torch.float32 if isinstance(v, np.floating) else torch.int32)
        else:
            raise TypeError(type(v))
        if dtype is not None:
            t = t.to(dtype.t)
        self.t = t

    # ---- attributes
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def dtype(self):
        return _dt(self.t.dtype)

    @property
    def size(self):
        return self.t.numel()

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def T(self):
        return array(self.t.permute(*reversed(range(self.t.dim()))))

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield array(self.t[i])

    def item(self):
        return self.t.item()

    def tolist(self):
        return self.t.tolist()

    def __array__(self, dtype=None, copy=None):
        a = self.t.float().numpy() if self.t.dtype in (torch.bfloat16, torch.float16) else self.t.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"array({self.t}, dtype={self.dtype})"

    # ---- shape ops
    def astype(self, dtype, stream=None):
        return array(self.t.to(dtype.t))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return array(self.t.reshape(*shape))

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(reversed(range(self.t.dim())))
        return array(self.t.permute(*axes))

    def swapaxes(self, a, b):
        return array(self.t.transpose(a, b))

    def flatten(self, start_axis=0, end_axis=-1):
        return array(self.t.flatten(start_axis, end_axis))

    def squeeze(self, axis=None):
        return array(self.t.squeeze() if axis is None else self.t.squeeze(axis))

    def split(self, indices_or_sections, axis=0):
        return split(self, indices_or_sections, axis)

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(_unwrap(i).long() if isinstance(i, array) else i for i in idx)
        elif isinstance(idx, array):
            idx = idx.t.long()
        return array(self.t[idx])

    def __setitem__(self, idx, val):
        if isinstance(idx, tuple):
            idx = tuple(_unwrap(i).long() if isinstance(i, array) else i for i in idx)
        v = _unwrap(val)
        self.t = self.t.clone()
        self.t[idx] = v.to(self.t.dtype) if isinstance(v, torch.Tensor) else v

    # ---- reductions
    def sum(self, axis=None, keepdims=False):
        return array(self.t.sum() if axis is None else self.t.sum(dim=axis, keepdim=keepdims))

    def mean(self, axis=None, keepdims=False):
        return array(self.t.mean() if axis is None else self.t.mean(dim=axis, keepdim=keepdims))

    def max(self, axis=None, keepdims=False):
        return array(self.t.max() if axis is None else self.t.amax(dim=axis, keepdim=keepdims))

    def min(self, axis=None, keepdims=False):
        return array(self.t.min() if axis is None else self.t.amin(dim=axis, keepdim=keepdims))

    def argmax(self, axis=None, keepdims=False):
        return array((self.t.argmax() if axis is None else self.t.argmax(dim=axis, keepdim=keepdims)).to(torch.int32))

    def square(self):
        return array(self.t * self.t)

    def abs(self):
        return array(self.t.abs())

    # ---- arithmetic
    def __add__(self, o):
        return _binary(self, o, torch.add)

    def __radd__(self, o):
        return _binary(o, self, torch.add)

    def __sub__(self, o):
        return _binary(self, o, torch.sub)

    def __rsub__(self, o):
        return _binary(o, self, torch.sub)

    def __mul__(self, o):
        return _binary(self, o, torch.mul)

    def __rmul__(self, o):
        return _binary(o, self, torch.mul)

    def __truediv__(self, o):
        return _binary(_float_if_int(self), o, torch.true_divide)

    def __rtruediv__(self, o):
        return _binary(o, _float_if_int(self), torch.true_divide)

    def __floordiv__(self, o):
        return _binary(self, o, torch.floor_divide)

    def __pow__(self, o):
        return _binary(self, o, torch.pow)

    def __rpow__(self, o):
        return _binary(o, self, torch.pow)

    def __neg__(self):
        return array(-self.t)

    def __matmul__(self, o):
        return _binary(self, o, torch.matmul)

    def __rmatmul__(self, o):
        return _binary(o, self, torch.matmul)

    def __lt__(self, o):
        return _binary(self, o, torch.lt)

    def __le__(self, o):
        return _binary(self, o, torch.le)

    def __gt__(self, o):
        return _binary(self, o, torch.gt)

    def __ge__(self, o):
        return _binary(self, o, torch.ge)

    def __eq__(self, o):
        return _binary(self, o, torch.eq)

    def __ne__(self, o):
        return _binary(self, o, torch.ne)

    __hash__ = None


def _float_if_int(a):
    return a if a.t.dtype.is_floating_point else array(a.t.to(torch.float32))


def _f(x):
    """unary float functions: integer inputs become float32, 16-bit inputs are evaluated in fp32 and rounded back"""
    x = x if isinstance(x, array) else array(x)
    t = x.t
    if not t.dtype.is_floating_point:
        t = t.to(torch.float32)
    return t


def _unary(fn):
    def g(x, stream=None):
        t = _f(x)
        return array(fn(t.float()).to(t.dtype))
    return g


exp, log, sin, cos, sqrt, rsqrt, tanh, erf, sigmoid = (_unary(f) for f in (
    torch.exp, torch.log, torch.sin, torch.cos, torch.sqrt, torch.rsqrt, torch.tanh, torch.erf, torch.sigmoid))


def abs(x):  # noqa: A001
    return array(_unwrap(x).abs())


def square(x):
    return array(_unwrap(x) * _unwrap(x))


def zeros(shape, dtype=float32):
    return array(torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=dtype.t))


def ones(shape, dtype=float32):
    return array(torch.ones(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=dtype.t))


def zeros_like(a):
    return array(torch.zeros_like(a.t))


def arange(start, stop=None, step=1, dtype=None):
    if stop is None:
        start, stop = 0, start
    is_f = any(isinstance(v, float) for v in (start, stop, step))
    t = torch.arange(start, stop, step, dtype=torch.float32 if is_f else torch.int32)
    return array(t if dtype is None else t.to(dtype.t))


def linspace(start, stop, num=50, dtype=float32):
    return array(torch.linspace(float(start), float(stop), int(num), dtype=torch.float32).to(dtype.t))


def concatenate(arrays, axis=0):
    ts = [_unwrap(a) for a in arrays]
    ct = ts[0].dtype
    for t in ts[1:]:
        ct = torch.promote_types(ct, t.dtype)
    return array(torch.cat([t.to(ct) for t in ts], dim=axis))


def stack(arrays, axis=0):
    ts = [_unwrap(a) for a in arrays]
    ct = ts[0].dtype
    for t in ts[1:]:
        ct = torch.promote_types(ct, t.dtype)
    return array(torch.stack([t.to(ct) for t in ts], dim=axis))


def split(a, indices_or_sections, axis=0):
    t = _unwrap(a)
    if isinstance(indices_or_sections, int):
        assert t.shape[axis] % indices_or_sections == 0
        return [array(x) for x in torch.chunk(t, indices_or_sections, dim=axis)]
    return [array(x) for x in torch.tensor_split(t, list(indices_or_sections), dim=axis)]


def repeat(a, repeats, axis=None):
    t = _unwrap(a)
    return array(t.flatten().repeat_interleave(repeats) if axis is None else t.repeat_interleave(repeats, dim=axis))


def expand_dims(a, axis):
    return array(_unwrap(a).unsqueeze(axis))


def broadcast_to(a, shape):
    return array(_unwrap(a).broadcast_to(tuple(shape)))


def pad(a, pad_width, constant_values=0):
    t = _unwrap(a)
    flat = []
    for lo, hi in reversed(list(pad_width)):
        flat += [lo, hi]
    return array(torch.nn.functional.pad(t, flat, value=constant_values))


def clip(a, a_min, a_max):
    return array(torch.clamp(_unwrap(a), a_min, a_max))


def maximum(a, b):
    return _binary(a if isinstance(a, array) else array(a), b, torch.maximum) if isinstance(b, array) else \
        array(torch.clamp(_unwrap(a), min=b))


def minimum(a, b):
    return _binary(a if isinstance(a, array) else array(a), b, torch.minimum) if isinstance(b, array) else \
        array(torch.clamp(_unwrap(a), max=b))


def where(c, a, b):
    a = a if isinstance(a, array) else array(a)
    b = b if isinstance(b, array) else array(b)
    ct = torch.promote_types(a.t.dtype, b.t.dtype)
    return array(torch.where(_unwrap(c).bool(), a.t.to(ct), b.t.to(ct)))


def softmax(a, axis=-1, precise=False):
    t = _unwrap(a)
    return array(torch.softmax(t.float(), dim=axis).to(t.dtype))


def einsum(eq, *ops):
    return array(torch.einsum(eq, *[_unwrap(o) for o in ops]))


def eval(*args):  # noqa: A001 — mx.eval: MLX is lazy, this stand-in is eager
    return None


class _Random:
    """mx.random: the reference seeds it (mlx/__init__.py:267) but draws its latent noise with numpy"""

    def __init__(self):
        self._g = torch.Generator().manual_seed(0)

    def seed(self, s):
        self._g.manual_seed(int(s))

    def normal(self, shape=(), dtype=float32, loc=0.0, scale=1.0, key=None):
        return array((torch.randn(tuple(shape), generator=self._g) * scale + loc).to(dtype.t))


random = _Random()


# ------------------------------------------------------------------------------------------------ mx.fast
def fast_layer_norm(x, weight, bias, eps, stream=None):
    t = _unwrap(x)
    f = t.float()
    mu = f.mean(-1, keepdim=True)
    var = (f - mu).pow(2).mean(-1, keepdim=True)
    y = (f - mu) * torch.rsqrt(var + eps)
    if weight is not None:
        y = y * _unwrap(weight).float()
    if bias is not None:
        y = y + _unwrap(bias).float()
    return array(y.to(t.dtype))


def fast_rms_norm(x, weight, eps, stream=None):
    t = _unwrap(x)
    f = t.float()
    y = f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + eps) * _unwrap(weight).float()
    return array(y.to(t.dtype))


def fast_sdpa(q, k, v, *, scale, mask=None, memory_efficient_threshold=None, stream=None):
    qt, kt, vt = _unwrap(q), _unwrap(k), _unwrap(v)
    s = (qt.float() * scale) @ kt.float().transpose(-1, -2)
    if mask is not None:
        s = s + _unwrap(mask).float()
    return array((torch.softmax(s, dim=-1) @ vt.float()).to(qt.dtype))


# ------------------------------------------------------------------------------------------------ mlx.utils
def tree_map(fn, tree, *rest):
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(fn, c, *(r[i] for r in rest)) for i, c in enumerate(tree))
    if isinstance(tree, dict):
        return {k: tree_map(fn, c, *(r[k] for r in rest)) for k, c in tree.items()}
    return fn(tree, *rest)


def tree_flatten(tree, prefix=""):
    out = []
    if isinstance(tree, (list, tuple)):
        for i, c in enumerate(tree):
            out += tree_flatten(c, f"{prefix}.{i}" if prefix else str(i))
    elif isinstance(tree, dict):
        for k, c in tree.items():
            out += tree_flatten(c, f"{prefix}.{k}" if prefix else k)
    else:
        out.append((prefix, tree))
    return out


def tree_unflatten(items):
    root = {}
    for key, val in items:
        parts = key.split(".")
        node = root
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = val

    def fix(n):
        if not isinstance(n, dict):
            return n
        if n and all(k.isdigit() for k in n):           # a list; indices without parameters become {} (as in MLX)
            size = max(int(k) for k in n) + 1
            return [fix(n[str(i)]) if str(i) in n else {} for i in range(size)]
        return {k: fix(v) for k, v in n.items()}

    return fix(root)


# ------------------------------------------------------------------------------------------------ mlx.nn
class Module(dict):
    """mlx.nn.Module: a dict whose items are the parameters / sub-modules; other attributes are plain attributes"""

    def __init__(self):
        dict.__init__(self)

    def __getattr__(self, key):
        if key in self:
            return self[key]
        raise AttributeError(f"{type(self).__name__} has no attribute {key}")

    def __setattr__(self, key, val):
        if isinstance(val, (array, dict, list, tuple)):
            if key in self.__dict__:                     # previously a plain attribute: the dict entry replaces it
                object.__delattr__(self, key)
            self[key] = val
        else:
            self.pop(key, None)
            object.__setattr__(self, key, val)

    def __delattr__(self, key):
        if key in self:
            del self[key]
        else:
            object.__delattr__(self, key)

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def _children(self):
        for k, v in self.items():
            yield k, v

    def parameters(self):
        """arrays reachable through dicts / lists / sub-modules, skipping keys that start with "_" (mlx
        valid_parameter_filter); list entries that are not containers or arrays become {} like in MLX"""
        def keep(k, v):
            return isinstance(v, (array, dict, list)) and not str(k).startswith("_")

        def rec(v):
            if isinstance(v, array):
                return v
            if isinstance(v, dict):                      # includes Module
                return {k: rec(c) for k, c in v.items() if keep(k, c)}
            return [rec(c) if keep("", c) else {} for c in v]
        return rec(self)

    def update(self, params):
        """mlx Module.update: only keys that already exist in the module tree are replaced — anything else in `params`
        is silently ignored (which is how the reference ends up without a k_proj bias for FLUX, model_io.py:776)"""
        def apply(dst, src):
            if isinstance(src, dict):
                for k, v in src.items():
                    if k not in dst:
                        continue
                    if isinstance(dst[k], (dict, list)) and isinstance(v, (dict, list)):
                        apply(dst[k], v)
                    elif isinstance(dst[k], array):
                        dst[k] = v
            elif isinstance(src, list):
                for i, v in enumerate(src):
                    if i >= len(dst):
                        continue
                    if isinstance(dst[i], (dict, list)) and isinstance(v, (dict, list)):
                        apply(dst[i], v)
                    elif isinstance(dst[i], array):
                        dst[i] = v
        apply(self, params)
        return self

    def load_weights(self, weights, strict=True):
        if strict:
            have = {k for k, _ in tree_flatten(self.parameters())}
            got = {k for k, _ in weights}
            if have != got:
                raise ValueError(f"load_weights: missing {sorted(have - got)[:5]} unexpected {sorted(got - have)[:5]}")
        self.update(tree_unflatten(list(weights)))
        return self

    def named_modules(self):
        out = [("", self)]

        def rec(prefix, v):
            if isinstance(v, Module):
                out.append((prefix, v))
                for k, c in v.items():
                    rec(f"{prefix}.{k}", c)
            elif isinstance(v, dict):
                for k, c in v.items():
                    rec(f"{prefix}.{k}", c)
            elif isinstance(v, (list, tuple)):
                for i, c in enumerate(v):
                    rec(f"{prefix}.{i}", c)
        for k, v in self.items():
            rec(k, v)
        return out

    def modules(self):
        return [m for _, m in self.named_modules()]


class Identity(Module):
    def __call__(self, x, *a, **k):
        return x


class Linear(Module):
    def __init__(self, input_dims, output_dims, bias=True):
        super().__init__()
        s = math.sqrt(1.0 / input_dims)
        self.weight = array(torch.empty(output_dims, input_dims).uniform_(-s, s))
        if bias:
            self.bias = array(torch.empty(output_dims).uniform_(-s, s))

    def __call__(self, x):
        y = x @ self["weight"].T
        return y + self["bias"] if "bias" in self else y


class Conv2d(Module):
    """NHWC input, weight (out, kh, kw, in)"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.weight = array(torch.randn(out_channels, k[0], k[1], in_channels) * 0.02)
        if bias:
            self.bias = array(torch.zeros(out_channels))
        self.stride = (stride, stride) if isinstance(stride, int) else tuple(stride)
        self.padding = (padding, padding) if isinstance(padding, int) else tuple(padding)

    def __call__(self, x):
        t, w = _unwrap(x), self["weight"].t
        ct = torch.promote_types(t.dtype, w.dtype)
        y = torch.nn.functional.conv2d(t.to(ct).permute(0, 3, 1, 2), w.to(ct).permute(0, 3, 1, 2),
                                       self["bias"].t.to(ct) if "bias" in self else None, stride=self.stride,
                                       padding=self.padding)
        return array(y.permute(0, 2, 3, 1))


class Embedding(Module):
    def __init__(self, num_embeddings, dims):
        super().__init__()
        self.weight = array(torch.randn(num_embeddings, dims) * math.sqrt(1.0 / dims))

    def __call__(self, x):
        return self["weight"][x]


class Sequential(Module):
    def __init__(self, *modules):
        super().__init__()
        self.layers = list(modules)

    def __call__(self, x):
        for m in self["layers"]:
            x = m(x)
        return x


class SiLU(Module):
    def __call__(self, x):
        return silu(x)


class GELU(Module):
    def __init__(self, approx="none"):
        super().__init__()
        object.__setattr__(self, "_approx", approx)

    def __call__(self, x):
        if self._approx == "none":
            return gelu(x)
        if self._approx == "fast":
            return gelu_fast_approx(x)
        return gelu_approx(x)


def silu(x):
    t = _unwrap(x)
    return array((t.float() * torch.sigmoid(t.float())).to(t.dtype))


def gelu(x):
    t = _unwrap(x)
    return array(torch.nn.functional.gelu(t.float()).to(t.dtype))


def gelu_approx(x):
    t = _unwrap(x)
    return array(torch.nn.functional.gelu(t.float(), approximate="tanh").to(t.dtype))


def gelu_fast_approx(x):
    t = _unwrap(x)
    return array((t.float() * torch.sigmoid(1.702 * t.float())).to(t.dtype))


def relu(x):
    return array(torch.relu(_unwrap(x)))


class MultiHeadAttention(Module):
    """mlx.nn.MultiHeadAttention: softmax((q * d^-1/2) k^T + mask) v; the four projections are Linears without bias
    unless the caller adds one (the reference's CLIP does, mlx/clip.py:36-41)"""

    def __init__(self, dims, num_heads, query_input_dims=None, key_input_dims=None, value_input_dims=None,
                 value_dims=None, value_output_dims=None, bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.query_proj = Linear(query_input_dims or dims, dims, bias=bias)
        self.key_proj = Linear(key_input_dims or dims, dims, bias=bias)
        self.value_proj = Linear(value_input_dims or dims, value_dims or dims, bias=bias)
        self.out_proj = Linear(value_dims or dims, value_output_dims or dims, bias=bias)

    def __call__(self, queries, keys, values, mask=None):
        q, k, v = self["query_proj"](queries), self["key_proj"](keys), self["value_proj"](values)
        H = self.num_heads
        B, L, _ = q.shape
        S = k.shape[1]
        q = q.reshape(B, L, H, -1).transpose(0, 2, 1, 3)
        k = k.reshape(B, S, H, -1).transpose(0, 2, 1, 3)
        v = v.reshape(B, S, H, -1).transpose(0, 2, 1, 3)
        scale = math.sqrt(1 / q.shape[-1])
        out = fast_sdpa(q, k, v, scale=scale, mask=mask)
        return self["out_proj"](out.transpose(0, 2, 1, 3).reshape(B, L, -1))


class RMSNorm(Module):
    def __init__(self, dims, eps=1e-5):
        super().__init__()
        self.weight = array(torch.ones(dims))
        self.eps = eps

    def __call__(self, x):
        return fast_rms_norm(x, self["weight"], self.eps)


class LayerNorm(Module):
    def __init__(self, dims, eps=1e-5, affine=True, bias=True):
        super().__init__()
        self.eps = eps
        if affine:
            self.weight = array(torch.ones(dims))
            if bias:
                self.bias = array(torch.zeros(dims))

    def __call__(self, x):
        return fast_layer_norm(x, self["weight"] if "weight" in self else None, self["bias"] if "bias" in self else None,
                               self.eps)


class GroupNorm(Module):
    """pytorch_compatible=True grouping on NHWC input"""

    def __init__(self, num_groups, dims, eps=1e-5, affine=True, pytorch_compatible=False):
        super().__init__()
        assert pytorch_compatible, "the reference only uses pytorch_compatible=True"
        self.num_groups, self.eps = num_groups, eps
        if affine:
            self.weight = array(torch.ones(dims))
            self.bias = array(torch.zeros(dims))

    def __call__(self, x):
        t = _unwrap(x)
        y = torch.nn.functional.group_norm(t.float().permute(0, 3, 1, 2), self.num_groups,
                                           self["weight"].t.float() if "weight" in self else None,
                                           self["bias"].t.float() if "bias" in self else None, self.eps)
        return array(y.permute(0, 2, 3, 1).to(t.dtype))


# ------------------------------------------------------------------------------------------------ registration
def install():
    """register `mlx`, `mlx.core`, `mlx.core.fast`, `mlx.nn`, `mlx.utils` (idempotent; refuses to shadow a real MLX)"""
    if "mlx" in sys.modules:
        if getattr(sys.modules["mlx"], _BACKING, False):
            return
        raise RuntimeError("a real mlx is importable here: use it instead of the stand-in")
    me = sys.modules[__name__]
    mlx = types.ModuleType("mlx")
    setattr(mlx, _BACKING, True)
    core = types.ModuleType("mlx.core")
    for name in ("array", "float32", "float16", "bfloat16", "int32", "int64", "int16", "uint32", "bool_", "exp", "log",
                 "sin", "cos", "sqrt", "rsqrt", "tanh", "erf", "sigmoid", "abs", "square", "zeros", "ones", "zeros_like",
                 "arange", "linspace", "concatenate", "stack", "split", "repeat", "expand_dims", "broadcast_to", "pad", "clip",
                 "maximum", "minimum", "where", "softmax", "einsum", "eval", "random"):
        setattr(core, name, getattr(me, name))
    core.Dtype = Dtype
    fast = types.ModuleType("mlx.core.fast")
    fast.layer_norm, fast.rms_norm, fast.scaled_dot_product_attention = fast_layer_norm, fast_rms_norm, fast_sdpa
    core.fast = fast
    metal = types.ModuleType("mlx.core.metal")          # memory-limit knobs the T5 encoder touches (mlx/t5.py:231-242)
    metal.set_memory_limit = lambda *a, **k: None
    metal.device_info = lambda: {"memory_size": 0}
    metal.set_cache_limit = lambda *a, **k: None
    core.metal = metal
    nn = types.ModuleType("mlx.nn")
    for name in ("Module", "Identity", "Linear", "Conv2d", "Embedding", "Sequential", "SiLU", "GELU", "RMSNorm",
                 "LayerNorm", "GroupNorm", "MultiHeadAttention", "silu", "gelu", "gelu_approx", "gelu_fast_approx",
                 "relu"):
        setattr(nn, name, getattr(me, name))
    utils = types.ModuleType("mlx.utils")
    utils.tree_map, utils.tree_flatten, utils.tree_unflatten = tree_map, tree_flatten, tree_unflatten
    mlx.core, mlx.nn, mlx.utils = core, nn, utils
    mods = {"mlx": mlx, "mlx.core": core, "mlx.core.fast": fast, "mlx.nn": nn, "mlx.utils": utils}
    import importlib.machinery

    for name, m in mods.items():       # a spec, so that importlib.util.find_spec("mlx") (transformers probes it) works
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=name == "mlx")
    mlx.__path__ = []
    sys.modules.update(mods)


def uninstall():
    if "mlx" in sys.modules and getattr(sys.modules["mlx"], _BACKING, False):
        for name in ("mlx", "mlx.core", "mlx.core.fast", "mlx.nn", "mlx.utils"):
            sys.modules.pop(name, None)

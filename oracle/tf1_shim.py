"""A stand-in ``tensorflow`` module for the ~15 TF 1.x symbols that the reference's
``compute_genetovec`` uses (/root/reference/G2Vec.py:231-257,264-283) -- TEST INFRASTRUCTURE ONLY.

TensorFlow 1.x (manual p.3: "Tensorflow (>=1.4)", unpinned) is not installed and cannot be (no
network).  With this module registered as ``sys.modules['tensorflow']`` the UNMODIFIED reference
function runs: its own shuffle/split (:219-226), graph construction (:231-251), session loop, eval order,
strict-drop early stop and snapshot (:255-286) are executed as written; only the TF ops underneath are
re-stated here, each from TF 1.x's published definition, in float32 on torch-CPU:

* ``tf.matmul``            -> torch.matmul; gradients by torch autograd (dA = dC.B^T, dB = A^T.dC, as
                              tensorflow/python/ops/math_grad.py:_MatMulGrad).
* ``tf.nn.sigmoid_cross_entropy_with_logits`` -> the composition of nn_impl.py:
                              where(x>=0, x, 0) - x*z + log1p(exp(where(x>=0, -x, x))), differentiated op by op.
* ``tf.reduce_mean``       -> mean over all elements.
* ``tf.train.AdamOptimizer(lr).minimize`` -> compute_gradients over every tf.Variable, then per variable
                              training_ops.cc ApplyAdam:  alpha = lr*sqrt(1-beta2_power)/(1-beta1_power);
                              m += (g-m)*(1-beta1); v += (g*g-v)*(1-beta2); var -= (m*alpha)/(sqrt(v)+eps);
                              then adam.py:_finish: beta1_power *= beta1, beta2_power *= beta2 (float32
                              variables initialised to beta1 / beta2).  Defaults 0.9 / 0.999 / 1e-8.
* ``tf.truncated_normal``  -> N(0, stddev) re-drawn while |x| > 2 stddev (random_ops.py), from the seeded
                              generator set with ``seed_initialisers`` -- the reference is unseeded, so the
                              oracle and the GPU path are given the very same initial tensors.
* ``tf.sigmoid`` / ``>`` / ``tf.cast`` / ``tf.equal`` -> elementwise, ``(sigmoid(O) > 0.5) == Y`` as written.

``trace`` records every Session.run / Tensor.eval in call order, as (kind, scalar value or None, perf_counter
time), so the golden generator can store the per-step accuracies the reference computes but prints only every
5th step, and bench.py can time the training loop of the unmodified function.
"""
import contextlib
import math
import sys
import time
import types

import numpy as np
import torch

float32 = torch.float32

_state = {"rng": None, "session": None, "variables": [], "inits": [], "trace": []}


def seed_initialisers(seed):
    """Seed for tf.truncated_normal: PCG64(seed) standard normals, first call first (the same draw
    order as g2vec_b200.cbow.init_weights: W_ih [V, D], then W_ho)."""
    _state["rng"] = np.random.Generator(np.random.PCG64(seed))


def reset():
    _state.update(session=None, variables=[], inits=[], trace=[])


def trace():
    return _state["trace"]


def initial_values():
    """The tensors drawn by tf.truncated_normal, in call order (float32 numpy arrays)."""
    return list(_state["inits"])


# --------------------------------------------------------------------------------- graph nodes
class Tensor:
    """A lazily evaluated graph node: fn(feed) -> torch tensor."""

    def __init__(self, fn, name=None):
        self._fn, self.name = fn, name

    def _value(self, feed):
        return self._fn(feed)

    def eval(self, feed_dict=None, session=None):
        sess = session or _state["session"]
        if sess is None:
            raise RuntimeError("Tensor.eval() needs a default session")
        return sess.run(self, feed_dict)

    def __gt__(self, other):
        return Tensor(lambda f: self._value(f) > other)

    def __neg__(self):
        return Tensor(lambda f: -self._value(f))


class Placeholder(Tensor):
    def __init__(self, dtype, shape, name):
        self.dtype, self.shape = dtype, shape
        Tensor.__init__(self, self._lookup, name)

    def _lookup(self, feed):
        if feed is None or self not in feed:
            raise ValueError("placeholder %r was not fed" % self.name)
        v = feed[self]
        if isinstance(v, torch.Tensor):
            return v.to(self.dtype)
        return torch.from_numpy(np.ascontiguousarray(v)).to(self.dtype)      # int32 0/1 rows are fed as float32


class Variable(Tensor):
    def __init__(self, initial_value, name=None):
        self.initial = np.array(initial_value, dtype=np.float32, copy=True)
        self.data = None                      # set by global_variables_initializer().run()
        Tensor.__init__(self, lambda f: self.data, name)
        _state["variables"].append(self)


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name)


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    if _state["rng"] is None:
        raise RuntimeError("call tf1_shim.seed_initialisers(seed) first (the reference is unseeded)")
    rng = _state["rng"]
    x = rng.standard_normal(size=tuple(shape))
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(x) > 2.0
    out = (x * stddev + mean).astype(np.float32)
    _state["inits"].append(out)
    return out


@contextlib.contextmanager
def name_scope(name):
    yield name


def matmul(a, b):
    return Tensor(lambda f: torch.matmul(a._value(f), b._value(f)))


def reduce_mean(x):
    return Tensor(lambda f: x._value(f).mean())


def sigmoid(x):
    return Tensor(lambda f: torch.sigmoid(x._value(f)))


def cast(x, dtype):
    return Tensor(lambda f: x._value(f).to(dtype))


def equal(a, b):
    return Tensor(lambda f: a._value(f) == b._value(f))


def _sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
    def fn(f):
        x, z = logits._value(f), labels._value(f)
        zeros = torch.zeros_like(x)
        cond = x >= zeros
        relu_logits = torch.where(cond, x, zeros)
        neg_abs_logits = torch.where(cond, -x, x)
        return (relu_logits - x * z) + torch.log1p(torch.exp(neg_abs_logits))
    return Tensor(fn)


nn = types.SimpleNamespace(sigmoid_cross_entropy_with_logits=_sigmoid_cross_entropy_with_logits)


# ------------------------------------------------------------------------------------ optimizer
class _TrainOp:
    def __init__(self, opt, loss):
        self.opt, self.loss = opt, loss

    def _run(self, feed):
        opt = self.opt
        vs = list(_state["variables"])
        leaves = []
        for v in vs:
            v.data = v.data.detach().requires_grad_(True)
            leaves.append(v.data)
        cost = self.loss._value(feed)
        grads = torch.autograd.grad(cost, leaves)
        f32 = np.float32
        b1, b2, eps, lr = f32(opt.beta1), f32(opt.beta2), f32(opt.epsilon), f32(opt.lr)
        alpha = float(f32(lr * f32(np.sqrt(f32(1) - opt.beta2_power)) / (f32(1) - opt.beta1_power)))
        with torch.no_grad():
            for v, g in zip(vs, grads):
                m, s = opt.slots.setdefault(id(v), (torch.zeros_like(g), torch.zeros_like(g)))
                m.add_((g - m) * float(f32(1) - b1))
                s.add_((g * g - s) * float(f32(1) - b2))
                v.data = (v.data.detach() - (m * alpha) / (s.sqrt() + float(eps))).contiguous()
        opt.beta1_power = f32(opt.beta1_power * b1)          # adam.py:_finish
        opt.beta2_power = f32(opt.beta2_power * b2)
        return None


class _AdamOptimizer:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False, name="Adam"):
        self.lr, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon
        self.beta1_power, self.beta2_power = np.float32(beta1), np.float32(beta2)
        self.slots = {}

    def minimize(self, loss):
        return _TrainOp(self, loss)


train = types.SimpleNamespace(AdamOptimizer=_AdamOptimizer)


# -------------------------------------------------------------------------------------- session
class ConfigProto:
    def __init__(self, **kw):
        self.kw = kw


class _InitOp:
    def run(self, feed_dict=None, session=None):
        for v in _state["variables"]:
            v.data = torch.from_numpy(v.initial.copy())


def global_variables_initializer():
    return _InitOp()


class Session:
    def __init__(self, target="", graph=None, config=None):
        self.config = config

    def __enter__(self):
        _state["session"] = self
        return self

    def __exit__(self, *exc):
        _state["session"] = None
        return False

    def run(self, fetches, feed_dict=None):
        if isinstance(fetches, _TrainOp):
            out = fetches._run(feed_dict)
            _state["trace"].append(("train", None, time.perf_counter()))
            return out
        if isinstance(fetches, _InitOp):
            return fetches.run()
        with torch.no_grad():
            val = fetches._value(feed_dict)
        out = val.detach().cpu().numpy()
        if out.ndim == 0:
            out = out[()]                                     # numpy scalar, as TF returns
        else:
            out = out.copy()
        _state["trace"].append(("var" if isinstance(fetches, Variable) else "eval", out if out.ndim == 0 else None,
                                time.perf_counter()))
        return out


def install():
    """Register this module as ``tensorflow`` (the reference does ``import tensorflow as tf``, G2Vec.py:3)."""
    sys.modules["tensorflow"] = sys.modules[__name__]
    return sys.modules[__name__]


# sqrt of the hidden size is taken by the reference with math.sqrt (G2Vec.py:6); nothing else is needed.
_ = math

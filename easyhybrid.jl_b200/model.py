"""Host-side mirror of EasyHybrid's model construction API.

Mirrors (names, argument meaning, error behaviour):
  constructHybridModel / SingleNNHybridModel / MultiNNHybridModel
      src/models/GenericHybridModel.jl:44-232
  ParameterContainer / build_parameters / default / lower / upper
      src/models/helpers_for_HybridModel.jl:39-102, GenericHybridModel.jl:330-345
  scale_single_param / inv_sigmoid / scale_single_param_minmax / hard_sigmoid
      src/models/GenericHybridModel.jl:9-18, 348-365
  initialparameters / initialstates (flat ComponentArray order)
      src/models/GenericHybridModel.jl:236-327, src/training/initialization.jl:17-51

This module holds no numerics of the training step: it only builds the descriptor the
CUDA library consumes.  The process-model function is traced once with a symbolic
number type (``Sym``) into a straight-line program (the plug-in boundary,
GenericHybridModel.jl:425) and matched against the built-in fused forms.
"""
from __future__ import annotations

import inspect
import os
import itertools
import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi

# ----------------------------------------------------------------------------------------
# small numeric helpers that are part of the reference's exported API
# ----------------------------------------------------------------------------------------


def sigmoid(x):
    x = np.asarray(x, dtype=np.float32)
    return (1.0 / (1.0 + np.exp(-x))).astype(np.float32)


def hard_sigmoid(x):
    """GenericHybridModel.jl:9-11."""
    return np.clip(0.2 * np.asarray(x, dtype=np.float64) + 0.5, 0.0, 1.0)


def inv_hard_sigmoid(y):
    """GenericHybridModel.jl:16-18."""
    return (np.asarray(y, dtype=np.float64) - 0.5) / 0.2


def inv_sigmoid(y):
    """GenericHybridModel.jl:354."""
    y = np.asarray(y, dtype=np.float32)
    return np.log(y / (np.float32(1) - y)).astype(np.float32)


class ParameterContainer:
    """(default, lower, upper) table, helpers_for_HybridModel.jl:95-102."""

    def __init__(self, values):
        if isinstance(values, ParameterContainer):
            values = values.values
        self.values = {str(k): tuple(np.float32(v) for v in t) for k, t in dict(values).items()}
        for k, t in self.values.items():
            if len(t) != 3:
                raise ValueError(f"parameter {k}: expected (default, lower, upper)")
        self.names = list(self.values.keys())
        self.table = np.array([self.values[k] for k in self.names], dtype=np.float32).reshape(len(self.names), 3)

    def column(self, j):
        return {k: self.table[i, j] for i, k in enumerate(self.names)}


def build_parameters(parameters, mechanistic_model=None):
    return parameters if isinstance(parameters, ParameterContainer) else ParameterContainer(parameters)


def default(p):
    return _container(p).column(0)


def lower(p):
    return _container(p).column(1)


def upper(p):
    return _container(p).column(2)


def _container(p):
    if isinstance(p, ParameterContainer):
        return p
    if hasattr(p, "parameters"):
        return p.parameters
    return ParameterContainer(p)


def scale_single_param(name, raw_val, hm):
    """ℓ + (u-ℓ)·sigmoid(raw), GenericHybridModel.jl:348-352."""
    lo, up = lower(hm)[name], upper(hm)[name]
    return (lo + (up - lo) * sigmoid(raw_val)).astype(np.float32)


def scale_single_param_minmax(name, hm):
    """inv_sigmoid((default-ℓ)/(u-ℓ)), GenericHybridModel.jl:361-365."""
    lo, up, de = lower(hm)[name], upper(hm)[name], default(hm)[name]
    return inv_sigmoid((de - lo) / (up - lo))


# ----------------------------------------------------------------------------------------
# symbolic tracing of the process model (plug-in boundary)
# ----------------------------------------------------------------------------------------


class Sym:
    """Symbolic float: arithmetic on it records a straight-line program."""

    __array_priority__ = 1000

    def __init__(self, tape, vid):
        self.tape, self.vid = tape, vid

    # binary
    def _bin(self, op, other, swap=False):
        o = self.tape.lift(other)
        a, b = (o, self) if swap else (self, o)
        return self.tape.emit(op, a.vid, b.vid)

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __pow__(self, o): return self._bin("pow", o)
    def __rpow__(self, o): return self._bin("pow", o, True)
    def __neg__(self): return self.tape.emit("neg", self.vid, 0)
    def __pos__(self): return self
    def __abs__(self): return self.tape.emit("abs", self.vid, 0)

    _UFUNCS = {
        "add": "add", "subtract": "sub", "multiply": "mul", "true_divide": "div", "divide": "div",
        "power": "pow", "minimum": "min", "maximum": "max", "negative": "neg", "exp": "exp", "log": "log",
        "sqrt": "sqrt", "tanh": "tanh", "absolute": "abs", "sin": "sin", "cos": "cos",
    }

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        op = self._UFUNCS.get(ufunc.__name__)
        if method != "__call__" or op is None or kwargs:
            return NotImplemented
        vals = [self.tape.lift(x) for x in inputs]
        return self.tape.emit(op, vals[0].vid, vals[1].vid if len(vals) > 1 else 0)


class Tape:
    def __init__(self):
        self.prog = []  # (opname, a, b, imm)

    def emit(self, op, a, b, imm=0.0):
        self.prog.append((op, int(a), int(b), float(imm)))
        return Sym(self, len(self.prog) - 1)

    def lift(self, x):
        if isinstance(x, Sym):
            return x
        if isinstance(x, (int, float, np.floating, np.integer)) or (isinstance(x, np.ndarray) and x.ndim == 0):
            return self.emit("const", 0, 0, float(np.float32(x)))
        raise TypeError(f"cannot trace value of type {type(x).__name__} inside a process model")


def sym_exp(x):
    return np.exp(x)


def trace_process_model(fn, forcing, param_names, targets):
    """Trace ``fn(**forcings, **params)`` -> (program, output value ids).

    Keyword arguments of ``fn`` that are neither forcings nor parameters keep their
    defaults (e.g. ``tref = 15.0`` of RbQ10, README.md:148)."""
    tape = Tape()
    kwargs = {}
    sig = inspect.signature(fn)
    for name, p in sig.parameters.items():
        if name in forcing:
            kwargs[name] = tape.emit("forcing", forcing.index(name), 0)
        elif name in param_names:
            kwargs[name] = tape.emit("param", param_names.index(name), 0)
        elif p.default is inspect.Parameter.empty and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
            raise ValueError(f"process model argument `{name}` is neither a forcing nor a parameter")
    out = fn(**kwargs)
    if hasattr(out, "_asdict"):
        out = out._asdict()
    if not isinstance(out, dict):
        raise TypeError("process model must return a dict / namedtuple containing the targets")
    outs = []
    for t in targets:
        if t not in out:
            raise KeyError(f"process model output has no target `{t}`")
        outs.append(tape.lift(out[t]).vid)
    return tape.prog, outs


def _canon(prog, vid, names):
    """canonical string of the expression rooted at vid (commutative ops sorted)."""
    op, a, b, imm = prog[vid]
    if op == "const":
        return f"c{np.float32(imm)!r}"
    if op == "forcing":
        return names["f"][a]
    if op == "param":
        return names["p"][a]
    if op in ("neg", "exp", "log", "sqrt", "tanh", "sigmoid", "abs", "sin", "cos"):
        return f"{op}({_canon(prog, a, names)})"
    sa, sb = _canon(prog, a, names), _canon(prog, b, names)
    if op in ("add", "mul", "min", "max") and sb < sa:
        sa, sb = sb, sa
    return f"{op}({sa},{sb})"


# built-in fused forms: canonical functions over (p0, p1, f0[, const0])
def _rbq10(p0, p1, f0, c0):
    return {"y0": p0 * p1 ** (0.1 * (f0 - c0))}


def _expo(p0, p1, f0, c0):
    return {"y0": p0 * np.exp(p1 * f0)}


def _linear(p0, p1, f0, c0):
    return {"y0": p0 * f0 + p1}


def _linear2(p0, p1, f0, c0):
    return {"y0": p0 * f0 + p1, "y1": 2.0 * p0 * f0 + p1}


def _expo2(p0, p1, f0, c0):
    y0 = p0 * np.exp(p1 * f0)
    return {"y0": y0, "y1": 2.0 * y0}


_BUILTINS = [("RBQ10", _rbq10, 1, True), ("EXPO", _expo, 1, False), ("LINEAR", _linear, 1, False),
             ("LINEAR2", _linear2, 2, False), ("EXPO2", _expo2, 2, False)]


def match_builtin(prog, outs, n_forc, n_params):
    """Find a built-in form and binding whose expressions equal the traced program.

    Returns (pm_name, (param_i, param_j, forcing_k), consts) or None."""
    consts = sorted({np.float32(imm) for (op, _, _, imm) in prog if op == "const"})
    names = {"f": [f"F{i}" for i in range(n_forc)], "p": [f"P{i}" for i in range(n_params)]}
    want = [_canon(prog, o, names) for o in outs]
    for pm_name, fn, nt, uses_const in _BUILTINS:
        if nt != len(outs):
            continue
        for pi, pj in itertools.permutations(range(n_params), 2):
            for fk in range(n_forc):
                for c0 in (consts if uses_const else [np.float32(0)]):
                    tape = Tape()
                    p0 = tape.emit("param", pi, 0)
                    p1 = tape.emit("param", pj, 0)
                    f0 = tape.emit("forcing", fk, 0)
                    out = fn(p0, p1, f0, float(c0))
                    got = [_canon(tape.prog, tape.lift(out[f"y{t}"]).vid, names) for t in range(nt)]
                    if got == want:
                        return pm_name, (pi, pj, fk), [float(c0)]
    return None


# ----------------------------------------------------------------------------------------
# ready-made process models (the reference's examples)
# ----------------------------------------------------------------------------------------


def RbQ10(*, ta, Q10, rb, tref=15.0):
    """README.md:148-151; test/test_split_data_train.jl:36-39."""
    reco = rb * Q10 ** (0.1 * (ta - tref))
    return {"reco": reco, "Q10": Q10, "rb": rb}


def Expo_resp_model(*, T, Resp0, k):
    """projects/ExpoHybrid/ExpoHybridEstim.jl:69-85."""
    Resp_obs = Resp0 * np.exp(k * T)
    return {"Resp_obs": Resp_obs, "Resp0": Resp0, "k": k}


def Expo_resp_model2(*, T, Resp0, k):
    """Two-target form of the Expo model (BASELINE config 5): the second target is twice the first."""
    Resp_obs = Resp0 * np.exp(k * T)
    return {"Resp_obs": Resp_obs, "Resp_obs2": 2.0 * Resp_obs, "Resp0": Resp0, "k": k}


def LinearModel(*, x1, a, b):
    """test/test_generic_hybrid_model.jl:10-12 (ŷ = a·x1 + b), src/models/LinearHM.jl:61-68."""
    return {"obs": a * x1 + b}


def LinearModel2(*, x1, a, b):
    """test/test_compute_loss.jl:209-211."""
    return {"var1": a * x1 + b, "var2": 2.0 * a * x1 + b}


# ----------------------------------------------------------------------------------------
# hybrid model structs
# ----------------------------------------------------------------------------------------

_ACT_NAMES = {"tanh": "tanh", "sigmoid": "sigmoid", "relu": "relu", "swish": "swish", "identity": "identity"}


def _act_name(activation):
    if callable(activation):
        activation = getattr(activation, "__name__", str(activation))
    name = str(activation).lower()
    if name in ("σ", "sigmoid_fast"):
        name = "sigmoid"
    if name == "tanh_fast":
        name = "tanh"
    if name not in _ACT_NAMES:
        raise ValueError(f"unsupported activation {activation!r}; supported: {sorted(_ACT_NAMES)}")
    return name


@dataclass
class _HybridModelBase:
    predictors: object
    forcing: list
    targets: list
    mechanistic_model: object
    parameters: ParameterContainer
    neural_param_names: list
    global_param_names: list
    fixed_param_names: list
    scale_nn_outputs: bool
    start_from_default: bool
    config: dict
    chains: list = field(default_factory=list)  # [{name, predictors, hidden, activation, n_out, input_batchnorm}]

    # ---- flat parameter layout (ComponentArray order) ----
    def layer_shapes(self):
        shapes = []
        for ch in self.chains:
            w = [len(ch["predictors"])] + list(ch["hidden"]) + [ch["n_out"]]
            shapes.append([(w[i + 1], w[i]) for i in range(len(w) - 1)])
        return shapes

    def num_params(self):
        n = sum(o * i + o for ch in self.layer_shapes() for (o, i) in ch)
        return n + len(self.global_param_names)

    def flat_index(self):
        """name -> slice of the flat vector (for round-tripping ``TrainResults.ps``)."""
        idx, off = {}, 0
        for ch, shapes in zip(self.chains, self.layer_shapes()):
            for li, (o, i) in enumerate(shapes, 1):
                idx[(ch["name"], f"layer_{li}", "weight")] = (off, off + o * i, (o, i)); off += o * i
                idx[(ch["name"], f"layer_{li}", "bias")] = (off, off + o, (o,)); off += o
        for g in self.global_param_names:
            idx[(g,)] = (off, off + 1, (1,)); off += 1
        return idx

    def initialparameters(self, rng):
        """LuxCore.initialparameters: Dense = glorot_uniform weight, zero bias (Lux defaults);
        global parameters start at inv_sigmoid((default-ℓ)/(u-ℓ)) or rand (GenericHybridModel.jl:236-256).
        The weight values come from numpy's generator, not Julia's (documented in DESIGN.md)."""
        flat = np.zeros(self.num_params(), dtype=np.float32)
        off = 0
        for shapes in self.layer_shapes():
            for (o, i) in shapes:
                lim = math.sqrt(6.0 / (i + o))
                w = rng.uniform(-lim, lim, size=(o, i)).astype(np.float32)
                flat[off:off + o * i] = w.reshape(-1, order="F"); off += o * i
                off += o
        for g in self.global_param_names:
            flat[off] = scale_single_param_minmax(g, self.parameters) if self.start_from_default else np.float32(rng.random())
            off += 1
        return flat

    def initialstates(self):
        """LuxCore.initialstates: fixed parameters hold their defaults (GenericHybridModel.jl:289-303)."""
        de = default(self.parameters)
        return {"fixed": {f: np.array([de[f]], dtype=np.float32) for f in self.fixed_param_names}}

    def unflatten(self, flat):
        flat = np.asarray(flat, dtype=np.float32)
        out = {}
        for key, (a, b, shape) in self.flat_index().items():
            v = flat[a:b]
            v = v.reshape(shape, order="F") if len(shape) == 2 else v.copy()
            d = out
            for k in key[:-1]:
                d = d.setdefault(k, {})
            d[key[-1]] = v
        return out


class SingleNNHybridModel(_HybridModelBase):
    pass


class MultiNNHybridModel(_HybridModelBase):
    pass


def constructHybridModel(predictors=None, forcing=None, targets=None, mechanistic_model=None, parameters=None,
                         neural_param_names=None, global_param_names=None, *, hidden_layers=(32, 32),
                         activation="tanh", scale_nn_outputs=False, input_batchnorm=False, start_from_default=True,
                         **kwargs):
    """Unified constructor dispatching on the type of ``predictors``
    (GenericHybridModel.jl:89-232): a list -> SingleNNHybridModel (one chain with one output
    row per neural parameter), a dict -> MultiNNHybridModel (one chain per neural parameter)."""
    if mechanistic_model is None or parameters is None or targets is None or forcing is None:
        raise TypeError("constructHybridModel needs predictors, forcing, targets, mechanistic_model, parameters")
    params = build_parameters(parameters, mechanistic_model)
    all_names = params.names
    forcing = [str(f) for f in forcing]
    targets = [str(t) for t in targets]
    cfg = dict(hidden_layers=hidden_layers, activation=activation, scale_nn_outputs=scale_nn_outputs,
               input_batchnorm=input_batchnorm, start_from_default=start_from_default, **kwargs)
    if isinstance(predictors, dict):
        if neural_param_names is not None and global_param_names is None:
            # positional form (predictors, forcing, targets, model, parameters, global_param_names)
            global_param_names, neural_param_names = neural_param_names, None
        global_param_names = [str(g) for g in (global_param_names or [])]
        neural = [str(k) for k in predictors.keys()]
        chains = []
        for name, preds in predictors.items():
            hl = hidden_layers[name] if isinstance(hidden_layers, dict) else hidden_layers
            act = activation[name] if isinstance(activation, dict) else activation
            chains.append(dict(name=str(name), predictors=[str(p) for p in preds], hidden=[int(h) for h in hl],
                               activation=_act_name(act), n_out=1, input_batchnorm=bool(input_batchnorm)))
        cls = MultiNNHybridModel
    elif isinstance(predictors, (list, tuple)):
        if neural_param_names is None:
            raise AssertionError("Provide neural_param_names for Vector predictors")
        neural = [str(n) for n in neural_param_names]
        global_param_names = [str(g) for g in (global_param_names or [])]
        if not all(n in all_names for n in neural):
            raise AssertionError("neural_param_names ⊆ param_names")
        chains = []
        if len(predictors) > 0 and len(neural) > 0:
            chains.append(dict(name="ps", predictors=[str(p) for p in predictors], hidden=[int(h) for h in hidden_layers],
                               activation=_act_name(activation), n_out=len(neural), input_batchnorm=bool(input_batchnorm)))
        predictors = [str(p) for p in predictors]
        cls = SingleNNHybridModel
    else:
        raise TypeError(f"predictors must be a list or a dict, got {type(predictors).__name__}")
    fixed = [n for n in all_names if n not in neural and n not in global_param_names]
    return cls(predictors=predictors, forcing=forcing, targets=targets, mechanistic_model=mechanistic_model,
               parameters=params, neural_param_names=neural, global_param_names=global_param_names,
               fixed_param_names=fixed, scale_nn_outputs=bool(scale_nn_outputs),
               start_from_default=bool(start_from_default), config=cfg, chains=chains)


# ----------------------------------------------------------------------------------------
# descriptor for the C ABI
# ----------------------------------------------------------------------------------------


def predictor_columns(model):
    """union of predictor columns in first-use order (the X matrix of prepare_data)."""
    cols = []
    for ch in model.chains:
        for p in ch["predictors"]:
            if p not in cols:
                cols.append(p)
    return cols


class WeightL2:
    """Declarative form of the reference's documented extra loss (src/utils/extract_weights.jl:55-91, hook
    src/losses/compute_loss.jl:31-34):

        extra_loss = (ŷ, ps) -> (; l2 = λ * weight_l2(ps.<branch>; normalize),)

    A closure cannot cross the C ABI; ``WeightL2(lam, branches, normalize)`` says the same thing as data.  ``branches``:
    names of the Dense chains whose `weight` arrays take part (MultiNNHybridModel: the neural parameter names; ``None`` =
    every chain, i.e. ``weight_l2(ps)``).  ONE extra term: loss = agg([L, λ·Σw² [/ n]])."""

    def __init__(self, lam, branches=None, normalize=False):
        # (λ travels as Float32, like every scalar of the reference's Float32 path)
        self.lam, self.branches, self.normalize = float(np.float32(lam)), (None if branches is None else [str(b) for b in branches]), bool(normalize)

    def chain_mask(self, model):
        if self.branches is None:
            return 0
        names = [ch["name"] for ch in model.chains]
        mask = 0
        for b in self.branches:
            if b not in names:
                raise KeyError(f"WeightL2: no Dense chain named {b!r} (chains: {names})")
            mask |= 1 << names.index(b)
        return mask

    def value(self, model, flat):
        """λ·weight_l2 of a flat parameter vector (host side: evaluation-mode losses, tests)"""
        import numpy as np
        s, n, off = 0.0, 0, 0
        names = [ch["name"] for ch in model.chains]
        for ch, shapes in zip(model.chains, model.layer_shapes()):
            for (o, i) in shapes:
                if self.branches is None or ch["name"] in self.branches:
                    w = np.asarray(flat[off:off + o * i], dtype=np.float64)
                    s += float((w * w).sum()); n += o * i
                off += o * i + o
        return self.lam * (s / n if (self.normalize and n) else s)


def build_desc(model, *, training_loss="mse", agg="sum", opt=None, device=0, flags=0, extra_loss=None):
    """eh_model_desc for ``model`` + the TrainConfig fields that select the path
    (training_loss, agg, opt: src/config/TrainingConfig.jl:43, 64, 77)."""
    from .config import Adam  # local import: config imports nothing from here

    opt = opt or Adam(0.01)
    pcols = predictor_columns(model)
    names = model.parameters.names
    roles, ridx = [], []
    for n in names:
        if n in model.neural_param_names:
            if isinstance(model, MultiNNHybridModel):
                c = [ch["name"] for ch in model.chains].index(n)
                roles.append(_abi.ROLE_NEURAL); ridx.append(c * 65536 + 0)
            else:
                roles.append(_abi.ROLE_NEURAL); ridx.append(model.neural_param_names.index(n))
        elif n in model.global_param_names:
            roles.append(_abi.ROLE_GLOBAL); ridx.append(model.global_param_names.index(n))
        else:
            roles.append(_abi.ROLE_FIXED); ridx.append(0)
    chains = [dict(in_cols=[pcols.index(p) for p in ch["predictors"]], hidden=ch["hidden"], n_out=ch["n_out"],
                   activation=_abi.ACT[ch["activation"]], input_batchnorm=ch["input_batchnorm"]) for ch in model.chains]
    prog, outs = trace_process_model(model.mechanistic_model, model.forcing, names, model.targets)
    # the library recognises built-in forms in traced programs itself (match_builtin_program, csrc/eh_lib.cu); the host-side
    # matcher is kept as the default so that the descriptor says what the model is, EH_PY_NO_MATCH=1 ships the raw trace
    m = None if os.environ.get("EH_PY_NO_MATCH") else match_builtin(prog, outs, len(model.forcing), len(names))
    if m is not None:
        pm_name, (pi, pj, fk), consts = m
        pm = dict(process_model=_abi.PM[pm_name], pm_args=[(0, pi), (0, pj), (1, fk)], pm_consts=consts)
    else:
        pm = dict(process_model=_abi.PM["PROGRAM"],
                  pm_prog=[(_abi.OPS[op], a, b, imm) for (op, a, b, imm) in prog], pm_outputs=outs)
    if isinstance(training_loss, PerTarget):
        losses = list(training_loss.losses)
        if len(losses) != len(model.targets):
            raise AssertionError("Length of targets and PerTarget losses tuple must match")
    else:
        losses = [training_loss] * len(model.targets)
    if extra_loss is not None and not isinstance(extra_loss, WeightL2):
        raise NotImplementedError("extra_loss closures cannot cross the C ABI; WeightL2(lam, branches, normalize) is the native form")
    for l in losses:
        if str(l) not in _abi.LOSS:
            raise ValueError(f"training loss {l!r} has no fused kernel (supported: {sorted(_abi.LOSS)})")
    agg_name = agg if isinstance(agg, str) else getattr(agg, "__name__", str(agg))
    if agg_name not in _abi.AGG:
        raise ValueError(f"agg must be sum or mean, got {agg!r}")
    tab = model.parameters.table
    return _abi.DescBundle(
        n_pred=len(pcols), n_forc=len(model.forcing), n_targ=len(model.targets), chains=chains,
        roles=roles, role_index=ridx, deflt=tab[:, 0], lower=tab[:, 1], upper=tab[:, 2],
        scale_nn_outputs=model.scale_nn_outputs, loss_per_target=[_abi.LOSS[str(l)] for l in losses],
        agg=_abi.AGG[agg_name], opt_kind=_abi.OPT[type(opt).__name__], eta=opt.eta, beta1=opt.beta[0],
        beta2=opt.beta[1], eps=opt.epsilon, lam=getattr(opt, "lambda_", 0.0),
        adamw_coupled=int(getattr(opt, "couple", True)), device=device, flags=flags,
        l2_lambda=(extra_loss.lam if extra_loss is not None else 0.0),
        l2_normalize=int(extra_loss.normalize) if extra_loss is not None else 0,
        l2_chain_mask=(extra_loss.chain_mask(model) if extra_loss is not None else 0), **pm)


class PerTarget:
    """src/losses/compute_loss_types.jl:33-45: one loss per target."""

    def __init__(self, *losses):
        if len(losses) == 1 and isinstance(losses[0], (tuple, list)):
            losses = tuple(losses[0])
        self.losses = tuple(losses)

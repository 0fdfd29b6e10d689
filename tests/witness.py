"""Independent float64 witness of the training objective, written directly from the reference
sources with torch autograd (no shared code with oracle/ or the CUDA library):
forward  src/models/GenericHybridModel.jl:370-431 / 458-530, chain src/models/NNModels.jl:225-230,
squash   GenericHybridModel.jl:348-352, losses src/losses/loss_fn.jl:58-81,
assembly src/losses/compute_loss.jl:50-53, 115-145, mask src/training/train.jl:221-232."""
import numpy as np
import torch

ACTS = {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "relu": torch.relu,
        "swish": lambda z: z * torch.sigmoid(z), "identity": lambda z: z}


def objective(model, flat, xf, y, idx, training_loss="mse", agg="sum", bn_eps=1e-5, extra_loss=None):
    """returns (loss, grad) in float64; flat is the reference-ordered parameter vector"""
    from easyhybrid_b200.model import PerTarget, predictor_columns
    X, forc = xf
    th = torch.tensor(np.asarray(flat, dtype=np.float64), requires_grad=True)
    pcols = predictor_columns(model)
    Xb = torch.tensor(np.asarray(X, dtype=np.float64)[idx])
    off = 0
    nn_out = {}
    for ch in model.chains:
        a = Xb[:, [pcols.index(p) for p in ch["predictors"]]].T  # (in, B)
        if ch["input_batchnorm"]:
            mu = a.mean(dim=1, keepdim=True)
            var = a.var(dim=1, unbiased=False, keepdim=True)
            a = (a - mu) / torch.sqrt(var + bn_eps)
        widths = [len(ch["predictors"])] + list(ch["hidden"]) + [ch["n_out"]]
        for l in range(len(widths) - 1):
            i, o = widths[l], widths[l + 1]
            W = th[off:off + o * i].reshape(i, o).T  # column-major (o x i)
            off += o * i
            b = th[off:off + o]
            off += o
            a = W @ a + b[:, None]
            if l < len(widths) - 2:
                a = ACTS[ch["activation"]](a)
        nn_out[ch["name"]] = a
    lo, up = model.parameters.column(1), model.parameters.column(2)
    de = model.parameters.column(0)
    vals = {}
    single = len(model.chains) == 1 and model.chains[0]["name"] == "ps"
    for n in model.parameters.names:
        if n in model.neural_param_names:
            z = nn_out["ps"][model.neural_param_names.index(n)] if single else nn_out[n][0]
            vals[n] = float(lo[n]) + float(up[n] - lo[n]) * torch.sigmoid(z) if model.scale_nn_outputs else z
        elif n in model.global_param_names:
            raw = th[off + model.global_param_names.index(n)]
            vals[n] = float(lo[n]) + float(up[n] - lo[n]) * torch.sigmoid(raw)
        else:
            vals[n] = torch.tensor(float(de[n]), dtype=torch.float64)
    kwargs = {f: torch.tensor(np.asarray(forc[f], dtype=np.float64)[idx]) for f in model.forcing}
    import inspect
    sig = inspect.signature(model.mechanistic_model)
    kwargs.update({n: v for n, v in vals.items() if n in sig.parameters})
    out = _call_torch(model.mechanistic_model, kwargs)
    losses = training_loss.losses if isinstance(training_loss, PerTarget) else [training_loss] * len(model.targets)
    terms = []
    for t, lt in zip(model.targets, losses):
        yt = torch.tensor(np.asarray(y[t], dtype=np.float64)[idx])
        m = ~torch.isnan(yt)
        yh, yv = out[t][m], yt[m]
        if lt == "mse":
            terms.append(((yh - yv) ** 2).mean())
        elif lt == "rmse":
            terms.append(torch.sqrt(((yh - yv) ** 2).mean()))
        elif lt == "mae":
            terms.append((yh - yv).abs().mean())
        elif lt == "nseLoss":
            terms.append(((yh - yv) ** 2).sum() / ((yv - yv.mean()) ** 2).sum())
        elif lt in ("pearsonLoss", "kgeLoss", "pbkgeLoss"):
            # loss_fn.jl:75-77, 104-127, 160-174: Statistics.cor / std (corrected) / mean
            ds, do = yh - yh.mean(), yv - yv.mean()
            r = (ds * do).sum() / torch.sqrt((ds ** 2).sum() * (do ** 2).sum())
            alpha = torch.sqrt((ds ** 2).sum() / (do ** 2).sum())
            beta = yh.mean() / yv.mean()
            if lt == "pearsonLoss":
                terms.append(1.0 - r)
            elif lt == "kgeLoss":
                terms.append(torch.sqrt((r - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2))
            else:
                terms.append(torch.sqrt((r - 1) ** 2 + (beta - 1) ** 2))
        else:
            raise ValueError(lt)
    L = sum(terms) if agg == "sum" else sum(terms) / len(terms)
    if extra_loss is not None:
        # extra_loss = (ŷ, ps) -> (; l2 = λ * weight_l2(ps.<branches>; normalize),): loss = agg([L, l2]) (compute_loss.jl:31-34)
        s, n, off = 0.0, 0, 0
        for ch, shapes in zip(model.chains, model.layer_shapes()):
            for (o, i) in shapes:
                if extra_loss.branches is None or ch["name"] in extra_loss.branches:
                    s = s + (th[off:off + o * i] ** 2).sum(); n += o * i
                off += o * i + o
        E = extra_loss.lam * (s / n if extra_loss.normalize else s)
        L = (L + E) if agg == "sum" else (L + E) / 2
    L.backward()
    return float(L.detach()), th.grad.numpy().copy()


def _call_torch(fn, kwargs):
    """evaluate the user's process model on torch tensors (np.exp -> torch.exp via __array_ufunc__ shim)"""
    class T:
        __array_priority__ = 2000

        def __init__(self, v): self.v = v
        def _w(self, o):
            # literals are Float32 in the reference (0.1f0, 15.0f0): round them the same way
            return o.v if isinstance(o, T) else float(np.float32(o))
        def __add__(self, o): return T(self.v + self._w(o))
        __radd__ = __add__
        def __sub__(self, o): return T(self.v - self._w(o))
        def __rsub__(self, o): return T(self._w(o) - self.v)
        def __mul__(self, o): return T(self.v * self._w(o))
        __rmul__ = __mul__
        def __truediv__(self, o): return T(self.v / self._w(o))
        def __rtruediv__(self, o): return T(self._w(o) / self.v)
        def __pow__(self, o): return T(self.v ** self._w(o))
        def __rpow__(self, o): return T(torch.as_tensor(self._w(o), dtype=torch.float64) ** self.v)
        def __neg__(self): return T(-self.v)

        def __abs__(self): return T(torch.abs(self.v))

        def __array_ufunc__(self, ufunc, method, *inputs, **kw):
            name = ufunc.__name__
            if name in ("minimum", "maximum"):
                a, b = (x.v if isinstance(x, T) else torch.as_tensor(float(np.float32(x)), dtype=torch.float64) for x in inputs)
                return T(torch.minimum(a, b) if name == "minimum" else torch.maximum(a, b))
            f = {"exp": torch.exp, "log": torch.log, "sqrt": torch.sqrt, "tanh": torch.tanh, "sin": torch.sin, "cos": torch.cos,
                 "absolute": torch.abs, "negative": torch.neg}[name]
            return T(f(inputs[0].v))
    out = fn(**{k: T(v) for k, v in kwargs.items()})
    return {k: (v.v if isinstance(v, T) else v) for k, v in out.items()}

"""Loss bookkeeping on the host: names, directions and the metric table evaluated from the
sufficient statistics the eval kernel returns.  Mirrors src/losses/loss_fn.jl:58-205 and
src/losses/compute_loss.jl:50-66 (the per-loss-type NamedTuple of per-target values + agg)."""
from __future__ import annotations

import math

LOSS_TYPES = ["mse", "rmse", "mae", "pearson", "r2", "pearsonLoss", "nseLoss", "nse", "kgeLoss", "kge",
              "pbkgeLoss", "pbkge", "α", "β"]
_MAXIMIZE = {"pearson", "r2", "nse", "kge"}


def bestdirection(loss_type):
    """loss_fn.jl:186-188."""
    return "Maximize" if str(loss_type) in _MAXIMIZE else "Minimize"


def isbetter(new, best, loss_type):
    """loss_fn.jl:190-194."""
    return new > best if bestdirection(loss_type) == "Maximize" else new < best


def check_training_loss(loss_type):
    """loss_fn.jl:196-205: a metric that is maximised cannot be a training loss."""
    from .model import PerTarget
    for lt in (loss_type.losses if isinstance(loss_type, PerTarget) else [loss_type]):
        if bestdirection(lt) == "Maximize":
            raise ValueError(
                f"Got a metric that is defined as `to be maximized` as a training loss: {lt}.\n"
                "For training you must use a true loss (to be minimized), e.g. "
                ":nseLoss (1-NSE), :kgeLoss (1-KGE), :pearsonLoss (1-Pearson), or :mse.")


def metrics_from_stats(st):
    """All of loss_fn.jl's Val methods from (n, Sy, Sh, Syy, Shh, Syh, SSE, SAE, shift):
    sums of y-shift, yhat-shift and their products over the valid entries of one target."""
    n, sy, sh, syy, shh, syh, sse, sae, shift = [float(x) for x in st]
    nan = float("nan")
    if n <= 0:
        return {k: nan for k in LOSS_TYPES}
    my, mh = sy / n, sh / n
    cyy, chh, cyh = syy - n * my * my, shh - n * mh * mh, syh - n * my * mh   # centred sums
    mean_y, mean_h = my + shift, mh + shift
    cor = cyh / math.sqrt(cyy * chh) if cyy > 0 and chh > 0 else nan
    sd_o = math.sqrt(cyy / (n - 1)) if n > 1 else nan
    sd_s = math.sqrt(chh / (n - 1)) if n > 1 else nan
    alpha = sd_s / sd_o if sd_o and sd_o == sd_o else nan
    beta = mean_h / mean_y if mean_y != 0 else nan
    nse_loss = sse / cyy if cyy > 0 else nan
    kge = math.sqrt((cor - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2)
    pbkge = math.sqrt((cor - 1) ** 2 + (beta - 1) ** 2)
    return {
        "mse": sse / n, "rmse": math.sqrt(sse / n), "mae": sae / n, "pearson": cor, "r2": 1.0 - nse_loss,
        "pearsonLoss": 1.0 - cor, "nseLoss": nse_loss, "nse": 1.0 - nse_loss, "kgeLoss": kge, "kge": 1.0 - kge,
        "pbkgeLoss": pbkge, "pbkge": 1.0 - pbkge, "α": alpha, "β": beta,
    }


def assemble_losses(stats_per_target, targets, loss_types, agg="sum"):
    """_compute_loss(..., loss_types::Vector, agg) -> {loss_type: {target...: v, agg: v}}
    (compute_loss.jl:55-66)."""
    agg_name = agg if isinstance(agg, str) else getattr(agg, "__name__", "sum")
    out = {}
    per_target = [metrics_from_stats(st) for st in stats_per_target]
    for lt in loss_types:
        vals = [m[str(lt)] for m in per_target]
        a = sum(vals) if agg_name == "sum" else sum(vals) / len(vals)
        d = {t: v for t, v in zip(targets, vals)}
        d[agg_name] = a
        out[str(lt)] = d
    return out

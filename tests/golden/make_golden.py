"""Generates tests/golden/reference_kat.json: the known answers the reference's OWN tests pin
for the hot path, restated in numpy float64 exactly as those tests state them.

The reference is Julia and cannot run here (no julia in the image), so these are not outputs
of the reference: they are the closed forms its tests assert, on the inputs those tests use.

  loss_fn table        test/test_loss_fn.jl:6-8 (inputs), :15-74 (no mask), :90-145 (mask [1,1,0,1])
  scale_single_param   test/test_generic_hybrid_model.jl:77-126
  hard_sigmoid         test/test_generic_hybrid_model.jl:24-35
  _compute_loss sums   test/test_compute_loss.jl:11-14 (inputs), :69-79, :90-94

Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np


def table(yh, y):
    r = yh - y
    my, mh = y.mean(), yh.mean()
    cor = np.corrcoef(yh, y)[0, 1]
    so, ss = y.std(ddof=1), yh.std(ddof=1)
    alpha, beta = ss / so, mh / my
    nse_loss = (r ** 2).sum() / ((y - my) ** 2).sum()
    kge = np.sqrt((cor - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2)
    pbkge = np.sqrt((cor - 1) ** 2 + (beta - 1) ** 2)
    return {
        "mse": np.mean(r ** 2), "rmse": np.sqrt(np.mean(r ** 2)), "mae": np.mean(np.abs(r)), "pearson": cor,
        "r2": 1 - nse_loss, "nse": 1 - nse_loss, "pearsonLoss": 1 - cor, "nseLoss": nse_loss, "kgeLoss": kge,
        "kge": 1 - kge, "pbkgeLoss": pbkge, "pbkge": 1 - pbkge, "α": alpha, "β": beta,
    }


def main():
    yh = np.array([1.0, 2.0, 3.0, 4.0])
    y = np.array([1.1, 1.9, 3.2, 3.8])
    mask = np.array([True, True, False, True])
    out = {
        "loss_fn": {
            "yhat": yh.tolist(), "y": y.tolist(),
            "all_valid": {k: float(v) for k, v in table(yh, y).items()},
            "mask": mask.tolist(),
            "masked": {k: float(v) for k, v in table(yh[mask], y[mask]).items()},
        },
        "scale_single_param": {
            "params": {"a": [1.0, 0.0, 2.0], "b": [2.0, 1.0, 3.0]},
            "raw": 0.0, "scaled": {"a": 1.0, "b": 2.0}, "minmax": {"a": 0.0, "b": 0.0},
        },
        "hard_sigmoid": {"x": [0.0, 1.0, 2.0, -1.0, 5.0], "y": [0.5, 0.7, 0.9, 0.3, 1.0]},
    }
    yh2 = {"var1": np.array([1.0, 2.0, 3.0]), "var2": np.array([2.0, 3.0, 4.0])}
    y2 = {"var1": np.array([1.1, 1.9, 3.2]), "var2": np.array([1.8, 3.1, 3.0])}
    m2 = np.array([True, False, True])
    out["compute_loss"] = {
        "yhat": {k: v.tolist() for k, v in yh2.items()}, "y": {k: v.tolist() for k, v in y2.items()},
        "mse_sum": float(sum(np.mean((yh2[k] - y2[k]) ** 2) for k in yh2)),
        "mae_sum": float(sum(np.mean(np.abs(yh2[k] - y2[k])) for k in yh2)),
        "mask": m2.tolist(),
        "mse_sum_masked": float(sum(np.mean((yh2[k][m2] - y2[k][m2]) ** 2) for k in yh2)),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    print("wrote", path)


if __name__ == "__main__":
    main()

"""TrainConfig / DataConfig / optimiser rules: mirror of src/config/TrainingConfig.jl:9-185,
src/config/DataConfig.jl:7-59 and the re-exported Optimisers.jl rules (src/EasyHybrid.jl:59)."""
from __future__ import annotations

from dataclasses import dataclass, field, fields, replace
from typing import Any, Optional

from .losses import check_training_loss


@dataclass(frozen=True)
class Adam:
    eta: float = 0.001
    beta: tuple = (0.9, 0.999)
    epsilon: float = 1e-8


@dataclass(frozen=True)
class AdamW:
    eta: float = 0.001
    beta: tuple = (0.9, 0.999)
    lambda_: float = 0.0
    epsilon: float = 1e-8
    couple: bool = True  # Optimisers >= 0.4: decay eta*lambda*x; False: lambda*x (Optimisers 0.3 OptimiserChain form)


@dataclass(frozen=True)
class RMSProp:
    eta: float = 0.001
    rho: float = 0.9
    epsilon: float = 1e-8

    @property
    def beta(self):
        return (0.0, self.rho)


@dataclass(frozen=True)
class Descent:
    eta: float = 0.1
    epsilon: float = 0.0

    @property
    def beta(self):
        return (0.0, 0.0)


def is_optimisers_rule(opt):
    """train.jl:20-22: only Optimisers.jl rules take the Lux.Training loop (the fused path)."""
    return isinstance(opt, (Adam, AdamW, RMSProp, Descent))


@dataclass
class TrainConfig:
    """TrainingConfig.jl:9-160 (fields that have no meaning without Makie / JLD2 are accepted and ignored)."""
    nepochs: int = 200
    batchsize: int = 64
    opt: Any = field(default_factory=lambda: Adam(0.01))
    patience: int = 2 ** 62
    autodiff_backend: Any = "FusedCUDA"
    return_gradients: bool = True
    gdev: Any = 0          # CUDA device ordinal
    cdev: Any = "cpu"
    training_loss: Any = "mse"
    loss_types: list = field(default_factory=lambda: ["mse", "r2"])
    extra_loss: Any = None
    agg: Any = "sum"
    train_from: Any = None
    random_seed: Optional[int] = 161803
    model_name: str = ""
    return_model: str = "best"
    keep_history: bool = True
    save_training: bool = False
    monitor_names: list = field(default_factory=list)
    output_folder: str = ""
    plotting: bool = False
    show_progress: bool = False
    yscale: Any = None
    tracked_params: tuple = ()
    full_batch: bool = False
    promote_f64: bool = False
    eval_every: int = 1
    inner_maxiters: int = 4


@dataclass
class DataConfig:
    """DataConfig.jl:7-59."""
    split_by_id: Any = None
    folds: Any = None
    val_fold: Optional[int] = None
    shuffleobs: bool = False
    split_data_at: float = 0.8
    sequence_kwargs: Any = None
    array_type: str = "KeyedArray"


def validate_config(cfg: TrainConfig):
    """TrainingConfig.jl:162-185."""
    if cfg.return_model not in ("best", "final"):
        raise ValueError(f"return_model must be :best or :final, got :{cfg.return_model}")
    if cfg.batchsize <= 0:
        raise ValueError(f"batchsize must be positive, got {cfg.batchsize}")
    if cfg.nepochs <= 0:
        raise ValueError(f"nepochs must be positive, got {cfg.nepochs}")
    if cfg.patience <= 0:
        raise ValueError(f"patience must be positive, got {cfg.patience}")
    if cfg.eval_every <= 0:
        raise ValueError(f"eval_every must be positive, got {cfg.eval_every}")
    if cfg.inner_maxiters <= 0:
        raise ValueError(f"inner_maxiters must be positive, got {cfg.inner_maxiters}")
    check_training_loss(cfg.training_loss)
    return cfg


def override_configs(train_cfg, data_cfg, kwargs):
    """train.jl:300-314: split flat kwargs by the field names of the two config structs."""
    tnames = {f.name for f in fields(TrainConfig)}
    dnames = {f.name for f in fields(DataConfig)}
    t_over = {k: v for k, v in kwargs.items() if k in tnames}
    d_over = {k: v for k, v in kwargs.items() if k in dnames}
    rest = {k: v for k, v in kwargs.items() if k not in tnames and k not in dnames}
    return replace(train_cfg, **t_over), replace(data_cfg, **d_over), rest

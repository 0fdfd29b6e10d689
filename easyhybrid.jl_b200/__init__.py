"""easyhybrid.jl_b200 -- B200-native fused training step for EasyHybrid hybrid models.

The compute path is ``libeasyhybrid_cuda.so`` (hand-written sm_100a CUDA behind a C ABI,
``include/easyhybrid_cuda.h``).  This package is the host-side mirror of the reference's
API for that path (constructHybridModel / train / TrainConfig ...), calling the library
through ctypes exactly as the Julia shim calls it through ccall.  No PyTorch, no CPU fallback.

Import it as ``easyhybrid_b200`` (the directory name contains a dot; the repo root ships
an alias module).
"""
from .config import Adam, AdamW, DataConfig, Descent, RMSProp, TrainConfig, validate_config
from .data import prepare_data, split_data, splitobs_indices, valid_mask
from .losses import (LOSS_TYPES, assemble_losses, bestdirection, check_training_loss, isbetter,
                     metrics_from_stats)
from .model import (Expo_resp_model, Expo_resp_model2, LinearModel, LinearModel2, MultiNNHybridModel, ParameterContainer,
                    PerTarget, RbQ10, SingleNNHybridModel, WeightL2, build_desc, build_parameters, constructHybridModel,
                    default, hard_sigmoid, inv_hard_sigmoid, inv_sigmoid, lower, scale_single_param,
                    scale_single_param_minmax, upper)
from .session import FusedSession
from .train import TrainResults, train
from ._lib import LIB_PATH, EasyHybridCudaError

__all__ = [n for n in dir() if not n.startswith("_")]

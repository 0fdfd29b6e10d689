"""ctypes mirror of include/easyhybrid_cuda.h (the C-ABI boundary).

Nothing here computes anything: it only declares the structs, enums and function
signatures of ``libeasyhybrid_cuda.so`` so that the host-side mirror of the
EasyHybrid API (``model.py``, ``train.py``) can call the CUDA library exactly the way
the Julia shim does with ``ccall``.
"""
from __future__ import annotations

import ctypes as C

EH_ABI_VERSION = 2

# eh_status
EH_OK, EH_EINVAL, EH_ENOMEM, EH_ECUDA, EH_ENCCL, EH_EUNSUPPORTED = range(6)
STATUS_NAMES = {0: "EH_OK", 1: "EH_EINVAL", 2: "EH_ENOMEM", 3: "EH_ECUDA", 4: "EH_ENCCL", 5: "EH_EUNSUPPORTED"}

ACT = {"identity": 0, "tanh": 1, "sigmoid": 2, "relu": 3, "swish": 4}
ROLE_NEURAL, ROLE_GLOBAL, ROLE_FIXED = 0, 1, 2
LOSS = {"mse": 0, "rmse": 1, "mae": 2, "nseLoss": 3, "pearsonLoss": 4, "kgeLoss": 5, "pbkgeLoss": 6}
AGG = {"sum": 0, "mean": 1}
OPT = {"Adam": 0, "AdamW": 1, "RMSProp": 2, "Descent": 3}
PM = {"RBQ10": 0, "EXPO": 1, "LINEAR": 2, "LINEAR2": 3, "EXPO2": 4, "PROGRAM": 100}

OPS = {
    "const": 0, "forcing": 1, "param": 2,
    "add": 10, "sub": 11, "mul": 12, "div": 13, "pow": 14, "min": 15, "max": 16,
    "neg": 20, "exp": 21, "log": 22, "sqrt": 23, "tanh": 24, "sigmoid": 25, "abs": 26, "sin": 27, "cos": 28,
}

EH_FLAG_NO_GRAPH = 1
EH_FLAG_NO_PDL = 2
EH_FLAG_NO_PERSIST = 4
EH_FLAG_TENSOR_PIPE = 16
EH_FLAG_JIT = 32
EH_SPLIT_TRAIN, EH_SPLIT_VAL = 0, 1
EH_EVAL_STATS = 9
EH_COMM_ID_BYTES = 128
EH_DP_MOMENTS = 37


class eh_pm_arg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_int32)]


class eh_pm_instr(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("imm", C.c_float)]


class eh_chain_desc(C.Structure):
    _fields_ = [
        ("n_in", C.c_int32),
        ("in_cols", C.POINTER(C.c_int32)),
        ("n_hidden", C.c_int32),
        ("hidden", C.POINTER(C.c_int32)),
        ("n_out", C.c_int32),
        ("activation", C.c_int32),
        ("input_batchnorm", C.c_int32),
    ]


class eh_model_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_pred", C.c_int32), ("n_forc", C.c_int32), ("n_targ", C.c_int32),
        ("n_chains", C.c_int32),
        ("chains", C.POINTER(eh_chain_desc)),
        ("n_params", C.c_int32),
        ("role", C.POINTER(C.c_int32)),
        ("role_index", C.POINTER(C.c_int32)),
        ("deflt", C.POINTER(C.c_float)),
        ("lower", C.POINTER(C.c_float)),
        ("upper", C.POINTER(C.c_float)),
        ("scale_nn_outputs", C.c_int32),
        ("process_model", C.c_int32),
        ("n_pm_args", C.c_int32),
        ("pm_args", C.POINTER(eh_pm_arg)),
        ("pm_consts", C.c_float * 4),
        ("pm_prog", C.POINTER(eh_pm_instr)),
        ("pm_len", C.c_int32),
        ("pm_outputs", C.POINTER(C.c_int32)),
        ("loss_per_target", C.POINTER(C.c_int32)),
        ("agg", C.c_int32),
        ("opt_kind", C.c_int32),
        ("eta", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("lambda_", C.c_float),
        ("adamw_decay_coupled_eta", C.c_int32),
        ("device", C.c_int32),
        ("flags", C.c_int32),
        ("l2_lambda", C.c_float), ("l2_normalize", C.c_int32), ("l2_chain_mask", C.c_uint32),
    ]


def _arr(ctype, values):
    values = list(values)
    return (ctype * max(len(values), 1))(*values)


class DescBundle:
    """An ``eh_model_desc`` plus the ctypes arrays it points into (kept alive here)."""

    def __init__(self, *, n_pred, n_forc, n_targ, chains, roles, role_index, deflt, lower, upper,
                 scale_nn_outputs, process_model, pm_args=(), pm_consts=(), pm_prog=(), pm_outputs=(),
                 loss_per_target, agg, opt_kind, eta, beta1, beta2, eps, lam, adamw_coupled=1, device=0, flags=0,
                 l2_lambda=0.0, l2_normalize=0, l2_chain_mask=0):
        self._keep = []
        d = eh_model_desc()
        d.abi_version = EH_ABI_VERSION
        d.n_pred, d.n_forc, d.n_targ = n_pred, n_forc, n_targ
        cds = (eh_chain_desc * max(len(chains), 1))()
        for i, ch in enumerate(chains):
            in_cols = _arr(C.c_int32, ch["in_cols"])
            hidden = _arr(C.c_int32, ch["hidden"])
            self._keep += [in_cols, hidden]
            cds[i].n_in = len(ch["in_cols"])
            cds[i].in_cols = in_cols
            cds[i].n_hidden = len(ch["hidden"])
            cds[i].hidden = hidden
            cds[i].n_out = ch["n_out"]
            cds[i].activation = ch["activation"]
            cds[i].input_batchnorm = int(ch["input_batchnorm"])
        d.n_chains = len(chains)
        d.chains = cds
        n_params = len(roles)
        d.n_params = n_params
        a_role, a_ri = _arr(C.c_int32, roles), _arr(C.c_int32, role_index)
        a_d, a_l, a_u = _arr(C.c_float, deflt), _arr(C.c_float, lower), _arr(C.c_float, upper)
        d.role, d.role_index, d.deflt, d.lower, d.upper = a_role, a_ri, a_d, a_l, a_u
        d.scale_nn_outputs = int(scale_nn_outputs)
        d.process_model = process_model
        args = (eh_pm_arg * max(len(pm_args), 1))()
        for i, (kind, index) in enumerate(pm_args):
            args[i].kind, args[i].index = kind, index
        d.n_pm_args = len(pm_args)
        d.pm_args = args
        for i in range(4):
            d.pm_consts[i] = float(pm_consts[i]) if i < len(pm_consts) else 0.0
        prog = (eh_pm_instr * max(len(pm_prog), 1))()
        for i, (op, a, b, imm) in enumerate(pm_prog):
            prog[i].op, prog[i].a, prog[i].b, prog[i].imm = op, a, b, imm
        d.pm_prog = prog
        d.pm_len = len(pm_prog)
        outs = _arr(C.c_int32, pm_outputs)
        d.pm_outputs = outs
        lpt = _arr(C.c_int32, loss_per_target)
        d.loss_per_target = lpt
        d.agg = agg
        d.opt_kind = opt_kind
        d.eta, d.beta1, d.beta2, d.eps, d.lambda_ = eta, beta1, beta2, eps, lam
        d.adamw_decay_coupled_eta = int(adamw_coupled)
        d.device = device
        d.flags = flags
        d.l2_lambda, d.l2_normalize, d.l2_chain_mask = float(l2_lambda), int(l2_normalize), int(l2_chain_mask)
        self._keep += [cds, a_role, a_ri, a_d, a_l, a_u, args, prog, outs, lpt]
        self.desc = d

    def byref(self):
        return C.byref(self.desc)


# name -> (restype, argtypes); every symbol declared in include/easyhybrid_cuda.h
_p = C.c_void_p
_fp = C.POINTER(C.c_float)
_fpp = C.POINTER(_fp)
_i64p = C.POINTER(C.c_int64)
SIGNATURES = {
    "eh_create": (C.c_int, [C.POINTER(_p), C.POINTER(eh_model_desc)]),
    "eh_destroy": (None, [_p]),
    "eh_last_error": (C.c_char_p, [_p]),
    "eh_num_params": (C.c_int64, [_p]),
    "eh_upload": (C.c_int, [_p, C.c_int32, C.c_int64, _fp, _fpp, _fpp]),
    "eh_set_params": (C.c_int, [_p, _fp, C.c_int64]),
    "eh_get_params": (C.c_int, [_p, _fp, C.c_int64]),
    "eh_set_opt_state": (C.c_int, [_p, _fp, _fp, C.c_int64, C.c_int64]),
    "eh_get_opt_state": (C.c_int, [_p, _fp, _fp, C.c_int64, _i64p]),
    "eh_set_bn_state": (C.c_int, [_p, C.c_int32, _fp, _fp, C.c_int32]),
    "eh_get_bn_state": (C.c_int, [_p, C.c_int32, _fp, _fp, C.c_int32]),
    "eh_loss_grad": (C.c_int, [_p, _i64p, C.c_int64, _fp, _fp]),
    "eh_step": (C.c_int, [_p, _i64p, C.c_int64, _fp, _fp]),
    "eh_step_host": (C.c_int, [_p, C.c_int64, _fp, _fpp, _fpp, _fp]),
    "eh_step_host_async": (C.c_int, [_p, C.c_int64, _fp, _fpp, _fpp, _fp]),
    "eh_sync": (C.c_int, [_p]),
    "eh_epoch": (C.c_int, [_p, _i64p, C.c_int64, C.c_int64, _fp]),
    "eh_set_perm": (C.c_int, [_p, _i64p, C.c_int64]),
    "eh_run_steps": (C.c_int, [_p, C.c_int64, C.c_int64, C.c_int64, _fp]),
    "eh_eval": (C.c_int, [_p, C.c_int32, _fp, C.POINTER(C.c_double), _fp]),
    "eh_comm_id": (C.c_int, [_p, _p]),
    "eh_comm_init": (C.c_int, [_p, C.c_int32, C.c_int32, _p]),
    "eh_last_timing": (C.c_int, [_p, _fp, _i64p, _fp]),
    "eh_set_profiling": (C.c_int, [_p, C.c_int32]),
    "eh_kernel_variant": (C.c_char_p, [_p]),
    "eh_jit_check": (C.c_int32, [_p, C.c_char_p, C.c_size_t]),
    "eh_epoch_variant": (C.c_char_p, [_p, C.c_int64]),
    "eh_dp_batch_moments": (C.c_int, [_p, C.c_int64, C.POINTER(C.c_double)]),
    "eh_dp_set_batch_moments": (C.c_int, [_p, C.c_int64, C.POINTER(C.c_double)]),
    "eh_selftest_wide_gemm": (C.c_int, [C.c_int32] * 6 + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]),
    "eh_host_alloc": (C.c_int, [C.POINTER(_p), C.c_size_t]),
    "eh_host_free": (C.c_int, [_p]),
}


def declare(lib):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib

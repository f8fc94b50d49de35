"""ctypes binding of libglmmrmcml_b200.so (the C-ABI declared in include/glmmrmcml_b200.h).

There is no CPU fallback: if the shared library is missing, or the machine has no CUDA device, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMB_LIB") or os.path.join(_HERE, "libglmmrmcml_b200.so")   # GMB_LIB: instrumented builds (csrc/Makefile)

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
vp = C.c_void_p


class GmbError(RuntimeError):
    """Raised for any non-zero status of the C-ABI; ``code`` is the GMB_E* value."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[GMB error {code}] {msg}")
        self.code = code


class HmcStats(C.Structure):
    _fields_ = [("accept_rate", C.c_double), ("step_size_mean", C.c_double), ("steps_mean", C.c_double),
                ("leapfrog_total", C.c_double), ("kernel_ms", C.c_double), ("n_chains", C.c_int),
                ("nsamp_per_chain", C.c_int), ("rows_used", C.c_int), ("kernel_variant", C.c_int), ("zl_nonzeros", C.c_double), ("component_groups", C.c_int), ("factored", C.c_int), ("lane_components", C.c_int)]


GMB_OK, GMB_EINVAL, GMB_EFAMILY, GMB_ECUDA, GMB_ENOTPD, GMB_ENCCL, GMB_ESTATE, GMB_ECOV = range(8)

# name -> (restype, argtypes); every symbol include/glmmrmcml_b200.h declares
_cov_args = [ip, C.c_int, dp, C.c_int, dp, C.c_int]
_fixed_u_args = _cov_args + [dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
PROTOTYPES = {
    "gmb_last_error": (C.c_char_p, []),
    "gmb_version": (C.c_char_p, []),
    "gmb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "gmb_ctx_destroy": (None, [vp]),
    "gmb_ctx_sync": (C.c_int, [vp]),
    "gmb_ctx_launch_count": (C.c_int64, [vp]),
    "gmb_ctx_stream": (vp, [vp]),
    "gmb_ctx_timer_start": (C.c_int, [vp]),
    "gmb_ctx_timer_stop": (C.c_int, [vp, dp]),
    "gmb_ctx_flush_l2": (C.c_int, [vp]),
    "gmb_comm_unique_id": (C.c_int, [vp]),
    "gmb_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "gmb_comm_rank": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gmb_comm_allreduce_host": (C.c_int, [vp, dp, C.c_int]),
    "gmb_comm_bcast_host": (C.c_int, [vp, dp, C.c_int]),
    "gmb_model_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_char_p, C.c_char_p, C.POINTER(vp)]),
    "gmb_model_create_prec": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]),
    "gmb_model_destroy": (None, [vp]),
    "gmb_model_flink": (C.c_int, [vp]),
    "gmb_model_set_u": (C.c_int, [vp, dp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gmb_model_use_device_u": (C.c_int, [vp, C.c_int]),
    "gmb_model_get_u": (C.c_int, [vp, C.c_int, C.c_int, dp]),
    "gmb_model_rebuild_zd": (C.c_int, [vp]),
    "gmb_model_loglik": (C.c_int, [vp, dp, C.c_double, dp]),
    "gmb_model_loglik_batch": (C.c_int, [vp, dp, dp, C.c_int, dp]),
    "gmb_model_mcnr": (C.c_int, [vp, dp, C.c_double, dp, dp, dp, dp]),
    "gmb_cov_create": (C.c_int, [vp] + _cov_args + [C.POINTER(vp)]),
    "gmb_cov_destroy": (None, [vp]),
    "gmb_cov_dims": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gmb_cov_gen": (C.c_int, [vp, dp, C.c_int, dp]),
    "gmb_cov_mvn_ll": (C.c_int, [vp, dp, dp, C.c_int, C.c_int, C.c_int, dp]),
    "gmb_cov_mvn_ll_model": (C.c_int, [vp, dp, vp, C.c_int, dp]),
    "gmb_cov_mvn_ll_model_batch": (C.c_int, [vp, dp, C.c_int, vp, C.c_int, dp]),
    "gmb_cov_logdet": (C.c_int, [vp, dp, dp]),
    "gmb_hmc_sample": (C.c_int, [vp, dp, dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int,
                                 C.c_int, C.c_uint32, C.c_uint64, C.c_int, dp, dp, C.POINTER(HmcStats)]),
    "gmb_hmc_set_variant": (C.c_int, [C.c_int]),
    "gmb_estep_set_multi": (C.c_int, [C.c_int]),
    "gmb_estep_set_tf32": (C.c_int, [C.c_int]),
    "gmb_estep_set_sparse_zd": (C.c_int, [C.c_int]),
    "gmb_hmc_set_components": (C.c_int, [C.c_int]),
    "gmb_hmc_set_factored": (C.c_int, [C.c_int]),
    "gmb_hmc_set_lane": (C.c_int, [C.c_int]),
    "gmb_hmc_set_cluster_size": (C.c_int, [C.c_int]),
    "gmb_estep_set_rowstats": (C.c_int, [C.c_int]),
    "gmb_hmc_set_row_aggregation": (C.c_int, [C.c_int]),
    "gmb_cov_set_gram": (C.c_int, [C.c_int]),
    "gmb_cov_set_block_classes": (C.c_int, [C.c_int]),
    "gmb_estep_set_row_aggregation": (C.c_int, [C.c_int]),
    "gmb_set_object_cache": (C.c_int, [C.c_int]),
    "gmb_model_estep_rows": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "gmb_cov_block_classes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "gmb_model_logprob_grad": (C.c_int, [vp, dp, dp, C.c_double, dp, C.c_int, dp, dp]),
    "gmb_set_default_ctx": (C.c_int, [vp]),
    "gmb_mcml_set_importance_form": (C.c_int, [C.c_int]),
    "gmb_cov_shape": (C.c_int, [ip, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gmb_mvn_ll": (C.c_int, _cov_args + [dp, C.c_int, dp, C.c_int, C.c_int, dp]),
    "gmb_mcmc_sample": (C.c_int, [dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64, dp]),
    "gmb_mcml_optim": (C.c_int, _fixed_u_args + [dp, C.c_int, C.c_int, C.c_int, dp, dp, dp]),
    "gmb_mcml_simlik": (C.c_int, _fixed_u_args + [dp, C.c_int, C.c_int, dp, dp, dp]),
    "gmb_mcml_hess": (C.c_int, _fixed_u_args + [dp, C.c_int, C.c_double, C.c_int, dp]),
    "gmb_aic_mcml": (C.c_int, _fixed_u_args + [dp, C.c_int, dp, C.c_int, dp]),
    "gmb_mcml_full": (C.c_int, _cov_args + [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, dp, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int,
                                            C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64,
                                            dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp]),
    "gmb_mcml_la": (C.c_int, _cov_args + [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, dp, C.c_int,
                                          C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]),
    "gmb_mcml_la_nr": (C.c_int, _cov_args + [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, dp, C.c_int,
                                             C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]),
    "gmb_la_objectives": (C.c_int, _cov_args + [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, dp, dp, C.c_int, dp,
                                                C.c_double, C.c_int, dp, dp, dp, dp]),
}

OBJECTIVE = C.CFUNCTYPE(C.c_int, dp, C.c_int, C.c_int, dp, C.c_void_p)
PROTOTYPES.update({
    "gmb_minimize_bounded": (C.c_int, [OBJECTIVE, C.c_void_p, C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_int, dp, C.POINTER(C.c_int)]),
    "gmb_fd_gradient": (C.c_int, [OBJECTIVE, C.c_void_p, C.c_int, dp, dp, dp, dp, C.c_int, dp]),
    "gmb_fd_hessian": (C.c_int, [OBJECTIVE, C.c_void_p, C.c_int, dp, dp, dp, dp, C.c_int, dp, C.POINTER(C.c_int)]),
})

_LIB = None


def lib():
    """Loads the shared library (once).  Raises if it has not been built — there is no fallback path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C glmmrmcml_b200/csrc`).  glmmrmcml_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)          # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != 0:
        raise GmbError(rc, lib().gmb_last_error().decode(errors="replace"))

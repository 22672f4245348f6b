"""``torch.ops.baysar.eval_logp`` — the batched posterior as a registered PyTorch operator.

    logp = torch.ops.baysar.eval_logp(posterior.op_handle, theta)      # theta: CUDA float64 (N, n_params) -> (N,)
    logp, status = torch.ops.baysar.eval_logp_status(posterior.op_handle, theta)

The operator is a dispatcher entry only: its CUDA implementation calls ``baysar_eval_logp`` of the C ABI
(include/baysar_b200.h) on the caller's current stream.  There is no CPU implementation — a CPU tensor fails in the
dispatcher, like every other entry into the path without a GPU.  The model is passed as an integer handle (a compiled
model lives in one GPU's HBM and is not a tensor); ``BaysarPosterior.op_handle`` registers it.
"""
import weakref

import torch

_LIB = torch.library.Library("baysar", "DEF")
_LIB.define("eval_logp(int model, Tensor theta) -> Tensor")
_LIB.define("eval_logp_status(int model, Tensor theta) -> (Tensor, Tensor)")
_MODELS = weakref.WeakValueDictionary()


def register_model(device_model):
    """-> integer handle of a ``runtime.DeviceModel`` for the operator calls (the registry holds a weak reference)."""
    handle = id(device_model)
    _MODELS[handle] = device_model
    return handle


def _model(handle):
    try:
        return _MODELS[handle]
    except KeyError:
        raise RuntimeError(f"baysar.eval_logp: unknown or destroyed model handle {handle}") from None


def _eval_logp_cuda(model, theta):
    return _model(model).eval_logp(theta)[0]


def _eval_logp_status_cuda(model, theta):
    logp, status, _ = _model(model).eval_logp(theta)
    return logp, status


def _eval_logp_meta(model, theta):
    return theta.new_empty((theta.shape[0],))


def _eval_logp_status_meta(model, theta):
    return theta.new_empty((theta.shape[0],)), theta.new_empty((theta.shape[0],), dtype=torch.int32)


_LIB.impl("eval_logp", _eval_logp_cuda, "CUDA")
_LIB.impl("eval_logp_status", _eval_logp_status_cuda, "CUDA")
_LIB.impl("eval_logp", _eval_logp_meta, "Meta")
_LIB.impl("eval_logp_status", _eval_logp_status_meta, "Meta")

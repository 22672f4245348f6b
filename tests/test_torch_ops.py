"""torch.ops.baysar.* registration (CPU part: schema, meta kernel, no CPU implementation)."""
import pytest
import torch

import baysar_b200.torch_ops  # noqa: F401


def test_operator_schema_and_meta_kernel():
    assert str(torch.ops.baysar.eval_logp.default._schema) == "baysar::eval_logp(int model, Tensor theta) -> Tensor"
    theta = torch.zeros((7, 10), dtype=torch.float64, device="meta")
    assert torch.ops.baysar.eval_logp(1, theta).shape == (7,)
    logp, status = torch.ops.baysar.eval_logp_status(1, theta)
    assert logp.shape == (7,) and status.dtype == torch.int32


def test_operator_has_no_cpu_implementation():
    with pytest.raises(NotImplementedError):
        torch.ops.baysar.eval_logp(1, torch.zeros((2, 10), dtype=torch.float64))

"""`-m gpu`: each CUDA kernel, called through the C ABI, against a PyTorch fp32 reference of the same op."""
import pytest

import kernel_checks


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(kernel_checks.CHECKS.keys()))
def test_kernel(name):
    kernel_checks.CHECKS[name]()

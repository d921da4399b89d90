import pytest


@pytest.fixture(scope="session")
def be():
    """The product backend (CUDA). Fails loudly -- never falls back to the oracle."""
    import torch  # noqa: F401  (device presence check only)
    from spectre_b200 import halo2
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    b = halo2.Backend([0])
    yield b
    b.close()

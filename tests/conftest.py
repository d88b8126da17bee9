import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def unet_weights():
    from oracle import sd15

    return sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)


@pytest.fixture(scope="session")
def vae_weights():
    from oracle import sd15

    return sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1)


@pytest.fixture(scope="session")
def contexts():
    import torch

    g = torch.Generator().manual_seed(5)
    return [torch.randn(77, 768, generator=g) for _ in range(3)]


@pytest.fixture(scope="session")
def engine(unet_weights, vae_weights, contexts):
    """one engine for the whole GPU session: synthetic SD-1.5 U-Net + VAE encoder, three context slots"""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from diff_mining_b200.engine import Engine
    from oracle import sd15

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    eng = Engine(0)
    eng.load_state_dict(unet_weights, "unet.")
    eng.load_state_dict(vae_weights, "vae.")
    eng.finalize()
    eng.set_schedule(*sd15.schedule_tables())
    for i, c in enumerate(contexts):
        eng.set_context(i, c)
    return eng


@pytest.fixture(scope="session")
def unet_weights_gpu(unet_weights):
    import torch

    return {k: v.to("cuda:0") for k, v in unet_weights.items()}


@pytest.fixture(scope="session")
def vae_weights_gpu(vae_weights):
    import torch

    return {k: v.to("cuda:0") for k, v in vae_weights.items()}

"""mvae_b200 - B200-native MVAE training step.

    from mvae_b200 import MVAE, MultimodalVAE, MVAETrainer                      # MNIST (mnist/)
    from mvae_b200.celeba import MultimodalVAE, CelebATrainer                    # CelebA (celeba/)
    from mvae_b200.multimnist import MultimodalVAE, MultiMNISTTrainer            # MultiMNIST (multimnist/)
"""
from . import _lib  # noqa: F401
from .mnist import MVAE, MultimodalVAE, MVAETrainer, HostPipeline, TERMS  # noqa: F401

from .parallel import DataParallelTrainer  # noqa: F401
from .functional import ProductOfExperts, elbo_loss, loss_function  # noqa: F401
from . import celeba, multimnist, evaluation, checkpoint  # noqa: F401

__all__ = ["MVAE", "MultimodalVAE", "MVAETrainer", "HostPipeline", "DataParallelTrainer", "ProductOfExperts", "elbo_loss", "loss_function", "TERMS", "_lib"]

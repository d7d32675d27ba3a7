"""b200ldm - B200-native latent-diffusion sampling (UNet x scheduler steps -> AutoencoderKL decode).

Drop-in surface (same names / kwargs / state_dict keys as the MONAI classes the reference uses,
3d_ldm/inference.py:23-24,71-99):

    from b200ldm import DiffusionModelUNet, AutoencoderKL, DDPMScheduler, DDIMScheduler,
                        LatentDiffusionInferer, define_instance
"""
from .config import ConfigParser, define_instance  # noqa: F401
from .inferer import DiffusionInferer, LatentDiffusionInferer  # noqa: F401
from .nets import AutoencoderKL, DiffusionModelUNet  # noqa: F401
from .schedulers import DDIMScheduler, DDPMScheduler  # noqa: F401
from .serving import LatentDiffusionGenerator  # noqa: F401

__version__ = "0.1.0"

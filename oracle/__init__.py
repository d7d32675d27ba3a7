"""CPU oracle for the latent-sampling hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product path (the b200ldm package) never does.

PARITY UNPINNED: the arithmetic of this path lives in un-vendored MONAI (>=1.4, no
pin: /root/reference/3d_ldm/requirements.txt:7); MONAI is not installable here and the
reference's only test (3d_ldm/test_losses.py) asserts nothing about this path.  The
restatement follows SURVEY.md Appendices A-E and the reference call sites
(3d_ldm/inference.py:71-99, 3d_ldm/train_diffusion.py:326-333).
"""
from .nets import DiffusionModelUNet, AutoencoderKL, randomize_parameters  # noqa: F401
from .schedulers import DDPMScheduler, DDIMScheduler  # noqa: F401
from .inferer import DiffusionInferer, LatentDiffusionInferer  # noqa: F401

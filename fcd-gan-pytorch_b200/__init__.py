"""fcd-gan-pytorch_b200 — B200-native hot path of FCD-GAN (networks + loss stack) behind the reference's
nn.Module call surface.  Import as `fcdgan_b200`.

    from fcdgan_b200 import Generator, Segmentor, Discriminator_SRGAN_simple      # Module.py
    from fcdgan_b200 import CNetLoss, CGeneratorLoss, region_loss, MS_SSIM, SSIM  # Loss.py / ssim.py
    from fcdgan_b200 import usss_step, rsss_step, wsss_step                       # Demo_*.py loop bodies

Everything numerical runs in hand-written sm_100a CUDA (libfcd_b200.so, C ABI in include/fcd_b200.h); there is
no CPU fallback — constructing modules works anywhere, calling them needs a B200 and the built library.
"""
__version__ = "0.1.0"

from .engine import (get_precision, get_streams, invalidate_weight_cache, set_batch_branches, set_precision,  # noqa: F401
                     set_rowpack, set_streams, set_sync_bn)
from .modules import (DoubleConv, Discriminator_SRGAN_simple, Down, Generator, OutConv, ResidualBlock,  # noqa: F401
                      Segmentor, Up)
from .losses import (CGeneratorLoss, CNetLoss, PerceptionLoss, mean, mean_abs, mean_sq, region_loss,  # noqa: F401
                     soft_mask)
from .ssim import MS_SSIM, SSIM, ms_ssim, ssim  # noqa: F401
from .steps import rsss_g_step, rsss_step, usss_g_step, usss_s_step, usss_step, wsss_step  # noqa: F401

"""codenerf_b200 -- B200-native CodeNeRF render path (drop-in for the reference's
src/model.py and the hot-path functions of src/utils.py).  See DESIGN.md."""
from .model import CodeNeRF, PE                                                    # noqa: F401
from .utils import get_rays, sample_from_rays, volume_rendering, volume_rendering_with_acc, make_z_vals  # noqa: F401
from .render import RayBundle, render, render_view                                 # noqa: F401
from . import trainer, optimizer, parallel, checkpoint, data                        # noqa: F401

__all__ = ["CodeNeRF", "PE", "get_rays", "sample_from_rays", "volume_rendering", "volume_rendering_with_acc",
           "make_z_vals", "RayBundle", "render", "render_view"]

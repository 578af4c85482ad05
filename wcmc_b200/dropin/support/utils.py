"""`support.utils` surface used by the hot path and by train_kpcn.py
(/root/reference/support/utils.py:24-42 crop_like, :70-100 BasicArgumentParser)."""
import argparse

from sbmc.modules import crop_like  # noqa: F401  (same rule: crop = max(d // 2, 0), crop2 = d - crop)


class BasicArgumentParser(argparse.ArgumentParser):
    """Base CLI flags shared by the reference's train_* scripts (utils.py:70-100)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        add = self.add_argument
        add("--sbmc", action="store_true", help="train the sample-based kernel-splatting network")
        add("--p_buf", action="store_true", help="use the multi-bounce path buffers for denoising")
        add("--model_name", type=str, default="tSUNet", help="name of the model")
        add("--data_dir", type=str, default="./data", help="directory of dataset")
        add("--visual", action="store_true", help="use the visualizer instead of the terminal")
        add("-b", "--batch_size", type=int, default=64, help="batch size")
        add("-e", "--num_epoch", type=int, default=100, help="number of epochs")
        add("-v", "--val_epoch", type=int, default=1, help="validate every val_epoch epochs")
        add("--vis_iter", type=int, default=4, help="visualise every vis_iter iterations")
        add("--start_epoch", type=int, default=0, help="epoch to resume from")
        add("--num_samples", type=int, default=8, help="number of samples to display")
        add("--save", type=str, default="./weights", help="directory to save the model")
        add("--overfit", action="store_true", help="launch the overfitting test")


# ---- display helpers the reference's own (fall-through) modules import from `support.utils` -------------------
# support/datasets.py:12 does `from support.utils import ToneMap, LinearToSrgb`; with this drop-in first on sys.path
# that import lands here, so the three numpy tone mappers of utils.py:44-67 are provided as well (host-side display
# code, never on the hot path).
def _luminance_scale(c, channel_axis, limit):
    import numpy as np
    r, g, b = (np.take(c, i, axis=channel_axis) for i in range(3))
    return 1.0 + (0.2126 * r + 0.7152 * g + 0.0722 * b) / limit


def ToneMap(c, limit=1.5):
    """(W,H,3) linear radiance -> Reinhard-style compression by the pixel's luminance (utils.py:44-51)."""
    import numpy as np
    return c / np.expand_dims(_luminance_scale(c, 2, limit), 2)


def LinearToSrgb(c):
    """Gamma 2.2 + clip to [0,1] (utils.py:53-56)."""
    import numpy as np
    return np.clip(c ** (1.0 / 2.2), 0.0, 1.0)


def ToneMapBatch(c):
    """(B,3,W,H): ToneMap with limit 1.5, negative values clipped, then LinearToSrgb (utils.py:58-67)."""
    import numpy as np
    col = np.clip(c / np.expand_dims(_luminance_scale(c, 1, 1.5), 1), 0, None)
    return LinearToSrgb(col)

"""Full-frame KPCN inference (BASELINE.json configs[3]: 1280x720 denoise on one B200).

The reference denoises a frame as overlapping 128x128 tiles with a 92x92 valid centre
(`FullImageDataset`, /root/reference/support/datasets.py:1174-1425, `pad_size = 32` :1208, tile
protocol :1277-1300; driver /root/reference/test_models.py:49-101).  That protocol cannot tile 720
rows at all (`(h-64) % 64 == 0` is asserted, :1277-1278) and recomputes every halo: 190 tiles x
124.7 GFLOP = 23.7 TFLOP for a frame whose valid-convolution cost is 11.1 TFLOP.  The network is
fully convolutional, so here the whole frame goes through the same kernels once: the inputs are
replicate-padded by the 18-pixel receptive-field shrink of each side (9 valid 5x5 convolutions),
the 21x21 kernels are applied to the un-padded radiance buffers (zero outside the frame, exactly
what each interior pixel sees in the tiled protocol) and the branches are recombined as in
`sbmc.KPCN.forward`.  Same semantics as `KPCNInterface.validate_batch` (interfaces.py:278-318)
with `use_llpm_buf=False`.
"""
import torch
import torch.nn.functional as F

SHRINK = 18  # 9 valid 5x5 convolutions: 4 * 9 / 2 pixels per side


def pad_frame(batch, shrink=SHRINK):
    """batch: dict with kpcn_{diffuse,specular}_in (B,C,H,W), kpcn_{diffuse,specular}_buffer,
    kpcn_albedo (B,3,H,W).  Returns the dict KPCN.forward expects: network inputs replicate-padded to
    (H+2*shrink, W+2*shrink); buffers / albedo zero-padded likewise (KPCN crops them back, so the
    padding values are never read)."""
    out = dict(batch)
    for k in ("kpcn_diffuse_in", "kpcn_specular_in"):
        out[k] = F.pad(batch[k], (shrink,) * 4, mode="replicate")
    for k in ("kpcn_diffuse_buffer", "kpcn_specular_buffer", "kpcn_albedo"):
        out[k] = F.pad(batch[k], (shrink,) * 4)
    return out


@torch.no_grad()
def denoise_frame(kpcn, batch, padded=False):
    """-> dict(radiance, diffuse, specular), each (B,3,H,W) at the frame's own resolution."""
    if not padded:
        batch = pad_frame(batch, (kpcn.depth * 4) // 2)
    return kpcn(batch)

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


@torch.no_grad()
def denoise_raw_frame(kpcn, raw):
    """raw (H,W,S,104) fp32 cuda: the renderer's per-sample buffer of one frame.  GPU preprocessing (SURVEY 8(f) N3:
    `DenoiseDataset._preprocess_kpcn`, datasets.py:487-582, as kernels) -> whole-frame KPCN denoise; nothing visits the
    host.  -> dict(radiance, diffuse, specular), each (1,3,H,W)."""
    from . import preprocess
    return denoise_frame(kpcn, preprocess.frame_batch(raw))


@torch.no_grad()
def denoise_stream(kpcn, host_frames, out_host=None, device=None):
    """Denoises a sequence of frames held in (pinned) host memory with the transfers overlapped: the host -> device
    copy of frame i+1 runs on a copy stream while frame i is computed (engine.DevicePrefetcher) and the radiance of
    frame i goes back to pinned host memory on a third stream, so a frame costs max(compute, transfer) instead of
    their sum (the reference's loop does blocking `.cuda()` / `.cpu()` per tile, test_models.py:55-75).
    host_frames: iterable of dicts (un-padded kpcn_* tensors, (B,C,H,W)); out_host: optional list of pinned
    (B,3,H,W) buffers, reused round-robin.  Returns the list of host radiance tensors in frame order; the caller
    synchronises (`torch.cuda.synchronize()`) before reading them."""
    from .engine import DevicePrefetcher
    device = device or torch.device("cuda", torch.cuda.current_device())
    pf = DevicePrefetcher(host_frames, device)
    out_stream = torch.cuda.Stream(device)
    cur = torch.cuda.current_stream(device)
    results = []
    slot_free = {}
    for i, batch in enumerate(pf):
        rad = denoise_frame(kpcn, batch)["radiance"]
        pf.release()
        done = torch.cuda.Event()
        done.record(cur)
        if out_host:
            host = out_host[i % len(out_host)]
            if id(host) in slot_free:
                slot_free[id(host)].synchronize()      # the previous copy into this pinned buffer has finished
        else:
            host = torch.empty(rad.shape, dtype=rad.dtype).pin_memory()
        with torch.cuda.stream(out_stream):
            out_stream.wait_event(done)
            host.copy_(rad, non_blocking=True)
            rad.record_stream(out_stream)
            ev = torch.cuda.Event()
            ev.record(out_stream)
            slot_free[id(host)] = ev
        results.append(host)
    cur.wait_stream(out_stream)
    return results


# ------------------------------------------------------------------------------------------------
# the reference's tile protocol (FullImageDataset + test_models.inference) and its
# one-pass equivalent
# ------------------------------------------------------------------------------------------------
PATCH_SIZE = 128   # datasets.py:1207
PAD_SIZE = 32      # datasets.py:1208


def tile_coords(h, w, patch_size=PATCH_SIZE, pad_size=PAD_SIZE):
    """[(i_start, j_start, i_end, j_end, i, j)] exactly as FullImageDataset builds them (datasets.py:1277-1296):
    tiles of `patch_size` at stride patch_size - 2*pad_size; a tile owns its centre, extended to the frame border
    for the first / last tile of a row or column."""
    stride = patch_size - 2 * pad_size
    assert (h - 2 * pad_size) % stride == 0 and (w - 2 * pad_size) % stride == 0, \
        "frame %dx%d does not tile with patch %d / pad %d" % (h, w, patch_size, pad_size)
    coords = []
    for i in range(0, h - 2 * pad_size, stride):
        for j in range(0, w - 2 * pad_size, stride):
            i_start, j_start = (0 if i == 0 else i + pad_size), (0 if j == 0 else j + pad_size)
            i_end = i + patch_size if i == h - patch_size else i + patch_size - pad_size
            j_end = j + patch_size if j == w - patch_size else j + patch_size - pad_size
            coords.append((i_start, j_start, i_end, j_end, i, j))
    return coords


class FrameTiles(torch.utils.data.Dataset):
    """The item contract of `FullImageDataset.__getitem__` (datasets.py:1417-1421) over an in-memory frame:
    (patch dict, i_start, j_start, i_end, j_end, i, j); tensors are sliced on their last two dimensions
    (`sample[k][..., i:i+P, j:j+P]`, :1298-1300).  `frame` holds un-batched tensors: (C,H,W) buffers and
    (S,C,H,W) paths.  Attributes h, w, PATCH_SIZE, has_hit as test_models.py:52-53, :227 read them."""

    def __init__(self, frame, patch_size=PATCH_SIZE, pad_size=PAD_SIZE, has_hit=None):
        t = next(v for v in frame.values() if torch.is_tensor(v))
        self.h, self.w = int(t.shape[-2]), int(t.shape[-1])
        self.PATCH_SIZE, self.pad_size = patch_size, pad_size
        self.frame = frame
        self.coords = tile_coords(self.h, self.w, patch_size, pad_size)
        self.has_hit = has_hit

    def __len__(self):
        return len(self.coords)

    def __getitem__(self, idx):
        i_start, j_start, i_end, j_end, i, j = self.coords[idx]
        p = self.PATCH_SIZE
        patch = {k: (v[..., i:i + p, j:j + p] if torch.is_tensor(v) else v) for k, v in self.frame.items()}
        return patch, i_start, j_start, i_end, j_end, i, j


@torch.no_grad()
def inference(interface, dataloader, spp=None, args=None, device=None):
    """Drop-in for `test_models.inference` (test_models.py:49-101): runs `interface.validate_batch` tile by tile,
    replicate-pads each output back to the tile size and stitches the owned region of every tile.
    Returns (radiance (3,H,W), p-buffers {name: (S,C,H,W)} or tensor or None) as torch tensors on the device
    (the reference converts to numpy HWC afterwards, :92-99 -- `to_numpy_hwc` does that)."""
    interface.to_eval_mode()
    ds = dataloader.dataset
    H, W, P = ds.h, ds.w, ds.PATCH_SIZE
    out_rad, out_path = None, None
    for batch, i_start, j_start, i_end, j_end, i, j in dataloader:
        if torch.cuda.is_available():   # (host-logic tests drive this loop with a stand-in interface on CPU)
            batch = {k: (v.cuda(device, non_blocking=True) if torch.is_tensor(v) and not v.is_cuda else v)
                     for k, v in batch.items()}
        out, p_buffers = interface.validate_batch(batch)
        pad_h, pad_w = P - out.shape[2], P - out.shape[3]
        if pad_h != 0 and pad_w != 0:
            out = F.pad(out, (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2), "replicate")
        if out_rad is None:
            out_rad = torch.zeros((3, H, W), device=out.device)
        if p_buffers is not None and out_path is None:
            if isinstance(p_buffers, dict):
                out_path = {k: torch.zeros((v.shape[1], v.shape[2], H, W), device=out.device) for k, v in p_buffers.items()}
            else:
                out_path = torch.zeros((p_buffers.shape[1], p_buffers.shape[2], H, W), device=out.device)
        for b in range(out.shape[0]):
            i0, i1, j0, j1, ti, tj = (int(i_start[b]), int(i_end[b]), int(j_start[b]), int(j_end[b]), int(i[b]), int(j[b]))
            out_rad[:, i0:i1, j0:j1] = out[b, :, i0 - ti:i1 - ti, j0 - tj:j1 - tj]
            if isinstance(p_buffers, dict):
                for k, v in p_buffers.items():
                    out_path[k][:, :, i0:i1, j0:j1] = v[b, :, :, i0 - ti:i1 - ti, j0 - tj:j1 - tj]
            elif p_buffers is not None:
                out_path[:, :, i0:i1, j0:j1] = p_buffers[b, :, :, i0 - ti:i1 - ti, j0 - tj:j1 - tj]
    return out_rad, out_path


def to_numpy_hwc(out_rad, out_path):
    """The layout `test_models.inference` returns (test_models.py:92-99): radiance (H,W,3), p-buffers (H,W,S,C)."""
    rad = out_rad.detach().cpu().numpy().transpose([1, 2, 0])
    if isinstance(out_path, dict):
        out_path = {k: v.detach().cpu().numpy().transpose([2, 3, 0, 1]) for k, v in out_path.items()}
    elif torch.is_tensor(out_path):
        out_path = out_path.detach().cpu().numpy().transpose([2, 3, 0, 1])
    return rad, out_path


@torch.no_grad()
def inference_one_pass(interface, frame, allow_pathnet=False):
    """The same frame through `interface.validate_batch` ONCE (batch of one, no tiling): the networks are fully
    convolutional, so a pixel at least `SHRINK + 10` (= 28, the crop `test_models.denoise` applies, :217-219)
    pixels away from the frame border sees exactly the inputs it sees in the tile that owns it -- the tiled
    protocol costs (H-64)/64 x (W-64)/64 tiles x 124.7 GFLOP, this costs the frame's own valid-convolution
    FLOPs once.  The (H-36, W-36) output is replicate-padded back to (H, W) as the tiles are.
    That equivalence holds for the KPCN alone: PathNet's U-Net pools and zero-pads, so with `use_llpm_buf` a
    tile's p-buffers depend on where the tile was cut; the one-pass result is then a (seam-free) different
    function of the frame and must be asked for explicitly (`allow_pathnet=True`).
    `frame`: un-batched tensors as for FrameTiles.  Returns (radiance (3,H,W), p-buffers or None)."""
    assert allow_pathnet or not getattr(interface, "use_llpm_buf", False), \
        "one-pass inference is exact only without the path-embedding network (see docstring)"
    interface.to_eval_mode()
    dev = "cuda" if torch.cuda.is_available() else None
    batch = {k: (v.unsqueeze(0).to(dev) if torch.is_tensor(v) else v) for k, v in frame.items()}
    out, p_buffers = interface.validate_batch(batch)
    h, w = next(v for v in batch.values() if torch.is_tensor(v)).shape[-2:]
    ph, pw = h - out.shape[2], w - out.shape[3]
    if ph or pw:
        out = F.pad(out, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), "replicate")
    if isinstance(p_buffers, dict):
        p_buffers = {k: v[0] for k, v in p_buffers.items()}
    elif p_buffers is not None:
        p_buffers = p_buffers[0]
    return out[0], p_buffers

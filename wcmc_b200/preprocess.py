"""GPU preprocessing of raw OptaGen sample buffers -> the tensors the hot path consumes.

Mirrors `DenoiseDataset._preprocess_kpcn` / `_preprocess_llpm` (/root/reference/support/datasets.py:487-582, :301-361)
and the slicing of `__getitem__` (:1078-1110, transposes :760-793).  STATUS: the kernels were written after round 1's
GPU budget was spent; they compile for sm_100a, the oracle they will be checked against is pinned to the reference
(tests/test_oracle.py::test_preprocess_matches_reference), but they have not run on a GPU yet
(tests/test_gpu_preprocess.py runs as non-strict xfail until then) and nothing on the product path calls them.
"""
import torch

from . import lib

EPS = 0.00316


def preprocess_kpcn(raw):
    """(H,W,S,104) fp32 cuda -> (H,W,44) fp32, the reference's channel order."""
    return lib.preprocess_kpcn(raw)


def preprocess_llpm(raw):
    """(H,W,S,104) fp32 cuda -> (H,W,S,37) fp32."""
    return lib.preprocess_llpm(raw)


def kpcn_batch_tensors(kpcn_buffer, llpm_buffer=None):
    """The un-batched tensors of the batch contract (datasets.py:1078-1110; CHW / SCHW): views and one small cat."""
    b = kpcn_buffer
    out = {"kpcn_diffuse_in": torch.cat([b[..., :10], b[..., 20:]], 2), "kpcn_specular_in": b[..., 10:],
           "kpcn_diffuse_buffer": b[..., :3], "kpcn_specular_buffer": b[..., 10:13], "kpcn_albedo": b[..., 34:37] + EPS}
    if llpm_buffer is not None:
        pw = llpm_buffer[..., :1].mean(2)
        out["kpcn_diffuse_in"] = torch.cat([out["kpcn_diffuse_in"], pw], 2)
        out["kpcn_specular_in"] = torch.cat([out["kpcn_specular_in"], pw], 2)
        out["paths"] = llpm_buffer[..., 1:]
    return {k: (v.permute(2, 0, 1) if v.dim() == 3 else v.permute(2, 3, 0, 1)).contiguous() for k, v in out.items()}

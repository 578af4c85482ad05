"""ORACLE (test infrastructure only; nothing under wcmc_b200/ may import this).

CPU restatement of the preprocessing that turns a raw OptaGen sample buffer (H,W,S,104) fp32 into what the hot
path consumes:

  sanitize            /root/reference/support/datasets.py:621-624   non-finite -> 1e38, clamp to 1e38
  gradients           :286-299    left / top zero-padded finite differences
  preprocess_kpcn     :487-582    per-pixel mean / variance over spp, albedo factorisation, log specular,
                                  depth normalised by the image maximum -> (H,W,44)
  preprocess_llpm     :301-361    per-sample path descriptor (log radiance / light / throughput, bounce types / 19,
                                  sqrt roughness) -> (H,W,S,37)
  kpcn_batch_tensors  :1078-1110 (+ transposes :760-793)  slices of the two buffers that become the batch dict

PINNED: tests/golden/make_golden_n3.py runs the reference's own DenoiseDataset methods on a seeded raw buffer and
tests/test_oracle.py::test_preprocess_* compares this file against those vectors.
Raw channel ranges follow datasets.py:223-266 with MAX_DEPTH = 5.
"""
import numpy as np

MAX_DEPTH = 5                 # datasets.py:83
_D = MAX_DEPTH + 1
IDX_NSY = {"radiance": (2, 5), "diffuse": (5, 8)}                                           # :228-231
IDX_G = {"albedo_at_diff": (24 + _D * 7, 27 + _D * 7), "normal_at_diff": (27 + _D * 7, 30 + _D * 7),
         "depth_at_diff": (30 + _D * 7, 31 + _D * 7)}                                        # :243-248
IDX_SBMC = {"bounce_types": (24 + _D * 6, 24 + _D * 7)}                                      # :253-254
IDX_LLPM = {"path_weight": (31 + _D * 7, 32 + _D * 7), "radiance_wo_weight": (32 + _D * 7, 35 + _D * 7),
            "light_intensity": (35 + _D * 7, 38 + _D * 7), "throughputs": (38 + _D * 7, 38 + _D * 10),
            "roughnesses": (38 + _D * 10, 38 + _D * 11)}                                      # :256-266
EPS = 0.00316


def sanitize(sample):
    """datasets.py:621-624."""
    sample = np.where(np.isfinite(sample), sample, 1.0e+38)
    return np.where(sample < 1.0e+38, sample, 1.0e+38).astype(np.float32)


def gradients(buf):
    """datasets.py:286-299: (h,w,c) -> (h,w,2c) = [dx | dy], zero in the first column / row."""
    dx = np.zeros_like(buf)
    dy = np.zeros_like(buf)
    dx[:, 1:] = buf[:, 1:] - buf[:, :-1]
    dy[1:] = buf[1:] - buf[:-1]
    return np.concatenate([dx, dy], 2)


def _sl(sample, rng):
    return sample[..., rng[0]:rng[1]]


def preprocess_kpcn(sample):
    """datasets.py:487-582.  sample (H,W,S,104) fp32 -> (H,W,44) fp32:
    diffuse 3 | var 1 | grad 6 | specular 3 | var 1 | grad 6 | normal 3 | var 1 | grad 6 | depth 1 | var 1 | grad 2 |
    albedo 3 | var 1 | grad 6."""
    spp = sample.shape[2]
    normal_s = _sl(sample, IDX_G["normal_at_diff"])
    normal = normal_s.mean(2)
    normal_v = normal_s.var(2).mean(2, keepdims=True) / spp
    depth_s = _sl(sample, IDX_G["depth_at_diff"])
    depth = depth_s.mean(2)
    depth_v = depth_s.var(2)
    max_depth = depth.max()
    if max_depth > 0:
        depth = depth / max_depth
        depth_v = depth_v / (max_depth * max_depth * spp)
    depth = np.clip(depth, 0, 1)
    albedo_s = _sl(sample, IDX_G["albedo_at_diff"])
    albedo = albedo_s.mean(2)
    albedo_v = albedo_s.var(2).mean(2, keepdims=True) / spp
    albedo_sqr = ((albedo + EPS) * (albedo + EPS)).mean(2, keepdims=True)
    diff_s = np.maximum(_sl(sample, IDX_NSY["diffuse"]), 0)
    diffuse = diff_s.mean(2)
    diffuse_v = diff_s.var(2).mean(2, keepdims=True) / spp
    spec_s = np.maximum(np.maximum(_sl(sample, IDX_NSY["radiance"]), 0) - diff_s, 0)
    specular = spec_s.mean(2)
    specular_v = spec_s.var(2).mean(2, keepdims=True) / spp
    specular_sqr = ((1 + specular) * (1 + specular)).mean(2, keepdims=True)
    diffuse = diffuse / (albedo + EPS)
    diffuse_v = diffuse_v / albedo_sqr
    specular = np.log(1 + specular)
    specular_v = specular_v / specular_sqr
    feats = [diffuse, diffuse_v, gradients(diffuse), specular, specular_v, gradients(specular), normal, normal_v,
             gradients(normal), depth, depth_v, gradients(depth), albedo, albedo_v, gradients(albedo)]
    return np.concatenate(feats, axis=2)


def preprocess_llpm(sample):
    """datasets.py:301-361.  sample (H,W,S,104) -> (H,W,S,37): path weight 1 | radiance 3 | light 3 | throughput 18 |
    bounce types 6 | roughness 6."""
    feats = [np.log(_sl(sample, IDX_LLPM["path_weight"]) + 1e-6) / 90.0,
             np.log(_sl(sample, IDX_LLPM["radiance_wo_weight"]) + 1e-6) / 30.0,
             np.log(_sl(sample, IDX_LLPM["light_intensity"]) + 1e-8) / 10.0,
             np.log(_sl(sample, IDX_LLPM["throughputs"]) + 1e-6) / 30.0,
             _sl(sample, IDX_SBMC["bounce_types"]) / 19.0,
             np.sqrt(_sl(sample, IDX_LLPM["roughnesses"]))]
    return np.concatenate(feats, axis=3)


def kpcn_batch_tensors(kpcn_buffer, llpm_buffer=None):
    """datasets.py:1078-1084 (KPCN slices), :1086-1110 (path-weight mean channel + path descriptor), transposed to the
    CHW / SCHW layout the dataset hands to the DataLoader (:760-793).  Un-batched tensors."""
    b = kpcn_buffer
    out = {"kpcn_diffuse_in": np.concatenate([b[..., :10], b[..., 20:]], axis=2),
           "kpcn_specular_in": b[..., 10:],
           "kpcn_diffuse_buffer": b[..., :3],
           "kpcn_specular_buffer": b[..., 10:13],
           "kpcn_albedo": b[..., 34:37] + EPS}
    if llpm_buffer is not None:
        pw = llpm_buffer[..., :1].mean(2)
        out["kpcn_diffuse_in"] = np.concatenate([out["kpcn_diffuse_in"], pw], axis=2)
        out["kpcn_specular_in"] = np.concatenate([out["kpcn_specular_in"], pw], axis=2)
        out["paths"] = np.array(llpm_buffer[..., 1:])
    res = {}
    for k, v in out.items():
        res[k] = np.ascontiguousarray(v.transpose([2, 0, 1]) if v.ndim == 3 else v.transpose([2, 3, 0, 1]))
    return res

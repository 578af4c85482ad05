"""CPU: host-side logic that needs no GPU -- stream fork/join is a no-op without CUDA, the fused optimiser's
eligibility rules, the synthetic batch contract, crop_like, and the all-pairs oracle against brute force."""
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_streams_are_noop_without_cuda():
    from wcmc_b200 import streams
    if torch.cuda.is_available():
        return
    with streams.fork("diffuse"):
        x = torch.ones(3) * 2
    streams.join()
    assert float(x.sum()) == 6.0 and not streams._open()


def test_fused_adam_eligibility():
    from wcmc_b200 import optim as wopt
    p = [torch.nn.Parameter(torch.zeros(4))]
    assert not wopt.supported(torch.optim.Adam(p, lr=1e-3))                    # CPU parameters
    assert not wopt.supported(torch.optim.SGD(p, lr=1e-3))
    assert not wopt.supported(torch.optim.AdamW(p, lr=1e-3))
    if torch.cuda.is_available():
        q = [torch.nn.Parameter(torch.zeros(4, device="cuda"))]
        assert wopt.supported(torch.optim.Adam(q, lr=1e-3))
        assert not wopt.supported(torch.optim.Adam(q, lr=1e-3, amsgrad=True))
        assert not wopt.supported(torch.optim.Adam(q, lr=1e-3, weight_decay=0.1))


def test_synthetic_batch_contract():
    """Tensor contract of DenoiseDataset.__getitem__ (/root/reference/support/datasets.py:1078-1126)."""
    from wcmc_b200.synth import make_batch
    b = make_batch(batch=2, spp=3, size=20, seed=1)
    assert tuple(b["kpcn_diffuse_in"].shape) == (2, 35, 20, 20) and tuple(b["paths"].shape) == (2, 3, 36, 20, 20)
    assert torch.equal(b["kpcn_diffuse_in"][:, :3], b["kpcn_diffuse_buffer"])
    assert float(b["kpcn_albedo"].min()) >= 0.00316
    want = b["kpcn_albedo"] * b["target_diffuse"] + torch.exp(b["target_specular"]) - 1.0
    torch.testing.assert_close(b["target_total"], want)
    v = make_batch(batch=1, size=16, paths=False)
    assert tuple(v["kpcn_specular_in"].shape) == (1, 34, 16, 16) and "paths" not in v
    again = make_batch(batch=2, spp=3, size=20, seed=1)
    assert all(torch.equal(b[k], again[k]) for k in b)


def test_crop_like_matches_reference_rule():
    """support/utils.py:24-42: crop = max(delta // 2, 0), crop2 = delta - crop (128 -> 92 = [18:110])."""
    from wcmc_b200 import dropin
    dropin.install()
    from support.utils import crop_like
    src = torch.arange(128 * 128.0).reshape(1, 1, 128, 128)
    out = crop_like(src, torch.empty(1, 1, 92, 92))
    assert torch.equal(out, src[..., 18:110, 18:110])
    odd = crop_like(torch.arange(49.0).reshape(1, 1, 7, 7), torch.empty(1, 1, 4, 4))
    assert torch.equal(odd, torch.arange(49.0).reshape(1, 1, 7, 7)[..., 1:5, 1:5])


def test_allpairs_oracle_vs_brute_force():
    spec = importlib.util.spec_from_file_location("oracle_allpairs_ref", os.path.join(ROOT, "oracle", "allpairs_ref.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(0)
    p, r = torch.rand(97, 5, generator=g).double(), (torch.rand(97, 3, generator=g) * 3).double()
    t = ref.tonemap_gamma(r)
    e = 0.5 * ((p[:, None] - p[None]) ** 2).sum(-1) - 0.5 * ((t[:, None] - t[None]) ** 2).sum(-1)
    off = ~torch.eye(97, dtype=torch.bool)
    torch.testing.assert_close(ref.allpairs_loss(p, r, chunk=10), (0.5 * e[off] ** 2).sum() / 97 ** 2)
    keep = off & (0.5 * ((t[:, None] - t[None]) ** 2).sum(-1) < 0.05)
    torch.testing.assert_close(ref.allpairs_loss(p, r, tau=0.05, chunk=33), (0.5 * e[keep] ** 2).sum() / 97 ** 2)
    x = 2.0 * torch.cat([e[off], -e[off], torch.zeros(1, dtype=torch.float64)])
    want = (torch.logsumexp(x, 0) - torch.log(torch.tensor(1.0 + 2 * int(off.sum())))) / 2 ** 0.5
    torch.testing.assert_close(ref.allpairs_loss(p, r, mode="lse", chunk=16), want)


def _ref_tile_coords(h, w, patch_size=128, pad_size=32):
    """Line-by-line restatement of /root/reference/support/datasets.py:1277-1296 (test infrastructure)."""
    stride = patch_size - 2 * pad_size
    assert (h - 2 * pad_size) % stride == 0 and (w - 2 * pad_size) % stride == 0
    coords = []
    for i in range(0, h - 2 * pad_size, stride):
        for j in range(0, w - 2 * pad_size, stride):
            i_start = i + pad_size
            j_start = j + pad_size
            i_end = i + patch_size - pad_size
            j_end = j + patch_size - pad_size
            if i == 0:
                i_start = 0
            if j == 0:
                j_start = 0
            if i == h - patch_size:
                i_end = i + patch_size
            if j == w - patch_size:
                j_end = j + patch_size
            coords.append((i_start, j_start, i_end, j_end, i, j))
    return coords


class _CropInterface:
    """Stand-in for KPCNInterface.validate_batch: 'denoises' by returning the 92x92 centre of the noisy buffer."""
    use_llpm_buf = False

    def to_eval_mode(self):
        pass

    def validate_batch(self, batch):
        return batch["kpcn_diffuse_buffer"][..., 18:-18, 18:-18], None


def test_tile_protocol_matches_reference_rule():
    """wcmc_b200.inference: tile coordinates of FullImageDataset and the stitching of test_models.inference
    (test_models.py:49-90), driven on CPU with a stand-in interface."""
    import pytest
    from wcmc_b200 import inference as inf
    for h, w in ((128, 128), (192, 256), (256, 320), (704, 1280)):
        coords = inf.tile_coords(h, w)
        assert coords == _ref_tile_coords(h, w)
        cover = torch.zeros(h, w)
        for i0, j0, i1, j1, _, _ in coords:   # owned regions partition the frame
            cover[i0:i1, j0:j1] += 1
        assert bool((cover == 1).all())
    with pytest.raises(AssertionError):
        inf.tile_coords(720, 1280)      # (720 - 64) % 64 != 0: the reference cannot tile 720 rows either
    g = torch.Generator().manual_seed(5)
    frame = {"kpcn_diffuse_buffer": torch.rand(3, 192, 256, generator=g), "paths": torch.rand(2, 4, 192, 256, generator=g),
             "note": "kept"}
    ds = inf.FrameTiles(frame)
    assert len(ds) == 6 and ds.h == 192 and ds.w == 256 and ds.PATCH_SIZE == 128
    patch, i0, j0, i1, j1, i, j = ds[4]
    assert tuple(patch["paths"].shape) == (2, 4, 128, 128) and (i, j) == (64, 64) and patch["note"] == "kept"
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False)
    rad, pb = inf.inference(_CropInterface(), loader)
    assert pb is None and tuple(rad.shape) == (3, 192, 256)
    src = frame["kpcn_diffuse_buffer"]
    # wherever the owning tile's 92x92 valid centre covers the pixel, the stitched frame is the source
    assert torch.equal(rad[:, 18:-18, 18:-18], src[:, 18:-18, 18:-18])
    # the outer 18 pixels come from the replicate padding of the border tiles (test_models.py:66-69)
    assert torch.equal(rad[:, 0, 0], src[:, 18, 18]) and torch.equal(rad[:, -1, 100], src[:, -19, 100])
    one, _ = inf.inference_one_pass(_CropInterface(), frame)
    assert torch.equal(one, rad)
    np_rad, _ = inf.to_numpy_hwc(rad, None)
    assert np_rad.shape == (192, 256, 3)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: the oracle port on the host cores, one JSON line with the base contract's keys;
    under torchrun only rank 0 works, the other ranks exit 0 silently."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "KPCN+WCMC train patches/s" and line["unit"] == "patches/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env["RANK"], env["WORLD_SIZE"], env["LOCAL_RANK"] = "1", "2", "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_preprocess_batch_tensor_bookkeeping_matches_oracle():
    """wcmc_b200.preprocess.kpcn_batch_tensors (pure slicing / permutes, runs on any device) against the oracle's
    restatement of datasets.py:1078-1110 on the reference-generated buffers."""
    import numpy as np
    from wcmc_b200 import preprocess
    spec = importlib.util.spec_from_file_location("oracle_preprocess_ref", os.path.join(ROOT, "oracle", "preprocess_ref.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_n3.npz"))
    kp, ll = g["kpcn_b"], g["llpm_b"]
    want = ref.kpcn_batch_tensors(kp, ll)
    got = preprocess.kpcn_batch_tensors(torch.from_numpy(kp), torch.from_numpy(ll))
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k].is_contiguous() and tuple(got[k].shape) == want[k].shape, k
        np.testing.assert_allclose(got[k].numpy(), want[k], rtol=1e-6, atol=1e-7, err_msg=k)
    no_paths = preprocess.kpcn_batch_tensors(torch.from_numpy(kp))
    assert "paths" not in no_paths and tuple(no_paths["kpcn_diffuse_in"].shape) == (34,) + kp.shape[:2]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_preprocess_kernel_arithmetic_on_host(tag):
    """The per-element functions the N3 CUDA kernels call (wcmc_b200/csrc/preprocess_math.cuh, __host__ __device__)
    compiled for the host (tests/native/preprocess_host_check.cu) against the vectors the REFERENCE's own
    _preprocess_kpcn / _preprocess_llpm produced -- channel indices, formulas, the NaN / inf clamp, the depth
    normalisation and the finite differences are checked here; only the launch mechanics are left for the GPU."""
    import ctypes
    import shutil
    import subprocess
    import numpy as np
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    so = os.path.join(ROOT, "build", "preprocess_host_check.so")
    src = os.path.join(ROOT, "tests", "native", "preprocess_host_check.cu")
    hdr = os.path.join(ROOT, "wcmc_b200", "csrc", "preprocess_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets", "-o", so,
                        src], check=True, capture_output=True)
    lib = ctypes.CDLL(so)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_n3.npz"))
    raw = np.ascontiguousarray(g["raw_" + tag], dtype=np.float32)      # un-sanitised: inf / NaN / 3e38 outliers in "a"
    h, w, s, _ = raw.shape
    fp = ctypes.POINTER(ctypes.c_float)
    ll = np.empty((h, w, s, 37), np.float32)
    lib.host_preprocess_llpm(raw.ctypes.data_as(fp), ctypes.c_long(h * w * s), ll.ctypes.data_as(fp))
    np.testing.assert_allclose(ll, g["llpm_" + tag], rtol=2e-6, atol=1e-7)
    kp = np.empty((h, w, 44), np.float32)
    ws = np.empty((h * w, 18), np.float32)
    lib.host_preprocess_kpcn(raw.ctypes.data_as(fp), h, w, s, ws.ctypes.data_as(fp), kp.ctypes.data_as(fp))
    want = g["kpcn_" + tag]
    assert np.array_equal(np.isnan(kp), np.isnan(want))     # NaN exactly where the reference has it (overflowed variance)
    np.testing.assert_allclose(kp, want, rtol=3e-5, atol=2e-6, equal_nan=True)


def test_bench_kernel_summary_from_profile():
    """bench.summarize_kernels: the `roofline` / `roofline_more` / `kernels` entries of the bench line from a profile
    of the eager timed pass.  Fed with the per-call table of the final round-1 bench (profiles/r01final_bench.json)."""
    import json
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    d = json.loads(open(os.path.join(ROOT, "profiles", "r01final_bench.json")).read().strip().splitlines()[-1])
    steps = 10
    prof = {}
    for name, v in d["kernels"].items():   # (calls, total ms, total work) as lib.profile_stop() returns them
        ms = v["ms_per_step"] * steps
        rate = v.get("tflops", 0.0) * 1e12 if "tflops" in v else v.get("gbs", 0.0) * 1e9
        prof[name] = (v["calls_per_step"] * steps, ms, rate * ms * 1e-3)
    pk = {"hbm_gbs": 6456.2, "bf16_tflops": 1699.4, "bf16_tflops_sustained": 1433.8}
    roofline, more, kernels = bench.summarize_kernels(prof, steps, pk, "measured")
    json.dumps({"roofline": roofline, "roofline_more": more, "kernels": kernels})     # serialisable
    # the dominant kernel is FOUND by measured time: in that profile the weight-gradient kernel (5x5 + 3x3 layers,
    # 1.34 + 1.04 ms) is ahead of the 5x5 forward / data-gradient launches (2.13 ms)
    ms_w = d["kernels"]["conv2d_wgrad_k5"]["ms_per_step"] + d["kernels"]["conv2d_wgrad_k3"]["ms_per_step"]
    assert ms_w > d["kernels"]["conv2d_k5"]["ms_per_step"]
    assert roofline["kernel"].startswith("conv_wgrad_group_kernel + wgrad_reduce_batch_kernel (weight gradients of all")
    assert abs(roofline["ms_per_step"] - ms_w) < 1e-3 and roofline["launches_per_step"] == 48
    # an eager, event-bracketed pass is not power-limited: burst denominator, the sustained fraction beside it
    assert roofline["bound"] == "tensor" and roofline["unit"] == "TFLOP/s" and roofline["peak"] == 1699.4
    assert abs(roofline["frac"] - roofline["achieved"] / 1699.4) < 1e-3 and 0.2 < roofline["frac"] < 0.45
    assert abs(roofline["frac_sustained"] - roofline["achieved"] / 1433.8) < 1e-3
    assert roofline["traffic"] and roofline["traffic"] > 1e6
    by = {m["kernel"]: m for m in more}
    ka = by["kernel_apply_fwd_kernel (8 x 92^2 pixels per launch)"]
    assert ka["bound"] == "hbm" and ka["peak"] == 6456.2 and ka["unit"] == "GB/s"
    k5w = [m for m in more if "5x5 layers: one launch" in m["kernel"]][0]
    k3w = [m for m in more if "3x3 layers: one launch" in m["kernel"]][0]
    assert abs(k5w["achieved"] - d["kernels"]["conv2d_wgrad_k5"]["tflops"]) < 1.0
    assert abs(k3w["achieved"] - d["kernels"]["conv2d_wgrad_k3"]["tflops"]) < 1.0
    k5 = [m for m in more if m["kernel"].startswith("conv_igemm_kernel<1,5>")][0]
    assert abs(k5["achieved"] - d["kernels"]["conv2d_k5"]["tflops"]) < 1.0 and k5["launches_per_step"] == 36
    assert all(0.0 < m["frac"] < 1.0 for m in more) and len(more) == 10
    assert set(kernels) == set(d["kernels"]) and abs(sum(k["share_of_kernel_time"] for k in kernels.values()) - 1.0) < 0.01
    # a pass that ran under load for >= 2 s is held against the sustained figure
    r2, _, _ = bench.summarize_kernels(prof, steps, pk, "measured", window_s=2.5)
    assert r2["peak"] == 1433.8
    sr = bench.step_roofline(prof, steps, 8.24, 0.16, pk, 8, 1)
    assert sr["peak_kind"] == "burst" and 4.0 < sr["tflop_per_step_per_gpu"] < 4.8 and 0.25 < sr["frac"] < 0.45
    assert abs(bench.mlp_flops_per_step(8, 8, 128, 3) - 3 * 2 * (22.0e9 + 35.2e9)) < 2e9
    # an empty profile (no conv launches) must not divide by zero
    r0, m0, k0 = bench.summarize_kernels({}, steps, pk, "fallback")
    assert r0["achieved"] == 0.0 and m0 == [] and k0 == {}


def test_exchange_host_helpers_without_a_gpu():
    """Host-side pieces of the multi-GPU path that need no device: 16-byte gradient slots, in-place detection,
    half-machine planning only on forked side streams, NUMA binding as a best-effort no-op."""
    import types

    import torch

    from wcmc_b200 import ddp, streams
    from wcmc_b200 import dropin
    dropin.install()
    from wcmc_b200.engine import bind_host_to_gpu
    assert streams.share() == 1                          # no CUDA, no fork: the whole machine
    assert bind_host_to_gpu(0) is None                   # no NVML device here: nothing changes, nothing raises
    ts = [torch.arange(5.0), torch.arange(3.0).reshape(3, 1), torch.arange(8.0)]
    assert ddp.GradAllReduce._slots(ts) == 8 + 4 + 8
    px = types.SimpleNamespace(buf=torch.full((32,), -1.0))
    views, n = ddp.GradAllReduce._gather(px, ts)
    assert n == 20 and [v.data_ptr() - px.buf.data_ptr() for v in views] == [0, 32, 48]    # 16-byte slots
    assert all(torch.equal(v, t) for v, t in zip(views, ts)) and float(px.buf[5]) == -1.0   # pad words untouched
    # gradients that already live in the buffer (zero_grad(set_to_none=False) + in-place accumulation): no copy
    px.buf[0] = 42.0
    again, _ = ddp.GradAllReduce._gather(px, views)
    assert float(again[0][0]) == 42.0
    sync = ddp.GradAllReduce(transport="nccl")
    assert sync.world == 1 and sync.describe() == {} and sync.peer_transport() == "nccl"
    import pytest
    with pytest.raises(AssertionError):
        ddp.GradAllReduce(transport="carrier-pigeon")

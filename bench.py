"""bench.py -- KPCN+WCMC training throughput on B200 (BASELINE.json: configs[1] at N=1, weak
scaling B=8 per GPU for N>1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = KPCNInterface.preprocess + train_batch on one batch of 8 synthetic 128x128 patches
with 8 spp (PathNet x2 -> p-buffer -> KPCN diffuse/specular -> L1 + path-disentangling loss ->
two backward passes -> clip -> Adam x3).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH, SPP, SIZE, OUTC = 8, 8, 128, 3
METRIC, UNIT = "KPCN+WCMC train patches/s", "patches/s"
WORKLOAD = "configs[1]: KPCN + path embedding network + path disentangling loss, batch 8 of 128x128 patches, 8 spp"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_oracle_models():
    import torch
    from tests._oracle_loader import load_oracle
    o = load_oracle()
    torch.manual_seed(0)
    models = {"dncnn": o.KPCN(35 + OUTC + 1), "backbone_diffuse": o.PathNet(36, outc=OUTC),
              "backbone_specular": o.PathNet(36, outc=OUTC)}
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    return o, models, optims


def cpu_reference_steps(sample_batch, warmup, steps):
    """The reference's CPU path for this step.  sbmc (KPCN / ConvChain / Autoencoder / KernelApply)
    is un-vendored, so the oracle port (pure PyTorch fp32, oneDNN) is what can run; it is driven on
    all host cores.  Returns (patches/s, threads, seconds per step)."""
    import torch
    from wcmc_b200.synth import make_batch
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm is meant to use every host core it can
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, ncpu))
    o, models, optims = build_oracle_models()
    batch = make_batch(batch=sample_batch, spp=SPP, size=SIZE, seed=1234)
    for _ in range(warmup):
        o.ref.kpcn_train_step(models, optims, batch, use_llpm_buf=True, manif_learn=True, w_manif=0.1)
    t0 = time.time()
    for _ in range(steps):
        o.ref.kpcn_train_step(models, optims, batch, use_llpm_buf=True, manif_learn=True, w_manif=0.1)
    dt = (time.time() - t0) / max(steps, 1)
    return sample_batch / dt, torch.get_num_threads(), dt


def run_reference(args, rank):
    if rank != 0:
        return
    sample = 1
    val, threads, dt = cpu_reference_steps(sample, args.warmup, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "oracle port of the reference step on host cores; each step is a "
                                                     "bounded sample of %d patch (same 128x128, 8 spp shapes)" % sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d step(s) of batch %d (128x128, 8 spp) after %d warm-up" %
                                       (args.steps, sample, args.warmup)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bench_720p(args, KPCN, make_batch):
    """BASELINE.json configs[3]: full-frame 1280x720 KPCN denoise (n_in = 34, no PathNet), ms/frame.
    `ms_per_frame`: padded inputs resident in HBM; `e2e_ms_per_frame`: un-padded fp32 frame buffers in
    pinned host memory -> device, replicate pad, both branches, radiance copied back to the host."""
    import torch
    from wcmc_b200 import inference, lib
    torch.manual_seed(0)
    net = KPCN(34).cuda().eval()
    host = {k: v.pin_memory() for k, v in make_batch(batch=1, size=0, height=720, width=1280, seed=7, paths=False,
                                                     llpm_channel=False).items() if k.startswith("kpcn")}
    dev = inference.pad_frame({k: v.cuda() for k, v in host.items()})
    out_host = torch.empty((1, 3, 720, 1280), dtype=torch.float32).pin_memory()

    def resident():
        return inference.denoise_frame(net, dev, padded=True)["radiance"]

    def e2e():
        b = {k: v.cuda(non_blocking=True) for k, v in host.items()}
        out_host.copy_(inference.denoise_frame(net, b)["radiance"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    from wcmc_b200 import streams
    for _ in range(3):
        resident()
    n = max(3, min(args.steps, 10))
    streams_on, streams.ENABLED = streams.ENABLED, False   # per-kernel brackets need one stream
    resident()
    lib.profile_start()
    ms = timed(resident, n)
    prof = lib.profile_stop()
    streams.ENABLED = streams_on
    ms_plain = timed(resident, n)          # the number reported: no per-launch events, two streams
    e2e()
    ms_latency = timed(e2e, n)             # one frame, strictly sequential: copy in, denoise, copy out
    # steady state of a frame SEQUENCE: transfers of neighbouring frames overlap the compute (inference.denoise_stream)
    outs = [torch.empty((1, 3, 720, 1280), dtype=torch.float32).pin_memory() for _ in range(2)]
    inference.denoise_stream(net, [host] * 2, outs)
    torch.cuda.synchronize()
    nseq = max(6, n)
    ms_e2e = timed(lambda: inference.denoise_stream(net, [host] * nseq, outs), 1) / nseq
    n_c, ms_c, fl_c = prof.get("conv2d_k5", (0, 1.0, 0.0))
    n_k, ms_k, by_k = prof.get("kernel_apply_fwd", (0, 1.0, 0.0))
    flops = fl_c / n
    return {"ms_per_frame": round(ms_plain, 3), "e2e_ms_per_frame": round(ms_e2e, 3),
            "e2e_single_frame_latency_ms": round(ms_latency, 3),
            "e2e_note": "e2e_ms_per_frame: %d-frame sequence from pinned host memory, H2D / D2H of neighbouring frames "
                        "overlapped with the compute on copy streams, every byte of every frame moved inside the timed "
                        "region; e2e_single_frame_latency_ms: one frame, copy in -> denoise -> copy out, no overlap" % nseq,
            "h2d_bytes_per_frame": sum(v.numel() * 4 for v in host.values()), "d2h_bytes_per_frame": out_host.numel() * 4,
            "conv_tflop_per_frame": round(flops / 1e12, 3), "conv_tflops": round(fl_c / (ms_c * 1e-3) / 1e12, 1),
            "conv_ms": round(ms_c / n, 3), "kernel_apply_ms": round(ms_k / n, 3),
            "kernel_apply_gbs": round(by_k / (ms_k * 1e-3) / 1e9, 1), "frames": n,
            "workload": "configs[3]: 1280x720, n_in 34, replicate pad 18, both branches + 21x21 kernel-apply, "
                        "whole frame in one pass (no tiling)"}


# (profile entry, label, bound, key of profiles/ncu_traffic.json "per_kernel" for the DRAM traffic)
ROOFLINE_ENTRIES = (
    ("conv2d_k5", "conv_igemm_kernel<1,5> (5x5 100-channel KPCN layers, forward + data gradient, tcgen05 cta_group::2)",
     "tensor", "conv_igemm_kernel<1, 5, 0>"),
    ("conv2d_k3", "conv_igemm_kernel<*,3> (3x3 U-Net layers of PathNet, forward + data gradient)", "tensor",
     "conv_igemm_kernel<0, 3, 0>"),
    ("conv2d_wgrad_all", "conv_wgrad_group_kernel + wgrad_reduce_batch_kernel (weight gradients of all 5x5 and 3x3 layers, one "
     "grouped launch per backward pass)", "tensor", "conv_wgrad_group_kernel"),
    ("conv2d_wgrad_k5", "conv_wgrad_group_kernel (5x5 layers: one launch per KPCN branch)", "tensor", "conv_wgrad_group_kernel"),
    ("conv2d_wgrad_k3", "conv_wgrad_group_kernel (3x3 layers: one launch per PathNet U-Net)", "tensor", "conv_wgrad_group_kernel"),
    ("kernel_apply_fwd", "kernel_apply_fwd_kernel (8 x 92^2 pixels per launch)", "hbm", "kernel_apply_fwd_kernel<3, 21, 16>"),
    ("kernel_apply_bwd", "kernel_apply_bwd_kernel", "hbm", "kernel_apply_bwd_kernel<3, 1, 21, 16>"),
    ("pathnet_embed_fwd", "pathnet_embed_fwd_kernel", "hbm", "pathnet_embed_fwd_kernel"),
    ("pathnet_final_fwd", "pathnet_final_fwd_kernel", "hbm", "pathnet_final_fwd_kernel"),
    ("pathnet_final_bwd", "pathnet_final_bwd_kernel + slab_reduce_kernel", "hbm", "pathnet_final_bwd_kernel<16>"),
    ("pathnet_embed_bwd", "pathnet_embed_bwd_kernel + slab_reduce_kernel", "hbm", "pathnet_embed_bwd_kernel"),
)


def mlp_flops_per_step(batch, spp, size, outc):
    """Algorithmic FLOPs of the two PathNets' 1x1 MLPs (networks.py:18-24), forward + backward (x3), which the
    fused K6-K9 kernels report in bytes: per sample-pixel 2*(36*64 + 64*64 + 64*64) + 2*(128*128 + 128*outc)."""
    per = 2.0 * (36 * 64 + 64 * 64 + 64 * 64) + 2.0 * (128 * 128 + 128 * outc)
    return 2 * 3 * per * batch * spp * size * size


def bench_preprocess(pk):
    """GPU preprocessing of a raw 1280x720, 4-spp OptaGen sample buffer (H,W,S,104) -> the 44-channel
    KPCN buffer and the 37-channel path descriptors, against the HBM roofline.  Algorithmic bytes = SURVEY 8(f)'s
    per-unit figure: the whole 416-byte raw row of every sample in (the 13 / 37 channels a kernel needs are scattered
    over the row; ncu: the statistics pass pulls 1.06 GB of the 1.53 GB buffer from DRAM, the descriptor pass 1.09 GB),
    176 B per pixel + the 72-byte statistics workspace written and re-read (kpcn), 148 B per sample out (llpm)."""
    import torch
    from wcmc_b200 import preprocess
    h, w, s = 720, 1280, 4
    g = torch.Generator(device="cuda").manual_seed(3)
    raw = torch.rand(h, w, s, 104, device="cuda", generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn):
        fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[2]
    ms_k = timed(lambda: preprocess.preprocess_kpcn(raw))
    ms_l = timed(lambda: preprocess.preprocess_llpm(raw))
    by_k = h * w * (s * 416.0 + 176.0 + 2 * 72.0)
    by_l = h * w * s * (416.0 + 148.0)
    return {"workload": "raw (720,1280,4,104) fp32 -> (720,1280,44) + (720,1280,4,37), L2 flushed between repetitions",
            "kpcn_ms": round(ms_k, 4), "kpcn_gbs": round(by_k / ms_k / 1e6, 1), "kpcn_frac_of_hbm": round(by_k / ms_k / 1e6 / pk["hbm_gbs"], 4),
            "llpm_ms": round(ms_l, 4), "llpm_gbs": round(by_l / ms_l / 1e6, 1), "llpm_frac_of_hbm": round(by_l / ms_l / 1e6 / pk["hbm_gbs"], 4)}


def summarize_kernels(prof, steps, pk, pk_src, window_s=0.0):
    """prof = {C-ABI call: (calls, total ms, total algorithmic work)} of the eager timed pass (lib.profile_stop()).
    -> (roofline of the dominant kernel, roofline_more, per-call table).  Pure host logic (tests/test_host_logic.py).

    Dominant kernel = the ROOFLINE_ENTRIES row with the largest MEASURED time in this very run (round 1 hard-coded
    conv2d_k5; the weight-gradient kernel was in fact larger).  `achieved` = algorithmic work / CUDA-event time;
    work of a data-gradient launch is the forward layer's 2*N*Ho*Wo*k^2*Cin*Cout (lib.conv2d(alg_hw=...)), not the
    launch's larger padded extent.  Denominator: the per-kernel pass launches every kernel eagerly from Python with
    an event bracket each -- the GPU idles between launches and is not power-limited -- so tensor-bound kernels are
    held against the BURST cuBLAS figure (`bf16_tflops`) unless the pass ran under load for >= 2 s (`window_s`);
    `frac_sustained` is given beside it."""
    total_ms = sum(v[1] for v in prof.values()) or 1.0
    sustained = window_s >= 2.0
    t_peak = pk["bf16_tflops_sustained"] if sustained else pk["bf16_tflops"]
    t_note = "sustained" if sustained else "burst: eager per-kernel pass, GPU not power-limited"
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("per_kernel", {})
    except Exception:  # noqa: BLE001
        traffic_tab = {}
    rows = []
    # one kernel serves both filter sizes: the merged row competes for "dominant", the per-size rows are detail
    parts = [prof[k] for k in ("conv2d_wgrad_k5", "conv2d_wgrad_k3") if k in prof]
    look = dict(prof)
    if parts:
        look["conv2d_wgrad_all"] = tuple(sum(x[i] for x in parts) for i in range(3))
    for name, label, bound, ncu_key in ROOFLINE_ENTRIES:
        cnt, ms, work = look.get(name, (0, 0.0, 0.0))
        if not (cnt and ms > 0 and work > 0):
            continue
        pkv = t_peak if bound == "tensor" else pk["hbm_gbs"]
        ach = work / (ms * 1e-3) / (1e12 if bound == "tensor" else 1e9)
        row = {"kernel": label, "bound": bound, "achieved": round(ach, 1), "peak": pkv,
               "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": round(ach / pkv, 4),
               "ms_per_step": round(ms / steps, 4), "launches_per_step": cnt // steps,
               "share_of_step_kernel_time": round(ms / total_ms, 4),
               "traffic": (traffic_tab.get(ncu_key) or {}).get("dram_bytes_per_launch"),
               "peak_source": "%s (%s)" % (pk_src, t_note if bound == "tensor" else "copy bandwidth")}
        if bound == "tensor":
            row["frac_sustained"] = round(ach / pk["bf16_tflops_sustained"], 4)
            row["algorithmic_tflop_per_launch"] = round(work / cnt / 1e12, 5)
        else:
            row["algorithmic_mb_per_launch"] = round(work / cnt / 1e6, 2)
        rows.append((ms if name not in ("conv2d_wgrad_k5", "conv2d_wgrad_k3") else 0.0, row))
    kernels = {k: {"calls_per_step": v[0] // steps, "ms_per_step": round(v[1] / steps, 4),
                   "share_of_kernel_time": round(v[1] / total_ms, 4)} for k, v in sorted(prof.items())}
    units = {"tflops": 1e12, "gbs": 1e9}
    for name, v in prof.items():
        key = "tflops" if name.startswith("conv2d") else ("gbs" if v[2] > 0 else None)
        if key and v[1] > 0:
            kernels[name][key] = round(v[2] / (v[1] * 1e-3) / units[key], 1)
    if not rows:
        return ({"kernel": None, "bound": "tensor", "achieved": 0.0, "peak": t_peak, "unit": "TFLOP/s", "frac": 0.0,
                 "traffic": None, "peak_source": pk_src}, [], kernels)
    rows.sort(key=lambda r: -r[0])
    roofline = dict(rows[0][1], chosen="largest measured time among the kernels of this run")
    return roofline, [r for _, r in rows[1:]], kernels


def step_roofline(prof, steps, ms_step, window_s, pk, batch, world):
    """Whole-step tensor-pipe fraction: algorithmic FLOPs of every contraction of one step (convolutions forward /
    data gradient / weight gradient from the per-call work, the PathNet MLPs analytically) / the graph-replayed step
    time; held against the sustained cuBLAS figure when the timed region ran >= 2 s, else the burst one."""
    conv = sum(v[2] for k, v in prof.items() if k.startswith("conv2d")) / max(steps, 1)
    flop = conv + mlp_flops_per_step(batch, SPP, SIZE, OUTC)
    sustained = window_s >= 2.0
    peak = pk["bf16_tflops_sustained"] if sustained else pk["bf16_tflops"]
    ach = flop / (ms_step * 1e-3) / 1e12
    return {"tflop_per_step_per_gpu": round(flop / 1e12, 3), "achieved": round(ach, 1), "unit": "TFLOP/s per GPU",
            "peak": peak, "frac": round(ach / peak, 4), "frac_burst": round(ach / pk["bf16_tflops"], 4),
            "frac_sustained": round(ach / pk["bf16_tflops_sustained"], 4),
            "peak_kind": "sustained" if sustained else "burst", "timed_window_s": round(window_s, 3)}


def _leave(world, graphed=None):
    """End of a rank under torchrun: the captured step graph (it holds NCCL kernels of the in-graph gradient
    all-reduce) is released BEFORE the communicator is destroyed -- destroying the process group with the graph still
    alive is what hung for minutes in round 1 (profiles/r01final_multi2.txt).  A watchdog bounds the teardown anyway:
    every collective of the run has completed by now (the timed regions end with a barrier + synchronize)."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Timer(30.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    if graphed is not None:
        graphed.release()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    t.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-720p", action="store_true", help="skip the 1280x720 full-frame denoise measurement")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python (no CUDA graph)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="BASELINE.json configs[2] as written: a FIXED global batch (64) split over the ranks "
                         "(32 / 16 / 8 per GPU at N = 2 / 4 / 8), reported as strong scaling.  Default 0 = weak scaling, "
                         "8 patches per GPU (configs[1] on every rank)")
    ap.add_argument("--per-gpu-batch", type=int, default=0, help="patches per GPU (overrides the default 8)")
    ap.add_argument("--kernel-pass-steps", type=int, default=0,
                    help="steps of the eager per-kernel timing pass (default: --steps)")
    ap.add_argument("--perm-rng", default="device", choices=["device", "cpu"],
                    help="pairing permutations of the path-disentangling loss: torch.randperm on the GPU "
                         "(default) or the reference's CPU default-generator contract (13 ms of host time per "
                         "541,696-element permutation, four per step)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    batch_per_gpu = BATCH
    scaling = "weak"
    if args.global_batch:
        assert args.global_batch % world == 0, "--global-batch must divide by the number of ranks"
        batch_per_gpu, scaling = args.global_batch // world, "strong"
    elif args.per_gpu_batch:
        batch_per_gpu = args.per_gpu_batch

    import torch
    import torch.distributed as dist
    from wcmc_b200 import ddp, dropin, lib
    from wcmc_b200.synth import make_batch
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dropin.install()
    lib.init(local)
    if world > 1 and os.environ.get("WCMC_NUMA_BIND", "1") != "0":
        from wcmc_b200.engine import bind_host_to_gpu
        numa_cpus = bind_host_to_gpu(local)      # before any pinned allocation: first touch on the GPU's node
    from sbmc import KPCN
    from support.interfaces import KPCNInterface
    from support.losses import FeatureMSE, RelativeMSE
    from support.networks import PathNet

    torch.manual_seed(0)
    models = {"dncnn": KPCN(35 + OUTC + 1).cuda(), "backbone_diffuse": PathNet(ic=36, outc=OUTC).cuda(),
              "backbone_specular": PathNet(ic=36, outc=OUTC).cuda()}
    ddp.broadcast_parameters(models)
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        l_manif = FeatureMSE(non_local=True, rng=args.perm_rng)
    loss_funcs = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
                  "l_test": RelativeMSE(), "l_manif": l_manif}
    itf = KPCNInterface(models, optims, loss_funcs, types.SimpleNamespace(model_name="bench"), use_llpm_buf=True,
                        manif_learn=True, w_manif=0.1, train_branches=True, disentanglement_option="m11r11")
    sync = ddp.GradAllReduce(overlap=os.environ.get("WCMC_DDP_OVERLAP", "1") != "0")
    if world > 1:
        itf.grad_sync = sync
    host = {k: v.pin_memory() for k, v in make_batch(batch=batch_per_gpu, spp=SPP, size=SIZE, seed=1234 + rank).items()}
    dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    itf.to_train_mode()

    use_graph = not args.no_graph   # rng="cpu" replays too: staged permutations (support.losses.PermStage)
    graphed = None
    if use_graph:
        from wcmc_b200.engine import GraphedTrainStep
        graphed = GraphedTrainStep(itf, dev)

    # The synthetic batch repeats, so a long run over-fits it and, at the reference's lr 1e-4, eventually diverges -- in
    # fp32 as well (profiles/r02_long_run_800_steps.txt: backend and fp32 oracle side by side).  A run of many steps
    # therefore puts the weights / Adam moments of the end of the warm-up back every RESET steps (eager per-kernel pass included) (three _foreach_copy_ launches inside
    # the timed region: extra work, nothing skipped); runs of <= RESET steps never see it.
    RESET = 100
    state = {"n": 0, "live": None, "init": None}

    def snapshot():
        live = [p_ for m in models.values() for p_ in m.parameters()]
        for o in optims.values():
            for st in o.state.values():
                live += [st["exp_avg"], st["exp_avg_sq"]]
        state["live"], state["init"] = live, [t.detach().clone() for t in live]
        steps = [st["step"] for o in optims.values() for st in o.state.values()]
        state["steps"], state["t0"] = steps, (float(steps[0]) if steps else 0.0)

    def maybe_reset():
        state["n"] += 1
        if state["n"] % RESET == 0 and state["init"] is not None:
            with torch.no_grad():
                torch._foreach_copy_(state["live"], state["init"])
                # ... and Adam's step count with them (bias correction with a later count on early moments would
                # take several times larger steps -- enough to overflow the 16-bit activations now and then)
                fa = itf._fused()
                if fa is not None:
                    fa.set_step(state["t0"])
                else:
                    for t in state["steps"]:
                        t.fill_(state["t0"])

    def step(batch):
        maybe_reset()
        if use_graph:
            graphed(batch)
        else:
            itf.preprocess(batch)
            itf.train_batch(batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, tail=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if tail is not None:
            tail()          # whatever the loop still owes (the last step's deferred flag / loss read-back)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(args.warmup):
        step(dev)
    snapshot()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # device-resident steps: the batch IS the graph's static input buffers (no staging copy); the finite flag of
    # step i is looked at after step i+1 has been enqueued (GraphedTrainStep(check="deferred")), the last one in `tail`
    resident = graphed.static if use_graph else dev
    ms_step = timed(lambda: step(resident), args.steps, tail=(graphed.finish if use_graph else None))
    clk = clocks.stop() if rank == 0 else None
    window_s = ms_step * args.steps * 1e-3

    # live per-kernel device time (CUDA events on the launching stream) over a second timed pass.  The
    # kernels are the same ones the graph replays; this pass launches them eagerly so that each launch
    # can be bracketed by events.
    def eager_step():
        maybe_reset()
        itf.preprocess(dev)
        itf.train_batch(dev)
    # one stream for this pass: with the diffuse / specular halves on two streams (wcmc_b200/streams.py) the event
    # brackets of concurrent kernels would overlap and every per-kernel time would be inflated
    from wcmc_b200 import streams
    streams_on, streams.ENABLED = streams.ENABLED, False
    eager_step()
    n0 = lib.LAUNCHES["count"]
    ksteps = args.kernel_pass_steps or args.steps
    lib.profile_start()
    timed(eager_step, ksteps)
    prof = lib.profile_stop()
    launches = (lib.LAUNCHES["count"] - n0) // ksteps
    streams.ENABLED = streams_on

    # end to end: every step's batch comes from pinned host memory (double-buffered copy stream, so the
    # PCIe transfer of step i+1 overlaps the compute of step i) and the step's loss is read back
    from wcmc_b200.engine import DevicePrefetcher

    def host_batches():
        while True:
            yield host

    pf = DevicePrefetcher(host_batches())

    # ... and every step's loss is read back: the running total goes to pinned host memory right behind the step and
    # is looked at one step later (the host stays one step ahead of the GPU, as in the device-resident loop)
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_seen = {"n": 0, "pending": None, "last": 0.0}

    def read_pending():
        pend, loss_seen["pending"] = loss_seen["pending"], None
        if pend is not None:
            pend[0].synchronize()
            loss_seen["last"] = float(loss_host[pend[1]])

    def e2e_step():
        batch = next(pf)
        step(batch)
        pf.release()
        slot = loss_seen["n"] & 1
        loss_seen["n"] += 1
        loss_host[slot].copy_(itf.m_losses["m_l_total"].reshape(1), non_blocking=True)   # device -> host, 4 bytes
        ev = torch.cuda.Event()
        ev.record()
        pend, loss_seen["pending"] = loss_seen["pending"], (ev, slot)
        if pend is not None:
            pend[0].synchronize()
            loss_seen["last"] = float(loss_host[pend[1]])

    def e2e_tail():
        read_pending()
        if use_graph:
            graphed.finish()

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps, tail=e2e_tail)
    assert loss_seen["last"] == loss_seen["last"], "non-finite running loss"

    if rank != 0:
        _leave(world, graphed)
        return
    pk, pk_src = peaks()
    frame = bench_720p(args, KPCN, make_batch) if (world == 1 and not args.no_720p) else None
    roofline, more, kernels = summarize_kernels(prof, ksteps, pk, pk_src)
    workload = WORKLOAD if batch_per_gpu == BATCH else WORKLOAD.replace("batch 8", "batch %d per GPU" % batch_per_gpu)
    if args.global_batch:
        workload = ("configs[2]: KPCN+WCMC data-parallel training, global batch %d of 128x128 patches, 8 spp, split over "
                    "%d GPU(s)" % (args.global_batch, world))
    line = {
        "metric": METRIC, "value": batch_per_gpu * world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": batch_per_gpu * world, "per_gpu_batch": batch_per_gpu, "spp": SPP,
                   "patch": SIZE, "pnet_out_size": OUTC, "parallelism": "dp%d" % world,
                   "l2": "inputs_exceed_l2 (%.0f MB of step inputs + %.0f MB of saved activations > 126 MB L2)"
                         % (h2d_bytes / 1e6, 700.0),
                   "precision": "fp16 operands (loss-scaled gradients), fp32 accumulate / master weights / losses",
                   "perm_rng": args.perm_rng, "cuda_graph": bool(use_graph), "branch_streams": bool(streams.ENABLED),
                   "step_sync": ("finite flag of step i read after step i+1 is enqueued (a non-finite step skips its update "
                                 "on the device); the last flag is read inside the timed region; the device-resident loop "
                                 "feeds the graph's static input buffers in place" if use_graph else "one host sync per step"),
                   "e2e_readback": "running loss copied to pinned host memory behind every step, read one step later; "
                                   "the last one inside the timed region",
                   "grad_exchange": (None if world == 1 else "%s in the step graph%s" % (
                       {"multimem": "own two-shot kernel over the NVSwitch multicast mapping (multimem.ld_reduce / st)",
                        "peer": "own two-shot kernel over NVLink peer loads / stores",
                        "nccl": "nccl all-reduce"}[sync.peer_transport()],
                       "; dncnn's half overlapped with the path-embedding networks' backward"
                       if sync.early is not None else ", after both backward passes")),
                   "grad_exchange_channels": None if world == 1 else sync.describe(),
                   "host_cores_bound_to_gpu_numa_node": None if numa_cpus is None else len(numa_cpus)},
        "e2e": {"value": batch_per_gpu * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
        "roofline_step": step_roofline(prof, ksteps, ms_step, window_s, pk, batch_per_gpu, world),
        "roofline_more": more,
        "kernels": kernels,
    }
    if frame is not None:
        line["denoise_720p"] = frame
        line["preprocess_n3"] = bench_preprocess(pk)
    if world == 1 and not args.no_cpu_baseline:
        val, threads, dt = cpu_reference_steps(2, 1, 2)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "2 timed steps of batch 2 (same 128x128, 8 spp patches) after 1 warm-up; "
                                          "%.1f s/step" % dt}
    print(json.dumps(line), flush=True)
    _leave(world, graphed)


if __name__ == "__main__":
    main()

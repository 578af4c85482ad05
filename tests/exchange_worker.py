"""Worker of tests/test_gpu_nccl.py::test_peer_exchange_* (run under torchrun, one rank per GPU; also with one rank):
the NVSwitch gradient-exchange kernel (`wcmc_grad_exchange` through wcmc_b200.ddp.PeerExchange) against NCCL's
all-reduce on the same data -- every transport the machine offers, ragged sizes, both channels in flight at once,
eager and replayed from a CUDA graph.  With `--bench` it also times the transports on the step's real message sizes
(46.9 MB and its two halves) over a sweep of grid sizes.  Writes one JSON record per rank into argv[1]."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    out_dir = sys.argv[1]
    bench = "--bench" in sys.argv
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from wcmc_b200 import ddp, lib
    lib.init(local)
    rec = {"rank": rank, "world": world, "transports": {}}
    sizes = [4, 1000, 4099, 65536 + 4, 1 << 20, 5_900_000 + 1]
    cap = max(sizes) + 8
    for want in ("multimem", "peer"):
        try:
            px = [ddp.PeerExchange(cap, multicast=(want == "multimem")) for _ in range(2)]
        except Exception as e:  # noqa: BLE001
            rec["transports"][want] = {"unavailable": repr(e)[:300]}
            continue
        if px[0].transport != want:
            rec["transports"][want] = {"unavailable": "no multicast mapping (multicast_ptr == 0)"}
            continue
        t = {"max_abs_err": 0.0, "bitwise_equal_to_nccl": True, "replicas_identical": True}
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        for it, n in enumerate(sizes * 2):
            x = torch.randn(n, device="cuda", generator=g) * (1.0 + it)
            want_sum = x.clone()
            dist.all_reduce(want_sum)
            want_sum /= world
            ch = it & 1
            px[ch].buf[:n].copy_(x)
            got = px[ch].all_reduce_(n, 1.0 / world, channel=ch).clone()
            t["max_abs_err"] = max(t["max_abs_err"], (got - want_sum).abs().max().item())
            t["bitwise_equal_to_nccl"] &= bool(torch.equal(got, want_sum))
            mine = got.clone()
            dist.broadcast(mine, 0)
            t["replicas_identical"] &= bool(torch.equal(mine, got))
        # both channels in flight at once on two streams, captured in a graph and replayed (the step's pattern)
        n0, n1 = 1 << 20, (1 << 20) + 4
        x0 = torch.randn(n0, device="cuda", generator=g)
        x1 = torch.randn(n1, device="cuda", generator=g)
        ref0, ref1 = x0.clone(), x1.clone()
        dist.all_reduce(ref0)
        dist.all_reduce(ref1)
        side = torch.cuda.Stream()
        out0, out1 = torch.empty_like(x0), torch.empty_like(x1)

        def both():
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                px[0].buf[:n0].copy_(x0)
                out0.copy_(px[0].all_reduce_(n0, 1.0, channel=0))
            px[1].buf[:n1].copy_(x1)
            out1.copy_(px[1].all_reduce_(n1, 1.0, channel=1))
            cur.wait_stream(side)

        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            both()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            both()
        worst = 0.0
        for _ in range(20):
            out0.zero_()
            out1.zero_()
            graph.replay()
            worst = max(worst, (out0 - ref0).abs().max().item(), (out1 - ref1).abs().max().item())
        torch.cuda.synchronize()
        t["graph_two_channels_max_abs_err"] = worst
        del graph
        rec["transports"][want] = t

        if bench:
            t["bench"] = {}
            big = ddp.PeerExchange(11_730_000, multicast=(want == "multimem"))
            for n in (11_727_112, 5_900_000):
                for blocks in (8, 16, 32, 64, 128):
                    lib.tuning_set("exchange_blocks", blocks)
                    for _ in range(5):
                        big.all_reduce_(n, 1.0 / world)
                    torch.cuda.synchronize()
                    dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(50):
                        big.all_reduce_(n, 1.0 / world)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1000 / 50
                    t["bench"]["%d_floats_%d_blocks" % (n, blocks)] = {
                        "us": round(us, 1), "algbw_GBs": round(n * 4 / us / 1e3, 1),
                        "busbw_GBs": round(n * 4 / us / 1e3 * 2 * (world - 1) / world, 1)}
            lib.tuning_set("exchange_blocks", 32)
            del big
    if bench:
        rec["nccl"] = {}
        for n in (11_727_112, 5_900_000):
            x = torch.randn(n, device="cuda")
            for _ in range(5):
                dist.all_reduce(x)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                dist.all_reduce(x)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1000 / 50
            rec["nccl"]["%d_floats" % n] = {"us": round(us, 1), "algbw_GBs": round(n * 4 / us / 1e3, 1),
                                            "busbw_GBs": round(n * 4 / us / 1e3 * 2 * (world - 1) / world, 1)}
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(rec, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

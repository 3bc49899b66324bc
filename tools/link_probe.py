#!/usr/bin/env python
"""Bare host<->device copy ceiling under the end-to-end number.

    python tools/link_probe.py [--images 1024] [--sub 32] [--steps 5]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... tools/link_probe.py --gpus N

Moves exactly the plane bytes BatchCodec.roundtrip_host moves per step -- n_sub host->device copies on one CUDA
stream and n_sub device->host copies on another, pinned host memory, no kernel in between -- on every rank at
once (the ranks share the host side of the links), and reports the aggregate GB/s per direction: h2d alone, d2h
alone, and both together (the pattern of the pipelined round trip).  `bench.py` calls `probe()` right after its
end-to-end leg, with the same pinned allocation policy and core binding, and puts the result into the bench line
as e2e.link_ceiling_gbs / e2e.frac_of_link.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def probe(device, n_bytes, n_sub, steps, barrier, max_over_ranks, h_src=None, h_dst=None, d_a=None, d_b=None):
    """Returns {"h2d_ms", "d2h_ms", "both_ms"} per step (max over ranks) for n_bytes each way.
    Buffers may be passed in (bench.py reuses the ones of its end-to-end leg)."""
    import torch
    with torch.cuda.device(device):
        if h_src is None:
            h_src = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
            h_src.zero_()
        if h_dst is None:
            h_dst = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
            h_dst.zero_()
        if d_a is None:
            d_a = torch.empty(n_bytes, dtype=torch.uint8, device=device)
        if d_b is None:
            d_b = torch.zeros(n_bytes, dtype=torch.uint8, device=device)
        h_src, h_dst, d_a, d_b = (t.view(-1)[:n_bytes] for t in (h_src.view(torch.uint8), h_dst.view(torch.uint8),
                                                                 d_a.view(torch.uint8), d_b.view(torch.uint8)))
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        bounds = [n_bytes * j // n_sub for j in range(n_sub + 1)]

        def run(do_in, do_out):
            cur = torch.cuda.current_stream()
            sa.wait_stream(cur)
            sb.wait_stream(cur)
            for j in range(n_sub):
                lo, hi = bounds[j], bounds[j + 1]
                if do_in:
                    with torch.cuda.stream(sa):
                        d_a[lo:hi].copy_(h_src[lo:hi], non_blocking=True)
                if do_out:
                    with torch.cuda.stream(sb):
                        h_dst[lo:hi].copy_(d_b[lo:hi], non_blocking=True)
            cur.wait_stream(sa)
            cur.wait_stream(sb)

        out = {}
        for name, di, do in (("h2d_ms", True, False), ("d2h_ms", False, True), ("both_ms", True, True)):
            run(di, do)                                    # warm-up
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                run(di, do)
            e1.record()
            barrier()
            out[name] = max_over_ranks(e0.elapsed_time(e1) / steps)
    return out


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--images", type=int, default=1024, help="1920x1080 3-band images in the whole job")
    ap.add_argument("--sub", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-bind", action="store_true", help="do not pin the rank to the cores next to its GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cores = None
    if not args.no_bind:
        import jpeg_b200 as jb
        cores = jb.sharding.bind_host_to_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_img = args.images // world
    n_bytes = n_img * 3 * 1080 * 1920
    res = probe(device, n_bytes, max(1, min(args.sub, n_img // 8)), args.steps, barrier, max_over_ranks)
    if rank == 0:
        total = n_bytes * world
        line = {"probe": "pinned host<->device copies, bytes of one BatchCodec.roundtrip_host step", "n_gpus": world,
                "bytes_per_direction": total, "host_cores_bound_to_gpu": len(cores) if cores else None,
                "h2d_gbs": total / res["h2d_ms"] / 1e6, "d2h_gbs": total / res["d2h_ms"] / 1e6,
                "both_gbs_per_direction": total / res["both_ms"] / 1e6, **res}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

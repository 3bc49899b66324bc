"""BASELINE.json config 5 on N GPUs: one 16384 x 16384 three-band image, DFT, table quantisation, split into block-row
bands.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_config5_bands.py [--side 16384] [--steps 10] [--verify]

Every rank builds its band of the synthetic image on its own GPU, compresses and decompresses it; no collective is on
the data path (NCCL: the timing barrier, the max over ranks and -- with --verify -- the gather of the sub-streams).
--verify: rank 0 also compresses the whole image on its GPU and checks that, per colour plane, the sub-streams
concatenated in rank order are byte-identical to the whole-image stream, and that the bands decode to the whole
decode.  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def band_planes(side, r0, r1, device):
    """Rows [r0, r1) of the three planes of the synthetic image (smooth field + deterministic hash noise, so that any
    band layout produces the same pixels)."""
    ys = torch.arange(r0, r1, device=device, dtype=torch.float32).view(-1, 1)
    xs = torch.arange(side, device=device, dtype=torch.float32).view(1, -1)
    out = torch.empty((3, r1 - r0, side), dtype=torch.uint8, device=device)
    yi = torch.arange(r0, r1, device=device, dtype=torch.int64).view(-1, 1)
    xi = torch.arange(side, device=device, dtype=torch.int64).view(1, -1)
    for b in range(3):
        v = 127.0 + 100.0 * torch.sin(xs / 37.0 + 0.7 * b) * torch.cos(ys / 53.0 + 0.7 * b)
        h = (yi * 1103515245 + xi * 12345 + b * 977) & 0x7FFFFFFF
        h = (h * 1664525 + 1013904223) & 0x7FFFFFFF
        noise = ((h >> 8) % 33).to(torch.float32) - 16.0
        out[b] = (v + noise).round().clamp_(0, 255).to(torch.uint8)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--verify", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    import jpeg_b200 as jb

    side, bs, d = args.side, 4, 8
    bands = jb.sharding.block_row_bands(side, bs, d, world)
    r0, r1 = bands[rank]
    q = jb.QuantizationMethod("qtable")
    planes = band_planes(side, r0, r1, device)
    cfg = jb.Configuration(width=side, height=r1 - r0, block_size=bs, dct_size=d, transform="DFT", quantization=q)
    bc = jb.BatchCodec(cfg, 3, device=device, flags=jb._lib.JB_FLAG_PDL)
    bc.d_planes.copy_(planes)
    comp = bc.compress_device()
    total = comp.total_bytes()
    out, status = bc.decompress_device(comp, total)
    jb.check_status(status)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        c = bc.compress_device()
        bc.decompress_device(c, total)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        ev[s][0].record()
        c = bc.compress_device()
        ev[s][1].record()
        bc.decompress_device(c, total)
        ev[s][2].record()
    barrier()
    t_c = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    t_d = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    # the same calls replayed as CUDA graphs (BatchCodec.capture_graphs): at ~100 MB per rank the direct calls are
    # bound by their launches from Python
    g_c, g_d, comp_g, status_g = bc.capture_graphs()
    for _ in range(args.warmup):
        g_c.replay(); g_d.replay()
    gev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        gev[s][0].record(); g_c.replay(); gev[s][1].record(); g_d.replay(); gev[s][2].record()
    barrier()
    jb.check_status(status_g)
    tg_c = sum(e[0].elapsed_time(e[1]) for e in gev) / args.steps
    tg_d = sum(e[1].elapsed_time(e[2]) for e in gev) / args.steps
    t = torch.tensor([t_c, t_d, float(total), tg_c, tg_d], dtype=torch.float64, device=device)
    tot = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_c, t_d, tg_c, tg_d = float(t[0]), float(t[1]), float(t[3]), float(t[4])
    stream_total = int(tot[2].item())

    verified = None
    if args.verify:
        streams = comp.to_bytes_list()
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(streams, gathered, dst=0)
            dec_parts = [torch.empty((3, b1 - b0, side), dtype=torch.uint8, device=device) for b0, b1 in bands] if rank == 0 else None
            # decoded bands travel through host memory (gather_object): verification only, far off the timed path
            dec_host = [None] * world if rank == 0 else None
            dist.gather_object(out.cpu().numpy(), dec_host, dst=0)
        else:
            gathered, dec_host = [streams], [out.cpu().numpy()]
        if rank == 0:
            import numpy as np
            whole_cfg = jb.Configuration(width=side, height=side, block_size=bs, dct_size=d, transform="DFT", quantization=q)
            del bc
            torch.cuda.empty_cache()
            whole = band_planes(side, 0, side, device)
            wc = jb.compress_planes(whole, whole_cfg)
            whole_streams = wc.to_bytes_list()
            cat = jb.sharding.concat_band_streams(gathered)
            lens = wc.offsets[1:] - wc.offsets[:-1]
            wdec, st = jb.decompress_planes(wc.data, wc.offsets[:-1], lens, whole_cfg, 3, in_bytes=wc.total_bytes())
            jb.check_status(st)
            wdec = wdec.cpu().numpy()
            verified = {"streams_identical_to_whole_image": all(a == b for a, b in zip(cat, whole_streams)) and len(cat) == 3,
                        "bands_decode_to_whole_decode": bool(np.array_equal(np.concatenate(dec_host, axis=1), wdec)),
                        "stream_bytes_whole": sum(len(x) for x in whole_streams)}
    if rank == 0:
        mp = side * side / 1e6
        a = 3.0 * side * side + stream_total
        line = {"config": "5: %dx%d DFT qtable, block-row bands over %d GPU(s)" % (side, side, world), "n_gpus": world,
                "bands_rows": [b1 - b0 for b0, b1 in bands], "ms_compress": t_c, "ms_decompress": t_d,
                "ms_compress_graph": tg_c, "ms_decompress_graph": tg_d,
                "compress_graph_GBps_aggregate": a / (tg_c * 1e-3) / 1e9, "decompress_graph_GBps_aggregate": a / (tg_d * 1e-3) / 1e9,
                "compress_MPps": mp / (t_c * 1e-3), "decompress_MPps": mp / (t_d * 1e-3),
                "compress_GBps_aggregate": a / (t_c * 1e-3) / 1e9, "decompress_GBps_aggregate": a / (t_d * 1e-3) / 1e9,
                "stream_bytes": stream_total, "timing": "CUDA events per rank, max over ranks", "verify": verified}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

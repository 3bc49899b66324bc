#!/usr/bin/env python
"""Device-timed compress / decompress throughput for the five BASELINE.json configurations on one GPU
(bench.py carries the headline config; this script is the per-config table kept under profiles/).

Each config: synthetic planes resident in HBM, W warm-up + K timed passes (CUDA events), MP/s of source
pixels, achieved algorithmic GB/s vs the measured HBM peak, and a parity spot check against the oracle
on the first plane(s).  Small configs (0.8 MB / 25 MB) are launch-latency bound: they are timed both as
a single image and as a batch that exceeds L2.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch

import jpeg_b200 as jb
from oracle import ref_port as rp
from parity import check_pixels, check_quantised


def synth(n_planes, h, w, device, seed):
    out = torch.empty((n_planes, h, w), dtype=torch.uint8, device=device)
    ys = torch.arange(h, device=device, dtype=torch.float32).view(h, 1)
    xs = torch.arange(w, device=device, dtype=torch.float32).view(1, w)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    for i in range(n_planes):
        ph = 0.7 * (i % 3) + 0.01 * (i // 3)
        v = 127.0 + 100.0 * torch.sin(xs / 37.0 + ph) * torch.cos(ys / 53.0 + ph)
        v = v + 8.0 * torch.randn((h, w), device=device, generator=gen)
        out[i] = v.round().clamp_(0, 255).to(torch.uint8)
    return out


def run(name, h, w, bs, d, transform, qname, qparam, n_images, steps=10, warmup=3, check_planes=1):
    dev = torch.device("cuda", 0)
    kw = {"keep": qparam} if qname == "discard" else {"divisor": qparam} if qname == "divide" else {}
    cfg = jb.Configuration(width=w, height=h, block_size=bs, dct_size=d, transform=transform,
                           quantization=jb.QuantizationMethod(qname, **kw))
    n_planes = 3 * n_images
    bc = jb.BatchCodec(cfg, n_planes, device=dev)
    bc.d_planes.copy_(synth(n_planes, h, w, dev, 1234))
    comp = bc.compress_device()
    total = comp.total_bytes()
    out, status = bc.decompress_device(comp, total)
    jb.check_status(status)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for _ in range(warmup):
        c = bc.compress_device(); bc.decompress_device(c, total)
    torch.cuda.synchronize()
    for s in range(steps):
        ev[s][0].record(); c = bc.compress_device(); ev[s][1].record()
        bc.decompress_device(c, total); ev[s][2].record()
    torch.cuda.synchronize()
    t_c = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    t_d = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    # the same two calls replayed as CUDA graphs (BatchCodec.capture_graphs): what a latency-bound caller does
    g_c, g_d, comp_g, status_g = bc.capture_graphs()
    for _ in range(warmup):
        g_c.replay(); g_d.replay()
    gev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    torch.cuda.synchronize()
    for s in range(steps):
        gev[s][0].record(); g_c.replay(); gev[s][1].record(); g_d.replay(); gev[s][2].record()
    torch.cuda.synchronize()
    jb.check_status(status_g)
    assert comp_g.total_bytes() == total and torch.equal(bc.d_decoded, out)
    tg_c = sorted(e[0].elapsed_time(e[1]) for e in gev)[steps // 2]
    tg_d = sorted(e[1].elapsed_time(e[2]) for e in gev)[steps // 2]
    mp = n_images * h * w / 1e6
    peak = 6550.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    a = n_planes * h * w + total
    # parity spot check
    ocfg = rp.OracleConfig(w, h, bs, d, transform, qname, qparam)
    planes = bc.d_planes[:check_planes].cpu().numpy()
    offs = comp.host_offsets()
    blob = comp.data[:int(offs[check_planes])].cpu().numpy().tobytes()
    coeffs = jb.stages.forward_coefficients(planes, cfg)
    exact = ties = 0
    maxerr = 0
    for i in range(check_planes):
        p64 = planes[i].astype(np.int64)
        zz = rp.quantised_zigzag(p64, ocfg)
        ties += check_quantised(coeffs[i], zz, rp.prerounding_zigzag(p64, ocfg), what=name)
        s = blob[int(offs[i]):int(offs[i + 1])]
        assert s == rp.pack_blocks(coeffs[i].reshape(-1, d * d))
        exact += int(s == rp.compress_band(p64, ocfg))
        ref = rp.decompress_band(s, ocfg)
        got = bc.d_decoded[i].cpu().numpy()
        check_pixels(got, ref, p64, what=name)
        maxerr = max(maxerr, int(np.abs(got.astype(np.int64) - ref).max()))
    res = {"config": name, "images": n_images, "h": h, "w": w, "block_size": bs, "dct_size": d,
           "transform": transform, "quantization": qname, "qparam": qparam,
           "input_MB": n_planes * h * w / 1e6, "stream_bytes": total,
           "ms_compress": t_c, "ms_decompress": t_d, "ms_compress_graph": tg_c, "ms_decompress_graph": tg_d,
           "compress_graph_frac_of_measured_hbm": a / (tg_c * 1e-3) / 1e9 / peak,
           "decompress_graph_frac_of_measured_hbm": a / (tg_d * 1e-3) / 1e9 / peak,
           "compress_MPps": mp / (t_c * 1e-3), "decompress_MPps": mp / (t_d * 1e-3),
           "compress_GBps": a / (t_c * 1e-3) / 1e9, "decompress_GBps": a / (t_d * 1e-3) / 1e9,
           "compress_frac_of_measured_hbm": a / (t_c * 1e-3) / 1e9 / peak,
           "decompress_frac_of_measured_hbm": a / (t_d * 1e-3) / 1e9 / peak,
           "parity": {"planes_checked": check_planes, "streams_byte_exact": exact, "tie_mismatches": ties,
                      "max_pixel_error": maxerr}}
    print(json.dumps(res), flush=True)
    del bc
    torch.cuda.empty_cache()
    return res


def measured_peak():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(pk))["hbm_gbs"]) if os.path.exists(pk) else 6550.1


def run_colour(n_img=512, h=1080, w=1920, steps=10, warmup=3, content="synthetic"):
    """RGB <-> YCbCr kernels: algorithmic bytes = 3 bytes per pixel read + 3 written."""
    import jpeg_b200 as jb
    if content == "random":                  # uniform random bytes: worst case for the table lookups (bank conflicts)
        gen = torch.Generator(device="cuda").manual_seed(5)
        rgb = torch.randint(0, 256, (n_img, h, w, 3), dtype=torch.uint8, device="cuda", generator=gen)
    else:                                    # the benchmark's smooth + noise planes, three per image, read as R, G, B
        import bench
        bench.H, bench.W = h, w
        rgb = bench.synth_planes_device(n_img, torch.device("cuda"), 0).view(n_img, 3, h, w).permute(0, 2, 3, 1).contiguous()
    planes = torch.empty((3 * n_img, h, w), dtype=torch.uint8, device="cuda")
    back = torch.empty_like(rgb)
    for _ in range(warmup):
        jb.rgb_to_ycbcr_planes(rgb, out=planes); jb.ycbcr_planes_to_rgb(planes, out=back)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        jb.rgb_to_ycbcr_planes(rgb, out=planes)
    ev[1].record()
    for _ in range(steps):
        jb.ycbcr_planes_to_rgb(planes, out=back)
    ev[2].record()
    torch.cuda.synchronize()
    t_f, t_i = ev[0].elapsed_time(ev[1]) / steps, ev[1].elapsed_time(ev[2]) / steps
    peak = measured_peak()
    a = 6.0 * n_img * h * w
    res = {"config": "colour: %d x %dx%d RGB <-> YCbCr planes, %s content" % (n_img, w, h, content), "ms_rgb_to_ycbcr": t_f, "ms_ycbcr_to_rgb": t_i,
           "rgb_to_ycbcr_GBps": a / (t_f * 1e-3) / 1e9, "ycbcr_to_rgb_GBps": a / (t_i * 1e-3) / 1e9,
           "rgb_to_ycbcr_frac_of_measured_hbm": a / (t_f * 1e-3) / 1e9 / peak,
           "ycbcr_to_rgb_frac_of_measured_hbm": a / (t_i * 1e-3) / 1e9 / peak,
           "MPps_rgb_to_ycbcr": n_img * h * w / 1e6 / (t_f * 1e-3), "MPps_ycbcr_to_rgb": n_img * h * w / 1e6 / (t_i * 1e-3)}
    print(json.dumps(res), flush=True)
    return res


def main():
    out = []
    if len(sys.argv) > 1 and sys.argv[1] == "colour":
        run_colour()
        run_colour(content="random")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "config5":
        run("5: 16384x16384 DFT qtable, single image", 16384, 16384, 4, 8, "DFT", "qtable", None, 1, steps=2, warmup=1)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "config3":
        run("3b: 3840x2160 bs5 d24 divide 1000, batch of 8", 2160, 3840, 5, 24, "DCT", "divide", 1000, 8, steps=2, warmup=1)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "single":
        # the single-image configurations only (launch lists of the latency-bound cases)
        run("1: 512x512 defaults, single image", 512, 512, 4, 8, "DCT", "qtable", None, 1, steps=2, warmup=1)
        run("2: 3840x2160 DCT qtable, single image", 2160, 3840, 4, 8, "DCT", "qtable", None, 1, steps=2, warmup=1)
        run("3: 3840x2160 bs5 d24 divide 1000, single image", 2160, 3840, 5, 24, "DCT", "divide", 1000, 1, steps=2, warmup=1)
        return
    out.append(run("1: 512x512 defaults, single image", 512, 512, 4, 8, "DCT", "qtable", None, 1, check_planes=3))
    out.append(run("1b: 512x512 defaults, batch of 1024", 512, 512, 4, 8, "DCT", "qtable", None, 1024))
    out.append(run("2: 3840x2160 DCT qtable, single image", 2160, 3840, 4, 8, "DCT", "qtable", None, 1))
    out.append(run("2b: 3840x2160 DCT qtable, batch of 64", 2160, 3840, 4, 8, "DCT", "qtable", None, 64))
    out.append(run("3: 3840x2160 bs5 d24 divide 1000, single image", 2160, 3840, 5, 24, "DCT", "divide", 1000, 1))
    out.append(run("3b: 3840x2160 bs5 d24 divide 1000, batch of 32", 2160, 3840, 5, 24, "DCT", "divide", 1000, 32))
    out.append(run("4: 1024 x 1920x1080 DCT qtable (bench.py workload)", 1080, 1920, 4, 8, "DCT", "qtable", None, 1024))
    out.append(run("5: 16384x16384 DFT qtable, single image", 16384, 16384, 4, 8, "DFT", "qtable", None, 1, steps=5))
    path = os.path.join(ROOT, "gpurun_out", "configs.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()

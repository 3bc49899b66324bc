#!/usr/bin/env python
"""Benchmark of the per-block codec hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the metric is quoted on at 1/2/4/8 GPUs): a batch
of 1024 synthetic 1920x1080 three-band images, --block_size 4 --dct_size 8 --transform DCT,
table quantisation, sharded by image across the ranks (strong scaling, no collective on the
data path).  One step = compress the rank's images, then decompress the streams just
produced.  `value` = source megapixels of the whole batch / step time (device-timed, inputs
resident in HBM); `e2e` = the same through BatchCodec.roundtrip_host with pinned host buffers,
host<->device copies inside the timed region.  `roofline.achieved` = algorithmic bytes of one launch /
the fused kernel's own duration, measured live: the library records caller-owned CUDA events right
before and after that kernel (jb_debug_kernel_events) in K direct launches after the timed region;
`whole_call_*` are the same bytes over the complete library call of the timed region.

--impl reference times the UNMODIFIED reference (the verbatim copy tools/install_reference.py leaves under
baseline/_ref/, which travels to the GPU box) on all host cores, on a bounded sample; the vectorised port
(oracle/ref_port.py) is timed beside it and is the fallback where no reference tree exists.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_IMAGES, H, W = 1024, 1080, 1920
BS, D, TRANSFORM, QNAME = 4, 8, "DCT", "qtable"
METRIC = "compress+decompress megapixels/s"


# ----------------------------------------------------------------------------------------------
# CPU side (oracle port) -- used for cpu_baseline and --impl reference only
# ----------------------------------------------------------------------------------------------
def _cpu_roundtrip_image(planes):
    """One image (3 planes, numpy uint8) through the oracle; returns (streams, decoded)."""
    from oracle import ref_port as rp
    cfg = rp.OracleConfig(planes.shape[2], planes.shape[1], BS, D, TRANSFORM, QNAME)
    streams = [rp.compress_band(p.astype("int64"), cfg) for p in planes]
    decoded = [rp.decompress_band(s, cfg) for s in streams]
    return streams, decoded


_CPU_IMAGES = None          # a few distinct synthetic images, built before the pool forks


def _cpu_images():
    global _CPU_IMAGES
    if _CPU_IMAGES is None:
        import numpy as np
        from golden_inputs import synth_plane
        _CPU_IMAGES = [np.stack([synth_plane(H, W, 1000 * i + b, phase=0.7 * b) for b in range(3)]).astype(np.uint8)
                       for i in range(4)]
    return _CPU_IMAGES


def _cpu_worker(i):
    _cpu_roundtrip_image(_CPU_IMAGES[i % len(_CPU_IMAGES)])
    return 0


_REF = None                 # the UNMODIFIED reference (oracle/load_reference.py), loaded before the pool forks


def reference_tree():
    """Path of the reference tree this run can import (baseline/_ref/reference on the GPU box), or None."""
    copy = os.path.join(ROOT, "baseline", "_ref", "reference")
    if os.path.isfile(os.path.join(copy, "pipeline", "__init__.py")):
        os.environ.setdefault("JB_REFERENCE_ROOT", copy)       # bench.py never reads /root/reference
    from oracle import load_reference as lr
    return lr.REFERENCE_ROOT if lr.reference_available() else None


def _ref_roundtrip_image(planes):
    """One image through the reference's own compress_band / decompress_band (pipeline/__init__.py:71-88)."""
    P = _REF.pipeline
    cfg = P.Configuration(width=planes.shape[2], height=planes.shape[1], block_size=BS, dct_size=D,
                          transform=TRANSFORM, quantization=P.QuantizationMethod(QNAME))
    streams = [P.compress_band(p.astype("int64"), cfg) for p in planes]
    decoded = [P.decompress_band(s, cfg) for s in streams]
    return streams, decoded


def _ref_worker(i):
    import warnings
    warnings.simplefilter("ignore")
    _ref_roundtrip_image(_CPU_IMAGES[i % len(_CPU_IMAGES)])
    return 0


def _warm_worker(i):
    import numpy  # noqa: F401
    return 0


def cpu_sample_throughput(n_images, procs, impl="port"):
    """Round-trip n_images synthetic images (compress + decompress, three planes each) over
    `procs` processes; returns (MP/s of the sample, seconds).  impl: "port" = oracle/ref_port.py,
    "reference" = the unmodified reference itself."""
    import multiprocessing as mp
    global _REF
    _cpu_images()
    if impl == "reference" and _REF is None:
        import warnings
        from oracle.load_reference import load_reference
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _REF = load_reference()
    worker = _ref_worker if impl == "reference" else _cpu_worker
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        # page in numpy in every worker (the port is fast enough to afford a full untimed pass)
        pool.map(_warm_worker if impl == "reference" else worker, range(procs), chunksize=1)
        t0 = time.perf_counter()
        pool.map(worker, range(n_images), chunksize=1)
        dt = time.perf_counter() - t0
    return n_images * H * W / 1e6 / dt, dt


REF_SAMPLE_NOTE = ("the UNMODIFIED reference (baseline/_ref/reference: pipeline.compress_band + decompress_band, "
                   "pipeline/__init__.py:71-88) under the two import shims of oracle/load_reference.py (pure-Python "
                   "bitarray stand-in, numpy aliases)")


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path on all host cores, on this arm's
    config; each step a bounded sample (the real reference needs ~5 s per 1080p image and core, so a step is one
    image per core; the vectorised port: two).  Falls back to the oracle port (kind "port") only where no
    reference tree travelled with the snapshot."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = "reference" if reference_tree() else "port"
    sample = cores if kind == "reference" else max(2 * cores, 8)
    vals = []
    t_begin = time.perf_counter()
    for step in range(args.warmup + args.steps):
        if kind == "reference" and step < args.warmup and step > 0:
            continue                                      # one warm-up pass pages everything in; each costs ~6 s
        v, dt = cpu_sample_throughput(sample, cores, kind)
        if step >= args.warmup:
            vals.append((v, dt))
        if time.perf_counter() - t_begin > 150 and vals:
            break
    v = sum(x for x, _ in vals) / len(vals)
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    port_v, port_dt = cpu_sample_throughput(max(2 * cores, 8), cores, "port") if kind == "reference" else (v, 0.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "MP/s", "cores": cores, "kind": kind,
                         "sample": "%d synthetic 1920x1080 images per step (compress + decompress, 3 bands each), one "
                                   "image per task, multiprocessing.Pool(%d); %s" % (
                                       sample, cores, REF_SAMPLE_NOTE if kind == "reference" else
                                       "oracle/ref_port.py (float64 numpy restatement; no reference tree on this box)"),
                         "port_value": port_v, "port_note": "oracle/ref_port.py (vectorised numpy restatement) on the "
                                                            "same images and cores, for comparison"},
        "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json_line(line)


def workload_config(n_gpus):
    return {"workload": "batch of %d synthetic %dx%d 3-band images, block_size %d, dct_size %d, %s, %s "
                        "quantisation; one step = compress then decompress" % (N_IMAGES, W, H, BS, D, TRANSFORM, QNAME),
            "images": N_IMAGES, "width": W, "height": H, "block_size": BS, "dct_size": D,
            "transform": TRANSFORM, "quantization": QNAME,
            "sharding": "by image, %d per rank, no collective" % (N_IMAGES // max(n_gpus, 1)),
            "l2": "inputs larger than L2 (%.0f MB per rank vs 126 MB), no flush needed"
                  % (N_IMAGES // max(n_gpus, 1) * 3 * H * W / 1e6)}


# ----------------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons of this rank's GPU during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            # "under load": samples drawing at least half of the peak power seen
            pmax = max(power)
            loaded = sorted(s for s, p in zip(sm, power) if p >= 0.5 * pmax) or sorted(sm)
            out.update(sm_mhz=loaded[len(loaded) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=pmax)
        return out


def synth_planes_device(n_images, device, first_image):
    """[n_images * 3, H, W] uint8 on the device: clip(127 + 100 sin(x/37 + ph) cos(y/53 + ph) + N(0, 8^2))
    (SURVEY.md section 8d), seeded per image so that any rank layout gives the same batch."""
    import torch
    out = torch.empty((n_images * 3, H, W), dtype=torch.uint8, device=device)
    ys = torch.arange(H, device=device, dtype=torch.float32).view(H, 1)
    xs = torch.arange(W, device=device, dtype=torch.float32).view(1, W)
    gen = torch.Generator(device=device)
    for i in range(n_images):
        gen.manual_seed(1000 * (first_image + i) + 7)
        for b in range(3):
            ph = 0.7 * b + 0.01 * ((first_image + i) % 97)
            v = 127.0 + 100.0 * torch.sin(xs / 37.0 + ph) * torch.cos(ys / 53.0 + ph)
            v = v + 8.0 * torch.randn((H, W), device=device, generator=gen)
            out[3 * i + b] = v.round().clamp_(0, 255).to(torch.uint8)
    return out


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import jpeg_b200 as jb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the codec path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    all_cores = os.sched_getaffinity(0)
    numa_cores = jb.sharding.bind_host_to_device(local_rank)   # pinned buffers on the GPU's own NUMA node
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner / debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    i0, i1 = jb.sharding.image_slice(N_IMAGES, rank, world)
    n_img = i1 - i0
    n_planes = 3 * n_img
    if args.streams == 0:
        args.streams = 4 if n_img <= 256 else 3
    cfg = jb.Configuration(width=W, height=H, block_size=BS, dct_size=D, transform=TRANSFORM,
                           quantization=jb.QuantizationMethod(QNAME))
    bc = jb.BatchCodec(cfg, n_planes, device=device, flags=(jb._lib.JB_FLAG_PDL if args.pdl else 0))
    bc.d_planes.copy_(synth_planes_device(n_img, device, i0))
    torch.cuda.synchronize()
    mp_rank = n_img * H * W / 1e6
    mp_total = N_IMAGES * H * W / 1e6

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

    # ---- untimed: first pass for the stream size, and the parity sample ----
    comp = bc.compress_device()
    total_bytes = comp.total_bytes()
    out, status = bc.decompress_device(comp, total_bytes)
    jb.check_status(status)
    serial_streams = int(status.cpu()[2].item())

    sampler = None
    if rank == 0:
        uuid = str(torch.cuda.get_device_properties(device).uuid)
        sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        if sampler.proc is None:
            sampler = None

    # ---- device-resident timing ----
    # The two library calls of a step are captured once into two CUDA graphs (same kernels, same buffers)
    # and replayed: per rank and step ~20 small launches otherwise cost more CPU and launch gaps than the
    # kernels themselves when the per-rank batch is small (8-GPU runs).  --no-graph launches them directly.
    K, Wm = args.steps, args.warmup
    launch_mode = "direct"
    state = {"comp": comp}
    def direct_c():
        state["comp"] = bc.compress_device()
    def direct_d():
        bc.decompress_device(state["comp"], total_bytes)
    run_c, run_d = direct_c, direct_d
    run_step = None
    two_streams = None
    launches_per_step, launches_source = 7, "known chain (3 compress + 4 decompress kernels): no graph was captured"
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                direct_c(); direct_d()                     # warm the allocator pools on the capture stream
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_c, g_d = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_c):
                comp_g = bc.compress_device()
            with torch.cuda.graph(g_d, pool=g_c.pool()):
                out_g, status_g = bc.decompress_device(comp_g, total_bytes)
            g_c.replay(); g_d.replay()
            torch.cuda.synchronize()
            jb.check_status(comp_g.status); jb.check_status(status_g)
            if comp_g.total_bytes() != total_bytes or not torch.equal(out_g[:3], out[:3]):
                raise RuntimeError("graph replay differs from the direct calls")
            state["comp"] = comp_g
            run_c, run_d = g_c.replay, g_d.replay
            launch_mode = "cuda graphs (one per library call)"
            if args.step_graph:
                # the step's two calls in ONE graph: no graph-launch gap between compress and decompress
                g_s = torch.cuda.CUDAGraph()
                lib_ = jb._lib.load()
                n_before = lib_.jb_debug_launch_count()
                with torch.cuda.graph(g_s, pool=g_c.pool()):
                    comp_s = bc.compress_device()
                    out_s, status_s = bc.decompress_device(comp_s, total_bytes)
                launches_per_step = int(lib_.jb_debug_launch_count() - n_before)
                launches_source = "jb_debug_launch_count() across the capture of the step's CUDA graph"
                g_s.replay()
                torch.cuda.synchronize()
                jb.check_status(comp_s.status); jb.check_status(status_s)
                if comp_s.total_bytes() != total_bytes or not torch.equal(out_s[:3], out[:3]):
                    raise RuntimeError("step graph replay differs from the direct calls")
                run_step = g_s.replay
                launch_mode = "cuda graph (one per step: compress + decompress)"
                if args.streams >= 2:
                    # Consecutive steps are independent batches: further codec objects (own buffers, own workspaces),
                    # each on its own CUDA stream, take the steps in turn, so the head of one step overlaps the tail of
                    # the step before it.  Every step still compresses and decompresses this rank's whole batch.
                    graphs, streams_, keep = [g_s], [torch.cuda.Stream()], []
                    for extra in range(args.streams - 1):
                        bcx = jb.BatchCodec(cfg, n_planes, device=device, flags=bc.flags)
                        bcx.d_planes.copy_(bc.d_planes)
                        compx = bcx.compress_device()
                        if compx.total_bytes() != total_bytes:
                            raise RuntimeError("another codec object disagrees on the stream size")
                        _, stx = bcx.decompress_device(compx, total_bytes)
                        jb.check_status(stx)
                        side.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(side):
                            bcx.compress_device(); bcx.decompress_device(compx, total_bytes)
                        torch.cuda.current_stream().wait_stream(side)
                        torch.cuda.synchronize()
                        g_x = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g_x):
                            comp_x = bcx.compress_device()
                            out_x, status_x = bcx.decompress_device(comp_x, total_bytes)
                        g_x.replay()
                        torch.cuda.synchronize()
                        jb.check_status(comp_x.status); jb.check_status(status_x)
                        if not torch.equal(out_x, out_s):
                            raise RuntimeError("another codec object decodes differently")
                        graphs.append(g_x); streams_.append(torch.cuda.Stream()); keep.append((bcx, comp_x, out_x, status_x))

                    def run_turn(step, _g=graphs, _s=streams_):
                        with torch.cuda.stream(_s[step % len(_s)]):
                            _g[step % len(_g)].replay()
                    two_streams = (streams_, run_turn, keep)
                    launch_mode = ("cuda graph (one per step: compress + decompress); steps take turns on %d codec objects, "
                                   "each on its own CUDA stream" % len(graphs))
        except Exception as exc:                          # noqa: BLE001 -- report and fall back to direct launches
            launch_mode = "direct (graph capture failed: %s)" % str(exc).splitlines()[0][:120]
            run_c, run_d, run_step, two_streams = direct_c, direct_d, None, None
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for step in range(Wm):
        run_c(); run_d()
    barrier()
    for step in range(K):
        ev[step][0].record()
        run_c()
        ev[step][1].record()
        run_d()
        ev[step][2].record()
    barrier()
    t_c = sum(e[0].elapsed_time(e[1]) for e in ev) / K           # ms per step, compress part
    t_d = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    t_all = ev[0][0].elapsed_time(ev[-1][2]) / K
    t_split = t_all
    if run_step is not None:
        # the timed region proper: K steps, one graph each; the split above (same kernels, one graph per call)
        # is kept for the compress / decompress breakdown
        cur = torch.cuda.current_stream()

        def run_steps(count):
            if two_streams is None:
                for step in range(count):
                    run_step()
                return
            streams_, run_turn, _keep = two_streams
            for st in streams_:
                st.wait_stream(cur)
            for step in range(count):
                run_turn(step)
            for st in streams_:
                cur.wait_stream(st)
        run_steps(Wm)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(K)
        e1.record()
        barrier()
        t_all = e0.elapsed_time(e1) / K
    t_c, t_d, t_all, t_split = max_over_ranks(t_c), max_over_ranks(t_d), max_over_ranks(t_all), max_over_ranks(t_split)
    comp = state["comp"]
    jb.check_status(comp.status)

    # ---- the two fused kernels alone (roofline.achieved): same steps launched directly, with the library's
    # measurement hook recording caller-owned CUDA events right before and after each fused kernel ----
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    for quad in kev:
        for e in quad:
            e.record()                                       # creates the handle
    torch.cuda.synchronize()
    lib = jb._lib.load()
    # (without programmatic dependent launch here: a kernel that may start under the tail of its predecessor would have
    # the wait for it counted into its own duration)
    flags_saved, bc.flags = bc.flags, bc.flags & ~jb._lib.JB_FLAG_PDL
    try:
        for quad in kev:
            if lib.jb_debug_kernel_events(*[e.cuda_event for e in quad]) != 0:
                raise RuntimeError("jb_debug_kernel_events refused the events")
            direct_c(); direct_d()
    finally:
        lib.jb_debug_kernel_events(None, None, None, None)
        bc.flags = flags_saved
    barrier()
    tk_c = max_over_ranks(sum(q[0].elapsed_time(q[1]) for q in kev) / K)      # ms per launch, fused forward kernel
    tk_d = max_over_ranks(sum(q[2].elapsed_time(q[3]) for q in kev) / K)      # fused inverse kernel
    if not args.no_graph and launch_mode.startswith("cuda graphs"):
        state["comp"] = comp_g

    # ---- end to end through host buffers ----
    e2e = None
    if not args.no_e2e:
        h_planes = torch.empty((n_planes, H, W), dtype=torch.uint8, pin_memory=True)
        h_planes.copy_(bc.d_planes)
        torch.cuda.synchronize()
        n_sub = max(1, min(args.e2e_sub, n_img // 8))
        for step in range(min(Wm, 2)):
            hs, offs = bc.compress_host(h_planes)
            bc.decompress_host(hs, offs)
            bc.roundtrip_host(h_planes, n_sub)
        # (a) the two calls one after the other: host -> device traffic, then device -> host traffic
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for step in range(K):
            hs, offs = bc.compress_host(h_planes)
            dec = bc.decompress_host(hs, offs)
        e1.record()
        barrier()
        t_seq = max_over_ranks(e0.elapsed_time(e1) / K)
        seq_ok = bool(torch.equal(dec[:3], bc.d_decoded[:3].cpu()))
        # (b) the same work as a two-stage pipeline over sub-batches: both directions of the link busy
        barrier()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for step in range(K):
            dec, parts = bc.roundtrip_host(h_planes, n_sub)
        e1.record()
        barrier()
        t_e2e = max_over_ranks(e0.elapsed_time(e1) / K)
        wall_e2e = max_over_ranks((time.perf_counter() - t0) / K * 1e3)
        # (c) the ceiling under (b): the same plane bytes over the same links, pinned buffers and streams, no kernels
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from link_probe import probe
        link = probe(device, n_planes * H * W, n_sub, min(K, 5), barrier, max_over_ranks,
                     h_src=h_planes, h_dst=bc.h_decoded, d_a=bc.d_planes, d_b=bc.d_decoded)
        job_bytes = N_IMAGES * 3 * H * W
        e2e = {"value": mp_total / (t_e2e * 1e-3), "unit": "MP/s",
               "link_ceiling_gbs": job_bytes / link["both_ms"] / 1e6, "link_ms_per_step": link["both_ms"],
               "frac_of_link": link["both_ms"] / t_e2e,
               "link_probe": {"what": "tools/link_probe.py: the step's plane bytes host->device and device->host at once, "
                                      "pinned memory, same sub-batches and streams, no kernels; aggregate GB/s per direction",
                              "h2d_alone_gbs": job_bytes / link["h2d_ms"] / 1e6, "d2h_alone_gbs": job_bytes / link["d2h_ms"] / 1e6,
                              "both_gbs_per_direction": job_bytes / link["both_ms"] / 1e6},
               "h2d_bytes_per_step": n_planes * H * W + total_bytes + 8 * (n_planes + n_sub),
               "d2h_bytes_per_step": total_bytes + 8 * (n_planes + n_sub) + 64 * n_sub + n_planes * H * W,
               "ms_per_step": t_e2e, "wall_ms_per_step": wall_e2e,
               "api": "BatchCodec.roundtrip_host (pinned host buffers in and out, streams pass through host memory; "
                      "%d sub-batches pipelined on two CUDA streams so that both link directions are busy), per rank" % n_sub,
               "host_cores_bound_to_gpu": len(numa_cores) if numa_cores else None,
               "sequential_ms_per_step": t_seq, "sequential_value": mp_total / (t_seq * 1e-3),
               "sequential_api": "BatchCodec.compress_host then decompress_host",
               "matches_device_path": bool(torch.equal(dec[:3], bc.d_decoded[:3].cpu())) and seq_ok
                                      and bool(torch.equal(dec[-3:], bc.d_decoded[-3:].cpu()))}

    clocks = sampler.stop() if sampler is not None else {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}

    # ---- gather totals ----
    stream_total = total_bytes
    if world > 1:
        t = torch.tensor([total_bytes], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        stream_total = int(t.item())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # per-rank algorithmic bytes of one launch of the fused kernels
        a_c = n_planes * H * W + total_bytes
        a_d = total_bytes + n_planes * H * W
        ach_c = a_c / (t_c * 1e-3) / 1e9
        ach_d = a_d / (t_d * 1e-3) / 1e9
        traffic = measured_traffic(n_img)
        os.sched_setaffinity(0, all_cores)              # the CPU leg uses every core again
        cpu = None if args.no_cpu else cpu_baseline_with_parity(bc, comp, n_img)
        line = {
            "metric": METRIC, "value": mp_total / (t_all * 1e-3), "unit": "MP/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": t_all, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "compress_mps": mp_total / (t_c * 1e-3), "decompress_mps": mp_total / (t_d * 1e-3),
            "ms_compress": t_c, "ms_decompress": t_d, "ms_per_step_two_graphs": t_split, "stream_bytes": stream_total,
            "pdl": bool(args.pdl), "streams": len(two_streams[0]) if two_streams is not None else 1,
            "decoder_serial_fallback_streams_rank0": serial_streams, "launch": launch_mode,
            "bytes_per_pixel": stream_total / (N_IMAGES * H * W),
            # achieved = algorithmic bytes of one launch / the fused kernel's own duration (CUDA events recorded by the
            # library around that kernel, averaged over K direct launches after the timed region); whole_call_* divides
            # by the time of the complete library call of the timed region (init, scan, gather / framing included)
            "roofline": {"bound": "hbm", "kernel": "jb_fwd_fast_kernel (fused compress)", "kernel_ms": tk_c,
                         "achieved": a_c / (tk_c * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": a_c / (tk_c * 1e-3) / 1e9 / peak,
                         "frac_of_nominal_8000": a_c / (tk_c * 1e-3) / 1e9 / 8000.0,
                         "whole_call_ms": t_c, "whole_call_achieved": ach_c, "whole_call_frac": ach_c / peak,
                         # dram__bytes_read + dram__bytes_write of one launch from `ncu --set full` (profiles/r2_traffic.json,
                         # recorded per image together with the build id of the kernels it was measured on; null when the
                         # sources have changed since), scaled to this rank's images
                         "traffic": traffic["fwd"], "traffic_source": traffic["source"],
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": a_c},
            "roofline_decompress": {"bound": "hbm", "kernel": "jb_inv_fast_kernel (fused decompress)", "kernel_ms": tk_d,
                                    "achieved": a_d / (tk_d * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": a_d / (tk_d * 1e-3) / 1e9 / peak,
                                    "frac_of_nominal_8000": a_d / (tk_d * 1e-3) / 1e9 / 8000.0,
                                    "whole_call_ms": t_d, "whole_call_achieved": ach_d, "whole_call_frac": ach_d / peak,
                                    "traffic": traffic["inv"], "traffic_source": traffic["source"],
                                    "algorithmic_bytes_per_launch": a_d},
            "cpu_baseline": cpu,
            "e2e": e2e,
            # kernels of ours inside the timed region: K steps x the kernel nodes of the step's CUDA graph (counted from
            # the graph's own DOT dump at capture time); per step: fused compress kernel, scan, gather; framing prep,
            # walk, stitch; fused decompress kernel.  The table builders run once per codec object, before it.
            "gpu_launches": K * launches_per_step, "gpu_launches_per_step": launches_per_step,
            "gpu_launches_source": launches_source,
            "clocks": clocks,
        }
        emit_json_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def library_build_id():
    """sha256 over the kernel sources (tools/sass_summary.py): ties a profile to the code it was measured on."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_summary
    return sass_summary.build_id()


def measured_traffic(n_img):
    """DRAM bytes of one launch of the two fused kernels, from profiles/r2_traffic.json (`ncu --set full`), scaled to
    n_img images; null values when the file is missing or belongs to other kernel sources."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    out = {"fwd": None, "inv": None, "source": "profiles/r2_traffic.json missing"}
    try:
        rec = json.load(open(path))
    except (OSError, ValueError):
        return out
    bid = library_build_id()
    if rec.get("build_id") != bid:
        out["source"] = "profiles/r2_traffic.json is for build %s, this is %s: not used" % (rec.get("build_id"), bid)
        return out
    return {"fwd": rec["fwd_dram_bytes_per_image"] * n_img, "inv": rec["inv_dram_bytes_per_image"] * n_img,
            "source": "profiles/r2_traffic.json (%s; build %s)" % (rec.get("how", "ncu --set full"), bid)}


def cpu_baseline_with_parity(bc, comp, n_img):
    """Rank 0, N-independent: the oracle on the first images of this rank's batch, all host
    cores; the oracle's outputs double as the parity check of the GPU results."""
    import numpy as np
    import torch
    from oracle import ref_port as rp
    from parity import check_pixels, check_quantised
    import jpeg_b200 as jb

    cores = os.cpu_count() or 1
    sample = max(2 * cores, 8)
    mps, dt = cpu_sample_throughput(sample, cores)
    ref_mps = ref_dt = None
    if reference_tree():
        ref_mps, ref_dt = cpu_sample_throughput(cores, cores, "reference")
    # parity of the GPU batch against the oracle on image 0 (three planes)
    planes = bc.d_planes[:3].cpu().numpy()
    offs = comp.host_offsets()
    blob = comp.data[:int(offs[3])].cpu().numpy().tobytes()
    streams = [blob[int(offs[i]):int(offs[i + 1])] for i in range(3)]
    ocfg = rp.OracleConfig(W, H, BS, D, TRANSFORM, QNAME)
    gpu_coeffs = jb.stages.forward_coefficients(planes, bc.config)
    exact, ties, maxerr = 0, 0, 0
    for i in range(3):
        zz = rp.quantised_zigzag(planes[i].astype(np.int64), ocfg)
        ties += check_quantised(gpu_coeffs[i], zz, rp.prerounding_zigzag(planes[i].astype(np.int64), ocfg),
                                what="bench plane %d" % i, strict_fraction=True)
        assert streams[i] == rp.pack_blocks(gpu_coeffs[i].reshape(-1, D * D))
        exact += int(streams[i] == rp.compress_band(planes[i].astype(np.int64), ocfg))
        ref_dec = rp.decompress_band(streams[i], ocfg)
        got = bc.d_decoded[i].cpu().numpy()
        check_pixels(got, ref_dec, planes[i], what="bench plane %d" % i)
        maxerr = max(maxerr, int(np.abs(got.astype(np.int64) - ref_dec).max()))
    port_sample = ("%d synthetic 1920x1080 images, one per task, multiprocessing.Pool(%d), %.1f s; "
                   "oracle/ref_port.py (float64 numpy restatement of the reference)" % (sample, cores, dt))
    if ref_mps is not None:
        head = {"value": ref_mps, "unit": "MP/s", "cores": cores, "kind": "reference",
                "sample": "%d synthetic 1920x1080 images (compress + decompress, 3 bands each), one per task, "
                          "multiprocessing.Pool(%d), %.1f s; %s" % (cores, cores, ref_dt, REF_SAMPLE_NOTE),
                "port_value": mps, "port_sample": port_sample}
    else:
        head = {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port", "sample": port_sample}
    return {**head,
            "parity_image0": {"streams_byte_exact": "%d/3" % exact, "tie_mismatches": ties,
                              "max_pixel_error": maxerr}}


_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly one JSON line: keep the real stdout for it and point file descriptor 1 at stderr,
    so that banners printed by libraries (NCCL's version line, for one) cannot end up in front of it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_json_line(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global N_IMAGES
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=N_IMAGES, help="batch size (profiling runs use a smaller one)")
    ap.add_argument("--e2e-sub", type=int, default=32, help="sub-batches of the pipelined end-to-end round trip")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels directly instead of replaying CUDA graphs")
    ap.add_argument("--no-pdl", dest="pdl", action="store_false",
                    help="launch the kernels without programmatic dependent launch (JB_FLAG_PDL is the default)")
    ap.add_argument("--call-graphs", dest="step_graph", action="store_false",
                    help="time one CUDA graph per library call instead of one per step (compress + decompress)")
    ap.add_argument("--streams", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="n > 1: consecutive steps take turns on n codec objects, each on its own CUDA stream; 1: one stream; "
                         "0 (default): 3, or 4 when a rank holds at most 256 images (measured: 2.38 against 2.41 ms per step "
                         "at 1024 images with 3 / 4 streams, 0.331 against 0.327 ms at 128)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    args = ap.parse_args()
    N_IMAGES = args.images
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3                       # timing rule: at least three warm-up steps
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

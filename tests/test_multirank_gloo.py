"""N > 1 path on CPU: two gloo ranks partition a batch / a tall image exactly like bench.py and the
sharding helpers do, encode their shard with the oracle (standing in for the GPU kernels, which need
a device), and rank 0 reassembles.  Checks the partition + concatenation contract of SURVEY.md 8(e):
no collective on the data path, per-shard streams concatenate to the single-process result."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, pickle
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import jpeg_b200 as jb
from oracle import ref_port as rp
from golden_inputs import synth_plane

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()

# (1) batch sharded by image
n_images, h, w = 5, 40, 56
cfg = rp.OracleConfig(w, h, 4, 8, "DCT", "qtable")
i0, i1 = jb.sharding.image_slice(n_images, rank, world)
mine = [[rp.compress_band(synth_plane(h, w, 10 * i + b), cfg) for b in range(3)] for i in range(i0, i1)]
gathered = [None] * world
dist.all_gather_object(gathered, (i0, i1, mine))          # results only; the data path has no exchange

# (2) one tall image sharded by block-row bands
H, W = 200, 64
plane = synth_plane(H, W, 99)
bands = jb.sharding.block_row_bands(H, 4, 8, world)
r0, r1 = bands[rank]
part = [rp.compress_band(plane[r0:r1], rp.OracleConfig(W, r1 - r0, 4, 8, "DCT", "qtable"))] if r1 > r0 else []
parts = [None] * world
dist.all_gather_object(parts, part)

# (3) max-over-ranks timing reduction used by bench.py
import torch
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)

if rank == 0:
    flat = []
    for a, b, streams in sorted(gathered):
        flat.extend(streams)
    with open(sys.argv[2], "wb") as f:
        pickle.dump({{"batch": flat, "bands": jb.sharding.concat_band_streams(parts), "tmax": float(t.item())}}, f)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_gloo_ranks_partition_and_reassemble(tmp_path):
    import pickle
    import socket
    sys.path.insert(0, ROOT)
    from oracle import ref_port as rp
    from golden_inputs import synth_plane
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    out = tmp_path / "out.pkl"
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(out)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = pickle.load(open(out, "rb"))
    n_images, h, w = 5, 40, 56
    cfg = rp.OracleConfig(w, h, 4, 8, "DCT", "qtable")
    want = [[rp.compress_band(synth_plane(h, w, 10 * i + b), cfg) for b in range(3)] for i in range(n_images)]
    assert res["batch"] == want
    H, W = 200, 64
    whole = rp.compress_band(synth_plane(H, W, 99), rp.OracleConfig(W, H, 4, 8, "DCT", "qtable"))
    assert res["bands"] == [whole]
    assert res["tmax"] == 2.0

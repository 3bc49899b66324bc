import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import jpeg_b200 as jb
from golden_inputs import synth_plane
flags = int(sys.argv[1]); h = int(sys.argv[2]); w = int(sys.argv[3]); what = sys.argv[4]
cfg = jb.Configuration(width=w, height=h, block_size=5, dct_size=24, transform="DCT", quantization=jb.QuantizationMethod("divide", divisor=1000))
a = synth_plane(h, w, 77).astype(np.uint8)
planes = torch.from_numpy(a).cuda().unsqueeze(0)
if what in ("fwd", "both"):
    comp = jb.compress_planes(planes, cfg, flags=flags)
    s = comp.to_bytes_list()
    print("fwd ok", flags, h, w, len(s[0]))
if what in ("coef",):
    c = jb.stages.forward_coefficients(a, cfg, flags=flags)
    print("coef ok", c.shape)
if what in ("both",):
    ref = jb.compress_planes(planes, cfg, flags=1).to_bytes_list()
    print("equal generic:", ref == s)
    rec = jb.decompress_bands(s, cfg, flags=flags)
    rec1 = jb.decompress_bands(s, cfg, flags=1)
    print("inv ok, max diff vs generic", int(np.abs(rec.astype(int) - rec1.astype(int)).max()))

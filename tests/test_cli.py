"""Command-line front ends (tools/jb_compress.py, tools/jb_decompress.py): the flags of the reference's
compress.py:20-62 / decompress.py:13-24, plus the batch mode of SURVEY.md section 8(f) row 4."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location("jbcli_" + name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_defaults_are_the_reference_defaults():
    args = _load("jb_compress").build_parser().parse_args(["in.png", "out.jb"])
    assert (args.block_size, args.dct_size, args.transform, args.quantization, args.qkeep, args.qdivisor) == \
        (4, 8, "DCT", "qtable", 2, 40)


@pytest.mark.parametrize("argv", [["in.png", "out.jb", "--transform", "DST"],
                                  ["in.png", "out.jb", "--quantization", "qtabel"],
                                  ["in.png", "out.jb", "--block_size", "0"],
                                  ["in.png"],
                                  ["in.png", "out.jb", "--batch", "a.png"]])
def test_bad_arguments_are_refused_before_any_work(argv):
    cli = _load("jb_compress")
    with pytest.raises(SystemExit) as e:
        cli.main(argv)
    assert e.value.code not in (0, None)


def test_decompress_needs_both_paths():
    with pytest.raises(SystemExit):
        _load("jb_decompress").main(["only_one"])


@pytest.mark.gpu
def test_cli_round_trip_and_batch_match_the_oracle(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    import jpeg_b200 as jb
    from oracle import ref_port as rp
    from golden_inputs import synth_plane
    h, w = 96, 160
    paths = []
    for k in range(3):
        rgb = np.stack([synth_plane(h, w, 40 + 3 * k + i) for i in range(3)], axis=-1).astype(np.uint8)
        paths.append(str(tmp_path / ("img%d.png" % k)))
        Image.fromarray(rgb, "RGB").save(paths[-1])
    comp, dec = _load("jb_compress"), _load("jb_decompress")
    out = str(tmp_path / "one.jb")
    assert comp.main([paths[0], out, "--quantization", "divide", "--qdivisor", "30", "--block_size", "2"]) == 0
    blob = open(out, "rb").read()
    # the container of the reference flow, with the oracle doing the band compression
    ycc = Image.open(paths[0]).convert("YCbCr")
    ocfg = rp.OracleConfig(w, h, 2, 8, "DCT", "divide", 30)
    cfg = jb.Configuration(width=w, height=h, block_size=2, dct_size=8, transform="DCT",
                           quantization=jb.QuantizationMethod("divide", divisor=30))
    want = [rp.compress_band(np.asarray(b, dtype=np.int64), ocfg) for b in ycc.split()]
    assert blob == jb.file_format.generate_data(cfg, jb.CompressedData(*want))
    png = str(tmp_path / "one.png")
    assert dec.main([out, png]) == 0
    restored = np.asarray(Image.open(png).convert("RGB"))
    assert restored.shape == (h, w, 3)
    assert np.array_equal(restored, np.asarray(jb.Jpeg.decompress(blob).convert("RGB")))
    # batch mode: one device batch, same bytes as one call per file
    outdir = str(tmp_path / "batch")
    assert comp.main(["--batch"] + paths + ["--outdir", outdir]) == 0
    for p in paths:
        single = str(tmp_path / "single.jb")
        assert comp.main([p, single]) == 0
        stem = os.path.splitext(os.path.basename(p))[0]
        assert open(os.path.join(outdir, stem + ".jb"), "rb").read() == open(single, "rb").read()
    assert dec.main(["--batch", os.path.join(outdir, "img0.jb"), os.path.join(outdir, "img2.jb"), "--outdir", outdir]) == 0
    assert Image.open(os.path.join(outdir, "img2.png")).size == (w, h)

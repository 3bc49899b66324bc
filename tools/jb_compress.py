"""Command-line front end with the flags of the reference's compress.py (compress.py:20-62), on the CUDA path.

    python tools/jb_compress.py in.png out.jb [--block_size 4 --dct_size 8 --transform DCT
                                            --quantization qtable --qkeep 2 --qdivisor 40]
    python tools/jb_compress.py --batch a.png b.png c.png --outdir compressed/      

The output file is the reference's container (file_format.py:67-93) and decodes through the reference's
decompress.py.  In batch mode images of equal size go through one device batch (colour conversion, compression
and container assembly on the GPU, `compress_images_rgb`); outputs are named <stem>.jb in --outdir.
Unlike the reference, an unknown --transform or --quantization is refused before any work is done.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

QUANTIZATIONS = ("none", "discard", "divide", "qtable")
TRANSFORMS = ("DCT", "DFT")


def build_parser():
    p = argparse.ArgumentParser(description="Given an image, compress it using the block-transform codec on a B200")
    p.add_argument("infile", nargs="?", help="image file to read (any format Pillow opens)")
    p.add_argument("outfile", nargs="?", help="container file to write")
    p.add_argument("--batch", nargs="+", metavar="IMAGE", help="compress several images in one device batch")
    p.add_argument("--outdir", default=".", help="destination directory of --batch outputs")
    p.add_argument("--block_size", type=int, default=4, help="edge of the box each band is averaged over before the transform")
    p.add_argument("--dct_size", type=int, default=8, help="edge of the transform blocks")
    p.add_argument("--transform", default="DCT", help="DCT or DFT (real part)")
    p.add_argument("--quantization", default="qtable", help="quantiser: none | discard | divide | qtable")
    p.add_argument("--qkeep", type=int, default=2, help="discard: rows and columns of coefficients that survive")
    p.add_argument("--qdivisor", type=int, default=40, help="divide: the common divisor")
    return p


def quantization_from_args(jb, args):
    """compress.py:53-60, except that an unknown name is an error here (the reference silently uses 'none')."""
    if args.quantization not in QUANTIZATIONS:
        raise SystemExit("unknown --quantization %r (expected one of %s)" % (args.quantization, ", ".join(QUANTIZATIONS)))
    if args.quantization == "discard":
        return jb.QuantizationMethod("discard", keep=args.qkeep)
    if args.quantization == "divide":
        return jb.QuantizationMethod("divide", divisor=args.qdivisor)
    if args.quantization == "qtable":
        return jb.QuantizationMethod("qtable")
    return None


def validate(args):
    if args.transform not in TRANSFORMS:
        raise SystemExit("unknown --transform %r (expected DCT or DFT)" % (args.transform,))
    if args.block_size < 1 or args.dct_size < 1:
        raise SystemExit("--block_size and --dct_size must be positive")
    if args.batch:
        if args.infile or args.outfile:
            raise SystemExit("--batch takes its inputs after the flag; give no positional infile/outfile")
    elif not (args.infile and args.outfile):
        raise SystemExit("usage: jb_compress.py infile outfile [options]  |  jb_compress.py --batch IMAGE... --outdir DIR")


BATCH_PIXEL_BUDGET = 256 * 1024 * 1024       # source pixels per device batch of --batch (0.75 GB of pinned RGB)


def main(argv=None):
    args = build_parser().parse_args(argv)
    validate(args)
    from PIL import Image
    import jpeg_b200 as jb

    quant = quantization_from_args(jb, args)

    def config_for(im):
        return jb.Configuration(width=im.width, height=im.height, block_size=args.block_size,
                                dct_size=args.dct_size, transform=args.transform, quantization=quant)

    if not args.batch:
        im = Image.open(args.infile)
        data = jb.Jpeg(config_for(im)).compress_rgb(im)
        with open(args.outfile, "wb") as f:
            f.write(data)
        return 0

    os.makedirs(args.outdir, exist_ok=True)
    stems = [os.path.splitext(os.path.basename(path))[0] for path in args.batch]
    clash = sorted({x for x in stems if stems.count(x) > 1})
    if clash:                                      # a/x.png and b/x.jpg would both become x.jb
        raise SystemExit("--batch: these inputs would overwrite each other's output: %s" % ", ".join(clash))
    groups = {}                                    # images of one size share a device batch
    for path in args.batch:
        with Image.open(path) as im:               # (the header only: sizes; the pixels are read per sub-batch below)
            groups.setdefault((im.width, im.height), []).append(path)
    for (w, h), paths in groups.items():
        per_batch = max(1, BATCH_PIXEL_BUDGET // (w * h))       # bounds the pinned staging buffer
        for i in range(0, len(paths), per_batch):
            part = paths[i:i + per_batch]
            images = [Image.open(path) for path in part]
            blobs = jb.compress_images_rgb(images, config_for(images[0]))
            for path, im, blob in zip(part, images, blobs):
                im.close()
                stem = os.path.splitext(os.path.basename(path))[0]
                with open(os.path.join(args.outdir, stem + ".jb"), "wb") as f:
                    f.write(blob)
    return 0


if __name__ == "__main__":
    sys.exit(main())

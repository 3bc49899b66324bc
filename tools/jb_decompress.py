"""Command-line front end with the interface of the reference's decompress.py (decompress.py:5-24), on the CUDA path.

    python tools/jb_decompress.py in.jb out.png
    python tools/jb_decompress.py --batch a.jb b.jb --outdir restored/        

Reads the reference's container (file_format.py:96-111); decoding and the YCbCr -> RGB conversion run on the GPU
(`Jpeg.decompress_rgb`), with the pixels `Jpeg.decompress(data).convert('RGB')` would give.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_parser():
    p = argparse.ArgumentParser(description="Restore an image from a container file")
    p.add_argument("infile", nargs="?", help="container file written by jb_compress.py or the reference")
    p.add_argument("outfile", nargs="?", help="image file to write (format from the extension)")
    p.add_argument("--batch", nargs="+", metavar="FILE", help="restore several files")
    p.add_argument("--outdir", default=".", help="destination directory of --batch outputs (PNG)")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.batch:
        if args.infile or args.outfile:
            raise SystemExit("--batch takes its inputs after the flag; give no positional infile/outfile")
    elif not (args.infile and args.outfile):
        raise SystemExit("usage: jb_decompress.py infile outfile  |  jb_decompress.py --batch FILE... --outdir DIR")
    import jpeg_b200 as jb

    def restore(src, dst):
        with open(src, "rb") as f:
            jb.Jpeg.decompress_rgb(f.read()).save(dst)

    if not args.batch:
        restore(args.infile, args.outfile)
        return 0
    os.makedirs(args.outdir, exist_ok=True)
    stems = [os.path.splitext(os.path.basename(path))[0] for path in args.batch]
    clash = sorted({x for x in stems if stems.count(x) > 1})
    if clash:
        raise SystemExit("--batch: these inputs would overwrite each other's output: %s" % ", ".join(clash))
    for path in args.batch:
        stem = os.path.splitext(os.path.basename(path))[0]
        restore(path, os.path.join(args.outdir, stem + ".png"))
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Regenerates tests/golden/: small seeded JPEGs and the pixels the oracle (X86 and SCALAR variants) produces
for them.  The reference is Rust and cannot run in this environment, so these pin the host stage + oracle pair
against regressions; they are not reference outputs.  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import jpeg_util  # noqa: E402
import oracle  # noqa: E402
from zune_jpeg_b200.decoder import ColorSpace, Decoder, ZuneJpegOptions  # noqa: E402

CASES = [  # name, w, h, subsampling, progressive, gray, restart_rows, out, variant
    ("c444_base", 72, 40, "444", False, False, 0, "RGB", "X86"),
    ("c420_base", 80, 48, "420", False, False, 0, "RGB", "X86"),
    ("c420_wide", 544, 32, "420", False, False, 0, "RGB", "X86"),  # chroma strip >= 500 samples: AVX2 HV form
    ("c420_rgba", 80, 48, "420", False, False, 1, "RGBA", "X86"),
    ("c422_prog", 88, 40, "422", True, False, 0, "RGB", "X86"),
    ("c420_scalar", 80, 48, "420", False, False, 0, "RGB", "SCALAR"),
    ("gray", 64, 24, "444", False, True, 0, "GRAYSCALE", "X86"),
    ("c444_ycc", 48, 16, "444", True, False, 0, "YCbCr", "SCALAR"),
]

if __name__ == "__main__":
    man = {"cases": []}
    for i, (name, w, h, sub, prog, gray, rst, out, variant) in enumerate(CASES):
        data = jpeg_util.synth_jpeg(900 + i, w, h, sub, 85, prog, gray, rst)
        open(os.path.join(HERE, name + ".jpg"), "wb").write(data)
        d = Decoder.new_with_options(ZuneJpegOptions().set_out_colorspace(ColorSpace[out]).set_use_unsafe(variant == "X86"))
        img, planes = d.decode_coefficients(data)
        for z in range(img.n_comp):
            img.comp[z].coeff = planes[z].ctypes.data
        px = oracle.reconstruct(img)
        px.tofile(os.path.join(HERE, name + ".bin"))
        man["cases"].append({"jpeg": name + ".jpg", "pixels": name + ".bin", "out": out, "variant": variant, "width": w, "height": h})
    json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
    print("wrote", len(CASES), "golden cases")

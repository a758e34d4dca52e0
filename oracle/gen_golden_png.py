"""Golden strings for the segmentation post-process tail: `binary_mask_to_base64` of the UNMODIFIED reference (focoos/utils/vision.py:270-293, OpenCV PNG) on
seeded masks -> tests/golden/png_masks.json.  Run in the build container only (needs /root/reference):  python -m oracle.gen_golden_png"""
import json
import os

import numpy as np

from oracle import ref_import


def masks():
    rng = np.random.default_rng(7)
    out = {"fixture_2x2": np.array([[1, 0], [0, 1]], dtype=bool)}  # tests/utils/conftest.py:11-15 of the reference
    out["blob_37x53"] = np.hypot(*np.mgrid[-18:19, -26:27]) < 15
    out["noise_64x48"] = rng.random((64, 48)) > 0.5
    out["row_1x200"] = rng.random((1, 200)) > 0.3
    out["full_16x16"] = np.ones((16, 16), dtype=bool)
    return out


def main():
    ref_import.install()
    import importlib
    import sys
    import cv2  # the real OpenCV (ref_import stubs it only when it is absent)
    assert not type(cv2).__name__.startswith("_Dummy"), "OpenCV is needed to generate the goldens"
    from focoos.utils.vision import binary_mask_to_base64
    g = {k: {"shape": list(m.shape), "bits": np.packbits(m).tolist(), "b64": binary_mask_to_base64(m)} for k, m in masks().items()}
    g["_meta"] = {"cv2": cv2.__version__, "source": "focoos.utils.vision.binary_mask_to_base64 (unmodified reference)"}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "png_masks.json")
    with open(path, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

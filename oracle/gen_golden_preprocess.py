"""Generates tests/golden/preprocess.npz by running the REFERENCE's own transform (clip/clip.py:73-80 `_transform`,
imported from /root/reference) on seeded synthetic RGB images of several sizes, stopping before ToTensor:
stores the input pixels and the uint8 crop the reference pipeline hands to ToTensor, plus the fp32 tensor for one.
    python oracle/gen_golden_preprocess.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402  (installs the ftfy / yacs shims and puts the reference on sys.path)

_T, ref_clip, _M = ref_loader.load_reference()
from PIL import Image  # noqa: E402
_transform = ref_clip.clip._transform       # the reference's function (clip/clip.py:73-80)



from oracle.preprocess_oracle import synth_rgb  # noqa: E402

CASES = [(64, 80, 107), (64, 150, 113), (64, 64, 64), (64, 40, 130), (64, 200, 150), (64, 65, 64), (64, 90, 300),
         (224, 256, 341)]

if __name__ == "__main__":
    out = {"cases": np.array(CASES, dtype=np.int32)}
    for i, (n_px, h, w) in enumerate(CASES):
        tf = _transform(n_px)
        img = synth_rgb(h, w, i)
        pil = Image.fromarray(img, mode="RGB")
        x = pil
        for t in tf.transforms[:3]:          # Resize, CenterCrop, convert("RGB")
            x = t(x)
        out[f"crop_{i}"] = np.asarray(x).transpose(2, 0, 1).copy()      # uint8 CHW [3, n_px, n_px]
        if i == 0:
            out["tensor_0"] = tf(pil).numpy()                            # fp32 after ToTensor + Normalize
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **out)
    print("wrote", len(CASES), "cases")

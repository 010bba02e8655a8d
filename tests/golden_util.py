import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_files(kind):
    out = []
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))):
        z = np.load(f)
        if str(z["kind"]) == kind:
            out.append(f)
    return out


def load_golden(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["params"] = eval(str(d["params"]), {"inf": np.inf, "np": np})   # written by make_golden.py with repr()
    d["name"] = os.path.basename(path)[:-4]
    if "cc" in d:
        d["cc"] = d["cc"].astype(np.float32)
    if "w" in d and d["w"].size == 0:
        d["w"] = None
    return d

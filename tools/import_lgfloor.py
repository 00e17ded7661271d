"""Re-encode the reference's map fixture (launch/lgfloor.pgm + lgfloor.yaml) as a compressed
.npz so the named benchmark configs can run where /root/reference does not exist (the GPU
box).  Only the pixel data and the six YAML scalars are stored; run once in the build
container:  python tools/import_lgfloor.py
"""
import re
import sys
from pathlib import Path

import numpy as np
import yaml

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/launch")
OUT = Path(__file__).resolve().parents[1] / "amclj_b200" / "data" / "lgfloor_map.npz"

raw = (REF / "lgfloor.pgm").read_bytes()
m = re.match(rb"P5\s+(?:#[^\n]*\n\s*)*(\d+)\s+(\d+)\s+(\d+)\s", raw)
w, h, maxval = (int(g) for g in m.groups())
img = np.frombuffer(raw[m.end():], dtype=np.uint8).reshape(h, w)
meta = yaml.safe_load((REF / "lgfloor.yaml").read_text())
np.savez_compressed(
    OUT,
    image=img,
    maxval=np.int32(maxval),
    resolution=np.float64(meta["resolution"]),
    origin=np.asarray(meta["origin"], dtype=np.float64),
    negate=np.int32(meta["negate"]),
    occupied_thresh=np.float64(meta["occupied_thresh"]),
    free_thresh=np.float64(meta["free_thresh"]),
)
print(OUT, OUT.stat().st_size, "bytes", img.shape)

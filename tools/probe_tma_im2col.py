"""GPU probe: semantics of TMA tiled loads with element strides (boxDim = traversal extent or element count?)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wenet_b200 import _lib  # noqa: E402
from wenet_b200._lib import ptr  # noqa: E402

lib = _lib.load()
d, F1, T1 = 128, 39, 40
x = torch.arange(T1 * F1 * d, dtype=torch.float32).reshape(T1, F1, d)
# encode (t, f, c) recognisably in bf16-exact integers: value = t*64 + f  for channel c<64 block, plus c parity
val = (torch.arange(T1).view(T1, 1, 1) * 64 + torch.arange(F1).view(1, F1, 1)).float().expand(T1, F1, d).clone()
buf = val.to(torch.bfloat16).cuda().contiguous()
out = torch.zeros(32768, dtype=torch.uint8, device="cuda")
status = torch.zeros(1, dtype=torch.int32, device="cuda")
dims = np.array([d, F1, T1], dtype=np.uint64)
strides = np.array([d * 2, F1 * d * 2], dtype=np.uint64)
estr = np.array([1, 2, 2], dtype=np.uint32)
for name, box, nrows in (("extent", (64, 37, 11), 114), ("extent38_12", (64, 38, 12), 114), ("count", (64, 19, 6), 114)):
    b = np.array(box, dtype=np.uint32)
    for expect in (nrows * 128,):
        status.zero_()
        rc = lib.wb_probe_tma3d(ptr(buf), ptr(dims), ptr(strides), ptr(b), ptr(estr), 64, 1, 3, expect, 16384, ptr(out), ptr(status),
                                None)
        torch.cuda.synchronize()
        msg = lib.wb_last_error().decode() if rc else ""
        rows = out[:16384].view(torch.bfloat16).float().view(128, 64).cpu()
        # un-swizzle not needed for the per-row constant value: every element of a row holds t*64+f
        got = rows[:, 0].tolist()
        print(name, "box", box, "rc", rc, msg, "barrier_completed", int(status.item()), "first rows (t*64+f):", got[:22], "... row113..115", got[112:116])

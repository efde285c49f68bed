"""Per-kernel CUDA-event times of one exposed forward / inverse NTT call (logN16, all 35 ordinary limbs, batch 64)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tiberate_fhe_b200 import Tb200Context, get_lib
from tiberate_fhe_b200.presets import PRESETS

lib = get_lib()
q, K = PRESETS[16]["q"], PRESETS[16]["K"]
ctx = Tb200Context(16, q, K)
no, N, B = ctx.num_ordinary, ctx.N, 64
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.stack([torch.randint(0, int(qi), (B, N), device="cuda", generator=gen) for qi in q[:no]], dim=1)
for what, fn in (("fwd enter", lambda: ctx.ntt(a, 0, True)), ("fwd plain", lambda: ctx.ntt(a, 0, False)),
                 ("inv exit_reduce", lambda: ctx.intt(a, 0, 2)), ("inv exit", lambda: ctx.intt(a, 0, 1))):
    fn()
    torch.cuda.synchronize()
    lib.tb200_prof_enable(1)
    fn()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.tb200_prof_collect(buf, len(buf))
    lib.tb200_prof_enable(0)
    tot = 0.0
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.split("\t")
        tot += float(ms)
        print(f"  {what:16s} {name:24s} x{cnt} {float(ms):8.3f} ms")
    print(f"{what}: {B * no * N / (tot / 1e3) / 1e9:.1f} Glimb/s")

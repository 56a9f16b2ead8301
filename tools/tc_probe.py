#!/usr/bin/env python
"""Bring-up ladder for csrc/self_attention_tc.cu: runs the kernel's probe modes one per subprocess
(each under its own timeout, so a device-side spin cannot take the whole run down) and prints what
came back.  Modes: 10 alloc/dealloc, 12 commit with no MMA, 11 first MMA + commit + wait,
13 same with the generic-address commit form, 1 logits, 0 full attention."""
import ctypes
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child(mode: int):
    import torch
    from pixelsplat_b200 import _lib
    dev = torch.device("cuda", 0)
    n, heads = 1, 4
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(n, 256, 3 * heads * 128, generator=g).to(dev)
    out = torch.zeros(n * heads * 256 * 256, device=dev)
    rc = _lib.lib.ps_self_attention_forward(n, 256, heads, 128, ctypes.c_void_p(qkv.data_ptr()),
                                            ctypes.c_float(128 ** -0.5), ctypes.c_void_p(out.data_ptr()), mode, None)
    print("mode", mode, "rc", rc, flush=True)
    torch.cuda.synchronize()
    if mode >= 10:
        vals = out[:2 * heads * n].cpu()
        print("  probe words:", [hex(x) for x in vals.view(torch.int32).tolist()], flush=True)
        return
    q, k, v = [t.reshape(n, 256, heads, 128).transpose(1, 2).double() for t in qkv.chunk(3, dim=-1)]
    s = q @ k.transpose(-1, -2)
    if mode == 1:
        got = out.reshape(n, heads, 256, 256).double()
        ref = s
    else:
        got = out[:n * 256 * heads * 128].reshape(n, 256, heads * 128).double()
        ref = (torch.softmax(s * 128 ** -0.5, -1) @ v).transpose(1, 2).reshape(n, 256, -1)
    err = float((got - ref).abs().max() / ref.abs().max())
    print("  rel err vs float64:", err, " got[0,0,:4]", got.flatten()[:4].tolist(), " ref", ref.flatten()[:4].tolist(),
          flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
    else:
        for mode in (10, 12, 11, 13, 1, 0):
            try:
                r = subprocess.run([sys.executable, "-u", __file__, str(mode)], timeout=90, capture_output=True, text=True)
                print(r.stdout.strip(), "\n  [exit", r.returncode, "]", r.stderr.strip()[-400:], flush=True)
            except subprocess.TimeoutExpired as e:
                print("mode", mode, "TIMEOUT (device-side spin);", (e.stdout or b"")[-300:], flush=True)

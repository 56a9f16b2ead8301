#!/usr/bin/env python
"""Second bring-up ladder: repeated / large launches of the self-attention kernel, one scenario per
subprocess under a timeout."""
import ctypes
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child(name: str):
    import torch
    from pixelsplat_b200 import _lib
    dev = torch.device("cuda", 0)
    heads = 4

    def launch(qkv, out, mode):
        rc = _lib.lib.ps_self_attention_forward(qkv.shape[0], 256, heads, 128, ctypes.c_void_p(qkv.data_ptr()),
                                                ctypes.c_float(128 ** -0.5), ctypes.c_void_p(out.data_ptr()), mode, None)
        assert rc == 0, _lib.lib.ps_last_error()

    def check(qkv, out):
        n = qkv.shape[0]
        q, k, v = [t.reshape(n, 256, heads, 128).transpose(1, 2).double() for t in qkv.chunk(3, dim=-1)]
        ref = (torch.softmax(q @ k.transpose(-1, -2) * 128 ** -0.5, -1) @ v).transpose(1, 2).reshape(n, 256, -1)
        got = out[:ref.numel()].reshape(ref.shape).double()
        return float((got - ref).abs().max() / ref.abs().max())

    n = {"two_synced": 1, "ten_back_to_back": 1, "n40": 40, "n40_x5": 40, "n2_then_matmul": 2, "mode1_then_mode0": 1}[name]
    qkv = torch.randn(n, 256, 3 * heads * 128, device=dev)
    out = torch.zeros(max(n * heads * 256 * 256, 1), device=dev)
    if name == "two_synced":
        launch(qkv, out, 0); torch.cuda.synchronize(); print(" first ok", flush=True)
        launch(qkv, out, 0); torch.cuda.synchronize(); print(" second ok", check(qkv, out), flush=True)
    elif name == "ten_back_to_back":
        for _ in range(10):
            launch(qkv, out, 0)
        torch.cuda.synchronize(); print(" ok", check(qkv, out), flush=True)
    elif name == "n40":
        launch(qkv, out, 0); torch.cuda.synchronize(); print(" ok", check(qkv, out), flush=True)
    elif name == "n40_x5":
        for _ in range(5):
            launch(qkv, out, 0)
        torch.cuda.synchronize(); print(" ok", check(qkv, out), flush=True)
    elif name == "n2_then_matmul":
        launch(qkv, out, 0)
        a = torch.randn(512, 512, device=dev); b = a @ a
        launch(qkv, out, 0); torch.cuda.synchronize(); print(" ok", check(qkv, out), float(b.sum()), flush=True)
    elif name == "mode1_then_mode0":
        launch(qkv, out, 1); torch.cuda.synchronize(); print(" logits ok", flush=True)
        launch(qkv, out, 0); torch.cuda.synchronize(); print(" attention ok", check(qkv, out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for name in ("two_synced", "mode1_then_mode0", "ten_back_to_back", "n40", "n40_x5", "n2_then_matmul"):
            try:
                r = subprocess.run([sys.executable, "-u", __file__, name], timeout=60, capture_output=True, text=True)
                print(name, ":", r.stdout.strip(), "[exit", r.returncode, "]", r.stderr.strip()[-300:], flush=True)
            except subprocess.TimeoutExpired as e:
                print(name, ": TIMEOUT;", (e.stdout or b"")[-300:], flush=True)

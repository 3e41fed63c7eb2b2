#!/usr/bin/env python
"""Debug aid: run a list of ad-hoc model shapes on the persistent engine, each in its own
process (a device fault poisons the context), and compare token ids with the graph engine."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASES = {
    # name: dim, hidden, layers, heads, kv_heads, vocab, seq_len
    "a_dim2048_full": (2048, 5632, 1, 32, 4, 1024, 64),
    "b_dim2048_hid2048": (2048, 2048, 1, 32, 4, 1024, 64),
    "c_dim1024_hid5632": (1024, 5632, 1, 16, 4, 1024, 64),
    "d_dim512_kvmul8": (512, 1024, 1, 8, 1, 1024, 64),
    "f_dim2048_vocab32000": (2048, 2048, 1, 32, 4, 32000, 64),
    "g_dim256_mha": (256, 512, 1, 4, 4, 1024, 64),
    "h_dim1024_mha16": (1024, 1024, 1, 16, 16, 1024, 64),
    "e_dim2048_mha": (2048, 2048, 1, 32, 32, 1024, 64),
}


def child(name):
    import torch
    from kuiperllama_b200 import Decoder, ModelShape, synth_weights
    d, h, L, nh, nkv, V, S = CASES[name]
    shape = ModelShape(name, d, h, L, nh, nkv, V, S)
    w = synth_weights(shape, "cuda", 5)
    os.environ["KLLM_ENGINE"] = "graph"
    ref = Decoder(shape, w).generate(1, 0, 24)
    os.environ["KLLM_ENGINE"] = "persistent"
    dec = Decoder(shape, w)
    got = dec.generate(1, 0, 24)
    print(name, "OK" if got == ref else f"MISMATCH {got[:6]} vs {ref[:6]}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=75)
                out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
                err = (r.stderr.strip().splitlines() or [""])[-1]
                print(f"{name:24s} rc={r.returncode} {out} {err[-120:] if r.returncode else ''}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"{name:24s} HANG (>75 s)", flush=True)

#!/usr/bin/env python
"""Phase timeline of one decode step of the persistent megakernel (kllm_decoder_profile).

    python tools/phase_timeline.py [--workload tinyllama-1.1b] [--pos 512] > profiles/rNN_phase_timeline.txt
"""
import argparse
import ctypes
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="tinyllama-1.1b")
    ap.add_argument("--pos", type=int, default=256)
    a = ap.parse_args()
    from kuiperllama_b200 import SHAPES, Decoder, check, synth_weights
    shape = SHAPES[a.workload]
    dec = Decoder(shape, synth_weights(shape, "cuda", 1235))
    assert dec.engine == "persistent"
    dec.generate(1, 0, 8)
    cap = 200 * 2000 * 8
    buf = np.zeros(cap, np.uint64)
    g, p = ctypes.c_int32(), ctypes.c_int32()
    n = a.pos + 1
    check(dec.lib.kllm_decoder_profile(dec.handle, 1, 0, n, a.pos, buf.ctypes.data_as(ctypes.c_void_p), cap,
                                       ctypes.byref(g), ctypes.byref(p)), "profile")
    G, P = g.value, p.value
    st = buf[: G * P * 8].reshape(G, P, 8).astype(np.int64)
    t0 = st[:, 0, 0].min()
    st = (st - t0) / 1e3  # us
    L = shape.layer_num
    print(f"# {shape.name}: phase timeline of the decode step at pos {a.pos} (us, globaltimer), grid {G}, {P} phases")
    print(f"# token time (first phase entered -> last barrier passed): {st[:, -1, 3].max():.1f} us")
    names = ["qkv", "attn", "wo", "w1w3", "w2"]
    stage = (st[:, :, 1] - st[:, :, 0])          # input staging (+norm)
    work = (st[:, :, 2] - st[:, :, 1])           # consuming ring stages (or attention)
    bar = (st[:, :, 3] - st[:, :, 2])            # waiting at the grid barrier
    dur = st[:, :, 3].max(axis=0) - st[:, :, 0].min(axis=0)
    print(f"{'phase':>10} {'count':>5} {'phase_us':>9} {'stage_x':>8} {'work_med':>9} {'work_max':>9} {'barrier_min':>11} {'barrier_med':>11}")
    for k, nm in enumerate(names):
        idx = [l * 5 + k for l in range(L)]
        print(f"{nm:>10} {len(idx):5d} {dur[idx].mean():9.2f} {np.median(stage[:, idx]):8.2f} {np.median(work[:, idx]):9.2f} "
              f"{work[:, idx].max(axis=0).mean():9.2f} {bar[:, idx].min(axis=0).mean():11.2f} {np.median(bar[:, idx]):11.2f}")
    idx = [P - 1]
    print(f"{'cls':>10} {1:5d} {dur[idx].mean():9.2f} {np.median(stage[:, idx]):8.2f} {np.median(work[:, idx]):9.2f} "
          f"{work[:, idx].max(axis=0).mean():9.2f} {bar[:, idx].min(axis=0).mean():11.2f} {np.median(bar[:, idx]):11.2f}")
    # with tagged hand-overs most phases have no barrier: `barrier_*` is then ~0 and the wait for the
    # previous phase's outputs shows up in `stage_x` of the consuming phase (the poll loop)
    heads = shape.head_num
    ai = [l * 5 + 1 for l in range(L)]
    attn = (st[:heads, ai, 2] - st[:heads, ai, 0])
    print(f"# attention on the {heads} head CTAs: median {np.median(attn):.2f} us, slowest head per layer (mean) "
          f"{attn.max(axis=0).mean():.2f} us")
    print(f"# sum of phase durations: {dur.sum():.1f} us; barrier_min = time the LAST arriving CTA spends in the barrier")


if __name__ == "__main__":
    main()

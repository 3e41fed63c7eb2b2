"""Helpers for the -m gpu parity tests: torch is device memory + RNG only."""
import ctypes

import numpy as np
import torch


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def bits(t):
    """Bit pattern view for exact comparisons (distinguishes -0.0 / NaN payloads)."""
    a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        fa = ba.view(np.float32).ravel(); fb = bb.view(np.float32).ravel()
        bad = np.nonzero(ba.ravel() != bb.ravel())[0]
        raise AssertionError(f"{what}: {bad.size}/{fa.size} elements differ bitwise; first at {bad[0]}: "
                             f"{fa[bad[0]]!r} vs {fb[bad[0]]!r}; max|d|={np.abs(fa - fb).max():.3e}")


def sync():
    torch.cuda.synchronize()

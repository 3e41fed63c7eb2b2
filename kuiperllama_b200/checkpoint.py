"""Read/write KuiperLLama checkpoint files (the flat .bin formats the reference's exporters
emit and its loader mmaps).  Layouts, exactly:

fp32 v0   tools/export.py:79-131, read by llama3.cpp:290-423
  int32[7] dim, hidden_dim, n_layers, n_heads, n_kv_heads, +-vocab (negative = separate
  classifier), max_seq_len; then fp32 tok_emb, attn_norm[L], wq[L], wk[L], wv[L], wo[L],
  ffn_norm[L], w1[L], w2[L], w3[L], final_norm, freqs_cos[seq,hs/2], freqs_sin[seq,hs/2],
  [wcls].   Qwen2 (export_qwen2.py:103-110, qwen2.cpp:304-333): each layer's wq/wk/wv is
  followed by its bias.
int8 v3   tools/export.py:134-210, read by llama3.cpp:184-288
  int32[8] = the 7 above + group_size; for each of wq wk wv wo w1 w2 w3, per layer: int8 q then
  fp32 scales; [wcls q + scales]; then fp32 tok_emb, attn_norm[L], ffn_norm[L], final_norm.
"""
from __future__ import annotations

import math
import struct

import numpy as np

from .decoder import ModelShape


def _np(t):
    return t.detach().cpu().contiguous().numpy() if hasattr(t, "detach") else np.ascontiguousarray(t)


def write_checkpoint(path: str, shape: ModelShape, w: dict) -> int:
    """Write `w` (dict from synth_weights, any device) in the reference's format for `shape`.
    Returns the number of bytes written."""
    s = shape
    L = s.layer_num
    vocab_field = s.vocab_size if s.shared_classifier else -s.vocab_size
    hdr = [s.dim, s.hidden_dim, L, s.head_num, s.kv_head_num, vocab_field, s.seq_len]
    with open(path, "wb") as f:
        if s.group_size == 0:
            f.write(struct.pack("7i", *hdr))
            f.write(_np(w["tok_emb"]).tobytes())
            f.write(_np(w["attn_norm"]).tobytes())
            qwen = s.flavour == "qwen2" and "bq" in w
            for name, bias in (("wq", "bq"), ("wk", "bk"), ("wv", "bv")):
                mat = _np(w[name])
                if qwen:
                    b = _np(w[bias])
                    for l in range(L):
                        f.write(mat[l].tobytes())
                        f.write(b[l].tobytes())
                else:
                    f.write(mat.tobytes())
            f.write(_np(w["wo"]).tobytes())
            f.write(_np(w["ffn_norm"]).tobytes())
            for name in ("w1", "w2", "w3"):
                f.write(_np(w[name]).tobytes())
            f.write(_np(w["final_norm"]).tobytes())
            # freqs_cos / freqs_sin: present in the file, skipped by the loader (llama3.cpp:367-368)
            hs = s.head_size
            freqs = 1.0 / (10000.0 ** (np.arange(0, hs, 2, dtype=np.float32)[: hs // 2] / hs))
            t = np.arange(s.seq_len, dtype=np.float32)
            ang = np.outer(t, freqs).astype(np.float32)
            f.write(np.cos(ang).astype(np.float32).tobytes())
            f.write(np.sin(ang).astype(np.float32).tobytes())
            if not s.shared_classifier:
                f.write(_np(w["wcls"]).tobytes())
        else:
            if s.shared_classifier:
                raise ValueError("int8 + shared classifier is not representable (reference defect)")
            f.write(struct.pack("8i", *hdr, s.group_size))
            for name in ("wq", "wk", "wv", "wo", "w1", "w2", "w3"):
                q, sc = _np(w[name]), _np(w["s" + name[1:]])
                for l in range(L):
                    f.write(q[l].tobytes())
                    f.write(sc[l].tobytes())
            f.write(_np(w["wcls"]).tobytes())
            f.write(_np(w["scls"]).tobytes())
            f.write(_np(w["tok_emb"]).tobytes())
            f.write(_np(w["attn_norm"]).tobytes())
            f.write(_np(w["ffn_norm"]).tobytes())
            f.write(_np(w["final_norm"]).tobytes())
        return f.tell()


def read_header(path: str, is_quant: bool):
    with open(path, "rb") as f:
        n = 8 if is_quant else 7
        vals = struct.unpack(f"{n}i", f.read(4 * n))
    return vals


def read_checkpoint(path: str, is_quant: bool = False, flavour: str = "llama2",
                    qkv_bias: bool | None = None, name: str | None = None):
    """Parse a reference-format checkpoint into (ModelShape, dict of numpy arrays) with the same
    keys synth_weights() produces.  Arrays are views on a read-only memmap."""
    hdr = read_header(path, is_quant)
    dim, hidden, L, heads, kv_heads, vocab_field, seq_len = hdr[:7]
    group = hdr[7] if is_quant else 0
    shared = vocab_field > 0
    V = abs(vocab_field)
    if qkv_bias is None:
        qkv_bias = flavour == "qwen2" and not is_quant
    shape = ModelShape(name or path, dim, hidden, L, heads, kv_heads, V, seq_len, shared,
                       flavour, group)
    kv = shape.kv_dim
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    off = 4 * len(hdr)

    def take_f32(*dims):
        nonlocal off
        n = int(np.prod(dims))
        a = mm[off:off + 4 * n].view(np.float32).reshape(dims)
        off += 4 * n
        return a

    def take_i8(*dims):
        nonlocal off
        n = int(np.prod(dims))
        a = mm[off:off + n].view(np.int8).reshape(dims)
        off += n
        return a

    w = {}
    if not is_quant:
        w["tok_emb"] = take_f32(V, dim)
        w["attn_norm"] = take_f32(L, dim)
        for nm, bn, rows in (("wq", "bq", dim), ("wk", "bk", kv), ("wv", "bv", kv)):
            if qkv_bias:
                mats, biases = [], []
                for _ in range(L):
                    mats.append(take_f32(rows, dim))
                    biases.append(take_f32(rows))
                w[nm], w[bn] = np.stack(mats), np.stack(biases)
            else:
                w[nm] = take_f32(L, rows, dim)
        w["wo"] = take_f32(L, dim, dim)
        w["ffn_norm"] = take_f32(L, dim)
        w["w1"] = take_f32(L, hidden, dim)
        w["w2"] = take_f32(L, dim, hidden)
        w["w3"] = take_f32(L, hidden, dim)
        w["final_norm"] = take_f32(dim)
        w["_freqs"] = take_f32(2, seq_len, shape.head_size // 2)
        w["wcls"] = None if shared else take_f32(V, dim)
    else:
        for nm, rows, cols in (("wq", dim, dim), ("wk", kv, dim), ("wv", kv, dim), ("wo", dim, dim),
                               ("w1", hidden, dim), ("w2", dim, hidden), ("w3", hidden, dim)):
            qs, ss = [], []
            for _ in range(L):
                qs.append(take_i8(rows, cols))
                ss.append(take_f32(rows * cols // group))
            w[nm], w["s" + nm[1:]] = np.stack(qs), np.stack(ss)
        if shared:
            raise ValueError("int8 + shared classifier checkpoints are not supported")
        w["wcls"] = take_i8(V, dim)
        w["scls"] = take_f32(V * dim // group)
        w["tok_emb"] = take_f32(V, dim)
        w["attn_norm"] = take_f32(L, dim)
        w["ffn_norm"] = take_f32(L, dim)
        w["final_norm"] = take_f32(dim)
    if off != mm.size:
        raise ValueError(f"{path}: parsed {off} of {mm.size} bytes -- wrong flavour/quant flag?")
    return shape, w


def to_device(w: dict, device="cuda"):
    """numpy dict from read_checkpoint -> contiguous torch tensors on `device`."""
    import torch
    out = {}
    for k, v in w.items():
        if k.startswith("_"):
            continue
        out[k] = None if v is None else torch.from_numpy(np.array(v, copy=True)).to(device)
    return out

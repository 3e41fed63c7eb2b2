"""Helpers for the tensor-parallel tests: a process spawner with a 127.0.0.1 rendezvous and a
numpy restatement of ONE tensor-parallel decode step built from the oracle's ops (test-side
checker: what the sharded decoder must compute, with the collective injected)."""
import os
import socket

import numpy as np


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, backend, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def spawn(fn, world, backend="gloo", args=()):
    """Run fn(rank, world, *args) in `world` processes; raises if any of them fails."""
    import torch.multiprocessing as mp
    mp.spawn(_entry, args=(world, free_port(), backend, fn, args), nprocs=world, join=True)


class OracleShardModel:
    """One rank of a tensor-parallel decoder, op by op through the oracle (llama3.cpp:600-745
    order), on numpy weights.  allreduce(vec) -> sum over ranks (rank-ordered)."""

    def __init__(self, oracle, full_shape, local, w, allreduce):
        self.o, self.s, self.l, self.w, self.allreduce = oracle, full_shape, local, w, allreduce
        s, l = full_shape, local
        self.hs = s.head_size
        self.kvd = l.kv_head_num * self.hs
        self.kc = np.zeros((s.layer_num, s.seq_len, self.kvd), np.float32)
        self.vc = np.zeros_like(self.kc)
        self.sin, self.cos = oracle.sincos(self.hs, s.seq_len, s.flavour)
        self.eps = oracle.eps(s.flavour)

    def mm(self, x, name, layer=None):
        w = self.w[name] if layer is None else self.w[name][layer]
        g = self.s.group_size
        if g:
            sc = self.w["s" + name[1:]] if layer is None else self.w["s" + name[1:]][layer]
            return self.o.matmul_w8(x, np.asarray(w), np.asarray(sc), g)
        return self.o.matmul(x, np.asarray(w))

    def step(self, token, pos):
        o, s, l, w = self.o, self.s, self.l, self.w
        x = np.array(w["tok_emb"][token], dtype=np.float32)
        for i in range(s.layer_num):
            h = o.rmsnorm(x, w["attn_norm"][i], self.eps)
            q, k, v = self.mm(h, "wq", i), self.mm(h, "wk", i), self.mm(h, "wv", i)
            if "bq" in w:
                q, k, v = o.add(q, w["bq"][i]), o.add(k, w["bk"][i]), o.add(v, w["bv"][i])
            q, k = o.rope(s.flavour, q, k, pos, self.sin, self.cos, self.hs)
            self.kc[i, pos], self.vc[i, pos] = k, v
            att, _ = o.mha(pos, l.head_num, i, s.seq_len, self.kvd, l.head_num // l.kv_head_num, self.hs,
                           q, self.kc, self.vc)
            x = o.add(x, self.allreduce(self.mm(att, "wo", i)))
            h = o.rmsnorm(x, w["ffn_norm"][i], self.eps)
            gate = o.swiglu(self.mm(h, "w1", i), self.mm(h, "w3", i))
            x = o.add(x, self.allreduce(self.mm(gate, "w2", i)))
        x = o.rmsnorm(x, w["final_norm"], self.eps)
        cls = w.get("wcls")
        if cls is None:
            logits = o.matmul(x, np.asarray(w["tok_emb"]))
        else:
            logits = self.mm(x, "wcls")
        return o.argmax(logits), logits


def numpy_weights(w):
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in w.items() if v is not None}

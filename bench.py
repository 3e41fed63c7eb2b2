#!/usr/bin/env python
"""bench.py -- decode tokens/s of the B200-native KuiperLLama hot path.

One "step" = one greedy decode position (one token) of the workload's model through the
device-resident decoder (libkllm_b200.so).  The timed region is EXACTLY --steps consecutive
positions starting at context 1 (pos 0), after --warmup untimed positions; every step streams
the full weight set (4.1 GB for TinyLlama-1.1B fp32) from HBM, which is >> the 126 MB L2, so no
L2 flush is needed between steps (config.l2 says so).

  value   tokens/s with everything resident in HBM: tokens fed back on the device, CUDA graph
          replays back to back, no host round trip inside the timed region.
  e2e     the same metric through the reference-facing call kllm_decoder_step() with HOST
          buffers: per step the token id + position go host->device (pinned, 16 B), the greedy
          id comes back device->host (16 B) and the host synchronises, like
          LLama2Model::predict + post_processing (llama3.cpp:642-650,733-745).
  roofline      dominant kernel (fused RMSNorm -> W1|W3 -> SiLU*gate GEMV) timed alone with CUDA
                events over every layer's weights (2 GB working set), algorithmic bytes / time vs
                the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU restatement of the reference path (oracle/, OpenBLAS sgemv when the
                bundled library is found, else OpenMP) on the box's host cores, bounded sample.

--impl reference times that CPU path alone (the reference has no other runnable build here:
its CMake needs Armadillo/glog/gtest/sentencepiece, none installed -- DESIGN.md "Oracle").
Multi-GPU (--gpus N under torchrun): tensor-parallel decode of the same model, heads / FFN
columns sharded, two all-reduces per layer (strong scaling).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "decode_tokens_per_s"
# the one number the reference publishes for this metric -- TinyLlama-1.1B fp32,
# batch 1, its CUDA backend on an RTX 3060 Laptop GPU (readme.md:25).  Other workloads: none.
PUBLISHED_TOK_S = {"tinyllama-1.1b": 60.34}


def vs_baseline(workload, tok_s):
    ref = PUBLISHED_TOK_S.get(workload)
    return tok_s / ref if ref else None
WORKLOAD_NAMES = {
    "tinyllama-1.1b": "TinyLlama-1.1B fp32 greedy decode, batch 1 (BASELINE.json configs[1])",
    "llama2-7b-int8": "Llama-2-7B int8 g64 (export.py --version 3) greedy decode, batch 1 (configs[2])",
    "qwen2.5-0.5b": "Qwen2.5-0.5B fp32 greedy decode, batch 1 (configs[3])",
    "llama2-7b": "Llama-2-7B fp32 greedy decode, batch 1 (configs[4])",
    "stories15m": "stories15M fp32 greedy decode, batch 1 (configs[0])",
    "small": "synthetic dim-288 3-layer model (debug)",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tinyllama-1.1b", choices=sorted(WORKLOAD_NAMES))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--seed", type=int, default=1235)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            try:
                sm.append(float(parts[0])); smax = max(smax, float(parts[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, kernel_key):
    """DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture)
    of a kernel, from the committed summary profiles/dominant_kernel_traffic.json, or None."""
    p = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(workload, {}).get(kernel_key)
        except Exception:
            return None
    return None


def dominant_kernel_roofline(lib, shape, w, stream_ptr, torch):
    """Time the fused RMSNorm->W1|W3->SiLU*gate GEMV alone, cycling over all layers' weights."""
    from kuiperllama_b200 import GemvJob
    L, dim, hid = shape.layer_num, shape.dim, shape.hidden_dim
    x = torch.empty(dim, device="cuda").normal_(0, 1)
    h = torch.empty(hid, device="cuda")
    jobs = []
    for l in range(L):
        j = GemvJob()
        j.x = x.data_ptr(); j.norm_w = w["ffn_norm"][l].data_ptr(); j.norm_eps = 1e-5
        j.in_dim = dim; j.group_size = shape.group_size; j.n_seg = 2; j.swiglu_pair = 1
        j.seg[0].w = w["w1"][l].data_ptr(); j.seg[0].out = h.data_ptr(); j.seg[0].rows = hid
        j.seg[1].w = w["w3"][l].data_ptr(); j.seg[1].rows = hid
        if shape.group_size:
            j.seg[0].scales = w["s1"][l].data_ptr(); j.seg[1].scales = w["s3"][l].data_ptr()
        jobs.append(j)
    reps = max(2, 66 // L)
    for j in jobs:  # warm-up pass
        lib.kllm_gemv_fused(ctypes.byref(j), stream_ptr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for j in jobs:
            lib.kllm_gemv_fused(ctypes.byref(j), stream_ptr)
    e1.record()
    torch.cuda.synchronize()
    n = reps * L
    sec = e0.elapsed_time(e1) / 1e3 / n
    wbytes = 2 * hid * dim * (4 if shape.group_size == 0 else 1)
    if shape.group_size:
        wbytes += 2 * hid * dim // shape.group_size * 4
    algo = wbytes + 2 * dim * 4 + hid * 4  # weights (+scales) + x + norm weight + output
    return algo, sec, n


def cpu_baseline(shape, w, budget_s, write_ckpt=True, first_token=1):
    """Time the CPU restatement of the reference path on this box's host cores."""
    from kuiperllama_b200.checkpoint import write_checkpoint
    from oracle.binding import Oracle, find_openblas
    ckpt_dir = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(ckpt_dir, f"kllm_bench_{os.getpid()}.bin")
    o = Oracle()
    blas = find_openblas()
    try:
        write_checkpoint(path, shape, w)
        ncores = os.cpu_count() or 1
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(min(ncores, 64)))
        o.use_fast_matmul(True, blas)
        m = o.open_model(path, shape.group_size > 0, shape.flavour)
        tok, pos = first_token, 0
        tok, _ = m.step(tok, pos, want_logits=False)  # warm (page in the mmap)
        pos, n, t0 = 1, 0, time.perf_counter()
        while True:
            tok, _ = m.step(tok, pos, want_logits=False)
            pos += 1; n += 1
            dt = time.perf_counter() - t0
            if dt >= budget_s or pos >= shape.seq_len - 1 or n >= 256:
                break
        m.close()
        o.use_fast_matmul(False)
    finally:
        if os.path.exists(path):
            os.remove(path)
    threads = min(ncores, 64) if blas else o.num_threads()
    return {"value": n / dt, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{n} decode positions (context 2..{pos}) of {shape.name}, "
                      f"{'OpenBLAS sgemv ' + os.path.basename(blas) if blas else 'OpenMP row-parallel'} matmuls, "
                      f"{dt:.1f} s; oracle/kuiper_oracle.c restating kuiper/source/op/kernels/cpu/*.cpp + llama3.cpp"}


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (restated port)."""
    if rank != 0:
        return
    import torch
    from kuiperllama_b200 import SHAPES, synth_weights
    shape = SHAPES[args.workload]
    w = synth_weights(shape, "cuda" if torch.cuda.is_available() else "cpu", args.seed)
    # K steps requested; bounded so the run ends within minutes
    budget = min(150.0, max(10.0, 0.25 * args.steps))
    res = cpu_baseline(shape, w, budget)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": vs_baseline(args.workload, res["value"]),
        "dtype": "f32" if shape.group_size == 0 else "int8w/f32",
        "data": "synthetic random-init weights (tools/model.py init), greedy decode from token 1",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "shape": shape.name},
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _leave_process_group():
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def run_ours(args, rank, world):
    import torch
    from kuiperllama_b200 import SHAPES, Decoder, load_library, synth_weights

    lib = load_library()
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    shape = SHAPES[args.workload]
    K, W = args.steps, args.warmup
    if W + 1 > shape.seq_len or K > shape.seq_len:
        raise SystemExit(f"--steps/--warmup exceed seq_len {shape.seq_len}")

    stream = torch.cuda.current_stream()
    stream_ptr = ctypes.c_void_p(stream.cuda_stream)
    comm = None
    local = shape  # this rank's share of the model (== shape on one GPU)
    if world > 1:
        from kuiperllama_b200.tensor_parallel import Comm, local_shape, make_tp_decoder
        comm = Comm(shape.dim)
        full = synth_weights(shape, "cuda", args.seed)  # same seed on every rank -> same model
        dec = make_tp_decoder(shape, full, comm, stream.cuda_stream)
        w, local = dec.weights, local_shape(shape, world, rank)
        del full
        torch.cuda.empty_cache()
    else:
        w = synth_weights(shape, "cuda", args.seed)
        dec = Decoder(shape, w, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------
    barrier()  # tensor parallel: every rank's kernel waits for its peers' partial sums, so start together
    dec.generate(1, 0, max(W, 3))  # warm-up positions (untimed)
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    launches0 = lib.kllm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    ids = dec.generate(1, 0, K)
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = lib.kllm_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- end to end through the host-buffer call ----------------------------------------
    tok = 1
    for pos in range(min(W, 3)):
        tok = dec.step(tok, pos)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_e2e0 = lib.kllm_launch_count()
    e2.record(stream)
    tok, ids_e2e = 1, []
    for pos in range(K):
        tok = dec.step(tok, pos)
        ids_e2e.append(tok)
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    launches_e2e = lib.kllm_launch_count() - launches_e2e0
    if ids_e2e != ids:
        raise SystemExit("e2e path produced different token ids than the device-resident loop")

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        dec.close()
        if comm:
            comm.close()  # collective (barrier): rank 0 does the same right below
        _leave_process_group()
        return
    if comm:
        dec_engine, dec_launches = dec.engine, dec.launches_per_step
        dec.close()
        comm.close()

    tok_s = K / (ms / 1e3)
    bytes_tok = shape.weight_bytes_per_token()
    # what ONE GPU streams per token: its shard of the matmuls + the replicated classifier,
    # embedding row and norm vectors
    from kuiperllama_b200.tensor_parallel import weight_bytes_per_token_per_gpu
    bytes_tok_gpu = weight_bytes_per_token_per_gpu(shape, world, rank)
    peak, peak_src = measured_peaks()
    line = {
        "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": vs_baseline(args.workload, tok_s),  # vs 60.34 tok/s (reference CUDA, RTX 3060 Laptop)
        "dtype": "f32" if shape.group_size == 0 else "int8w/f32",
        "data": "synthetic random-init weights (tools/model.py init, seed %d), greedy decode from token 1" % args.seed,
        "config": {"workload": WORKLOAD_NAMES[args.workload], "shape": shape.name,
                   "context": f"1->{K}", "batch": 1,
                   "parallelism": "single GPU" if world == 1 else f"tp{world}",
                   "l2": "no flush: every step streams %.2f GB of weights >> 126 MB L2" % (bytes_tok / 1e9),
                   "weight_bytes_per_token": bytes_tok,
                   "launches_per_step": dec_launches if comm else dec.launches_per_step},
        "e2e": {"value": K / (ms_e2e / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": 16,
                "d2h_bytes_per_step": 16},
        # kernels of libkllm_b200 launched inside the timed region of `value`: the persistent engine
        # decodes all K positions in ONE cooperative launch (the graph engine: K x launches_per_step);
        # the e2e region launches once per token
        "gpu_launches": int(launches),
        "gpu_launches_e2e": int(launches_e2e),
        "clocks": clocks,
    }
    if world > 1:
        line["config"]["tp_comm"] = comm.backend
        line["config"]["weight_bytes_per_token_per_gpu"] = bytes_tok_gpu
    engine = dec_engine if comm else dec.engine
    line["config"]["engine"] = engine
    algo, sec, n = dominant_kernel_roofline(lib, local, w, stream_ptr, torch)
    gemv = {"bound": "hbm", "achieved": algo / sec / 1e9, "peak": peak, "unit": "GB/s",
            "frac": algo / sec / 1e9 / peak, "traffic": ncu_traffic(args.workload, "gemv"),
            "kernel": "gemv_kernel<*,swiglu> (RMSNorm->W1|W3->SiLU*gate), timed alone",
            "algorithmic_bytes_per_launch": algo, "avg_launch_us": sec * 1e6,
            "launches_timed": n, "peak_source": peak_src}
    if engine == "persistent":
        # ONE launch of the persistent megakernel decodes all K positions: the dominant (only)
        # kernel of the step.  Algorithmic bytes per launch = K x bytes per token; its duration is
        # the event-timed region above (the launch is the only work between the two events).
        ach = bytes_tok_gpu * K / (ms / 1e3) / 1e9
        traffic = ncu_traffic(args.workload, "megakernel")
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": traffic["dram_bytes_per_token"] * K if traffic else None,
                            "kernel": "decode_megakernel (persistent, whole forward + argmax, %d positions per launch)" % K,
                            "algorithmic_bytes_per_launch": bytes_tok_gpu * K, "avg_launch_us": ms * 1e3,
                            "launches_timed": 1, "peak_source": peak_src,
                            "traffic_source": traffic}
        line["roofline_fused_gemv_alone"] = gemv
    else:
        line["roofline"] = gemv
        line["step_hbm_frac"] = {"algorithmic_gbs": bytes_tok_gpu * tok_s / 1e9, "peak_gbs": peak,
                                 "frac": bytes_tok_gpu * tok_s / 1e9 / peak,
                                 "note": "whole decode step (all launches), per GPU"}
    if world == 1 and not args.no_cpu_baseline:
        dec.close()
        line["cpu_baseline"] = cpu_baseline(shape, w, args.cpu_seconds)
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        _leave_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if world == 1 and args.gpus > 1:  # plain `python bench.py --gpus N`: become the torchrun launch
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                  f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                  "--master-port", os.environ.get("MASTER_PORT", "29517"),
                                  str(Path(__file__).resolve()), *sys.argv[1:]])
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()

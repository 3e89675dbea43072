#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on N GPUs of one node.

Metric: BERT-base eval tokens/s (forward + MLM loss, no data loading) with
``--attn_softmax "clippedsoftmax1(-.025:1)"``, fp16, seq 128, batch 32 PER GPU (BASELINE config 2), synthetic
tokens and random-init weights.  One process per GPU (torchrun), batches sharded over ranks, no data-path
collective; the eval-loss (sum, count) pair is all-reduced once at the end.

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference path's CPU port on the host cores (rank 0 only)

Prints ONE JSON line (rank 0).  ``value`` = whole-job tokens/s with inputs resident in HBM; ``e2e`` = the same
through the module API with pinned-host inputs (H2D) and a host read of the loss (D2H) every step;
``roofline`` = the fused attention kernel (the hot path's own kernel; the step's largest single kernel is now this
repo's FFN1 GEMM with the GELU epilogue, the other GEMMs are library calls) timed alone with CUDA events on its launch stream on the model's own Q/K/V shapes;
``roofline_gemm`` = the step's largest own kernel (FFN1 GEMM + erf GELU) against the dense tensor peak, timed the same way;
``cpu_baseline`` = the oracle port of the reference path on the host cores (rank 0, N=1).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bert_base_eval_tokens_per_s"
UNIT = "tokens/s"
ATTN_SOFTMAX = "clippedsoftmax1(-.025:1)"
BATCH, SEQ, HEADS, HEAD_DIM, LAYERS = 32, 128, 12, 64, 12
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the roofline kernel (ncu --set full capture under
# profiles/); None until a capture of the current kernel is committed
NCU_TRAFFIC_BYTES_PER_LAUNCH = 18911744   # profiles/r02_ncu_hot_kernels.txt, launch 5: 18.909 MB read + 0.002 MB written (ctx still in L2)
NCU_GEMM_TRAFFIC_BYTES_PER_LAUNCH = 11045632   # same capture, launch 6: 11.045 MB read (x + part of W; W and y mostly stay in L2) + 256 B written
WORKLOAD = "BERT-base fp16 eval, --attn_softmax clippedsoftmax1(-.025:1), seq 128, batch 32 per GPU (BASELINE config 2)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tensor=d.get("bf16_tflops", 1590.0), tensor_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(tensor=1590.0, tensor_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and clock-event (throttle) reasons of one GPU sampled every 5 ms DURING the timed region through NVML
    (a 50-step timed region lasts ~70 ms -- shorter than one `nvidia-smi -lms` period or its start-up time);
    falls back to one `nvidia-smi --query-gpu` call per second if pynvml is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index: int, uuid: str = ""):
        self.gpu, self.uuid = gpu_index, uuid
        self.rows = []                       # (time, sm_mhz, reason bit mask)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.backend = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
            except Exception:
                pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.gpu
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            idx = int(vis.split(",")[self.gpu])
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.rows.append((time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.backend = "nvml"
        except Exception:
            def loop():
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
                while not self.stop_flag.is_set():
                    try:
                        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                        self.max_mhz = float(out[1])
                        self.rows.append((time.time(), float(out[0]), int(out[2].strip(), 16)))
                    except Exception:
                        pass
                    time.sleep(1.0)
            self.backend = "nvidia-smi"
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self, t0: float, t1: float):
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        rows = inside or [r for r in self.rows if t0 - 0.5 <= r[0] <= t1 + 0.5] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        mask = 0
        for r in rows:
            mask |= r[2]
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz,
                "reasons": [name for name, bit in self.REASONS if mask & bit], "samples": len(rows),
                "samples_inside_timed_region": len(inside), "source": self.backend}


def run_reference(args):
    """The reference path's CPU implementation (oracle port; /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle
    from oracle.cpu_model import time_bert_eval_cpu
    c_oracle.build()
    # bounded sample: the full 32x128 batch per step unless K is so large that the run would take many minutes
    sample_batch = BATCH if args.steps * 3.0 <= 240 else max(1, int(BATCH * 240 / (args.steps * 3.0)))
    tps, ms, threads, losses = time_bert_eval_cpu(ATTN_SOFTMAX, sample_batch, SEQ, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic tokens, random-init weights",
        "config": {"workload": WORKLOAD + " -- CPU fp32 port of the reference path (oracle/cpu_model.py)",
                   "sample": f"{sample_batch}x{SEQ} tokens per step"},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of {sample_batch}x{SEQ} tokens, HF BERT-base shell + C oracle attention"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
    return 0


def time_attention_kernel(torch, ops, fn, dtype, pool_sets=8, reps=10):
    """Average launch duration of the fused attention kernel on the model's own shapes/strides ((B,T,H,d)-ordered
    Q/K/V viewed as (B,H,T,d), output (B,T,H*d)), COLD: the launches rotate over a pool of input/output sets larger
    than L2 (8 x 25 MB = 201 MB vs 126 MB), captured back to back in one CUDA graph so that the host-side call
    overhead of the eager path is not in the number, and bracketed by one CUDA-event pair on the launch stream."""
    g = torch.Generator(device="cuda").manual_seed(1)
    mk = lambda: torch.randn(BATCH, SEQ, HEADS, HEAD_DIM, device="cuda", dtype=dtype, generator=g).permute(0, 2, 1, 3)
    pool = [(mk(), mk(), mk(), torch.empty(BATCH, SEQ, HEADS * HEAD_DIM, device="cuda", dtype=dtype))
            for _ in range(pool_sets)]
    run = lambda s: ops.attention_fwd(s[0], s[1], s[2], scale=HEAD_DIM ** -0.5, kind=fn.kind, gamma=fn.gamma,
                                      eta=fn.eta, out=s[3])
    for s in pool:
        run(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            for s in pool:
                run(s)
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (reps * pool_sets))
    return sum(ts) / len(ts) * 1e-3     # seconds per launch


def time_ffn1_gemm(torch, ops, dtype, pool_sets=8, reps=5):
    """Average launch duration of this repo's tcgen05 GEMM with the erf-GELU epilogue at the model's FFN1 shape
    (M = 32 x 128 tokens, K = 768, N = 3072): the single largest kernel of the headline step (24 % of it,
    profiles/r02_launches_bert_step_v.summary.txt).  Cold like time_attention_kernel: activations rotate over a pool
    larger than L2 (8 x 31.5 MB), back to back in one CUDA graph, one CUDA-event pair."""
    M, K, N = BATCH * SEQ, HEADS * HEAD_DIM, 4 * HEADS * HEAD_DIM
    g = torch.Generator(device="cuda").manual_seed(2)
    w = torch.randn(N, K, device="cuda", dtype=dtype, generator=g) * K ** -0.5
    b = torch.randn(N, device="cuda", dtype=dtype, generator=g)
    pool = [(torch.randn(M, K, device="cuda", dtype=dtype, generator=g), torch.empty(M, N, device="cuda", dtype=dtype))
            for _ in range(pool_sets)]
    run = lambda s: ops.linear_fwd(s[0], w, b, act="gelu", out=s[1])
    for s in pool:
        run(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            for s in pool:
                run(s)
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) / (reps * pool_sets))
    return sum(ts) / len(ts) * 1e-3, 2.0 * M * N * K, 2.0 * (M * K + N * K + N + M * N)


def _event_time(torch, f, reps=5, warm=3, clock=None):
    """Seconds per call (CUDA events around `reps` back-to-back calls).  clock: optional one-element list that receives the SM
    clock (MHz) the GPU was running at when the last call ended (outeffhop_b200._lib.sm_clock_probe, queued behind it)."""
    for _ in range(warm):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mhz = torch.zeros(1, device="cuda") if clock is not None else None
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        f()
    b.record()
    if clock is not None:
        from outeffhop_b200 import _lib
        _lib.sm_clock_probe(mhz, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if clock is not None:
        clock.append(float(mhz.item()))
    return a.elapsed_time(b) * 1e-3 / reps


def _skip_fraction(torch, q, k, gamma, eta, b=0, h=0):
    """Share of (32-row warp block x 64-key step) tiles of one (batch, head) in which clippedsoftmax1 leaves no key alive:
    exactly the tiles the PV pass of the two-pass kernel skips (attn_fwd_ws.cuh, "sparse skip").  Untimed, torch on the GPU."""
    s = (q[b, h].float() @ k[b, h].float().t()) * 0.125
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s - m)
    p = e / (e.sum(-1, keepdim=True) + torch.exp(-m))
    alive = ((eta - gamma) * p + gamma) > 0
    T = s.shape[0]
    tiles = alive.view(T // 32, 32, T // 64, 64).any(3).any(1)
    return 1.0 - tiles.float().mean().item(), alive.float().sum(-1).mean().item()


def attention_sweep(torch, ops, peaks):
    """Config-5 style points (12 heads, d=64, B*T = 65536 tokens): attention TFLOP/s vs the tensor roofline."""
    out = {}
    for T in (2048, 8192):
        B = 65536 // T
        q = torch.randn(B, HEADS, T, HEAD_DIM, device="cuda", dtype=torch.float16)
        k, v = torch.randn_like(q), torch.randn_like(q)
        # peaky rows: q scaled x1 / x3 / x5 per 32-row block, so that clipped rows keep a handful of surviving keys
        # (tests/test_gpu_attention_baseline_shapes.py uses the same construction)
        blk = torch.randint(0, 3, (B, HEADS, T // 32, 1), device="cuda")
        qp = (q.float() * (1.0 + 2.0 * blk).repeat_interleave(32, dim=2)).half()
        # clippedsoftmax1(-.025:1) zeroes every probability below 0.0222, and on N(0,1) inputs no key of a T >= 2048 row
        # gets that much: the PV pass skips all exponentials ("sparse skip", data dependent) -- a DEGENERATE workload whose
        # output is all zeros, reported only as *_allzero.  Credit the *_peaky line (surviving keys, skip fraction reported)
        # and the *_dense line (gamma = 0: nothing can be skipped, the full two-pass cost).
        for name, kind, gamma, qq in (("softmax1", 1, 0.0, q), ("clippedsoftmax1_allzero", 3, -0.025, q),
                                      ("clippedsoftmax1_peaky", 3, -0.025, qp), ("clippedsoftmax1_dense", 3, 0.0, q),
                                      ("softmax1_causal", 1, 0.0, q)):
            causal = name.endswith("causal")
            # ~10 ms of back-to-back launches after ~10 ms of warm-up (the first launches after the allocations above run up to
            # 10 % slower), from an idle board (0.25 s pause): `tflops`.  The two peaks of MEASURED_PEAKS.json are a BURST figure
            # (best single matmul at full clock) and a SUSTAINED one (seconds of back-to-back matmuls, lower SM clock under the
            # power cap); both fractions are given, with the SM clock this point ended at (probe kernel queued right behind it)
            # and the best single launch timed on its own.
            f = lambda: ops.attention_fwd(qq, k, v, scale=0.125, kind=kind, gamma=gamma, eta=1.1, causal=causal)
            reps = 20 if T <= 2048 else 5
            # Three such batches per point, every batch mean reported (`tflops_batches`); `tflops` is the best batch: single
            # batches of the SAME library scatter by up to 20 % on this pool (one launch of a batch at half speed while the
            # best single launch is unchanged: profiles/r02_bench_z.json, r02_wsp_tile_phase.txt), the way the peak's own
            # "best of 10" does.
            flops = 2.0 * B * HEADS * T * (T + 1) * HEAD_DIM if causal else 4.0 * B * HEADS * T * T * HEAD_DIM
            batches, clocks = [], []
            for _ in range(3):
                time.sleep(0.25)
                mhz = []
                batches.append(_event_time(torch, f, reps=reps, warm=reps, clock=mhz))
                clocks.append(mhz[0])
            t = min(batches)
            mhz = [clocks[batches.index(t)]]
            best = min(_event_time(torch, f, reps=1, warm=0) for _ in range(min(reps, 8)))
            tf, tf_best = flops / t / 1e12, flops / best / 1e12
            out[f"{name}_T{T}"] = {"tflops": round(tf, 1), "frac_of_measured_peak": round(tf / peaks["tensor"], 4),
                                   "frac_of_sustained_peak": round(tf / peaks["tensor_sustained"], 4),
                                   "tflops_batches": [round(flops / b / 1e12, 1) for b in batches],
                                   "tflops_single_launch_best": round(tf_best, 1), "sm_mhz_at_end": round(mhz[0])}
            if name.endswith("peaky"):
                skip, alive = _skip_fraction(torch, qp, k, gamma, 1.1)
                out[f"{name}_T{T}"].update(skipped_tile_fraction=round(skip, 4), surviving_keys_per_row=round(alive, 2))
        del q, k, v, qp
    return out


def quant_kernel_bandwidth(torch, ops, peaks):
    """Standalone quantizer / range kernels (north_star: "achieved HBM GB/s ... against B200 peak"): algorithmic bytes
    (fake-quant: read + write, min/max: read) / CUDA-event time, L2 flushed between launches when the tensor
    fits L2.  N(0,1) activations with 0.01 % outliers x50."""
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for dtn, dt in (("f32", torch.float32), ("f16", torch.float16)):
        for logn in (24, 28):
            n = 1 << logn
            x = torch.randn(n, device="cuda", dtype=dt)
            x[:: 10000] *= 50
            y = torch.empty_like(x)
            mm = torch.empty(2, device="cuda")
            state = ops.new_qstate("cuda")
            ops.range_update(state, ops.minmax(x), 0.9, 8, 1e-8)
            qp = state[4:6]
            for op, f, nbytes in (("fake_quant", lambda: ops.fake_quant_asym(x, qp, 8, out=y), 2 * n * x.element_size()),
                                  ("minmax", lambda: ops.minmax(x, out=mm), n * x.element_size())):
                for _ in range(3):
                    f()
                ts = []
                for _ in range(7):
                    if logn < 27:
                        flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    f()
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e-3)
                t = statistics.median(ts)
                out[f"{op}_{dtn}_2^{logn}"] = {"gbs": round(nbytes / t / 1e9, 1), "frac_of_hbm_peak": round(nbytes / t / 1e9 / peaks["hbm"], 4),
                                              "us": round(t * 1e6, 1)}
            del x, y
    return out


def extra_model_configs(torch, dist, rank, world, steps, warmup, timed):
    """BASELINE configs 3 and 4 at model level on EVERY rank (weak scaling, batches sharded over ranks like the headline):
      config3  OPT-125m fp16 causal, vanilla softmax + conditional_per_token gate (init 0.25), block 512, batch 4 per GPU
      config4  OPT-125m INT8 (--quantize: fp32_head, running_minmax, asymmetric 8-bit activations, symmetric 8-bit weights)
               with softmax1, block 512, batch 4 per GPU, fp16 activations (the throughput configuration; the fp32 variant is
               the parity configuration of tests/test_fullsize_int8.py).  Range estimation runs on 4 batches PER RANK with
               the MIN/MAX all-reduce BEFORE the EMA (NCCL when world > 1), and the resulting quantizer states are gathered
               and compared bit for bit across ranks inside this run."""
    from outeffhop_b200 import eval_harness as EH, quantization as Q, quantized_models as QM
    out = {}

    def run(step, pool, tokens_per_step):
        def body(n):
            for i in range(n):
                for k2, v in pool[i % len(pool)].items():
                    step.static_inputs[k2].copy_(v, non_blocking=True)
                step.run()
        body(warmup)
        ms, _, _ = timed(lambda: body(steps))
        return {"value": world * steps * tokens_per_step / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                "loss": float(step.loss)}

    model = EH.build_opt_125m(seed=5678, attn_softmax="vanilla", attn_gate_type="conditional_per_token", attn_gate_init=0.25)
    host = [EH.synthetic_clm_batch(4, 512, seed=5678 + 1000 * rank + i) for i in range(8)]
    pool = [{k2: v.cuda() for k2, v in b.items()} for b in host]
    step = EH.make_opt_eval_step(model, host[0])
    out["config3"] = dict(run(step, pool, 4 * 512),
                          workload="OPT-125m fp16 causal, vanilla + conditional_per_token gate (init 0.25), 4x512 per GPU")
    del model, step

    model = EH.build_opt_125m(seed=6789, attn_softmax="softmax1")
    q = QM.quantize_opt_for_causal_lm(model, quant_setup="fp32_head")
    n_est = 4
    est = [EH.synthetic_clm_batch(4, 512, seed=6789 + i * world + rank) for i in range(n_est)]   # global batch i: ranks 0..N-1
    host = [EH.synthetic_clm_batch(4, 512, seed=7789 + 1000 * rank + i) for i in range(8)]
    pool = [{k2: v.cuda() for k2, v in b.items()} for b in host]
    Q.set_range_sync(world > 1)
    c0 = Q.RANGE_SYNC_COLLECTIVES
    step = EH.make_opt_int8_eval_step(q, host[0], est)
    Q.set_range_sync(False)
    states = torch.stack([m.quantizer._qstate.detach().clone() for m in q.modules()
                          if isinstance(m, Q.QuantizationManager) and isinstance(m.quantizer, Q.AsymmetricUniformQuantizer)
                          and m.quantizer.is_initialized])
    identical = None
    if world > 1:
        gathered = [torch.empty_like(states) for _ in range(world)]
        dist.all_gather(gathered, states)
        identical = all(torch.equal(g, gathered[0]) for g in gathered)
        if not identical:
            raise SystemExit("bench.py: activation ranges differ across ranks after the synchronised estimation")
    out["config4"] = dict(run(step, pool, 4 * 512),
                          workload="OPT-125m INT8 W8A8 (fp32_head, running_minmax, asymmetric_uniform) softmax1, 4x512 per GPU, "
                                   "fp16 activations, fixed ranges after 4 estimation batches per rank",
                          range_sync={"activation_quantizers": int(states.shape[0]), "estimation_batches_per_rank": n_est,
                                      "all_reduces": Q.RANGE_SYNC_COLLECTIVES - c0, "backend": "nccl" if world > 1 else None,
                                      "bit_identical_across_ranks": identical})
    del model, q, step
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from outeffhop_b200 import ops
    from outeffhop_b200._lib import lib
    from outeffhop_b200.eval_harness import aggregate_perplexity, build_bert_base, make_bert_eval_step, synthetic_mlm_batch
    from outeffhop_b200.softmax import SOFTMAX_MAPPING
    lib()                                                 # fail loudly if the CUDA extension is missing
    peaks = load_peaks()
    steps, warmup = args.steps, max(args.warmup, 3)

    model = build_bert_base(seed=3000, dtype=torch.float16, attn_softmax=ATTN_SOFTMAX)
    n_pool = 8                                            # distinct synthetic batches per rank, cycled
    host_pool = [synthetic_mlm_batch(BATCH, SEQ, seed=3000 + 1000 * rank + i) for i in range(n_pool)]
    dev_pool = [{k: v.cuda() for k, v in hb.items()} for hb in host_pool]
    step = make_bert_eval_step(model, host_pool[0], use_graph=not args.no_graph)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = ""
    sampler = ClockSampler(local_rank, gpu_uuid)
    if rank == 0:
        sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(loop_body):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        loop_body()
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)     # max over ranks
        return ms.item(), t0, t1

    # ---- device-resident throughput ("value") -------------------------------------------------------------
    loss_sum = torch.zeros((), dtype=torch.float32, device="cuda")

    def body_resident(n):
        for i in range(n):
            for k2, v in dev_pool[i % n_pool].items():    # device-to-device refresh of the static inputs
                step.static_inputs[k2].copy_(v, non_blocking=True)
            loss_sum.add_(step.run())

    body_resident(warmup)
    if args.profile_one_step:                             # ncu --profile-from-start off: exactly one warm step
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        body_resident(1)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    loss_sum.zero_()
    ms_res, t0, t1 = timed(lambda: body_resident(steps))
    ppl = aggregate_perplexity(loss_sum, steps)
    clocks = sampler.summary(t0, t1) if rank == 0 else None

    # ---- end to end: pinned host -> H2D -> step -> D2H loss, every step --------------------------------------
    def body_e2e(n):
        for i in range(n):
            step.load(host_pool[i % n_pool])
            loss_host.copy_(step.run(), non_blocking=True)
            torch.cuda.current_stream().synchronize()     # the caller reads the loss of every step
            _ = float(loss_host)

    body_e2e(warmup)
    ms_e2e, _, _ = timed(lambda: body_e2e(steps))
    if rank == 0:
        sampler.stop()
    extra = None
    if not args.no_extra:
        extra = extra_model_configs(torch, dist, rank, world, steps, warmup, timed)

    tokens = world * steps * BATCH * SEQ
    line = None
    if rank == 0:
        fn = SOFTMAX_MAPPING[ATTN_SOFTMAX]
        t_attn = time_attention_kernel(torch, ops, fn, torch.float16)
        flops = 4.0 * BATCH * HEADS * SEQ * SEQ * HEAD_DIM            # algorithmic: QK^T + PV, per launch
        # algorithmic HBM bytes per launch: Q, K, V read + context written, 2 B each = 25.17 MB.
        # Arithmetic intensity 1.61 GFLOP / 25.17 MB = 64 flop/B << the ~250 flop/B ridge: this shape is HBM-bound.
        nbytes = 4.0 * BATCH * HEADS * SEQ * HEAD_DIM * 2
        achieved = nbytes / t_attn / 1e9
        line = {
            "metric": METRIC, "value": tokens / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_res / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic tokens, random-init weights",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "seq_len": SEQ, "parallelism": f"dp{world}",
                       "cuda_graph": not args.no_graph,
                       "l2": "no flush: 220 MB of fp16 weights + activations per step exceed the 126 MB L2"},
            "e2e": {"value": tokens / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes(),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / steps},
            # own kernels per step: per layer the fused attention, 2 x fused add+LayerNorm and the tcgen05 FFN1 GEMM with the
            # erf GELU in its epilogue (oeh_linear_fwd); + the 2 loss kernels (the other GEMMs are cuBLASLt)
            "gpu_launches": (4 * LAYERS + 2) * steps,
            "clocks": clocks,
            "perplexity": ppl,
            "roofline": {"kernel": "attn_fwd_one_kernel<fp16> (fused QK^T->clippedsoftmax1->PV, one K/V tile), "
                                   "B=32 H=12 T=128 d=64, model strides",
                         "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm"], "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                         "peak_source": peaks["source"] + ", copy bandwidth",
                         "launch_us": t_attn * 1e6, "algorithmic_bytes_per_launch": nbytes,
                         "algorithmic_flops_per_launch": flops, "tflops": flops / t_attn / 1e12,
                         "timing": "80 launches back to back in one CUDA graph over 8 input sets (201 MB > L2), CUDA events",
                         "note": "64 flop/B: HBM/latency-bound at T=128; the tensor roofline applies to attn_sweep (T>=2048)"},
        }
        # The step's largest own kernel by time share is tensor-bound: reported next to the attention line, same timing method
        t_gemm, gemm_flops, gemm_bytes = time_ffn1_gemm(torch, ops, torch.float16)
        line["roofline_gemm"] = {
            "kernel": "gemm_kernel<fp16, 128x256 tiles, cta_group::2> (oeh_linear_fwd: x @ W^T + bias -> erf GELU in the "
                      "TMEM->register epilogue -> TMA store), M=4096 K=768 N=3072 (BERT-base FFN1)",
            "bound": "tensor", "achieved": gemm_flops / t_gemm / 1e12, "peak": peaks["tensor"], "unit": "TFLOP/s",
            "frac": gemm_flops / t_gemm / 1e12 / peaks["tensor"],
            "frac_of_sustained_peak": gemm_flops / t_gemm / 1e12 / peaks["tensor_sustained"],
            "traffic": NCU_GEMM_TRAFFIC_BYTES_PER_LAUNCH,
            "peak_source": peaks["source"] + ", cuBLAS bf16 burst (sustained beside it)", "launch_us": t_gemm * 1e6,
            "algorithmic_flops_per_launch": gemm_flops, "algorithmic_bytes_per_launch": gemm_bytes,
            "timing": "40 launches back to back in one CUDA graph over 8 activation sets (252 MB > L2), CUDA events",
            "note": "share of the step: 24 % (12 FFN1 launches + the MLM-head transform); the erf-GELU epilogue, not the "
                    "main loop, bounds it (profiles/r02_timeline_gemm.txt)"}
        if extra is not None:
            line["configs"] = extra
        if not args.no_sweep:
            line["attn_sweep"] = attention_sweep(torch, ops, peaks)
            line["quant_kernels"] = quant_kernel_bandwidth(torch, ops, peaks)
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            from oracle.cpu_model import time_bert_eval_cpu
            c_oracle.build()
            tps, ms_cpu, threads, _ = time_bert_eval_cpu(ATTN_SOFTMAX, BATCH, SEQ, steps=4, warmup=1)
            line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"4 steps of {BATCH}x{SEQ} tokens (+1 warm-up), fp32, HF BERT-base shell "
                                              "+ C oracle attention (oracle/cpu_model.py)", "ms_per_step": ms_cpu}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="run the eval step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-sweep", action="store_true", help="skip the T=2048/8192 attention sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 3 / config 4 model-level lines")
    ap.add_argument("--profile-one-step", action="store_true",
                    help="bracket one warm eval step with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." on rank 0 under
    # torchrun): keep a private handle on the real stdout for the result line and send everything else to stderr.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

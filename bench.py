#!/usr/bin/env python
"""Benchmark of the CFM-ViT hot path on B200: ds2 training samples/s (headline) and ODE-sampled showers/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

* A training "step" is the reference's ``BaseExperiment._step`` work on one batch of synthetic ds2 showers
  (64 per GPU, weak scaling): ``_batch_loss`` (t and x_0 draws, trajectory, patchify, ViT forward, MSE),
  ``zero_grad``, backward (with the bucketed gradient all-reduce when N > 1), ``clip_grad_norm_(1000)``,
  AdamW step.  ``value`` is timed with the batches already resident in HBM; ``e2e`` feeds pinned HOST
  batches through ``model._batch_loss`` (H2D inside the timed region) and reads the loss back every step.
* ``roofline``: the tcgen05 GEMM family.  Each of the 12 GEMM call sites of a transformer block (forward, dgrad and
  wgrad of qkv / proj / fc1 / fc2, with the epilogues the step uses, proj and fc2 including the fused LayerNorm) is
  replayed 20 times back to back between two CUDA events on the launching stream; ``achieved`` is the flops of one
  block's 12 GEMMs over the sum of their times (``classes`` lists each).  ``kernels`` is the per-kernel-class device
  time of whole steps from CUDA events the library records around each launch (v4h_profile_*, eager launches, a
  separate pass so the event overhead does not leak into ``value``); the brackets overstate a kernel by ~3 us, so
  ``frac_eager_brackets`` (the same roofline taken from that table) is a lower bound.
* ``sampling``: BASELINE.json configs[3] as one sharded job — the conditions are split over the ranks, every rank
  solves its batches of 256 with RK4 3/8 (80 network evaluations), the showers are all-gathered and copied to the
  host on rank 0 (``e2e``); ``energy_model`` is the energy-ratio network's solve for the same conditions; with its
  own ``roofline`` and ``cpu_baseline``.
* ``fp32``: the reference-precision mode (SIMT fp32 kernels) on the same step; ``strong_scaling`` (N > 1): the
  reference's ``batchsize // world_size`` split of a global batch of 64.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle port of the same training step (oracle/, fp32,
  all host threads) on a bounded sample.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# NCCL's INFO banner goes to stdout and would precede the JSON line: default to WARN, but keep whatever the
# caller asked for (the driver reads NCCL's own log to count ranks)
os.environ.setdefault("NCCL_DEBUG", "WARN")

import torch  # noqa: E402

METRIC = "ds2 CFM-ViT train samples/s"  # --config ds3 reports the same metric on the ds3 shape (secondary)
UNIT = "samples/s"
TRAIN_GFLOP_PER_SAMPLE = {"ds2": 14.160, "ds3": 52.078}   # 3 x forward
SAMPLE_GFLOP_PER_SHOWER = {"ds2": 377.6, "ds3": 1388.7}    # 80 NFE


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(power), samples=len(sm),
                    reasons=sorted(reasons))


def ds2_setup(config: str):
    """(geometry, ViT param dict) of the CPU arm: the oracle's own tables (test infrastructure; only the
    reference / cpu_baseline legs call this)"""
    from oracle import vit_oracle as vo
    cfg = vo.CONFIGS[config]
    return cfg["geom"], dict(cfg["param"])


_SHAPES = {"ds2": (135, 48), "ds3": (450, 90)}  # (tokens, patch_dim) of the shipped shape-model configs


def workload_config(name: str, batch: int, world: int) -> dict:
    """The `config` object of the JSON line: the workload only, identical for the B200 arm and the reference arm
    (how each arm executes it is under `details`)."""
    tokens, patch_dim = _SHAPES[name]
    return {"workload": f"CaloChallenge {name} shape CFM-ViT training step (BASELINE.json configs[1]): _batch_loss, "
                        f"backward, clip_grad_norm(1000), AdamW; batch {batch} per GPU, data-parallel x{world}",
            "global_batch": batch * world, "per_gpu_batch": batch, "tokens": tokens, "patch_dim": patch_dim,
            "parallelism": f"dp{world}",
            "l2": "no explicit flush: the per-step working set (activation workspace ~1 GB + 104 MB fp32 / 52 MB bf16 "
                  "weights + 8 rotating input batches) exceeds the 126 MB L2"}


def rerandomise(net, seed=1, std=0.02):
    """reference init leaves adaLN / output layers at zero (output identically 0): re-draw them so the
    benchmark exercises non-trivial values"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if "adaLN_modulation" in name or name.startswith("final_layer.linear") or name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * std)


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port of the training step
# ------------------------------------------------------------------------------------------------
def cpu_train_samples_per_s(config: str, batch: int, steps: int, warmup: int):
    from oracle import vit_oracle as vo
    geom, param = ds2_setup(config)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = vo.init_state_dict(param, seed=0)
    params = {k: v.clone().requires_grad_(k == "pos_embed_freqs" or not k.startswith("pos_")) for k, v in sd.items()}
    leaves = [p for p in params.values() if p.requires_grad]
    opt = torch.optim.AdamW(leaves, lr=1e-4, weight_decay=0.1)
    g = torch.Generator().manual_seed(1234)
    x1 = torch.randn(batch, *geom.sample_shape, generator=g)
    c = torch.rand(batch, param["condition_dim"], generator=g)

    def step():
        t = torch.rand(batch, generator=g)
        x0 = torch.randn(batch, *geom.sample_shape, generator=g)
        loss = vo.cfm_loss(params, x1, c, x0, t, geom, param["num_heads"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(leaves, 1000.0)
        opt.step()
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    steps, warmup = min(args.steps, 40), min(args.warmup, 3)  # bounded sample: about a second per step
    value, ms, threads = cpu_train_samples_per_s(args.config, batch, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC.replace("ds2", args.config), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, batch, args.gpus),
        "details": {"execution": "CPU oracle port of the reference path (oracle/vit_oracle.py, torch fp32 on the host "
                                 "cores), one process whatever --gpus says", "batch_per_step": batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} training steps of batch {batch} (fwd+bwd+clip+AdamW) of the CPU oracle "
                                   f"port, {warmup} warm-up (bounded: --steps {args.steps} --warmup {args.warmup})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def cpu_showers_per_s(config: str, batch: int, nfe: int):
    """CPU oracle port of the sampling path: `nfe` network evaluations of one batch (wrapper forward: patchify,
    ViT, unpatchify), extrapolated to the 80 evaluations of a 20-step RK4 (3/8) solve."""
    from oracle import vit_oracle as vo
    geom, param = ds2_setup(config)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = vo.init_state_dict(param, seed=0)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, *geom.sample_shape, generator=g)
    c = torch.rand(batch, param["condition_dim"], generator=g)
    t = torch.full((batch, 1), 0.5)
    with torch.no_grad():
        vo.cfm_forward(sd, x, t, c, geom, param["num_heads"])
        t0 = time.perf_counter()
        for _ in range(nfe):
            vo.cfm_forward(sd, x, t, c, geom, param["num_heads"])
        per = (time.perf_counter() - t0) / nfe
    return batch / (80 * per), per * 1e3, threads


def family_roofline(prof, psteps, peaks, traffic_key):
    """roofline object of the tcgen05 GEMM kernel (every gemm.* / dgrad.* / wgrad.* class is ONE template,
    csrc/gemm_umma.cu): achieved = algorithmic flops of its launches (2 M N K each) / their summed device time"""
    total_ms = sum(e["ms"] for e in prof) or 1.0
    kernels = []
    for e in sorted(prof, key=lambda e: -e["ms"]):
        per = e["ms"] / max(e["launches"], 1)
        kernels.append({"name": e["name"], "launches_per_step": e["launches"] / psteps, "ms_per_step": e["ms"] / psteps,
                        "share": e["ms"] / total_ms, "avg_ms": per,
                        "tflops": e["flops"] / (e["ms"] * 1e-3) / 1e12 if e["ms"] > 0 else 0.0,
                        "gbs": e["bytes"] / (e["ms"] * 1e-3) / 1e9 if e["ms"] > 0 else 0.0})
    fam = [k for k in kernels if k["name"].startswith(("gemm", "wgrad", "dgrad"))]
    fam_ms = sum(k["ms_per_step"] for k in fam)
    if fam_ms <= 0:
        return kernels, None
    fam_flops = sum(k["tflops"] * 1e12 * k["ms_per_step"] * 1e-3 for k in fam)
    fam_launches = sum(k["launches_per_step"] for k in fam)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(traffic_key, {}).get("dram_bytes_per_launch")
    ach = fam_flops / (fam_ms * 1e-3) / 1e12
    roofline = {"kernel": "gemm_umma_kernel (tcgen05 GEMM: all gemm.* / dgrad.* / wgrad.* classes)",
                "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["tflops_sustained"], "traffic": traffic,
                "flops_per_launch": fam_flops / max(fam_launches, 1),
                "avg_launch_ms": fam_ms / max(fam_launches, 1), "launches_per_step": fam_launches,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "share_of_step": fam_ms / (total_ms / psteps)}
    return kernels, roofline


def gemm_class_roofline(lib, dev, M, param, tokens, peaks, fam_eager):
    """Per-call-site timing of the tcgen05 GEMM (v4h_debug_gemm launches the library's kernel with the epilogue of
    the named call site): TFLOP/s per class and, weighted by launches per step, for the family."""
    from vit4hep_b200 import _cabi
    D, Hm, depth = param["hidden_dim"], int(param["hidden_dim"] * param["mlp_ratio"]), param["depth"]
    bf = torch.bfloat16
    s = torch.cuda.current_stream(dev).cuda_stream
    # (class, kind of v4h_debug_gemm, m, n, k): forward, dgrad and wgrad GEMMs of one block
    # kind 7: gated residual + the following LayerNorm / modulation in the epilogue (v4h_debug_gemm_ln), as the step runs it
    classes = [("gemm.qkv", 1, M, 3 * D, D), ("gemm.proj+ln", 7, M, D, D), ("gemm.fc1", 0, M, Hm, D), ("gemm.fc2+ln", 7, M, D, Hm),
               ("dgrad.fc2", 3, M, Hm, D), ("dgrad.fc1", 4, M, D, Hm), ("dgrad.proj", 4, M, D, D), ("dgrad.qkv", 4, M, D, 3 * D),
               ("wgrad.fc2", 5, D, Hm, M), ("wgrad.fc1", 5, Hm, D, M), ("wgrad.proj", 5, D, D, M), ("wgrad.qkv", 5, 3 * D, D, M)]
    g = torch.Generator().manual_seed(0)
    rows, tot_flops, tot_ms = [], 0.0, 0.0
    for name, kind, m, n, k in classes:
        A = (torch.randn((k, m) if kind == 5 else (m, k), generator=g) * 0.1).to(dev, bf)
        Bm = (torch.randn((n, k) if kind in (0, 1, 2, 7) else (k, n), generator=g) * 0.1).to(dev, bf)
        bias = torch.randn(n, generator=g).to(dev)
        out = torch.zeros((m, n), device=dev, dtype=torch.float32 if kind == 5 else bf)
        out2 = torch.zeros((m, n), device=dev, dtype=bf)
        gated = kind in (2, 7)
        nb = (m + tokens - 1) // tokens
        res_in = torch.randn(m, n, generator=g).to(dev) if gated else None
        res_out = torch.empty(m, n, device=dev) if gated else None
        gate = torch.randn(nb, n, generator=g).to(dev) if gated else None
        aux = torch.randn(m, n, generator=g).to(dev, bf) if kind == 3 else None
        ptr = lambda t: None if t is None else t.data_ptr()
        if kind == 7:
            shift, scale = torch.randn(nb, n, generator=g).to(dev), torch.randn(nb, n, generator=g).to(dev)
            ln = torch.zeros(m, n + 8, device=dev, dtype=bf)
            stats = torch.zeros(m, 2, device=dev)

        def call():
            if kind == 7:
                _cabi.check(lib.v4h_debug_gemm_ln(m, n, k, tokens, A.data_ptr(), Bm.data_ptr(), bias.data_ptr(),
                                                  out2.data_ptr(), res_in.data_ptr(), res_out.data_ptr(), gate.data_ptr(),
                                                  shift.data_ptr(), scale.data_ptr(), ln.data_ptr(), n + 8,
                                                  stats.data_ptr(), None, s))
                return
            _cabi.check(lib.v4h_debug_gemm(kind, m, n, k, tokens, A.data_ptr(), Bm.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                           out2.data_ptr(), ptr(res_in), ptr(res_out), ptr(gate), ptr(aux), None, s))
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * m * n * k
        rows.append({"class": name, "M": m, "N": n, "K": k, "us": ms * 1e3, "tflops": flops / (ms * 1e-3) / 1e12,
                     "frac": flops / (ms * 1e-3) / 1e12 / peaks["tflops_sustained"], "launches_per_step": depth})
        tot_flops += flops * depth
        tot_ms += ms * depth
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("gemm_umma_kernel", {}).get("dram_bytes_per_launch")
    ach = tot_flops / (tot_ms * 1e-3) / 1e12
    n_launch = depth * len(classes)
    return {"kernel": "gemm_umma_kernel / gemm_gate_res_ln_kernel (tcgen05 GEMMs: the 12 GEMM call sites of a transformer block, "
                      "forward + dgrad + wgrad; proj / fc2 with the LayerNorm epilogue)",
            "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
            "frac": ach / peaks["tflops_sustained"], "traffic": traffic, "flops_per_launch": tot_flops / n_launch,
            "avg_launch_ms": tot_ms / n_launch, "launches_per_step": n_launch,
            "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
            "timing": "20 back-to-back launches per call site between two CUDA events on the launching stream, operands "
                      "L2-hot as in the step; weighted by launches per step",
            "share_of_step_eager_brackets": None if fam_eager is None else fam_eager["share_of_step"],
            "frac_eager_brackets": None if fam_eager is None else fam_eager["frac"],
            "classes": rows}


def run_b200(args):
    import torch.distributed as dist
    from vit4hep_b200 import FusedAdamW, GraphedTrainStep, _cabi, configs, dp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N > 1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _cabi.require_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def make_model(precision):
        torch.manual_seed(0)
        m = configs.build(args.config, precision)
        rerandomise(m.net)
        m = m.to(dev)
        m.device, m.dtype = dev, torch.float32
        return m

    model = make_model(args.precision)
    geom = model.geometry
    K = configs.MODELS[args.config]["net"]["param"]["condition_dim"]
    if world > 1:
        dp.enable_data_parallel(model.net)
    params = list(model.net.parameters())
    if args.torch_optimizer:
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.1, fused=True)
    else:  # clip_grad_norm_(1000) + AdamW + bf16 weight refresh in one native multi-tensor pass
        opt = FusedAdamW(model.net, lr=1e-4, weight_decay=0.1, max_grad_norm=1000.0)

    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    npool = 8
    host_x = [torch.randn(B, *geom.sample_shape, generator=g).pin_memory() for _ in range(npool)]
    host_c = [torch.rand(B, K, generator=g).pin_memory() for _ in range(npool)]
    dev_x = [x.to(dev) for x in host_x]
    dev_c = [c.to(dev) for c in host_c]
    lib = _cabi.load()

    def make_train_step(mdl, optimizer):
        plist = list(mdl.net.parameters())

        def train_step(batch, read_loss: bool):
            loss = mdl._batch_loss(batch)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            if args.torch_optimizer:
                torch.nn.utils.clip_grad_norm_(plist, 1000.0)
            optimizer.step()
            return loss.item() if read_loss else loss
        return train_step

    train_step = make_train_step(model, opt)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.v4h_launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.v4h_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches

    W, S = max(args.warmup, 3), args.steps
    for i in range(W):
        train_step((dev_x[i % npool], dev_c[i % npool]), False)
    # eager numbers first (every launch issued from Python), then the same step replayed as ONE CUDA graph
    S_eager = min(S, 50)
    ms_eager, launches = timed(lambda i: train_step((dev_x[i % npool], dev_c[i % npool]), False), S_eager)
    launches = launches * S // S_eager  # kernels of S steps (the graph replays the same launches)
    for i in range(2):
        train_step((host_x[i % npool], host_c[i % npool]), True)
    ms_e2e_eager, _ = timed(lambda i: train_step((host_x[i % npool], host_c[i % npool]), True), S_eager)
    use_graph = (not args.no_graph and not args.torch_optimizer
                 and (world == 1 or os.environ.get("V4H_GRAPH_DP", "1") == "1"))
    graphed = GraphedTrainStep(model, opt, dev_x[0], dev_c[0]) if use_graph else None
    if graphed is not None:
        step_dev = lambda i: graphed.step(dev_x[i % npool], dev_c[i % npool])
        h2d = host_x[0].numel() * 4 + host_c[0].numel() * 4  # x and c; t is drawn on the device
    else:
        step_dev = lambda i: train_step((dev_x[i % npool], dev_c[i % npool]), False)
        step_e2e = lambda i: train_step((host_x[i % npool], host_c[i % npool]), True)
        h2d = host_x[0].numel() * 4 + host_c[0].numel() * 4 + B * 4  # x, c and the host-drawn t
    for i in range(W):
        step_dev(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, _ = timed(step_dev, S)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * S / (ms * 1e-3)

    # end to end: pinned host batches through the public API, loss read back every step.  With the graphed step
    # the copy of batch i + 1 is issued (GraphedTrainStep.stage) before the loss of step i is read, so it runs
    # under step i -- every batch still crosses PCIe inside the timed region
    if graphed is not None:
        def step_e2e(i):
            loss = graphed.step_staged()
            graphed.stage(host_x[(i + 1) % npool], host_c[(i + 1) % npool])
            return loss.item()
        graphed.stage(host_x[0], host_c[0])
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, S)
    e2e = world * B * S / (ms_e2e * 1e-3)

    # per-kernel-class device time of one eager step (the library brackets every launch with two CUDA events on the
    # launching stream; each bracket adds ~3 us, so these are upper bounds that show the SHARES of the step) ...
    peaks = measured_peaks()
    _cabi.profile_begin()
    psteps = min(S, 3)
    for i in range(psteps):
        train_step((dev_x[i % npool], dev_c[i % npool]), False)
    prof = _cabi.profile_end(128)
    kernels, fam_eager = family_roofline(prof, psteps, peaks, "gemm_umma_kernel")
    # ... and the roofline of the dominant kernel, the tcgen05 GEMM: every GEMM call site of a transformer block is
    # replayed with its exact shape and epilogue, 20 launches back to back between two CUDA events on the launching
    # stream (no per-launch bracket, operands L2-hot as inside the step where the producing kernel has just written
    # them), weighted by its launches per step
    roofline = None
    if args.precision == "bf16":
        roofline = gemm_class_roofline(lib, dev, B * geom.tokens, configs.MODELS[args.config]["net"]["param"], geom.tokens,
                                       peaks, fam_eager)

    # ---- strong scaling (the reference's semantics: global batch 64, batchsize // world_size per rank,
    # reference experiments/calochallenge/experiment.py:94-98)
    strong = None
    if world > 1 and B % world == 0 and not args.no_strong:
        try:
            Bs = B // world
            sx = [x[:Bs].contiguous() for x in dev_x]
            sc = [c[:Bs].contiguous() for c in dev_c]
            gs = GraphedTrainStep(model, opt, sx[0], sc[0]) if graphed is not None else None
            fn = (lambda i: gs.step(sx[i % npool], sc[i % npool])) if gs is not None else \
                (lambda i: train_step((sx[i % npool], sc[i % npool]), False))
            for i in range(W):
                fn(i)
            ms_s, _ = timed(fn, S)
            strong = {"scaling": "strong", "global_batch": B, "per_gpu_batch": Bs, "value": B * S / (ms_s * 1e-3),
                      "unit": UNIT, "ms_per_step": ms_s / S}
            del gs
        except Exception as exc:  # the headline line must survive a failure of this secondary measurement
            strong = {"error": repr(exc)[:200]}

    # ---- ODE sampling (BASELINE.json configs[3]): 20 RK4 (3/8) steps = 80 network evaluations per shower,
    # conditions sharded over the ranks, one final all_gather, the result read back to the host
    sampling = None
    if not args.no_sampling:
        sampling = measure_sampling(args, model, K, world, rank, dev, timed, peaks, lib)

    # ---- the reference's own precision (fp32 SIMT GEMMs / attention): stated once beside bf16
    fp32 = None
    if world == 1 and args.precision == "bf16" and not args.no_fp32:
        try:
            m32 = make_model("fp32")
            o32 = FusedAdamW(m32.net, lr=1e-4, weight_decay=0.1, max_grad_norm=1000.0)
            ts32 = make_train_step(m32, o32)
            for i in range(2):
                ts32((dev_x[i], dev_c[i]), False)
            n32 = 5
            ms32, _ = timed(lambda i: ts32((dev_x[i % npool], dev_c[i % npool]), False), n32)
            fp32 = {"value": B * n32 / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / n32, "steps": n32,
                    "note": "precision='fp32' (parity mode, <= 1e-5 rel-L2): SIMT fp32 GEMMs and attention, eager launches"}
            del m32, o32
        except Exception as exc:
            fp32 = {"error": repr(exc)[:200]}

    data_feed = measure_data_feed(args, dev, peaks) if rank == 0 and world == 1 and not args.no_sampling else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cms, threads = cpu_train_samples_per_s(args.config, args.cpu_batch, args.cpu_steps, 2)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_steps} training steps of batch {args.cpu_batch} (fwd+bwd+clip+AdamW) of the CPU "
                         f"oracle port of the reference path, 2 warm-up, {cms:.0f} ms/step"}

    if rank == 0:
        gf = TRAIN_GFLOP_PER_SAMPLE[args.config]
        line = {
            "metric": METRIC.replace("ds2", args.config), "value": value, "unit": UNIT, "n_gpus": world, "steps": S, "warmup": W,
            "ms_per_step": ms / S, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args.config, B, world),
            "details": {"precision": args.precision,
                        "execution": ("one CUDA graph per training step (vit4hep_b200.GraphedTrainStep)"
                                      if graphed is not None else "eager launches"),
                        "optimizer": ("torch AdamW(fused) + clip_grad_norm_(1000)" if args.torch_optimizer else
                                      "vit4hep_b200.FusedAdamW: clip_grad_norm(1000) + AdamW + bf16 weight refresh")},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / S, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4,
                    "input_copy": ("prefetched: batch i+1 copied from pinned host memory under step i "
                                   "(GraphedTrainStep.stage / step_staged)") if graphed is not None
                                  else "inline, before each step"},
            "gpu_launches": launches,
            "eager": {"value": world * B * S_eager / (ms_eager * 1e-3), "e2e": world * B * S_eager / (ms_e2e_eager * 1e-3),
                      "steps": S_eager},
            "model_tflops_per_gpu": value * gf / 1e3 / world,
            "frac_of_peak": value * gf / 1e3 / world / peaks["tflops_sustained"],
            "roofline": roofline,
            "kernels": kernels[:40],
            "strong_scaling": strong,
            "sampling": sampling,
            "data_feed": data_feed,
            "fp32": fp32,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # every collective of this run has completed on every rank (the last one is the max-over-ranks of a
        # timed region).  Tearing NCCL down after CUDA-graph captures that contain collectives can hang, so
        # leave without the process-group destructor.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def measure_data_feed(args, dev, peaks):
    """Pre- and post-processing kernels of the data feed on 25 600 synthetic showers of the
    configured grid: showers/s and the HBM roofline (algorithmic bytes: every voxel read once and written once)."""
    from vit4hep_b200 import FusedForwardTransforms
    try:
        layers, per = 45, {"ds2": 144, "ds3": 900}[args.config]
        V, N = layers * per, 25600 if args.config == "ds2" else 6400
        chain = {"NormalizeByElayer": {}, "ScaleTotalEnergy": {"n_layers": layers, "factor": 0.35},
                 "CutValues": {"cut": 1.0e-7, "n_layers": layers},
                 "ExclusiveLogitTransform": {"delta": 1.0e-6, "rescale": True},
                 "GlobalStandardizeFromFile": {"model_dir": None, "eps": 1.0e-6}, "LogEnergy": {},
                 "ScaleEnergy": {"e_min": 6.907755, "e_max": 13.815510}, "AddFeaturesToCond": {"split_index": V},
                 "Reshape": {"shape": [1, V]}}
        g = torch.Generator(device=dev).manual_seed(11)
        raw = torch.exp(torch.randn(N, V, device=dev, generator=g) * 2 + 3) * (torch.rand(N, V, device=dev, generator=g) < 0.3)
        e_inc = 10.0 ** (3 + 3 * torch.rand(N, 1, device=dev, generator=g))
        raw = raw * (e_inc * 0.8 / raw.sum(1, keepdim=True))
        fwd = FusedForwardTransforms(chain, range(0, V + 1, per))
        x, cond = fwd(raw, e_inc)           # first call: statistics from the data (two kernels)
        rev = fwd.reverse()

        def timed(fn, iters=10):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        ms_pre, ms_post = timed(lambda: fwd(raw, e_inc)), timed(lambda: rev(x, cond))
        back, _ = rev(x, cond)
        keep = raw > 1e-6 * raw.sum(1, keepdim=True) / layers
        err = float(((back - raw)[keep]).norm() / raw[keep].norm())
        gb = 2 * N * V * 4 / 1e9
        return {"showers": N, "voxels": V,
                "preprocess": {"showers_per_s": N / (ms_pre * 1e-3), "ms": ms_pre, "gbs": gb / (ms_pre * 1e-3),
                               "frac_of_hbm": gb / (ms_pre * 1e-3) / peaks["hbm_gbs"]},
                "postprocess": {"showers_per_s": N / (ms_post * 1e-3), "ms": ms_post, "gbs": gb / (ms_post * 1e-3),
                                "frac_of_hbm": gb / (ms_post * 1e-3) / peaks["hbm_gbs"]},
                "round_trip_rel_l2": err,
                "note": "v4h_preprocess_showers (statistics known: one kernel) / v4h_postprocess_showers, timings include "
                        "the output allocations of the host classes; 8 B of HBM traffic per voxel"}
    except Exception as exc:  # secondary measurement
        return {"error": repr(exc)[:200]}


def measure_sampling(args, model, K, world, rank, dev, timed, peaks, lib):
    """showers/s of the sharded ODE sampling job (dp.sample_sharded): device-resident conditions (`value`), and
    end to end from pinned host conditions to host showers: H2D of the conditions, the solves, the final
    all_gather over the ranks and the .cpu() of reference experiments/calochallenge/experiment.py:223 (`e2e`)."""
    import torch.distributed as dist
    from vit4hep_b200 import _cabi, dp
    SB, total = args.sample_batch, args.sample_showers
    g = torch.Generator().manual_seed(4321)  # identical conditions on every rank; each integrates its shard
    conds_host = torch.rand(total, K, generator=g).pin_memory()
    conds = conds_host.to(dev)
    geom = model.geometry
    voxels = geom.voxels
    model.graph_sampling = False
    model.sample_batch(conds[: min(SB, 32)])
    nb_eager = 1
    ms_eager, launches = timed(lambda i: model.sample_batch(conds[:SB]), nb_eager)
    if not args.no_graph:
        model.graph_sampling = True  # the whole 80-evaluation solve of a batch replayed as one CUDA graph
        model.sample_batch(conds[:SB])
        b, e = dp.shard_range(total, rank, world)
        if (e - b) % SB:
            model.sample_batch(conds[: (e - b) % SB])  # the shard's last, shorter batch has its own graph
    ms, _ = timed(lambda i: dp.sample_sharded(model, conds, SB, gather=False), 1)
    showers = total / (ms * 1e-3)
    out_host = {}

    def e2e_job(i):
        c = conds_host.to(dev, non_blocking=True)
        full = dp.sample_sharded(model, c, SB, gather=True)
        if rank == 0:
            out_host["showers"] = full.cpu()
    ms_e2e, _ = timed(e2e_job, 1)
    ok = True
    if rank == 0:
        sh = out_host["showers"]
        ok = tuple(sh.shape) == (total, *geom.sample_shape) and bool(torch.isfinite(sh).all())
    gf = SAMPLE_GFLOP_PER_SHOWER[args.config]
    # the dominant kernel of a network evaluation, timed on the device around eager launches of 3 evaluations
    x = torch.randn(SB, geom.tokens, geom.patch_dim, device=dev)
    tt = torch.full((1,), 0.5, device=dev)
    model.net(x, tt, conds[:SB], shared_t=True)
    _cabi.profile_begin()
    for _ in range(3):
        model.net(x, tt, conds[:SB], shared_t=True)
    prof = _cabi.profile_end(128)
    kernels, roofline = family_roofline(prof, 3, peaks, "gemm_umma_kernel_sampling")
    if roofline is not None:
        roofline["timing"] = "CUDA events around eager launches of 3 network evaluations at the sampling batch"
    # the energy-ratio CFM that the reference samples for the same conditions right before the shape model
    # (experiments/calochallenge/experiment.py:225-247): same sharding, its own line
    energy = None
    try:
        from vit4hep_b200 import configs
        torch.manual_seed(1)
        em = configs.build(args.config + "_energy", args.precision).to(dev)
        with torch.no_grad():
            for p in em.net.parameters():
                if p.requires_grad and p.dim() == 1:
                    p.add_(0.05 * torch.randn_like(p))
        em.graph_sampling = not args.no_graph
        econd = conds[:, -1:].contiguous()
        em.sample_batch(econd[:SB])
        b, e = dp.shard_range(total, rank, world)
        if (e - b) % SB:
            em.sample_batch(econd[: (e - b) % SB])
        ms_en, _ = timed(lambda i: dp.sample_sharded(em, econd, SB, gather=False), 1)
        energy = {"metric": f"{args.config} energy-ratio CFM ODE-sampled vectors/s", "value": total / (ms_en * 1e-3),
                  "unit": "showers/s", "ms_total": ms_en, "batch": SB, "nfe_per_shower": 80,
                  "share_of_pipeline": ms_en / (ms_en + ms)}
        # the network is launch-latency bound at the reference's sample batch (configs/training/default.yaml:3,
        # batchsize_sample 256, shared with the shape model): the same job with a sample batch of 4096
        try:
            SBL = 4096
            shard = e - b
            em.sample_batch(econd[:min(SBL, shard)])
            if shard > SBL and shard % SBL:
                em.sample_batch(econd[: shard % SBL])
            ms_big, _ = timed(lambda i: dp.sample_sharded(em, econd, SBL, gather=False), 1)
            energy["large_batch"] = {"batch": SBL, "value": total / (ms_big * 1e-3), "ms_total": ms_big,
                                     "share_of_pipeline": ms_big / (ms_big + ms)}
        except Exception as exc:
            energy["large_batch"] = {"error": repr(exc)[:200]}
        del em
    except Exception as exc:  # secondary measurement: never take the main line down
        energy = {"error": repr(exc)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, per_ms, threads = cpu_showers_per_s(args.config, 64, 3)
        cpu = {"value": v, "unit": "showers/s", "cores": threads, "kind": "port",
               "sample": f"3 network evaluations of a 64-shower batch by the CPU oracle port ({per_ms:.0f} ms each), "
                         f"extrapolated to the 80 evaluations of a shower"}
    return {"metric": f"{args.config} ODE-sampled showers/s", "value": showers, "unit": "showers/s",
            "n_gpus": world, "scaling": "strong (fixed job sharded over the ranks)",
            "config": {"workload": f"CaloChallenge {args.config} ODE sampling of {total} showers (a stated subset of the "
                                   f"100k of BASELINE.json configs[3]), RK4 3/8 rule, 20 steps = 80 network evaluations, "
                                   f"sample batch {SB}, conditions sharded over {world} rank(s)",
                       "showers": total, "batch": SB, "nfe_per_shower": 80,
                       "execution": "one CUDA graph per batch solve" if not args.no_graph else "eager launches"},
            "ms_total": ms, "ms_per_batch": ms / max(1, -(-(total // world) // SB)),
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "showers/s", "ms_total": ms_e2e,
                    "h2d_bytes_per_step": conds_host.numel() * 4, "d2h_bytes_per_step": total * voxels * 4,
                    "includes": "H2D of the conditions, all solves, final all_gather over the ranks, .cpu() on rank 0",
                    "output_ok": ok},
            "eager_showers_per_s": world * SB * nb_eager / (ms_eager * 1e-3), "gpu_launches_per_batch": launches // nb_eager,
            "model_tflops_per_gpu": showers * gf / 1e3 / world,
            "frac_of_peak": showers * gf / 1e3 / world / peaks["tflops_sustained"],
            "roofline": roofline, "kernels": kernels[:12], "energy_model": energy, "cpu_baseline": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="ds2", choices=["ds2", "ds3"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=64, help="training batch per GPU")
    ap.add_argument("--sample-batch", type=int, default=256)
    ap.add_argument("--sample-showers", type=int, default=25600,
                    help="size of the sharded sampling job (BASELINE.json configs[3] is 100000)")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32 (reference precision) line")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling line at N > 1")
    ap.add_argument("--cpu-batch", type=int, default=64, help="batch of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=12, help="timed steps of the cpu_baseline leg (2 warm-up)")
    ap.add_argument("--torch-optimizer", action="store_true", help="clip_grad_norm_ + torch.optim.AdamW(fused)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA graph replay")
    ap.add_argument("--no-sampling", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

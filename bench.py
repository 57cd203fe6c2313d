#!/usr/bin/env python
"""Benchmark of the CFM-ViT hot path on B200: ds2 training samples/s (headline) and ODE-sampled showers/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

* A training "step" is the reference's ``BaseExperiment._step`` work on one batch of synthetic ds2 showers
  (64 per GPU, weak scaling): ``_batch_loss`` (t and x_0 draws, trajectory, patchify, ViT forward, MSE),
  ``zero_grad``, backward (with the bucketed gradient all-reduce when N > 1), ``clip_grad_norm_(1000)``,
  AdamW step.  ``value`` is timed with the batches already resident in HBM; ``e2e`` feeds pinned HOST
  batches through ``model._batch_loss`` (H2D inside the timed region) and reads the loss back every step.
* ``roofline`` / ``kernels``: per-kernel-class device time from CUDA events recorded by the library on the
  launching stream around each launch (v4h_profile_*), in a separate pass of the same steps so that the
  event overhead does not leak into ``value``.
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

os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout and would precede the JSON line

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
    from oracle import vit_oracle as vo  # geometry / hyper-parameter tables only
    cfg = vo.CONFIGS[config]
    return cfg["geom"], dict(cfg["param"])


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
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CaloChallenge {args.config} shape CFM-ViT training step (CPU oracle port of the "
                               f"reference path, batch {batch} per step)", "batch_per_step": batch},
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
def run_b200(args):
    import torch.distributed as dist
    from vit4hep_b200 import CaloChallengeCFM, FusedAdamW, GraphedTrainStep, ViT, _cabi, dp

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

    geom, param = ds2_setup(args.config)
    param["precision"] = args.precision
    torch.manual_seed(0)
    net = ViT(param)
    rerandomise(net)
    seg = geom.segments[0]
    model = CaloChallengeCFM(net, list(seg.patch), 1, "uniform", "linear",
                             dict(method="rk4", options=dict(step_size=0.05)), shape=list(seg.shape)).to(dev)
    model.device, model.dtype = dev, torch.float32
    if world > 1:
        dp.enable_data_parallel(model.net)
    params = list(model.net.parameters())
    if args.torch_optimizer:
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.1, fused=True)
    else:  # clip_grad_norm_(1000) + AdamW + bf16 weight refresh in one native multi-tensor pass
        opt = FusedAdamW(model.net, lr=1e-4, weight_decay=0.1, max_grad_norm=1000.0)

    B = args.batch
    K = param["condition_dim"]
    g = torch.Generator().manual_seed(1234 + rank)
    npool = 8
    host_x = [torch.randn(B, *geom.sample_shape, generator=g).pin_memory() for _ in range(npool)]
    host_c = [torch.rand(B, K, generator=g).pin_memory() for _ in range(npool)]
    dev_x = [x.to(dev) for x in host_x]
    dev_c = [c.to(dev) for c in host_c]
    lib = _cabi.load()

    def train_step(batch, read_loss: bool):
        loss = model._batch_loss(batch)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if args.torch_optimizer:
            torch.nn.utils.clip_grad_norm_(params, 1000.0)
        opt.step()
        return loss.item() if read_loss else loss

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
    ms_eager, launches = timed(lambda i: train_step((dev_x[i % npool], dev_c[i % npool]), False), S)
    for i in range(2):
        train_step((host_x[i % npool], host_c[i % npool]), True)
    ms_e2e_eager, _ = timed(lambda i: train_step((host_x[i % npool], host_c[i % npool]), True), S)
    use_graph = (not args.no_graph and not args.torch_optimizer
                 and (world == 1 or os.environ.get("V4H_GRAPH_DP", "1") == "1"))
    graphed = GraphedTrainStep(model, opt, dev_x[0], dev_c[0]) if use_graph else None
    if graphed is not None:
        step_dev = lambda i: graphed.step(dev_x[i % npool], dev_c[i % npool])
        step_e2e = lambda i: graphed.step(host_x[i % npool], host_c[i % npool]).item()
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

    # per-kernel-class device time (separate pass: the bracketing events cost launch overhead)
    peaks = measured_peaks()
    kernels, roofline = [], None
    if rank == 0 or world == 1:
        pass
    _cabi.profile_begin()
    psteps = min(S, 3)
    for i in range(psteps):
        train_step((dev_x[i % npool], dev_c[i % npool]), False)
    prof = _cabi.profile_end(128)
    total_ms = sum(e["ms"] for e in prof) or 1.0
    for e in sorted(prof, key=lambda e: -e["ms"]):
        per = e["ms"] / max(e["launches"], 1)
        kernels.append({"name": e["name"], "launches_per_step": e["launches"] / psteps, "ms_per_step": e["ms"] / psteps,
                        "share": e["ms"] / total_ms, "avg_ms": per,
                        "tflops": e["flops"] / (e["ms"] * 1e-3) / 1e12 if e["ms"] > 0 else 0.0,
                        "gbs": e["bytes"] / (e["ms"] * 1e-3) / 1e9 if e["ms"] > 0 else 0.0})
    # roofline of the dominant kernel: every gemm.* / dgrad.* / wgrad.* class is ONE kernel template
    # (gemm_umma_kernel, csrc/gemm_umma.cu), which together takes the largest share of the step.
    # achieved = algorithmic flops of those launches (2 M N K each) / their summed device time;
    # traffic = DRAM bytes per launch of the same kernel from the committed ncu capture (profiles/traffic.json)
    if kernels:
        fam = [k for k in kernels if k["name"].startswith(("gemm", "wgrad", "dgrad"))]
        fam_ms = sum(k["ms_per_step"] for k in fam)
        fam_flops = sum(k["tflops"] * 1e12 * k["ms_per_step"] * 1e-3 for k in fam)
        fam_launches = sum(k["launches_per_step"] for k in fam)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get("gemm_umma_kernel", {}).get("dram_bytes_per_launch")
        if fam_ms > 0:
            ach = fam_flops / (fam_ms * 1e-3) / 1e12
            roofline = {"kernel": "gemm_umma_kernel (tcgen05 GEMM: all gemm.* / dgrad.* / wgrad.* classes)",
                        "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tflops_sustained"], "traffic": traffic,
                        "flops_per_launch": fam_flops / max(fam_launches, 1),
                        "avg_launch_ms": fam_ms / max(fam_launches, 1), "launches_per_step": fam_launches,
                        "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                        "share_of_step": fam_ms / (total_ms / psteps)}

    # ODE sampling: 20 RK4 (3/8) steps = 80 network evaluations per shower, batches sharded over ranks
    sampling = None
    if not args.no_sampling:
        SB = args.sample_batch
        conds = torch.rand(SB, K, generator=g).to(dev)
        model.sample_batch(conds[: min(SB, 32)])
        nb = args.sample_batches
        ms_s_eager, launches_s = timed(lambda i: model.sample_batch(conds), nb)
        if not args.no_graph:
            model.graph_sampling = True  # the whole 80-evaluation solve replayed as one CUDA graph
            model.sample_batch(conds)
        ms_s, _ = timed(lambda i: model.sample_batch(conds), nb)
        showers = world * SB * nb / (ms_s * 1e-3)
        sampling = {"metric": f"{args.config} ODE-sampled showers/s", "value": showers, "unit": "showers/s",
                    "batch": SB, "batches": nb, "nfe_per_shower": 80, "ms_per_batch": ms_s / nb,
                    "gpu_launches": launches_s, "cuda_graph": not args.no_graph,
                    "eager_showers_per_s": world * SB * nb / (ms_s_eager * 1e-3),
                    "model_tflops": showers * SAMPLE_GFLOP_PER_SHOWER[args.config] / 1e3 / world,
                    "frac_of_peak": showers * SAMPLE_GFLOP_PER_SHOWER[args.config] / 1e3 / world
                    / peaks["tflops_sustained"]}

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
            "config": {"workload": f"CaloChallenge {args.config} shape CFM-ViT {args.precision} training, "
                                   f"batch {B} per GPU, data-parallel x{world} (BASELINE.json configs[1])",
                       "global_batch": B * world, "per_gpu_batch": B, "tokens": geom.tokens,
                       "patch_dim": geom.patch_dim, "parallelism": f"dp{world}",
                       "execution": ("one CUDA graph per training step (vit4hep_b200.GraphedTrainStep)"
                                     if graphed is not None else "eager launches"),
                       "optimizer": ("torch AdamW(fused) + clip_grad_norm_(1000)" if args.torch_optimizer else
                                     "vit4hep_b200.FusedAdamW: clip_grad_norm(1000) + AdamW + bf16 weight refresh"),
                       "l2": "no explicit flush: the per-step working set (activation workspace ~1 GB + 104 MB "
                             "fp32 / 52 MB bf16 weights + 8 rotating input batches) exceeds the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / S, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4,
                    "input_copy": ("prefetched: batch i+1 copied from pinned host memory under step i "
                                   "(GraphedTrainStep.stage / step_staged)") if graphed is not None
                                  else "inline, before each step"},
            "gpu_launches": launches,
            "eager": {"value": world * B * S / (ms_eager * 1e-3), "e2e": world * B * S / (ms_e2e_eager * 1e-3)},
            "model_tflops_per_gpu": value * gf / 1e3 / world,
            "frac_of_peak": value * gf / 1e3 / world / peaks["tflops_sustained"],
            "roofline": roofline,
            "kernels": kernels[:32],
            "sampling": sampling,
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="ds2", choices=["ds2", "ds3"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=64, help="training batch per GPU")
    ap.add_argument("--sample-batch", type=int, default=256)
    ap.add_argument("--sample-batches", type=int, default=2)
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

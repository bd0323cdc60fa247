#!/usr/bin/env python
"""Benchmark of the mFormer hot path on B200 (contract: see the task brief / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the host CPU (baseline/_ref)

Headline workload (BASELINE.json configs[1]): mFormerV1_sm training step (fwd + 6-rank hierarchical CE loss + bwd + clip 5.0 +
AdamW), bf16 compute, batch 256 per GPU, 224x224, 3 metadata components, synthetic inputs, random-init weights.  Weak scaling:
every rank runs the same per-GPU batch; gradients are averaged with a bucketed NCCL all-reduce captured inside the step's CUDA
graph (overlapping the rest of backward), the 1/world factor folded into the AdamW kernel.

The same JSON line carries, under "sub", the other BASELINE.json configurations measured in the same run: config 3 (md, global
batch 2048 with gradient accumulation below 8 GPUs), config 4 (xl at 384^2), and at N = 1 the inference numbers (V1 sm at batch
256 and batch 1, config 5 = mFormerV0 sweep) and the sm step at the reference's own drop-path rate (0.2).  "roofline_top" lists
the five largest (kernel, shape) items of the step by measured share, each on its own bound.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP = {"sm": 8.659, "md": 12.505, "xl": 448.16}  # per image (SURVEY.md section 6, 224^2; xl at 384^2)
FWD_GFLOP_V0 = {"sm": 8.94}
FMA_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12  # fp32 FMA pipe: 148 SMs x 128 lanes x 2 flop at 1965 MHz (no measured figure)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures under profiles/ (bf16, B = 256 shapes)
DW_FWD_TRAFFIC, DW_DGRAD_TRAFFIC, DW_WGRAD_TRAFFIC = 157.2e6 + 135.5e6, 312.6e6 + 143.8e6, 311.1e6 + 5.9e6  # rows 32-37 of r02_kernels_summary.md
NCU_TRAFFIC_BYTES = {  # profiles/r02_kernels_summary.md (read + write)
    "dwconv7_fwd_mma_kernel 256x56x56x96 (forward)": DW_FWD_TRAFFIC,
    "dwconv7_fwd_mma_kernel 256x56x56x96 (data gradient + fused skip gradient)": DW_DGRAD_TRAFFIC,
    "dwconv7_wgrad_mma_kernel 256x56x56x96": DW_WGRAD_TRAFFIC,
    "mlp_fused_fwd_kernel C=96 M=802816": 308.5e6 + 126.1e6,
    "mlp_fused_bwd_kernel C=96 M=802816": 308.5e6 + 1335.0e6,
    "wgrad_tc_kernel dW1 384x96 K=802816": 771.0e6 + 4.2e6,
    "wgrad_tc_kernel dW2 96x384 K=802816": 771.0e6 + 4.4e6,
    "ln_bwd_bf16_kernel 802816x96": 314.7e6 + 126.2e6,
    "attn_bwd_tc2_kernel 256x6x200x64": 197.9e6 + 83.8e6,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="sm")
    ap.add_argument("--arch", default="v1", choices=["v1", "v0"], help="mFormerV1 (train / infer) or mFormerV0 (RelativeAttention variant, infer only)")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU (micro-)batch")
    ap.add_argument("--accum", type=int, default=1, help="gradient accumulation steps per optimizer step")
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--drop-path", type=float, default=0.0)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (other BASELINE configs) and the roofline_top table")
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the CPU sample (0: chosen to fit the time limit)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML every 10 ms; nvidia-smi polling if pynvml is missing)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run_nvml(self) -> bool:
        """10 ms polling through NVML (same counters nvidia-smi prints); False if pynvml is unusable."""
        try:
            import pynvml as N

            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
        except Exception:
            return False
        bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self._stop.is_set():
            try:
                row = [str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx), str(N.nvmlDeviceGetPowerUsage(h) / 1e3), "", "", "", ""]
                mask = int(get_reasons(h))
                for bit, col in bits:
                    row[col] = "Active" if mask & bit else "Not Active"
                self.rows.append(row)
            except Exception:
                pass
            self._stop.wait(0.01)
        return True

    def _run(self):
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def workload_name(args) -> str:
    """One name for both arms (the driver pairs their lines by metric and config)."""
    S = args.img
    if args.mode == "train":
        return f"mFormerV1_{args.variant} train step (fwd+6-rank CE loss+bwd+clip+AdamW) {S}x{S}, 3 meta comps, random init"
    return f"mFormer{args.arch.upper()}_{args.variant} inference (eval forward, 6 rank heads) {S}x{S}, 3 meta comps, random init"


# ----------------------------------------------------------------------------- CPU baselines
def cpu_port_train(variant: str, img: int, batch: int, steps: int = 3, warmup: int = 1):
    """The reference's CPU path restated by the oracle (fp32), timed on this box's host cores."""
    from linnaeus_b200.config import make_synthetic_config
    from oracle import mformer_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    cfg, nc = make_synthetic_config(variant, img)
    a = O.arch_from_config(cfg, nc)
    P = O.synth_state_dict(O.param_shapes(a), 0)
    x, meta, tg = O.synth_batch(a, batch, 0)
    state = {}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(P, a, x, meta, tg, state, i + 1, 1e-4 * batch / 512)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med, torch.get_num_threads()


def cpu_port_infer(variant: str, img: int, batch: int, steps: int = 5, warmup: int = 2, arch: str = "v1"):
    from linnaeus_b200.config import make_synthetic_config, make_synthetic_config_v0

    torch.set_num_threads(os.cpu_count() or 1)
    if arch == "v0":
        from oracle import mformer_v0_oracle as O

        cfg, nc = make_synthetic_config_v0(variant, img)
        a = O.arch_from_config(cfg, nc)
        P = O.synth_state_dict(a, 0)
        x, meta = O.synth_batch(a, batch, 0)
    else:
        from oracle import mformer_oracle as O

        cfg, nc = make_synthetic_config(variant, img)
        a = O.arch_from_config(cfg, nc)
        P = O.synth_state_dict(O.param_shapes(a), 0)
        x, meta, _ = O.synth_batch(a, batch, 0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.forward(P, a, x, meta)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med, torch.get_num_threads()


def cpu_arm(args, batch: int, steps: int, warmup: int):
    """-> (img/s, s/step, threads, kind, what).  The unmodified reference from baseline/_ref when it is installed, else the oracle port."""
    from baseline import ref_arm

    if ref_arm.available():
        if args.mode == "train":
            v, med, cores = ref_arm.time_train(args.variant, args.img, batch, steps, warmup)
        else:
            v, med, cores = ref_arm.time_infer(args.arch, args.variant, args.img, batch, steps, warmup)
        return v, med, cores, "reference", "unmodified reference (baseline/_ref) through build_model / weighted_hierarchical_loss / build_optimizer, fp32"
    if args.mode == "train":
        v, med, cores = cpu_port_train(args.variant, args.img, batch, steps, warmup)
    else:
        v, med, cores = cpu_port_infer(args.variant, args.img, batch, steps, warmup, arch=args.arch)
    return v, med, cores, "port", "oracle port of the reference path (oracle/), fp32"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.arch == "v0" and args.mode != "infer":
        raise SystemExit("--arch v0 is the inference benchmark (config 5): use --mode infer")
    metric = "train_img_per_s" if args.mode == "train" else "infer_img_per_s"
    t0 = time.perf_counter()
    b = args.cpu_batch
    if b <= 0:
        # the same per-GPU batch as the GPU arm when the whole run (steps + warm-up) then ends within ~4 minutes, else the largest
        # power of two that does; calibrated with two steps at batch 8
        _, med8, _, _, _ = cpu_arm(args, 8, 1, 1)
        per_img = med8 / 8
        b = args.batch
        while b > 8 and per_img * b * (args.steps + args.warmup) > 240.0:
            b //= 2
    val, med, cores, kind, what = cpu_arm(args, b, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": metric, "value": val, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus,
                   "parallelism": f"dp{args.gpus}", "reference_arm": f"host CPU, {what}", "cpu_sample_batch": b},
        "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps of batch {b} (median), {args.warmup} warm-up; {what}"},
        "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
class Ctx:
    """Process-wide state of the B200 arm."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from linnaeus_b200 import _lib

        _lib.load()
        self.lib = _lib
        self.l2_flush = None

    def flush_buf(self):
        if self.l2_flush is None:
            self.l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        return self.l2_flush

    def timed(self, fn, steps, warmup, flush=False):
        """-> (ms per call, C-ABI launches).  CUDA events on the current stream, barrier + synchronize on both sides, max over ranks.
        flush=True: a 256 MB fill between calls (working set fits the 126 MB L2), per-call events so the fill is not timed."""
        dist, world = self.dist, self.world
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = self.lib.launch_count
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            total = e0.elapsed_time(e1)
        else:
            buf = self.flush_buf()
            evs = []
            for _ in range(steps):
                buf.zero_()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                fn()
                a1.record()
                evs.append((a0, a1))
            torch.cuda.synchronize()
            total = sum(a0.elapsed_time(a1) for a0, a1 in evs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([total], device=self.dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, (self.lib.launch_count - l0)


def host_batch(nc, B, S, rank):
    """Synthetic host data (pinned), per-rank seed like main.py:2625."""
    g = torch.Generator().manual_seed(42 + rank)
    h_img = torch.randn(B, 3, S, S, generator=g).pin_memory()
    h_meta = torch.randn(B, 15, generator=g).pin_memory()
    h_tg = {k: torch.randint(0, c, (B,), generator=g).pin_memory() for k, c in nc.items()}
    return h_img, h_meta, h_tg


def measure_train(cx: Ctx, variant, S, B, accum, drop_path, dtype, steps, warmup, use_graph=True, sample_clocks=False):
    """One training configuration.  A "step" is one OPTIMIZER step = ``accum`` micro-batches of B images per GPU."""
    import linnaeus_b200 as L
    from linnaeus_b200.engine import TrainStep
    from linnaeus_b200.optim import FlatAdamW
    from linnaeus_b200.parallel import DataParallel

    world, dev = cx.world, cx.dev
    cfg, nc = L.make_synthetic_config(variant, S, drop_path=drop_path)
    torch.manual_seed(0)
    model = L.build_model(cfg, nc).to(dev)
    cd = torch.bfloat16 if dtype == "bf16" else torch.float32
    model.set_compute_dtype(cd).train()
    keys = list(nc.keys())
    h_img, h_meta, h_tg = host_batch(nc, B, S, cx.rank)
    d_img, d_meta = h_img.to(dev), h_meta.to(dev)
    d_tg = {k: v.to(dev) for k, v in h_tg.items()}
    h2d = h_img.numel() * 4 + h_meta.numel() * 4 + sum(v.numel() * 8 for v in h_tg.values())
    lr = 1e-4 * (B * accum * world) / 512.0  # linear LR scaling, schedule_utils.py:517-523
    # gradient averaging (DDP semantics) rides on the AdamW kernel: the all-reduce sums, grad_scale = 1 / world
    opt = FlatAdamW(model.named_parameters(), lr=lr, weight_decay=0.05, clip_grad=5.0, grad_scale=1.0 / world)
    dp = DataParallel(model, opt.flat, average=False) if world > 1 else None
    ts = TrainStep(model, opt, keys, nc, kind="ce", config=cfg, dp=dp, accum_steps=accum)
    if use_graph:
        ts.capture(d_img, d_meta, d_tg, warmup=2)

        def resident():
            for _ in range(accum):
                ts.replay()

        # end to end = what a prefetching loader gives train.py: the pinned-host batch of iteration i+1 is copied (H2D, side stream)
        # while the graph of iteration i runs; every iteration's copy and the loss read-back are inside the timed region
        ts.prefetch(h_img, h_meta, h_tg)

        def e2e_step():
            for _ in range(accum):
                ts.replay()                          # waits for the staged batch, moves it into the graph inputs, runs
                ts.prefetch(h_img, h_meta, h_tg)     # next iteration's H2D overlaps this iteration's graph
            return float(ts.loss)                    # D2H read of the loss (host sync)
    else:
        def resident():
            for _ in range(accum):
                ts.step(d_img, d_meta, d_tg)

        def e2e_step():
            for _ in range(accum):
                loss = ts.step(h_img.to(dev, non_blocking=True), h_meta.to(dev, non_blocking=True),
                               {k: v.to(dev, non_blocking=True) for k, v in h_tg.items()})
            return float(loss)

    sampler = ClockSampler(cx.local) if sample_clocks else None
    if sampler:
        sampler.start()
    ms, launches = cx.timed(resident, steps, warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = cx.timed(e2e_step, max(3, steps // 2), 2)
    if use_graph:
        launches = ts.launches_per_step * steps * accum
    # replicas must still hold identical parameters after the timed loops (the all-reduce covered every gradient)
    chk = torch.stack([f.p.double().sum() for f in opt.flat if f is not None] + [f.p.double().abs().sum() for f in opt.flat if f is not None])
    replicas_equal = True
    if world > 1:
        allc = [torch.empty_like(chk) for _ in range(world)]
        cx.dist.all_gather(allc, chk)
        replicas_equal = all(torch.equal(allc[0], c) for c in allc)
        if not replicas_equal:
            raise SystemExit(f"rank {cx.rank}: parameter checksums differ across ranks after the timed loop: {[c.tolist() for c in allc]}")
    imgs = B * accum * world
    out = {"img_per_s": imgs / (ms / 1e3), "ms_per_step": ms, "e2e_img_per_s": imgs / (ms_e2e / 1e3), "ms_per_step_e2e": ms_e2e,
           "per_gpu_batch": B, "accum_steps": accum, "global_batch": imgs, "h2d_bytes_per_step": h2d * accum, "launches": launches,
           "clocks": clocks, "replicas_equal": replicas_equal, "cfg": cfg, "cd": cd,
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    del ts, dp, opt, model
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)
    return out


def measure_infer(cx: Ctx, arch, variant, S, B, dtype, steps, warmup, use_graph=True, sample_clocks=False):
    import linnaeus_b200 as L

    dev = cx.dev
    cfg, nc = (L.make_synthetic_config_v0 if arch == "v0" else L.make_synthetic_config)(variant, S)
    torch.manual_seed(0)
    model = L.build_model(cfg, nc).to(dev)
    cd = torch.bfloat16 if dtype == "bf16" else torch.float32
    model.set_compute_dtype(cd).eval()
    h_img, h_meta, _ = host_batch(nc, B, S, cx.rank)
    d_img, d_meta = h_img.to(dev), h_meta.to(dev)
    flush = B * S * S < 64 * 224 * 224  # small-batch inference keeps its whole working set inside the 126 MB L2
    s_img, s_meta = d_img.clone(), d_meta.clone()

    def fwd(img, meta):
        with torch.no_grad():
            return model(img, meta)

    if use_graph:  # one CUDA graph per forward over static input buffers (same mechanism as TrainStep.capture)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fwd(s_img, s_meta)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = cx.lib.launch_count
        with torch.cuda.graph(graph):
            s_out = fwd(s_img, s_meta)
        per_fwd = cx.lib.launch_count - n0
        resident = lambda: graph.replay()  # noqa: E731
        # double-buffered input staging: the H2D copy of the next batch overlaps the graph of this one
        g_img, g_meta = torch.empty_like(s_img), torch.empty_like(s_meta)
        copy_stream = torch.cuda.Stream(device=dev)
        copy_done, stage_free = torch.cuda.Event(), torch.cuda.Event()
        stage_free.record()

        def prefetch():
            copy_stream.wait_event(stage_free)
            with torch.cuda.stream(copy_stream):
                g_img.copy_(h_img, non_blocking=True)
                g_meta.copy_(h_meta, non_blocking=True)
                copy_done.record(copy_stream)

        prefetch()

        def e2e_step():
            cur = torch.cuda.current_stream()
            cur.wait_event(copy_done)
            s_img.copy_(g_img, non_blocking=True)
            s_meta.copy_(g_meta, non_blocking=True)
            stage_free.record(cur)
            graph.replay()
            prefetch()
            return s_out.cat[:, :8].float().cpu()
    else:
        per_fwd = None
        resident = lambda: fwd(d_img, d_meta)  # noqa: E731

        def e2e_step():
            out = fwd(h_img.to(dev, non_blocking=True), h_meta.to(dev, non_blocking=True))
            return out.cat[:, :8].float().cpu()

    sampler = ClockSampler(cx.local) if sample_clocks else None
    if sampler:
        sampler.start()
    ms, launches = cx.timed(resident, steps, warmup, flush=flush)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = cx.timed(e2e_step, max(3, steps // 2), 2)
    if per_fwd is not None:
        launches = per_fwd * steps
    out = {"img_per_s": B * cx.world / (ms / 1e3), "ms_per_step": ms, "e2e_img_per_s": B * cx.world / (ms_e2e / 1e3), "ms_per_step_e2e": ms_e2e,
           "per_gpu_batch": B, "h2d_bytes_per_step": h_img.numel() * 4 + h_meta.numel() * 4, "launches": launches, "clocks": clocks,
           "l2_flush": flush, "cfg": cfg, "cd": cd}
    del model
    gc.collect()
    torch.cuda.empty_cache()
    return out


def roofline_table(cx: Ctx, B: int, step_ms: float, hbm: float, tf_sus: float):
    """The hot kernels of the mFormerV1_sm train step at this per-GPU batch, each timed live in isolation (CUDA events, inputs larger than
    L2 or L2 flushed) on its own bound; share = time x launches per step / measured step time.  Sorted by share."""
    import linnaeus_b200.functional as F

    dev = cx.dev
    bf = torch.bfloat16
    M0, C0 = B * 56 * 56, 96
    M1, C1 = B * 28 * 28, 192
    T3, D3, heads3, N3 = B * 200, 384, 6, 200
    rows = []

    def add(name, fn, count, bound, alg_bytes=None, flops=None, note=None):
        ms, _ = cx.timed(fn, 10, 3, flush=True)
        item = {"kernel": name, "launches_per_step": count, "ms_per_launch": ms, "share_of_step": ms * count / step_ms, "bound": bound}
        if bound == "hbm":
            item.update(achieved=alg_bytes / (ms / 1e3) / 1e9, peak=hbm, unit="GB/s", algorithmic_bytes=alg_bytes)
        elif bound == "tensor":
            item.update(achieved=flops / (ms / 1e3) / 1e12, peak=tf_sus, unit="TFLOP/s", flops=flops)
        else:  # fp32 FMA pipe (no measured peak in MEASURED_PEAKS.json: nominal lanes x clock)
            item.update(achieved=flops / (ms / 1e3) / 1e12, peak=FMA_TFLOPS_NOMINAL, unit="TFLOP/s (fp32 FMA, nominal peak)", flops=flops,
                        hbm_gbs=alg_bytes / (ms / 1e3) / 1e9, hbm_frac=alg_bytes / (ms / 1e3) / 1e9 / hbm)
        item["frac"] = item["achieved"] / item["peak"]
        item["traffic"] = NCU_TRAFFIC_BYTES.get(name)
        if note:
            item["note"] = note
        rows.append(item)

    # ---- stage 0 (802 816 rows at B = 256): depthwise 7x7 family
    x = torch.randn(B, 56, 56, C0, device=dev).to(bf)
    g = torch.randn(B, 56, 56, C0, device=dev).to(bf)
    w49 = torch.randn(49, C0, device=dev)
    bias = torch.randn(C0, device=dev)
    y = torch.empty_like(x)
    dw = torch.zeros(49, C0, device=dev)
    db = torch.zeros(C0, device=dev)
    conv_flops = 2.0 * 49 * M0 * C0
    dw_note = ("banded matrix products on mma.sync (tensor pipe); bounded by HBM in the limit (49 MAC per 4 bytes is below the tensor ridge), "
               "measured limiters: shared-memory wavefronts and issue slots of the register-level NHWC -> channel-planar transposition, "
               "see profiles/r02_dwconv_mma.md")
    add(f"dwconv7_fwd_mma_kernel {B}x56x56x96 (forward)",
        lambda: cx.lib.call("lnx_dwconv7_fwd", x.data_ptr(), w49.data_ptr(), 0, bias.data_ptr(), None, y.data_ptr(), B, 56, 56, C0, 1), 3, "hbm",
        alg_bytes=2 * M0 * C0 * 2, flops=conv_flops, note=dw_note)
    add(f"dwconv7_fwd_mma_kernel {B}x56x56x96 (data gradient + fused skip gradient)",
        lambda: cx.lib.call("lnx_dwconv7_fwd", g.data_ptr(), w49.data_ptr(), 0, None, x.data_ptr(), y.data_ptr(), B, 56, 56, C0, 1), 3, "hbm",
        alg_bytes=3 * M0 * C0 * 2, flops=conv_flops, note=dw_note)
    add(f"dwconv7_wgrad_mma_kernel {B}x56x56x96",
        lambda: cx.lib.call("lnx_dwconv7_wgrad", x.data_ptr(), g.data_ptr(), dw.data_ptr(), 0, db.data_ptr(), B, 56, 56, C0, 1), 3, "hbm",
        alg_bytes=2 * M0 * C0 * 2, flops=conv_flops, note=dw_note)
    # ---- stage 0: fused pointwise pair
    x2 = x.view(M0, C0)
    g2 = g.view(M0, C0)
    w1 = (torch.randn(4 * C0, C0, device=dev) * C0 ** -0.5).to(bf)
    w2 = (torch.randn(C0, 4 * C0, device=dev) * (4 * C0) ** -0.5).to(bf)
    b1 = torch.randn(4 * C0, device=dev) * 0.1
    b2 = torch.randn(C0, device=dev) * 0.1
    gamma = torch.rand(C0, device=dev) + 0.5
    out = torch.empty_like(x2)
    add(f"mlp_fused_fwd_kernel C=96 M={M0}", lambda: F.mlp_fused_fwd(x2, w1, b1, w2, b2, gamma=gamma, residual=g2, out=out), 3, "hbm",
        alg_bytes=3 * M0 * C0 * 2, note="x, residual read, y written; its own tile pipeline is the limiter (weights resident: ~56 KB of loads in flight per SM), see profiles/r02_microbench.md")
    h, dpre, _ = F.mlp_fused_bwd(x2, g2, w1, b1, w2)
    add(f"mlp_fused_bwd_kernel C=96 M={M0}", lambda: F.mlp_fused_bwd(x2, g2, w1, b1, w2), 3, "hbm", alg_bytes=11 * M0 * C0 * 2,
        note="x, dY read; h, dPre (4C wide) and dX written: write-bandwidth bound")
    dw1 = torch.zeros(4 * C0, C0, device=dev)
    db1 = torch.zeros(4 * C0, device=dev)
    add(f"wgrad_tc_kernel dW1 384x96 K={M0}", lambda: F.wgrad(dpre, x2, out=dw1, db_out=db1), 3, "hbm", alg_bytes=5 * M0 * C0 * 2)
    dw2 = torch.zeros(C0, 4 * C0, device=dev)
    db2 = torch.zeros(C0, device=dev)
    add(f"wgrad_tc_kernel dW2 96x384 K={M0}", lambda: F.wgrad(g2, h, out=dw2, db_out=db2), 3, "hbm", alg_bytes=5 * M0 * C0 * 2)
    del h, dpre
    # ---- stage 0: LayerNorm backward
    lw = torch.ones(C0, device=dev)
    mean = torch.zeros(M0, device=dev)
    rstd = torch.ones(M0, device=dev)
    dxl = torch.empty_like(x2)
    add(f"ln_bwd_bf16_kernel {M0}x96", lambda: cx.lib.call("lnx_layernorm_bwd", g2.data_ptr(), x2.data_ptr(), lw.data_ptr(), mean.data_ptr(),
                                                             rstd.data_ptr(), None, dxl.data_ptr(), dw[0].data_ptr(), db.data_ptr(), M0, C0, 1),
        4, "hbm", alg_bytes=3 * M0 * C0 * 2, note="launch count: the 3 block norms + the stem norm at this shape")
    del x, g, y, x2, g2, out, dxl
    # ---- stage 1 (C = 192): two-GEMM pointwise pair (forward expand GEMM saves gelu')
    xs1 = torch.randn(M1, C1, device=dev).to(bf)
    w1s = (torch.randn(4 * C1, C1, device=dev) * C1 ** -0.5).to(bf)
    b1s = torch.randn(4 * C1, device=dev) * 0.1
    hs1 = torch.empty(M1, 4 * C1, device=dev, dtype=bf)
    aux1 = torch.empty_like(hs1)
    add(f"gemm_tc2_kernel<GELU_DG,aux> pwconv1 stage 1: {M1}x768x192",
        lambda: F.gemm(xs1, w1s, M1, 4 * C1, C1, out=hs1, bias=b1s, act=3, aux_out=aux1), 3, "hbm", alg_bytes=(M1 * C1 + 2 * M1 * 4 * C1) * 2)
    del xs1, hs1, aux1
    # ---- stage 3: attention backward and the fc1 GEMM (tensor bound)
    q = torch.randn(3, B, heads3, N3, 64, device=dev).to(bf) * 0.5
    o = torch.randn(B, N3, D3, device=dev).to(bf)
    do = torch.randn(B, N3, D3, device=dev).to(bf)
    lse = torch.zeros(B, heads3, N3, device=dev)
    cx.lib.call("lnx_attn_fwd", q[0].data_ptr(), q[1].data_ptr(), q[2].data_ptr(), o.data_ptr(), lse.data_ptr(), B, heads3, N3, 64, 1, 0)
    dq = torch.empty_like(q)
    delta = torch.empty(B * heads3 * N3 * 65 + 4, device=dev)
    add(f"attn_bwd_tc2_kernel {B}x6x200x64",
        lambda: cx.lib.call("lnx_attn_bwd", q[0].data_ptr(), q[1].data_ptr(), q[2].data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                            dq[0].data_ptr(), dq[1].data_ptr(), dq[2].data_ptr(), delta.data_ptr(), B, heads3, N3, 64, 1, 0),
        5, "tensor", flops=5 * 2.0 * B * heads3 * N3 * N3 * 64)
    xt = torch.randn(T3, D3, device=dev).to(bf)
    wt = (torch.randn(4 * D3, D3, device=dev) * D3 ** -0.5).to(bf)
    ht = torch.empty(T3, 4 * D3, device=dev, dtype=bf)
    add(f"gemm_tc2_kernel fc1 stage 3: {T3}x1536x384 (plain epilogue)", lambda: F.gemm(xt, wt, T3, 4 * D3, D3, out=ht), 10, "tensor",
        flops=2.0 * T3 * 4 * D3 * D3, note="launch count: fc1 fwd + dPre bwd of the 5 stage-3 blocks (same shape class)")
    rows.sort(key=lambda r: -r["share_of_step"])
    gc.collect()
    torch.cuda.empty_cache()
    return rows


def _sub_train(m, world, gflop, tf_sus, what):
    tf = m["img_per_s"] * gflop * 3.0 / 1e3
    return {"workload": what, "train_img_per_s": m["img_per_s"], "ms_per_optimizer_step": m["ms_per_step"], "e2e_img_per_s": m["e2e_img_per_s"],
            "per_gpu_batch": m["per_gpu_batch"], "accum_steps": m["accum_steps"], "global_batch": m["global_batch"], "n_gpus": world,
            "model_tflops_per_gpu": tf / world, "frac_of_bf16_sustained_per_gpu": tf / world / tf_sus, "peak_mem_gb": m["peak_mem_gb"],
            "replicas_equal_after_run": m["replicas_equal"]}


def _sub_infer(i):
    return {"infer_img_per_s": i["img_per_s"], "latency_ms": i["ms_per_step"], "e2e_img_per_s": i["e2e_img_per_s"], "batch": i["per_gpu_batch"],
            "l2_flush_between_iterations": i["l2_flush"]}


def single_kernel_roofline(cx: Ctx, args, r, hbm, peak_src):
    """Non-default workloads: the stage-0 pointwise kernel exactly as that workload launches it (HBM bound)."""
    import linnaeus_b200.functional as F

    cfg, cd, B, S, dev = r["cfg"], r["cd"], args.batch, args.img, cx.dev
    if args.arch == "v0":  # stage_1 expand 1x1 conv (+folded BN, swish): one output
        M, K, N = B * (S // 4) ** 2, cfg.MODEL.CONV_STAGES.EMBED_DIMS[0], 4 * cfg.MODEL.CONV_STAGES.EMBED_DIMS[0]
        act_code, n_out, kname = 5, 1, "gemm_tc2_kernel<SWISH> (MBConv expand stage 1: M=%d K=%d N=%d, folded BN + swish)"
    else:
        K = cfg.MODEL.CONVNEXT_STAGES.DIMS[0]
        if K == 96 and cd == torch.bfloat16:  # the fused pointwise-pair kernel covers this stage: report it
            M = B * (S // 4) ** 2
            x = torch.randn(M, K, device=dev).to(cd)
            res = torch.randn(M, K, device=dev).to(cd)
            w1 = (torch.randn(4 * K, K, device=dev) * K ** -0.5).to(cd)
            w2 = (torch.randn(K, 4 * K, device=dev) * (4 * K) ** -0.5).to(cd)
            b1, b2 = torch.zeros(4 * K, device=dev), torch.zeros(K, device=dev)
            out = torch.empty_like(x)
            kms, _ = cx.timed(lambda: F.mlp_fused_fwd(x, w1, b1, w2, b2, residual=res, out=out), 20, 3, flush=True)
            alg = 3 * M * K * 2
            ach = alg / (kms / 1e3) / 1e9
            name = f"mlp_fused_fwd_kernel C=96 M={M}"
            return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": NCU_TRAFFIC_BYTES.get(name), "peak_source": peak_src, "algorithmic_bytes": alg, "ms_per_launch": kms}
        M, N = B * (S // 4) ** 2, 4 * K
        act_code, n_out, kname = 3, 2, "gemm_tc2_kernel<GELU_DG, aux> (pwconv1 stage 0: M=%d K=%d N=%d, bias+GELU, saves gelu')"
        if args.mode == "infer":
            act_code, n_out, kname = 1, 1, "gemm_tc2_kernel<GELU> (pwconv1 stage 0: M=%d K=%d N=%d, bias+GELU)"
    esz = 2 if cd == torch.bfloat16 else 4
    a_ = torch.randn(M, K, device=dev).to(cd)
    w_ = torch.randn(N, K, device=dev).to(cd)
    b_ = torch.randn(N, device=dev)
    o_ = torch.empty(M, N, device=dev, dtype=cd)
    aux_ = torch.empty(M, N, device=dev, dtype=cd) if n_out == 2 else None
    kms, _ = cx.timed(lambda: F.gemm(a_, w_, M, N, K, out=o_, bias=b_, act=act_code, aux_out=aux_), 20, 3, flush=True)
    alg_bytes = (M * K + N * K + n_out * M * N) * esz
    ach = alg_bytes / (kms / 1e3) / 1e9
    return {"bound": "hbm", "kernel": kname % (M, K, N), "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes": alg_bytes, "ms_per_launch": kms}


def run_b200(args):
    cx = Ctx()
    world, rank = cx.world, cx.rank
    hbm, tf_burst, tf_sus, peak_src = peaks()
    B, S = args.batch, args.img
    if args.arch == "v0" and args.mode != "infer":
        raise SystemExit("--arch v0 is the inference benchmark (config 5): use --mode infer")
    use_graph = not args.no_graph
    if args.mode == "train":
        r = measure_train(cx, args.variant, S, B, args.accum, args.drop_path, args.dtype, args.steps, args.warmup, use_graph, sample_clocks=True)
        metric, flop_mult = "train_img_per_s", 3.0
        d2h = 4 * args.accum
    else:
        r = measure_infer(cx, args.arch, args.variant, S, B, args.dtype, args.steps, args.warmup, use_graph, sample_clocks=True)
        metric, flop_mult = "infer_img_per_s", 1.0
        d2h = B * 8 * 4
    value, e2e = r["img_per_s"], r["e2e_img_per_s"]
    gflop = (FWD_GFLOP_V0 if args.arch == "v0" else FWD_GFLOP).get(args.variant, 0.0)
    if args.variant == "xl" and S != 384:
        gflop = 0.0  # the xl FLOP count is for 384^2 only
    model_tflops = value * gflop * flop_mult / 1e3
    line = {
        "metric": metric, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "per_gpu_batch": B, "global_batch": r.get("global_batch", B * world), "accum_steps": args.accum if args.mode == "train" else None,
                   "parallelism": f"dp{world}", "cuda_graph": bool(use_graph),
                   "l2_policy": ("256 MB fill between timed iterations (working set fits in L2); per-iteration CUDA events" if r.get("l2_flush")
                                 else "inputs+activations per step exceed the 126 MB L2; no explicit flush"),
                   "drop_path": args.drop_path,
                   "all_reduce": ("bucketed NCCL all-reduce captured inside the step graph, overlapped with backward; 1/world folded into AdamW"
                                  if (world > 1 and args.mode == "train") else None)},
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": d2h,
                "ms_per_step": r["ms_per_step_e2e"]},
        "gpu_launches": r["launches"],
        "clocks": r["clocks"],
        "model": {"model_tflops": model_tflops, "model_tflops_per_gpu": model_tflops / world,
                  "model_frac_of_bf16_sustained_per_gpu": model_tflops / world / tf_sus, "gflop_per_img_fwd": gflop, "peak_source": peak_src},
    }
    if args.mode == "train":
        line["replicas_equal_after_run"] = r["replicas_equal"]
        line["peak_mem_gb"] = r["peak_mem_gb"]

    default_workload = (args.mode == "train" and args.arch == "v1" and args.variant == "sm" and S == 224 and args.dtype == "bf16" and use_graph
                        and args.accum == 1 and args.drop_path == 0.0)
    # ---- roofline: the five largest (kernel, shape) items of the step by measured share; "roofline" = the largest
    if default_workload and not args.no_sub:
        table = roofline_table(cx, B, r["ms_per_step"], hbm, tf_sus)
        line["roofline_top"] = table[:5]
        top = table[0]
        tensor = top["bound"] == "tensor"
        line["roofline"] = {"bound": "tensor" if tensor else "hbm", "kernel": top["kernel"],
                            "achieved": top["achieved"] if (tensor or top["bound"] == "hbm") else top["hbm_gbs"],
                            "peak": tf_sus if tensor else hbm, "unit": "TFLOP/s" if tensor else "GB/s",
                            "frac": top["frac"] if (tensor or top["bound"] == "hbm") else top["hbm_frac"],
                            "traffic": top["traffic"], "peak_source": peak_src, "ms_per_launch": top["ms_per_launch"],
                            "share_of_step": top["share_of_step"],
                            "limiting_pipe": ("fp32 FMA: %.1f TFLOP/s = %.2f of the nominal %.1f" % (top["achieved"], top["frac"], FMA_TFLOPS_NOMINAL))
                            if top["bound"] == "fp32-fma" else top["bound"]}
    else:
        line["roofline"] = single_kernel_roofline(cx, args, r, hbm, peak_src)
    line["roofline"]["model_tflops"] = model_tflops
    line["roofline"]["model_frac_of_bf16_sustained"] = model_tflops / world / tf_sus  # per GPU

    # ---- the other BASELINE.json configurations, in the same run (all ranks take part in the data-parallel ones)
    if default_workload and not args.no_sub:
        sub = {}
        acc = max(1, 2048 // (256 * world))
        m = measure_train(cx, "md", 224, 256, acc, 0.0, "bf16", 2 if acc > 2 else 4, 1)
        sub["config3_md_global2048"] = _sub_train(m, world, FWD_GFLOP["md"], tf_sus,
                                                  f"mFormerV1_md, global batch {m['global_batch']} = {world} GPU x 256 x {acc} accumulation steps, 224^2, meta on")
        xb = 48
        m = measure_train(cx, "xl", 384, xb, 1, 0.0, "bf16", 3, 2)
        sub["config4_xl384"] = _sub_train(m, world, FWD_GFLOP["xl"], tf_sus, f"mFormerV1_xl at 384^2, {xb} per GPU (activations kept, no checkpointing)")
        if world == 1:
            m = measure_train(cx, "sm", 224, 256, 1, 0.2, "bf16", 5, 2)
            sub["sm_drop_path_0.2"] = _sub_train(m, world, FWD_GFLOP["sm"], tf_sus, "the headline workload at the reference's configured DROP_PATH_RATE 0.2")
            sub["infer_v1_sm_b256"] = _sub_infer(measure_infer(cx, "v1", "sm", 224, 256, "bf16", 10, 3))
            sub["infer_v1_sm_b1"] = _sub_infer(measure_infer(cx, "v1", "sm", 224, 1, "bf16", 20, 5))
            for b in (1, 16, 256, 1024):
                sub[f"config5_v0_sm_b{b}"] = _sub_infer(measure_infer(cx, "v0", "sm", 224, b, "bf16", 10 if b < 1024 else 5, 3))
        line["sub"] = sub

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cb = args.cpu_batch if args.cpu_batch > 0 else 16
        v, med, cores, kind, what = cpu_arm(args, cb, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "img/s", "cores": cores, "kind": kind,
                                "sample": f"3 steps of batch {cb} (median, 1 warm-up) of the same workload; {what}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

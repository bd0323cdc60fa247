#!/usr/bin/env python
"""Benchmark of the mFormerV1 hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

Workload (BASELINE.json configs[1]): mFormerV1_sm training step (fwd + 6-rank
hierarchical CE loss + bwd + clip 5.0 + AdamW), bf16 compute, batch 256 per GPU, 224x224,
3 metadata components, synthetic inputs, random-init weights.  Weak scaling: every rank
runs the same per-GPU batch; gradients are averaged with bucketed NCCL all-reduce.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel (profiles/r01_kernels_summary.md),
# keyed by (arch, variant, per-GPU batch, image size, dtype); None when that exact shape was not captured
NCU_TRAFFIC_BYTES = {("v1", "sm", 256, 224, "bf16"): 154.3e6 + 1174.6e6}  # rows 0-1 of r01_kernels_summary.md


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="sm")
    ap.add_argument("--arch", default="v1", choices=["v1", "v0"], help="mFormerV1 (train / infer) or mFormerV0 (RelativeAttention variant, infer only)")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=8)
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
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


def cpu_train_baseline(variant: str, img: int, batch: int, steps: int = 3, warmup: int = 1):
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


def cpu_infer_baseline(variant: str, img: int, batch: int, steps: int = 5, warmup: int = 2, arch: str = "v1"):
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


# ----------------------------------------------------------------------------- reference arm
def workload_name(args) -> str:
    """One name for both arms (the driver pairs their lines by metric and config)."""
    S = args.img
    if args.mode == "train":
        return f"mFormerV1_{args.variant} train step (fwd+6-rank CE loss+bwd+clip+AdamW) {S}x{S}, 3 meta comps, random init"
    return f"mFormer{args.arch.upper()}_{args.variant} inference (eval forward, 6 rank heads) {S}x{S}, 3 meta comps, random init"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    metric = "train_img_per_s" if args.mode == "train" else "infer_img_per_s"
    fn = cpu_train_baseline if args.mode == "train" else cpu_infer_baseline
    b = args.cpu_batch
    t0 = time.perf_counter()
    if args.arch == "v0":
        if args.mode != "infer":
            raise SystemExit("--arch v0 is the inference benchmark (config 5): use --mode infer")
        val, med, cores = cpu_infer_baseline(args.variant, args.img, b, steps=args.steps, warmup=args.warmup, arch="v0")
    else:
        val, med, cores = fn(args.variant, args.img, b, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference",
        "metric": metric, "value": val, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus,
                   "parallelism": f"dp{args.gpus}", "reference_arm": "host CPU, oracle port of the reference path, fp32", "cpu_sample_batch": b},
        "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {b} (median), {args.warmup} warm-up; oracle/mformer_oracle.py (fp32, torch CPU ops)"},
        "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist

    import linnaeus_b200 as L
    from linnaeus_b200 import _lib
    from linnaeus_b200.engine import TrainStep
    from linnaeus_b200.optim import FlatAdamW
    from linnaeus_b200.parallel import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B, S = args.batch, args.img
    if args.arch == "v0":
        if args.mode != "infer":
            raise SystemExit("--arch v0 is the inference benchmark (config 5): use --mode infer")
        cfg, nc = L.make_synthetic_config_v0(args.variant, S)
    else:
        cfg, nc = L.make_synthetic_config(args.variant, S)
    torch.manual_seed(0)
    model = L.build_model(cfg, nc).to(dev)
    cd = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    model.set_compute_dtype(cd)
    keys = list(nc.keys())

    # synthetic host data (pinned), per-rank seed like main.py:2625
    g = torch.Generator().manual_seed(42 + rank)
    h_img = torch.randn(B, 3, S, S, generator=g).pin_memory()
    h_meta = torch.randn(B, 15, generator=g).pin_memory()
    h_tg = {k: torch.randint(0, c, (B,), generator=g).pin_memory() for k, c in nc.items()}
    d_img, d_meta = h_img.to(dev), h_meta.to(dev)
    d_tg = {k: v.to(dev) for k, v in h_tg.items()}
    h2d = h_img.numel() * 4 + h_meta.numel() * 4 + sum(v.numel() * 8 for v in h_tg.values())

    sampler = ClockSampler(local)
    hbm, tf_burst, tf_sus, peak_src = peaks()

    # small-batch inference keeps its whole working set inside the 126 MB L2: flush it between timed iterations then
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if (args.mode == "infer" and B * S * S < 64 * 224 * 224) else None

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launch_count
        if l2_flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            total = e0.elapsed_time(e1)
        else:  # per-step events around fn only; the flush (a 256 MB fill) runs between them
            evs = []
            for _ in range(steps):
                l2_flush.zero_()
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
        ms = torch.tensor([total], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, (_lib.launch_count - l0)

    if args.mode == "train":
        lr = 1e-4 * (B * world) / 512.0  # linear LR scaling, schedule_utils.py:517-523
        # gradient averaging (DDP semantics) rides on the AdamW kernel: the all-reduce sums, grad_scale = 1 / world
        opt = FlatAdamW(model.named_parameters(), lr=lr, weight_decay=0.05, clip_grad=5.0, grad_scale=1.0 / world)
        dp = DataParallel(model, opt.flat, average=False) if world > 1 else None
        model.train()
        ts = TrainStep(model, opt, keys, nc, kind="ce", config=cfg, dp=dp)
        use_graph = not args.no_graph
        if use_graph:
            ts.capture(d_img, d_meta, d_tg, warmup=2)
            resident = lambda: ts.replay()  # noqa: E731
            # end to end = what a prefetching loader gives train.py: the pinned-host batch of step i+1 is copied (H2D,
            # side stream) while the graph of step i runs; every step's copy and loss read-back is inside the timed region
            ts.prefetch(h_img, h_meta, h_tg)
            def e2e_step():
                ts.replay()                          # waits for the staged batch, moves it into the graph inputs, runs
                ts.prefetch(h_img, h_meta, h_tg)     # next step's H2D overlaps this step's graph
                return float(ts.loss)                # D2H read of the loss (host sync)
        else:
            resident = lambda: ts.step(d_img, d_meta, d_tg)  # noqa: E731
            def e2e_step():
                return float(ts.step(h_img.to(dev, non_blocking=True), h_meta.to(dev, non_blocking=True),
                                     {k: v.to(dev, non_blocking=True) for k, v in h_tg.items()}))
        metric = "train_img_per_s"
        flop_mult = 3.0
    else:
        model.eval()
        use_graph = not args.no_graph
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
            n0 = _lib.launch_count
            with torch.cuda.graph(graph):
                s_out = fwd(s_img, s_meta)
            infer_launches = _lib.launch_count - n0
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
            resident = lambda: fwd(d_img, d_meta)  # noqa: E731
            def e2e_step():
                out = fwd(h_img.to(dev, non_blocking=True), h_meta.to(dev, non_blocking=True))
                return out.cat[:, :8].float().cpu()
        metric = "infer_img_per_s"
        flop_mult = 1.0

    sampler.start()
    ms, launches = timed(resident, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, _ = timed(e2e_step, max(3, args.steps // 2), 2)
    value = B * world / (ms / 1e3)
    e2e = B * world / (ms_e2e / 1e3)
    if use_graph and args.mode == "train":
        launches = ts.launches_per_step * args.steps if hasattr(ts, "launches_per_step") else launches
    elif use_graph:
        launches = infer_launches * args.steps

    # roofline of the dominant kernel, timed live: the stage-0 pointwise-expand GEMM of the ConvNeXt blocks exactly as the
    # training step launches it (gemm_tc2_kernel<GELU_DG, aux>: bias + GELU, gelu'(pre) saved as second output).  K = 96:
    # HBM bound, algorithmic bytes = A + W + 2 outputs.  `traffic` = dram bytes per launch of this kernel from the ncu
    # --set full capture committed under profiles/ (B = 256 shape only).
    import linnaeus_b200.functional as F
    if args.arch == "v0":  # stage_1 expand 1x1 conv (+folded BN, swish): one output
        M, K, N = B * (S // 4) ** 2, cfg.MODEL.CONV_STAGES.EMBED_DIMS[0], 4 * cfg.MODEL.CONV_STAGES.EMBED_DIMS[0]
        act_code, n_out, kname = 5, 1, "gemm_tc2_kernel<SWISH> (MBConv expand stage 1: M=%d K=%d N=%d, folded BN + swish)"
    else:
        M, K, N = B * (S // 4) ** 2, cfg.MODEL.CONVNEXT_STAGES.DIMS[0], 4 * cfg.MODEL.CONVNEXT_STAGES.DIMS[0]
        act_code, n_out, kname = 3, 2, "gemm_tc2_kernel<GELU_DG, aux> (pwconv1 stage 0: M=%d K=%d N=%d, bias+GELU, saves gelu')"
        if args.mode == "infer":  # the eval forward saves nothing for a backward: plain GELU epilogue, one output
            act_code, n_out, kname = 1, 1, "gemm_tc2_kernel<GELU> (pwconv1 stage 0: M=%d K=%d N=%d, bias+GELU)"
    esz = 2 if cd == torch.bfloat16 else 4
    a_ = torch.randn(M, K, device=dev).to(cd)
    w_ = torch.randn(N, K, device=dev).to(cd)
    b_ = torch.randn(N, device=dev)
    o_ = torch.empty(M, N, device=dev, dtype=cd)
    aux_ = torch.empty(M, N, device=dev, dtype=cd) if n_out == 2 else None
    def gemm():
        F.gemm(a_, w_, M, N, K, out=o_, bias=b_, act=act_code, aux_out=aux_)
    kms, _ = timed(gemm, 20, 3)
    alg_bytes = (M * K + N * K + n_out * M * N) * esz
    ach = alg_bytes / (kms / 1e3) / 1e9
    traffic = NCU_TRAFFIC_BYTES.get((args.arch, args.variant, B, S, args.dtype)) if args.mode == "train" else None
    gflop = (FWD_GFLOP_V0 if args.arch == "v0" else FWD_GFLOP).get(args.variant, 0.0)
    roofline = {"bound": "hbm", "kernel": kname % (M, K, N),
                "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes": alg_bytes, "ms_per_launch": kms,
                "model_tflops": value * gflop * flop_mult / 1e3,
                "model_frac_of_bf16_sustained": value * gflop * flop_mult / 1e3 / tf_sus}

    line = {
        "metric": metric, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": bool(use_graph),
                   "l2_policy": ("inputs+activations per step exceed the 126 MB L2; no explicit flush" if l2_flush is None
                                 else "256 MB fill between timed iterations (working set fits in L2); per-iteration CUDA events"),
                   "drop_path": 0.0},
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": h2d if args.mode == "train" else h_img.numel() * 4 + h_meta.numel() * 4,
                "d2h_bytes_per_step": 4 if args.mode == "train" else B * 8 * 4, "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        if args.mode == "train":
            v, med, cores = cpu_train_baseline(args.variant, S, args.cpu_batch, steps=3, warmup=1)
        else:
            v, med, cores = cpu_infer_baseline(args.variant, S, args.cpu_batch, steps=3, warmup=1, arch=args.arch)
        line["cpu_baseline"] = {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
                                "sample": f"3 steps of batch {args.cpu_batch} (median) of the same workload in fp32; "
                                          + ("oracle/mformer_v0_oracle.py" if args.arch == "v0" else "oracle/mformer_oracle.py")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

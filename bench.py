#!/usr/bin/env python
"""Benchmark of the hot path: batched MultimodalClassifier forward (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU forward (oracle port)

Workload ("step"): one full multimodal inference pass over a GLOBAL batch of 4096 synthetic samples
(224x224 fp32 images + 128 tokens with padding masks L ~ U{16..128}), sharded data-parallel over the
N GPUs of one box (4096/N samples per rank, replicated random-init weights), followed by the NCCL
all-gather of the [4096,10] logits.  Total work is fixed as N grows -> "scaling": "strong".

value  = samples/s with the inputs already resident in HBM, timed with CUDA events over exactly K
         steps (barrier + synchronize on both sides, max over ranks).  The per-rank inputs
         (>= 308 MB even at N=8) exceed the 126 MB L2, so no L2 flush is needed between steps.
e2e    = the same metric through the public module API starting from pinned HOST buffers: every step
         copies this rank's images/ids/mask host->device and reads the gathered logits back.
roofline = the dominant kernel family (tcgen05 implicit-GEMM: all convolutions and linears), measured
         in a separate profiled step with CUDA events around every launch on the launching stream.
cpu_baseline = the oracle port of the reference forward (oracle/forward_oracle.py, fp32, all host
         threads) timed on a bounded sample of the same workload, rank 0, N=1 only.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fused fwd samples/sec (224² img+128 tok) @1/2/4/8 B200; % tensor-pipe peak"
UNIT = "samples/s"
FLOP_PER_SAMPLE = 30.53e9  # full multimodal forward, S=128
SEQ = 128
IMG = 224


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"src": "measured", "bf16_burst": p.get("bf16_tflops", 1590.0),
                "bf16_sustained": p.get("bf16_tflops_sustained", 1400.0), "hbm": p.get("hbm_gbs", 6650.0)}
    return {"src": "fallback", "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel family, from the
    committed ncu capture of this same command (profiles/r01_traffic.json); None when absent."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        return json.load(fh)


def make_shard(n, seed, device=None, pin=False):
    """This rank's synthetic shard (SURVEY.md 8(d) cfg 4): randn images, ids with [CLS] first and 0 on
    the padded tail, prefix masks with L ~ U{16..128}."""
    import torch

    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n, 3, IMG, IMG, generator=g)
    ids = torch.randint(1, 28996, (n, SEQ), generator=g)
    lengths = torch.randint(16, SEQ + 1, (n,), generator=g)
    mask = (torch.arange(SEQ).unsqueeze(0) < lengths.unsqueeze(1)).long()
    ids = ids * mask
    ids[:, 0] = 101
    if pin:
        return images.pin_memory(), ids.pin_memory(), mask.pin_memory()
    if device is not None:
        return images.to(device), ids.to(device), mask.to(device)
    return images, ids, mask


def cpu_forward_timer(samples: int, iters: int, warmup: int):
    """Times the oracle port (the reference's algorithm, fp32, ATen CPU kernels, all host threads)."""
    import torch

    import synth
    from oracle import forward_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = synth.build_model(0)
    sd = {k: v.float() for k, v in model.state_dict().items() if v.is_floating_point()}
    images, ids, mask = make_shard(samples, 4321)
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        oracle.multimodal_forward(sd, images, ids, mask)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return samples / statistics.mean(times), statistics.mean(times), cores, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    samples = args.cpu_samples
    value, sec, cores, threads = cpu_forward_timer(samples, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{samples} samples of the workload per step (oracle/forward_oracle.py, "
                                   f"fp32, {threads} threads of {cores} host cores)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _config(args, world):
    return {"workload": "configs[3]: full multimodal inference (ResNet50 + BioBERT-base + attention fusion + "
                        f"head), global batch {args.global_batch} data-parallel, 224x224 fp32 images, "
                        f"{SEQ} tokens, padding masks L~U{{16..128}}, logits all-gather",
            "global_batch": args.global_batch, "per_gpu_batch": -(-args.global_batch // world),
            "seq_len": SEQ, "image": IMG, "parallelism": f"dp{world}",
            "l2": "per-rank inputs exceed L2 (>= 308 MB vs 126 MB); no flush between steps",
            "img_chunk": args.img_chunk, "tok_chunk": args.tok_chunk}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--global-batch", type=int, default=4096)
    ap.add_argument("--img-chunk", type=int, default=0)
    ap.add_argument("--tok-chunk", type=int, default=0)
    ap.add_argument("--cpu-samples", type=int, default=128,
                    help="samples per CPU pass: the reference arm's step, and (x2) the cpu_baseline leg")
    ap.add_argument("--micro-batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    # dump all stacks and exit instead of hanging a GPU box (a healthy run takes 1-2 minutes)
    import faulthandler

    faulthandler.dump_traceback_later(int(os.environ.get("MRD_BENCH_WATCHDOG", "1500")), exit=True)
    import torch
    import torch.distributed as dist

    import mrd_b200
    import synth

    # torchrun pins OMP_NUM_THREADS=1; the synthetic host data (randn of the image shard) and the
    # model construction are CPU work outside the timed region - give them the rank's share of cores
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total = args.global_batch
    lo, hi = mrd_b200.shard_bounds(total, world, rank)
    n_local = hi - lo
    model = synth.build_model(0).to(dev)
    model.configure_b200(args.img_chunk, args.tok_chunk)
    eng = model._engine()
    dp = mrd_b200.DataParallelForward(
        lambda im, i, m, out: model(im, i, m, logits_out=out), model.num_classes)

    # ---------------------------------------------------------------- value: inputs resident in HBM
    images, ids, mask = make_shard(n_local, 1234 + rank, device=dev)

    def step():
        with torch.no_grad():
            return dp.forward_shard(images, ids, mask, total)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        logits = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if sampler else None
    # tokens the packed BERT path actually processes (mask != 0, CLS always kept) vs the dense count
    live_frac = float(((mask != 0) | (torch.arange(SEQ, device=dev) == 0)).float().mean().item())
    t = torch.tensor([ms, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, launches = tmax[0].item(), int(t[1].item())
    ms_per_step = ms / args.steps
    value = total / (ms_per_step * 1e-3)
    assert logits.shape == (total, model.num_classes) and bool(torch.isfinite(logits).all())

    # ---------------------------------------------------------------- e2e: host buffers -> logits on host
    e2e = None
    if not args.no_e2e:
        h_images, h_ids, h_mask = make_shard(n_local, 1234 + rank, pin=True)
        h_logits = torch.empty(total, model.num_classes, dtype=torch.float32).pin_memory()

        def e2e_step():
            with torch.no_grad():
                out = dp.forward_shard_host(
                    lambda im, i, m, o: model.forward_host(im, i, m, micro_batch=args.micro_batch, logits_out=o,
                                                           next_batch=(im, i, m)),
                    h_images, h_ids, h_mask, total, dev)
                h_logits.copy_(out, non_blocking=True)

        for _ in range(args.warmup):
            e2e_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ems = tt.item()
        h2d = sum(x.numel() * x.element_size() for x in (h_images, h_ids, h_mask))
        e2e = {"value": total / (ems / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h_logits.numel() * 4,
               "ms_per_step": ems / args.steps,
               "path": "pinned host tensors -> MultimodalClassifier.forward_host (H2D of micro-batch i+1 on a copy "
                       f"stream under the kernels of micro-batch i, micro_batch={args.micro_batch}; the first micro-batch of the next "
                       f"step is copied under the last one of this step, as a loader that holds the next batch would) -> logits "
                       "all-gather -> pinned host"}

    # ---------------------------------------------------------------- roofline: profiled step (rank 0)
    roofline, families = None, None
    peaks = _peaks()
    if rank == 0:
        eng.profile(True)
        with torch.no_grad():
            model(images, ids, mask)  # local shard only: no collective, the other ranks are not in this step
        rows = eng.profile_report()
        eng.profile(False)
        tens = [r for r in rows if r["cat"] == "tensor"]
        t_ms = sum(r["ms"] for r in tens)
        t_fl = sum(r["flops"] for r in tens)
        n_l = sum(r["launches"] for r in tens)
        all_ms = sum(r["ms"] for r in rows)
        achieved = t_fl / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM: 53 convs + all linears)",
                    "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_sustained"], "peak_source": peaks["src"] + " (sustained: timed inside a long step)",
                    "launches_per_step": n_l, "avg_launch_ms": t_ms / max(n_l, 1),
                    "flops_per_launch_avg": t_fl / max(n_l, 1), "share_of_step": t_ms / all_ms if all_ms else None,
                    # DRAM bytes per launch of this kernel family (dram__bytes_read.sum + dram__bytes_write.sum, ncu);
                    # the capture it comes from is described in traffic_detail
                    "traffic": (_ncu_traffic() or {}).get("dram_bytes_per_launch"),
                    "traffic_detail": _ncu_traffic(),
                    "note": "achieved counts ALGORITHMIC flops (dense, as the reference executes; SURVEY 8(d)). "
                            f"The BERT launches skip padded tokens (live fraction {live_frac:.3f} of B*S) and the last "
                            "layer's post-attention half runs on CLS rows only, so executed FLOP/s are lower: "
                            "see profiles/ and DESIGN.md section 5."}
        fam = {}
        for r in rows:
            f = fam.setdefault(r["cat"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            f["ms"] += r["ms"]; f["flops"] += r["flops"]; f["bytes"] += r["bytes"]; f["launches"] += r["launches"]
        families = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                        "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["ms"] else 0,
                        "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] else 0}
                    for k, v in fam.items()}
        if args.profile_out:
            with open(args.profile_out, "w") as fh:
                fh.write("label,category,launches,total_ms,tflops,gbs,share\n")
                for r in sorted(rows, key=lambda r: -r["ms"]):
                    fh.write(f'{r["label"]},{r["cat"]},{r["launches"]},{r["ms"]:.4f},'
                             f'{r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] else 0:.1f},'
                             f'{r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] else 0:.1f},'
                             f'{r["ms"] / all_ms:.4f}\n')

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = 2 * args.cpu_samples   # ~6 s per pass on 16 cores: 1 warm-up + 2 timed passes = ~20 s of CPU work
        v, sec, cores, threads = cpu_forward_timer(n_cpu, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_cpu} samples of the same workload, 1 warm-up + 2 timed passes, "
                         f"{sec:.2f} s per pass (oracle/forward_oracle.py, fp32, {threads} threads of {cores} host cores)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": _config(args, world),
            "tensor_peak_frac": value * FLOP_PER_SAMPLE / (world * peaks["bf16_burst"] * 1e12),
            "tensor_peak_tflops": peaks["bf16_burst"], "flop_per_sample": FLOP_PER_SAMPLE,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "kernel_families": families, "cpu_baseline": cpu, "bert_live_token_fraction": live_frac,
            "device_bytes": eng.device_bytes,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

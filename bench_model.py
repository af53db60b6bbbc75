#!/usr/bin/env python
"""Full-model legs of the benchmark (BASELINE.json configs[0], [2], [3]); imported by bench.py, runnable on its own.

    python bench_model.py train  [--batch 32] [--img 128] [--steps 10] [--warmup 3] [--variant dropin|reference] [--fp32]
    python bench_model.py infer  [--batch 64] [--img 256] [--steps 5]  [--warmup 2] [--variant dropin|reference]
    python bench_model.py breakdown [--batch 32] [--img 128]          (torch.profiler kernel table of one training step)
    torchrun --nproc-per-node N bench_model.py train ...               (batch-sharded DP, one process per GPU)

train : one step = train.py:132-146 of the reference (forward, enRainfallLoss, backward, clip 0.025, AdamW, zero_grad) on a
        synthetic Shanghai-shaped batch (B, 25, 1, img, img) -> 5 input / 20 target frames, bf16 autocast with fp32 master
        weights, through adnm_unet_b200.trainer.DataParallelTrainer (bucketed NCCL all-reduce overlapped with backward).
        seq/s = B * world / step time (CUDA events, max over ranks).
infer : validate.py:96-106 minus its per-batch .cpu().numpy(): eval + no_grad forward of (B, 5, 1, img, img), then the
        device threshold counts of the predictions (adn_threshold_counts) - the on-device evaluation path.
The host network is the UNMODIFIED reference from git-ignored baseline/_ref (adnm_unet_b200.refhost); `--variant reference`
runs the same step with the reference's own Mamba2 / WTConv2d (eager PyTorch on the B200: the like-for-like GPU baseline).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# full-model roofline per sample (probe numbers): unit-boundary activation traffic fwd+bwd and dense FLOPs
TRAIN_BYTES_PER_SAMPLE = {128: 445e6, 256: 1.78e9}
TRAIN_FLOP_PER_SAMPLE = {128: 46.4e9, 256: 188.1e9}


def _host_note(variant):
    if variant != "dropin":
        return "the unmodified reference network (baseline/_ref), eager PyTorch"
    dead = os.environ.get("ADNM_KEEP_DEAD_BRIDGES", "0") != "1"
    return ("unmodified reference network with Mamba2 / WTConv2d / Block / RMSNorm / FeedForward / StandardAttention / WTLayer / PatchEmbed / OutProj bound to the "
            "sm_100a modules" + ("; the four EncoderToDecoder bridges whose outputs never reach the network output (e2ds[3..6]) "
                                 "are skipped - bit-identical outputs and gradients, refhost.prune_dead_bridges" if dead else ""))


def _dist_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def _timed(torch, dist, world, dev, fn, steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def train_bench(batch=32, img=128, steps=10, warmup=3, variant="dropin", bf16=True, e2e=True, init_dist=True, peaks=None):
    """Returns a dict (rank 0) / None (other ranks).  The process group may already be initialised by the caller."""
    import torch
    import torch.distributed as dist
    from adnm_unet_b200 import _lib, refhost
    from adnm_unet_b200.trainer import DataParallelTrainer

    world, rank, local = _dist_env()
    if world > 1:
        os.environ.setdefault("ADN_SM_RESERVE", "1")      # one SM for the NCCL CTAs beside the persistent kernels (read once by the library)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized() and init_dist:
        dist.init_process_group("nccl", device_id=dev)
    model = refhost.build_adnm_unet(img, dropin=(variant == "dropin"), seed=0).to(dev)     # identical weights on every rank
    use_graph = variant == "dropin" and os.environ.get("ADNM_TRAIN_GRAPH", "1") != "0"
    trainer = DataParallelTrainer(model, refhost.reference_loss(), autocast_dtype=torch.bfloat16 if bf16 else None, graph=use_graph)
    g = torch.Generator().manual_seed(1000 + rank)                                     # per-rank data shard
    NSETS = 2
    host = [torch.rand(batch, 25, 1, img, img, generator=g).pin_memory() for _ in range(NSETS)]
    resident = [h.to(dev) for h in host]
    losses = torch.zeros(max(steps, warmup, 3) + 3, device=dev)
    host_loss = torch.zeros(1).pin_memory()

    def step_resident(i):
        d = resident[i % NSETS]
        losses[i] = trainer.step(d[:, :5], d[:, 5:])

    def step_e2e(i):
        d = host[i % NSETS].to(dev, non_blocking=True)          # H2D of the step's batch from pinned memory
        loss = trainer.step(d[:, :5], d[:, 5:])
        host_loss.copy_(loss.reshape(1), non_blocking=True)     # D2H of the step's loss (what train.py:146 reads)

    step_resident(0)                       # discovery step (flat buffers), always eager
    n0 = _lib.launch_count()
    step_resident(1)                       # one eager step on the flat buffers: the library launches a step consists of
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - n0
    for i in range(2, max(warmup, 3) + 2):  # capture (graph mode) + warm replays
        step_resident(i)
    torch.cuda.synchronize()
    mem = torch.cuda.max_memory_allocated(dev)
    ms = _timed(torch, dist, world, dev, step_resident, steps)
    ms_e2e = _timed(torch, dist, world, dev, step_e2e, steps) if e2e else None
    final_loss = float(losses[steps - 1])
    if rank != 0:
        return None
    step_ms = ms / steps
    seq_s = batch * world * steps / (ms * 1e-3)
    hbm, tf = peaks if peaks else (6524.9, 1400.6)
    t_star = max(TRAIN_BYTES_PER_SAMPLE.get(img, 0) / (hbm * 1e9), TRAIN_FLOP_PER_SAMPLE.get(img, 0) / (tf * 1e12))
    res = {
        "metric": "adnm_unet_train_seq_per_s", "value": seq_s, "unit": "seq/s", "n_gpus": world, "ms_per_step": step_ms,
        "steps": steps, "warmup": max(warmup, 3), "variant": variant, "dtype": "bf16 autocast, fp32 master weights" if bf16 else "f32",
        "config": {"workload": f"ADNM-UNet training step (BASELINE configs[2]): B={batch}/GPU, 5->20 frames at {img}x{img}, "
                               "enRainfallLoss, clip 0.025, AdamW", "global_batch": batch * world, "parallelism": f"dp{world}",
                   "host": _host_note(variant),
                   "launch": ("cuda_graph_replay: forward + loss + backward captured once (static input buffers), gradient all-reduce and the "
                              "clip / AdamW kernels launched per step" if trainer._graph is not None else
                              "eager" + (f" (graph capture abandoned: {trainer.graph_error})" if trainer.graph_error else "")),
                   "grad_allreduce": f"{len(trainer.buckets)} fp32 buckets ({trainer.n_live_elements()} live elements of "
                                     f"{sum(p.numel() for p in model.parameters())}), NCCL sum " + ("after the replayed backward" if trainer._graph is not None else "overlapped with backward") if world > 1 else "none (1 GPU)"},
        "live_param_tensors": len(trainer.live), "final_loss": final_loss, "grad_norm": float(trainer.grad_norm()),
        "lib_launches_per_step": launches_per_step, "peak_mem_gb": mem / 2**30,
        "roofline": {"per_sample_t_star_us": t_star * 1e6, "ceiling_seq_per_s_per_gpu": (1 / t_star) if t_star else None,
                     "frac": (seq_s / world * t_star) if t_star else None,
                     "note": "t* = max(445 MB / HBM, 46.4 GFLOP / bf16 sustained) per sample at 128^2 (SURVEY 8(d) full-model row)"},
    }
    if ms_e2e is not None:
        res["e2e"] = {"value": batch * world * steps / (ms_e2e * 1e-3), "unit": "seq/s", "ms_per_step": ms_e2e / steps,
                      "h2d_bytes_per_step": world * batch * 25 * img * img * 4, "d2h_bytes_per_step": world * 4}
    return res


def infer_bench(batch=64, img=256, steps=5, warmup=2, variant="dropin", bf16=True, cpu_eval_sample=2):
    """validate.py:92-118 on the device: eval + no_grad forward, then SimplifiedEvaluator.evaluate of the predictions - the
    on-device evaluation path (adn_eval_batch: threshold counts + per-lead-time MSE in one kernel, no .cpu().numpy()).
    `eval_baseline`: the reference's own evaluator (float2int + the Python loops over batch x frame x threshold around
    _cal_frame, datasets/Shanghai_metrics.py:61-80, LPIPS / SSIM skipped) timed on the host on `cpu_eval_sample` samples."""
    import numpy as np
    import torch
    from adnm_unet_b200 import refhost
    from adnm_unet_b200.evaluator import SimplifiedEvaluator
    world, rank, local = _dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model = refhost.build_adnm_unet(img, dropin=(variant == "dropin"), seed=0).to(dev).eval()
    g = torch.Generator().manual_seed(7 + rank)
    x = torch.rand(batch, 5, 1, img, img, generator=g).to(dev)
    tgt = torch.rand(batch, 20, img, img, generator=g).to(dev)
    ev = SimplifiedEvaluator(seq_len=20, value_scale=90, thresholds=[20, 30, 35, 40], device=dev)
    last = {}

    def fwd():
        with torch.no_grad():
            if bf16:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return model(x)
            return model(x)

    graph, graph_note = None, "eager"

    def step(i):
        if graph is not None:
            graph.replay()
            out = last["static_out"]
        else:
            out = fwd()
        last["out"] = out
        with torch.no_grad():
            ev.evaluate(tgt, out.squeeze(2).float())

    for i in range(warmup):
        step(i)
    if variant == "dropin" and os.environ.get("ADNM_INFER_GRAPH", "1") != "0":
        # the eval forward captured once into a CUDA graph (static input buffer `x`), replayed per batch; the evaluator kernel stays eager
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                fwd()
            torch.cuda.current_stream(dev).wait_stream(side)
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                last["static_out"] = fwd()
            graph, graph_note = g_, "cuda_graph_replay of the eval forward (static input buffer); evaluator kernel launched per batch"
            step(0)
        except Exception as ex:
            graph, graph_note = None, f"eager (graph capture abandoned: {type(ex).__name__}: {ex})"[:300]
            torch.cuda.synchronize()
    ev.reset()
    ms = _timed(torch, None, 1, dev, step, steps)
    table = ev.counts().clone()
    res = ev.done()
    # the evaluation alone, device vs the reference's host loops
    pred = last["out"].squeeze(2).float().contiguous()
    ev.reset()
    ms_eval = _timed(torch, None, 1, dev, lambda i: ev.evaluate(tgt, pred), 5) / 5
    nb = min(cpu_eval_sample, batch)
    p_np, t_np = pred[:nb].cpu().numpy(), tgt[:nb].cpu().numpy()
    t0 = time.perf_counter()
    pi, ti = (np.clip(p_np, 0.0, 1.0) * 90).astype(np.uint16), (np.clip(t_np, 0.0, 1.0) * 90).astype(np.uint16)
    for thr in (20, 30, 35, 40):
        for b in range(nb):
            for f in range(20):
                ob, sb = (ti[b][f] >= thr).astype(int), (pi[b][f] >= thr).astype(int)
                _ = (np.sum((ob == 1) & (sb == 1)), np.sum((ob == 1) & (sb == 0)), np.sum((ob == 0) & (sb == 1)), np.sum((ob == 0) & (sb == 0)))
    cpu_s = time.perf_counter() - t0
    csi = [res["threshold_metrics"][t]["CSI"] for t in (20, 30, 35, 40)]
    hss = [res["threshold_metrics"][t]["HSS"] for t in (20, 30, 35, 40)]
    return {"metric": "adnm_unet_infer_seq_per_s", "value": batch * steps / (ms * 1e-3), "unit": "seq/s", "ms_per_step": ms / steps,
            "variant": variant, "dtype": "bf16 autocast" if bf16 else "f32",
            "config": {"workload": f"ADNM-UNet inference (BASELINE configs[3], validate.py:92-118): B={batch}, 5->20 frames at {img}x{img}, "
                                   "eval + no_grad forward + on-device SimplifiedEvaluator (threshold counts + RMSE)",
                       "host": _host_note(variant), "launch": graph_note},
            "counts_table": table.cpu().tolist(), "csi": csi, "hss": hss, "rmse": res["RMSE"], "far": res["FAR"],
            "evaluator": {"device_ms_per_batch": ms_eval, "device_samples_per_s": batch / (ms_eval * 1e-3),
                          "eval_baseline": {"kind": "reference loops (float2int + _cal_frame over batch x frame x threshold) on the host",
                                            "samples": nb, "samples_per_s": nb / cpu_s}},
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}


def breakdown(batch=32, img=128, variant="dropin", top=45):
    """Kernel-time table of one training step (torch.profiler, CUDA activities), grouped by kernel name."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    from adnm_unet_b200 import refhost
    from adnm_unet_b200.trainer import DataParallelTrainer
    dev = torch.device("cuda", 0)
    model = refhost.build_adnm_unet(img, dropin=(variant == "dropin"), seed=0).to(dev)
    trainer = DataParallelTrainer(model, refhost.reference_loss())
    d = torch.rand(batch, 25, 1, img, img).to(dev)
    for _ in range(3):
        trainer.step(d[:, :5], d[:, 5:])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        trainer.step(d[:, :5], d[:, 5:])
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if e.device_type == torch.autograd.DeviceType.CUDA and t > 0:
            rows.append((e.key[:110], t, e.count))
    rows.sort(key=lambda r: -r[1])
    total = sum(r[1] for r in rows)
    return {"variant": variant, "total_kernel_ms": total / 1e3, "n_kernel_names": len(rows), "n_launches": sum(r[2] for r in rows),
            "top": [{"kernel": k, "ms": t / 1e3, "count": c, "share": t / total} for k, t, c in rows[:top]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["train", "infer", "breakdown"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--img", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variant", default="dropin", choices=["dropin", "reference"])
    ap.add_argument("--fp32", action="store_true")
    a = ap.parse_args()
    if a.mode == "train":
        r = train_bench(a.batch or 32, a.img or 128, a.steps, a.warmup, a.variant, not a.fp32)
    elif a.mode == "infer":
        r = infer_bench(a.batch or 64, a.img or 256, a.steps, a.warmup, a.variant, not a.fp32)
    else:
        r = breakdown(a.batch or 32, a.img or 128, a.variant)
    if r is not None:
        print(json.dumps(r))
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

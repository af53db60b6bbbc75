#!/usr/bin/env python
"""bench.py - ADN-SSD mixer fwd+bwd tokens/s (BASELINE.json configs[1]) on N B200s of one node, plus the full-model legs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--d-model 32] [--no-model]

One "step" = one forward + backward of one ADN-SSD mixer (models/ADNssd.py::Mamba2 of the reference) over a
synthetic batch B=16 of 128x128 token grids (262 144 tokens per GPU), bf16 I/O, random-init weights.
  value    : tokens/s with u / dout already resident in HBM (CUDA events, max over ranks)
  e2e      : the same through the public module call (adnm_unet_b200.Mamba2) with HOST buffers: pinned-host u and
             dout copied H2D and out / du copied D2H inside the timed region, every step
  roofline : SURVEY.md 8(d): the step's compulsory HBM traffic (10*D bytes per token) / the step time vs MEASURED_PEAKS.json
             (`frac`), the tensor-or-HBM bound t* next to it, the per-step DRAM traffic of the committed ncu capture
             (`traffic`), and the dominant kernel's own DRAM throughput under `dominant_kernel`
  cpu_baseline : the UNMODIFIED reference Mamba2 (git-ignored baseline/_ref, shims of adnm_unet_b200.refhost) on the box's
             host cores, fp32, bounded sample;  gpu_baseline : the same module eagerly on the B200 (fp32 and bf16 autocast)
  train / infer : BASELINE configs[2] / [3] through bench_model.py - full ADNM-UNet training step (B=32/GPU at 128x128,
             batch-sharded DP over all N ranks, bucketed NCCL all-reduce overlapped with backward) and, at N=1, the
             validate.py-shaped inference at 256x256 with on-device threshold counts; each next to the eager reference model.
N > 1 (torchrun): batch-sharded data parallel, one process per GPU, weak scaling; the mixer's parameter gradients are
all-reduced over NCCL every step (the only exchange the path has) on a side stream that overlaps the next step's kernels.
`--impl reference` times the reference's own CPU implementation of the path (the unmodified modules; the closed-form oracle
port only if baseline/_ref is absent) on the host cores for the same metric and config, plus configs[0] (full model on CPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "adnssd_fwd_bwd_tokens_per_s"
UNIT = "tokens/s"
B_PER_GPU, GRID = 16, 128
HEADDIM, D_STATE = 4, 16
N_INPUT_SETS = 8   # rotating (u, dout) sets: 8 x 33.5 MB (bf16) > L2, on top of ~1 GB of intermediates rewritten each step


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # under load = upper half of the samples (the sampler also sees the idle edges of the region)
        med = sm[(len(sm) * 3) // 4] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step_fn(d_model, batch, threads):
    """The CPU algorithm of the reference (oracle port), fwd + autograd bwd, fp32.  Fallback when baseline/_ref is absent."""
    import torch
    from oracle import adnssd_oracle as AO
    torch.set_num_threads(threads)
    p = {k: v.clone().requires_grad_(k not in AO.UNUSED_PARAMS) for k, v in AO.init_params(d_model, HEADDIM, D_STATE, seed=0).items()}
    g = torch.Generator().manual_seed(1234)
    u = torch.randn(batch, GRID * GRID, d_model, generator=g).requires_grad_(True)
    dout = torch.randn(batch, GRID * GRID, d_model, generator=g)

    def step():
        for v in p.values():
            v.grad = None
        u.grad = None
        AO.mixer_forward(p, u, GRID, GRID, HEADDIM, D_STATE).backward(dout)
    return step


def reference_step_fn(d_model, batch, threads):
    """The UNMODIFIED reference module models/ADNssd.py::Mamba2 on the host cores, fp32, fwd + autograd bwd."""
    import torch
    from adnm_unet_b200 import refhost
    torch.set_num_threads(threads)
    ns = refhost.load_reference()
    torch.manual_seed(0)
    m = ns.ref_Mamba2(d_model=d_model, headdim=HEADDIM, d_state=D_STATE)
    g = torch.Generator().manual_seed(1234)
    u = torch.randn(batch, GRID * GRID, d_model, generator=g).requires_grad_(True)
    dout = torch.randn(batch, GRID * GRID, d_model, generator=g)

    def step():
        m.zero_grad(set_to_none=True)
        u.grad = None
        with refhost.cuda_to_is_noop(force=True):
            m(u, GRID, GRID).backward(dout)
    return step


def cpu_kind():
    from adnm_unet_b200 import refhost
    return "reference" if refhost.reference_available() else "port"


def time_cpu(d_model, batch, steps, warmup, budget_s=None):
    """tokens/s of the CPU arm; with `budget_s` the timed loop stops early once the budget is spent (>= 2 steps)."""
    threads = os.cpu_count() or 1
    kind = cpu_kind()
    step = (reference_step_fn if kind == "reference" else oracle_step_fn)(d_model, batch, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if budget_s is not None and done >= 2 and time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return batch * GRID * GRID / dt, dt, threads, kind, done


def full_model_cpu(steps=3, warmup=1):
    """BASELINE configs[0]: the unmodified ADNM-UNet fwd + enRainfallLoss + bwd on the host cores, B=1, 5->20 frames, 128x128."""
    import torch
    from adnm_unet_b200 import refhost
    if not refhost.reference_available():
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    with refhost.cuda_to_is_noop(force=True):
        model = refhost.build_adnm_unet(128, dropin=False, seed=0)
        loss_fn = refhost.reference_loss()
        data = torch.rand(1, 25, 1, 128, 128, generator=torch.Generator().manual_seed(0))
        ts = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model.zero_grad(set_to_none=True)
            loss_fn(model(data[:, :5]), data[:, 5:]).backward()
            ts.append(time.perf_counter() - t0)
    ts = sorted(ts[warmup:])
    med = ts[len(ts) // 2]
    return {"metric": "adnm_unet_fwd_bwd_seq_per_s_cpu", "value": 1.0 / med, "unit": "seq/s", "s_per_step": med,
            "cores": os.cpu_count() or 1, "config": {"workload": "ADNM-UNet fwd + enRainfallLoss + bwd on CPU (BASELINE configs[0]): B=1, "
                                                                 "5->20 frames at 128x128, standalone RMSNorm, fp32, unmodified reference"}}


CPU_BATCH = B_PER_GPU        # the CPU arm runs the SAME batch as one GPU (B=16): same config, per-token metric


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = CPU_BATCH
    steps, warmup = max(2, min(args.steps, 40)), max(1, min(args.warmup, 3))
    val, dt, threads, kind, done = time_cpu(args.d_model, batch, steps, warmup, budget_s=90.0)     # a few minutes at most
    what = "unmodified reference Mamba2 (baseline/_ref)" if kind == "reference" else "oracle port (baseline/_ref absent)"
    sample = f"{what}, PyTorch fp32 CPU, B={batch} x {GRID}x{GRID} tokens per step, {done} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ADN-SSD mixer fwd+bwd (BASELINE configs[1]): D={args.d_model}, headdim={HEADDIM}, d_state={D_STATE}, "
                               f"B={B_PER_GPU}/GPU, {GRID}x{GRID} tokens",
                   "note": "CPU arm: the reference's own CPU path on all host cores, same batch as one GPU of the GPU arm"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_model:
        fm = full_model_cpu()
        if fm is not None:
            line["full_model_cpu"] = fm
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import adnm_unet_b200 as A
    from adnm_unet_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # leave one SM to the NCCL all-reduce that runs beside the persistent (one CTA per SM) kernels; read once by the library
        os.environ.setdefault("ADN_SM_RESERVE", "1")
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its banner there)
        # NCCL prints its version banner on stdout when the communicator is created; stdout must stay the ONE JSON line
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))      # communicator fully set up before any CUDA-graph capture
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.load()
    assert lib.adn_device_supported() == 1

    D = args.d_model
    torch.manual_seed(0)                                  # identical weights on every rank
    mixer = A.Mamba2(d_model=D, headdim=HEADDIM, d_state=D_STATE).to(dev)
    params = [p for n, p in mixer.named_parameters() if n not in ("scale", "shift", "alpha2")]
    B, L = B_PER_GPU, GRID * GRID
    tokens = B * L
    g = torch.Generator().manual_seed(1234 + rank)        # per-rank data shard
    host_u = [torch.randn(B, L, D, generator=g).bfloat16().pin_memory() for _ in range(N_INPUT_SETS)]
    host_g = [torch.randn(B, L, D, generator=g).bfloat16().pin_memory() for _ in range(N_INPUT_SETS)]
    dev_u = [t.to(dev) for t in host_u]
    dev_g = [t.to(dev) for t in host_g]
    from adnm_unet_b200.dp import GradAllReducer
    reducer = GradAllReducer(params) if world > 1 else None

    def allreduce_grads():
        if reducer is not None:
            reducer()       # NCCL all-reduce(sum) / world of the flat parameter-gradient bucket

    def step_eager(i):
        u = dev_u[i % N_INPUT_SETS].requires_grad_(True)
        u.grad = None
        for p in params:
            p.grad = None
        out = mixer(u, GRID, GRID)
        out.backward(dev_g[i % N_INPUT_SETS])
        allreduce_grads()
        return out, u.grad

    # ---- CUDA graphs: one captured fwd+bwd of the public module per static (u, dout) pair, replayed in the timed loops
    # (the C ABI enqueues on the caller's stream without host synchronisation, so the whole step is capturable; the host
    # then issues one cudaGraphLaunch per step instead of ~15 launches through Python / autograd).
    pool = torch.cuda.graph_pool_handle() if args.graph else None

    # Data-parallel step as ONE cudaGraphLaunch with the gradient all-reduce overlapped: graph k copies its flat gradient
    # buffer (50 KB) into a pre-allocated staging buffer at its end, and all-reduces the staging buffer of the PREVIOUS step
    # on a forked branch that runs beside its own kernels (an all-reduce placed after the kernels of the same graph measured
    # 0.312 ms / step at N = 2: its latency is exposed; a Python-issued side-stream all-reduce 0.294 ms).
    n_grad = sum(p.numel() for p in params)
    stage = [torch.zeros(n_grad, dtype=torch.float32, device=dev) for _ in range(2)] if world > 1 else None
    cap_count = [0]

    def capture(u_static, g_static):
        u_static.requires_grad_(True)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                       # warm-up outside capture (PyTorch CUDA-graph recipe)
            for _ in range(2):
                u_static.grad = None
                for p in params:
                    p.grad = None
                mixer(u_static, GRID, GRID).backward(g_static)
        torch.cuda.current_stream().wait_stream(s)
        u_static.grad = None
        for p in params:
            p.grad = None
        graph = torch.cuda.CUDAGraph()
        k = cap_count[0]
        cap_count[0] += 1
        with torch.cuda.graph(graph, pool=pool):
            in_graph = world > 1 and args.graph_nccl
            if in_graph:      # forked branch: all-reduce of the previous step's staged gradients, beside this step's kernels
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    dist.all_reduce(stage[(k + 1) % 2], op=dist.ReduceOp.AVG)
            out = mixer(u_static, GRID, GRID)
            out.backward(g_static)
            # the 18 parameter gradients are views of ONE flat fp32 buffer (adnm_unet_b200/mixer.py)
            flat = torch.empty(0, dtype=torch.float32, device=dev).set_(params[0].grad.untyped_storage())
            if in_graph:
                cur.wait_stream(side)                      # join: the previous all-reduce is done before ...
                stage[k % 2].copy_(flat)                   # ... this step's gradients are staged for the next graph
        assert flat.numel() == sum(p.numel() for p in params), "flat gradient buffer layout changed"
        return graph, out.detach(), u_static.grad, flat

    n0 = _lib.launch_count()
    step_eager(0)
    launches_per_step = _lib.launch_count() - n0
    graphs = [capture(dev_u[k], dev_g[k]) for k in range(N_INPUT_SETS)] if args.graph else None

    # Gradient all-reduce on its own stream: every captured graph owns its flat gradient buffer, so the NCCL call of step
    # i overlaps the kernels of step i+1 (it only has to finish before the same graph is replayed again).
    s_comm = torch.cuda.Stream(device=dev) if world > 1 else None
    ev_step = [torch.cuda.Event() for _ in range(N_INPUT_SETS)]
    ev_comm = [torch.cuda.Event() for _ in range(N_INPUT_SETS)]

    def allreduce_async(k, flat, cur):
        ev_step[k].record(cur)
        with torch.cuda.stream(s_comm):
            s_comm.wait_event(ev_step[k])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            ev_comm[k].record(s_comm)

    side_allreduce = world > 1 and not args.graph_nccl

    def step_resident(i):
        if graphs is None:
            return step_eager(i)
        k = i % N_INPUT_SETS
        graph, out, du, flat = graphs[k]
        cur = torch.cuda.current_stream()
        if side_allreduce:
            cur.wait_event(ev_comm[k])        # the previous all-reduce of this graph's gradient buffer has finished
        graph.replay()
        if side_allreduce:
            allreduce_async(k, flat, cur)
        return out, du

    def drain_resident():
        if world > 1:
            torch.cuda.current_stream().wait_stream(s_comm)
            if args.graph and args.graph_nccl:             # the last step's staged gradients (inside the timed region)
                dist.all_reduce(stage[0], op=dist.ReduceOp.AVG)
                dist.all_reduce(stage[1], op=dist.ReduceOp.AVG)

    # ---- end-to-end: host buffers in, host buffers out, every step.  Copies run on their own streams so that the H2D of
    # step i+1 and the D2H of step i-1 overlap the kernels of step i (double-buffered device staging slots).
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    NSLOT = 2
    slot_u = [torch.empty(B, L, D, dtype=torch.bfloat16, device=dev) for _ in range(NSLOT)]
    slot_g = [torch.empty(B, L, D, dtype=torch.bfloat16, device=dev) for _ in range(NSLOT)]
    host_outs = [torch.empty(B, L, D, dtype=torch.bfloat16).pin_memory() for _ in range(NSLOT)]
    host_dus = [torch.empty(B, L, D, dtype=torch.bfloat16).pin_memory() for _ in range(NSLOT)]
    ev_in = [torch.cuda.Event() for _ in range(NSLOT)]
    ev_done = [torch.cuda.Event() for _ in range(NSLOT)]
    ev_out = [torch.cuda.Event() for _ in range(NSLOT)]
    keep = [None] * NSLOT

    slot_graphs = [capture(slot_u[k], slot_g[k]) for k in range(NSLOT)] if args.graph else None

    def step_e2e(i):
        k = i % NSLOT
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_done[k])            # the kernels that last read this staging slot have finished
            with torch.no_grad():
                slot_u[k].copy_(host_u[i % N_INPUT_SETS], non_blocking=True)
                slot_g[k].copy_(host_g[i % N_INPUT_SETS], non_blocking=True)
            ev_in[k].record(s_in)
        cur.wait_event(ev_in[k])
        cur.wait_event(ev_out[k])                  # the previous D2H out of this slot's result tensors has finished
        if slot_graphs is None:
            u = slot_u[k].detach().requires_grad_(True)
            for p in params:
                p.grad = None
            out = mixer(u, GRID, GRID)
            out.backward(slot_g[k])
            allreduce_grads()
            res = (out.detach(), u.grad)
        else:
            graph, out, du, flat = slot_graphs[k]
            if side_allreduce:
                cur.wait_event(ev_comm[k])
            graph.replay()
            if side_allreduce:
                allreduce_async(k, flat, cur)      # side stream, as in step_resident (round 1 blocked the compute stream here)
            res = (out, du)
        ev_done[k].record(cur)
        keep[k] = res
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[k])
            host_outs[k].copy_(keep[k][0], non_blocking=True)
            host_dus[k].copy_(keep[k][1], non_blocking=True)
            if slot_graphs is None:
                keep[k][0].record_stream(s_out)
                keep[k][1].record_stream(s_out)
            ev_out[k].record(s_out)

    def drain_e2e():
        cur = torch.cuda.current_stream()
        for k in range(NSLOT):
            cur.wait_event(ev_out[k])
        if side_allreduce:
            cur.wait_stream(s_comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if drain:
            drain()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    n_warm = max(args.warmup, 2 * N_INPUT_SETS)      # every captured graph / input set replayed at least twice before timing
    for i in range(n_warm):
        step_resident(i)
        step_e2e(i)
    # clocks / throttle reasons sampled DURING the timed region: nvidia-smi needs ~0.2 s to produce its first line, so it is
    # started under load (a second warm-up burst) and stopped right after the timed steps
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for i in range(1500):                 # the SAME number of steps on every rank (the steps contain collectives)
        step_resident(i)
    drain_resident()
    torch.cuda.synchronize()
    ms = timed(step_resident, args.steps, drain_resident)
    launches = launches_per_step * args.steps      # kernels of this library per step (counted on an eager step) x steps
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(step_e2e, args.steps, drain_e2e)

    # per-kernel device times (CUDA events recorded by the library on the launching stream), a few extra steps
    barrier()
    with _lib.profile() as prof:
        for i in range(min(args.steps, 5)):
            step_eager(i)
    per = {}
    for name, t in prof.records:
        a = per.setdefault(name, [0.0, 0])
        a[0] += t; a[1] += 1
    nprof = min(args.steps, 5)
    tot = sum(v[0] for v in per.values())
    top = max(per, key=lambda k: per[k][0])
    hbm_peak, tf_peak, peak_src = load_peaks()
    # algorithmic bytes of the whole mixer fwd+bwd: 10*D bytes per token: u, out, dout, u, du in bf16
    alg_bytes_step = 10 * D * tokens
    step_ms = ms / args.steps
    kernels = {k: {"ms_per_step": v[0] / nprof, "launches_per_step": v[1] / nprof, "share": v[0] / tot} for k, v in per.items()}

    # ---- H2D-only probe: what the host side can feed this rank while every other rank does the same (explains e2e scaling)
    def h2d_only(i):
        with torch.no_grad():
            slot_u[i % NSLOT].copy_(host_u[i % N_INPUT_SETS], non_blocking=True)
            slot_g[i % NSLOT].copy_(host_g[i % N_INPUT_SETS], non_blocking=True)
    ms_h2d = timed(h2d_only, min(args.steps, 50))
    h2d_gbs = 2 * tokens * D * 2 * min(args.steps, 50) / (ms_h2d * 1e-3) / 1e9      # per rank, max-over-ranks time

    # ---- like-for-like GPU baseline: the UNMODIFIED reference module, eager PyTorch on this B200 (rank 0 of a 1-GPU run)
    gpu_base = None
    if world == 1 and not args.no_model:
        gpu_base = eager_reference_on_gpu(torch, dev, D, B, L)

    # ---- full-model legs (BASELINE configs[2] at N ranks, configs[3] at N = 1)
    model_legs = {}
    if not args.no_model:
        del graphs, slot_graphs
        torch.cuda.empty_cache()
        model_legs = full_model_legs(world, rank, load_peaks())

    if rank == 0:
        top_ms = per[top][0] / per[top][1]
        # dominant kernel: achieved = bytes it must move per launch / its mean launch time
        top_bytes = KERNEL_ALG_BYTES.get(top, lambda D, T: None)(D, tokens)
        # roofline.achieved = the step's compulsory HBM traffic (10*D bytes per token: u, out, dout, u, du in
        # bf16) / the measured step time; t* = max(HBM time of those bytes, tensor time of 3*F_fwd flops per token)
        Di, GN, nh, CC, dip = _dims(D)
        f_fwd = 2 * D * dip + 2 * GN * Di + 4 * Di * D + 15 * (Di + 2 * GN) + 18 * Di
        t_hbm, t_tc = alg_bytes_step / (hbm_peak * 1e9), 3 * f_fwd * tokens / (tf_peak * 1e12)
        achieved = alg_bytes_step / (step_ms * 1e-3) / 1e9
        traffic = load_step_traffic()
        roof = {"bound": "hbm", "kernel": f"whole fwd+bwd chain ({launches_per_step} launches per step)", "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src, "frac": achieved / hbm_peak,
                "algorithmic_bytes": alg_bytes_step, "traffic": traffic,
                "traffic_over_algorithmic": (traffic / alg_bytes_step) if traffic else None,
                "t_star_us": max(t_hbm, t_tc) * 1e6, "t_star_bound": "tensor" if t_tc > t_hbm else "hbm",
                "frac_of_t_star": max(t_hbm, t_tc) / (step_ms * 1e-3),
                "dominant_kernel": {"name": top, "ms_per_launch": top_ms, "share_of_step": per[top][0] / tot,
                                    "own_bytes_per_launch": top_bytes, "own_GBps": (top_bytes / (top_ms * 1e-3) / 1e9) if top_bytes else None,
                                    "own_frac_of_hbm_peak": (top_bytes / (top_ms * 1e-3) / 1e9 / hbm_peak) if top_bytes else None,
                                    "ncu_dram_bytes_per_launch": load_traffic(top),
                                    "note": "DRAM throughput of the kernel on the bytes its own design moves - NOT the roofline fraction"}}
        # bounded CPU sample (~10 s): the oracle port on every host core, CPU_BATCH samples per step
        cpu_val = None
        if world == 1 and not args.no_cpu:
            cpu_val, cpu_dt, cpu_threads, cpu_kind_, cpu_done = time_cpu(D, CPU_BATCH, 40, 1, budget_s=20.0)
        line = {
            "metric": METRIC, "value": world * tokens * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ADN-SSD mixer fwd+bwd (BASELINE configs[1]): D={D}, headdim={HEADDIM}, d_state={D_STATE}, "
                                   f"B={B}/GPU, {GRID}x{GRID} tokens", "global_batch": B * world, "tokens_per_step": tokens * world,
                       "parallelism": f"dp{world}", "grad_allreduce": ("none (1 GPU)" if world == 1 else "NCCL AVG of one flat fp32 buffer per step inside the step's CUDA graph, pipelined: a forked branch all-reduces the previous step's staged gradients beside this step's kernels" if (args.graph and args.graph_nccl) else "NCCL AVG of one flat fp32 buffer per step on a side stream (overlaps the next step)"), "launch": "cuda_graph_replay" if args.graph else "eager", "l2": f"{N_INPUT_SETS} rotating input sets (> L2) + ~1 GB of intermediates rewritten per step"},
            "e2e": {"value": world * tokens * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": world * 2 * tokens * D * 2, "d2h_bytes_per_step": world * 2 * tokens * D * 2,
                    "ms_per_step": ms_e2e / args.steps,
                    "h2d_only_GBps_per_rank": h2d_gbs,
                    "note": "bound by the 2 x 33.5 MB per step and rank that cross PCIe in each direction; h2d_only is the host->device "
                            "rate one rank sustains while all ranks copy at once (host memory / PCIe root share)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kernels,
        }
        if cpu_val is not None:
            what = "unmodified reference Mamba2 (baseline/_ref)" if cpu_kind_ == "reference" else "oracle port (baseline/_ref absent)"
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": cpu_threads, "kind": cpu_kind_,
                                    "sample": f"{what}, PyTorch fp32, fwd + autograd bwd, B={CPU_BATCH} x {GRID}x{GRID} tokens per step, "
                                              f"{cpu_done} steps after 1 warm-up, {cpu_dt * 1e3:.1f} ms/step"}
        if gpu_base is not None:
            line["gpu_baseline"] = gpu_base
        line.update(model_legs)
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        # leave without tearing NCCL down: a communicator destroyed while CUDA graphs / side streams still reference it has hung
        # at exit (seen with captured collectives); every rank has printed / finished its work after this barrier
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def eager_reference_on_gpu(torch, dev, D, B, L, steps=10):
    """The unmodified reference Mamba2 run eagerly on the B200 (its ten `.to('cuda')` index uploads per forward are real
    here): the like-for-like GPU baseline of SURVEY.md 8(d).  fp32 (what the reference trains in) and bf16 autocast."""
    from adnm_unet_b200 import refhost
    if not refhost.reference_available():
        return {"unavailable": "baseline/_ref not on this box"}
    ns = refhost.load_reference()
    torch.manual_seed(0)
    m = ns.ref_Mamba2(d_model=D, headdim=HEADDIM, d_state=D_STATE).to(dev)
    u = torch.randn(B, L, D, device=dev, requires_grad=True)
    g = torch.randn(B, L, D, device=dev)
    res = {"unit": UNIT, "what": "unmodified reference models/ADNssd.py::Mamba2, eager PyTorch on the same B200, same shape"}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        def step():
            m.zero_grad(set_to_none=True)
            u.grad = None
            if ac:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    o = m(u, GRID, GRID)
                o.float().backward(g)
            else:
                m(u, GRID, GRID).backward(g)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[name] = {"value": B * L / (ms * 1e-3), "ms_per_step": ms}
    del m, u, g
    torch.cuda.empty_cache()
    return res


def full_model_legs(world, rank, peaks):
    """BASELINE configs[2] (training step, all ranks) and configs[3] (inference, 1 GPU) through bench_model.py."""
    import torch
    import bench_model
    from adnm_unet_b200 import refhost
    if not refhost.reference_available():
        return {"train": {"unavailable": "baseline/_ref not on this box (python baseline/fetch_ref.py in the build container)"}}
    pk = (peaks[0], load_sustained_tf())
    out = {}
    tr = bench_model.train_bench(batch=32, img=128, steps=8, warmup=3, variant="dropin", peaks=pk, init_dist=False)
    if rank == 0:
        out["train"] = tr
    torch.cuda.empty_cache()
    if world == 1:
        ref = bench_model.train_bench(batch=32, img=128, steps=4, warmup=3, variant="reference", e2e=False, peaks=pk)
        out["train"]["gpu_baseline"] = {"value": ref["value"], "unit": "seq/s", "ms_per_step": ref["ms_per_step"],
                                        "what": "the same step with the reference's own Mamba2 / WTConv2d (eager PyTorch, bf16 autocast) on the same B200"}
        torch.cuda.empty_cache()
        inf = bench_model.infer_bench(batch=64, img=256, steps=3, warmup=1, variant="dropin")
        torch.cuda.empty_cache()
        iref = bench_model.infer_bench(batch=64, img=256, steps=2, warmup=1, variant="reference")
        inf["gpu_baseline"] = {"value": iref["value"], "unit": "seq/s", "ms_per_step": iref["ms_per_step"],
                               "what": "the reference's own modules, eager PyTorch bf16 autocast, same B200"}
        out["infer"] = inf
        torch.cuda.empty_cache()
    return out


def load_sustained_tf():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f).get("bf16_tflops_sustained", 1400.0)
    return 1400.0


# algorithmic HBM bytes per launch of each kernel (bf16 elements = 2 bytes), D = d_model, T = tokens; see DESIGN.md
def _dims(D):
    Di = 2 * D
    GN = 2 * D_STATE
    nh = Di // HEADDIM
    CC = 2 * Di + 2 * GN
    return Di, GN, nh, CC, CC + nh


# bytes each kernel must move once per token (bf16 = 2 bytes)
def _alg(D):
    Di, GN, nh, CC, dip = _dims(D)
    return {
        # tile kernels (token grids that are not 128 wide)
        "k_inproj": 2 * (D + dip),
        "k_conv_fwd_tile": 2 * 3 * CC,
        "k_state": 2 * (Di + GN + nh),
        "k_readout": 2 * (2 * Di + GN) + 2 * D,
        "k_bwd1": 2 * D + 4 * (2 * Di + GN),
        "k_bwd2": 2 * (2 * Di + GN + nh) + 2 * (Di + GN + nh),
        "k_conv_bwd_tile": 2 * 4 * CC,
        "k_bwd4": 2 * (dip + 2 * D),
        # row kernels + warp-specialised backward (128-wide grids: the benchmark shape)
        "k_fconv": 2 * D + 2 * (2 * CC + nh) + 2 * D,                       # read u; write act, SiLU', dt, TL copy of u
        "k_bwd1_ws": 2 * D + 2 * (2 * Di + GN) + 2 * (Di + GN) + 2 * (2 * Di + GN),   # dout, act z|x|C, SiLU' z|C; dpre z, dy, dpre C
        "k_bwd2_ws": 2 * (Di + GN) + 2 * Di + 2 * nh + 2 * (Di + GN) + 2 * (Di + GN + nh),   # act x|B, dy, dt, SiLU' x|B; dpre x|B, ddt
        "k_bconv_du": 2 * (CC + nh) + 2 * D,                                # dpre, ddt; du
        "k_bconv_wg": 2 * (CC + nh) + 2 * D,                                # dpre, ddt, TL copy of u
    }


class _AlgBytes(dict):
    def get(self, name, default=None):
        return (lambda D, T: _alg(D)[name] * T) if name in _alg(32) else default


KERNEL_ALG_BYTES = _AlgBytes()


def _ncu_summary():
    for name in ("r02_mixer_kernels.json", "r01_mixer_kernels.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(p):
            with open(p) as f:
                return json.load(f)
    return None


def load_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full summary."""
    d = _ncu_summary()
    e = d.get("kernels", {}).get(kernel) if d else None
    return ((e["dram_read_bytes"] + e["dram_write_bytes"]) / max(1, e.get("launches", 1))) if e else None


def load_step_traffic():
    """Sum of the ncu DRAM bytes over every launch of one fwd+bwd step (same committed capture)."""
    d = _ncu_summary()
    if not d:
        return None
    # every kernel of the chain launches once per step; a capture window may hold a second launch of one of them
    return float(sum((e["dram_read_bytes"] + e["dram_write_bytes"]) / max(1, e.get("launches", 1)) for e in d.get("kernels", {}).values()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--d-model", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="issue every step through Python / autograd instead of replaying a captured CUDA graph")
    ap.add_argument("--graph-nccl", dest="graph_nccl", action="store_true",
                    help="N > 1: capture the gradient all-reduce inside the step graph (pipelined on a forked branch) instead of "
                         "issuing it on a side stream after each replay.  Measured SLOWER at N = 2 (0.309 vs 0.294 ms / step; a "
                         "plain in-graph all-reduce after the kernels 0.312) and the process hung in destroy_process_group with the "
                         "captured collectives alive, so the side stream is the default")
    ap.add_argument("--no-model", action="store_true", help="skip the full-model legs (train / infer / gpu_baseline / configs[0])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

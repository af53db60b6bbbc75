#!/usr/bin/env python
"""bench.py - ADN-SSD mixer fwd+bwd tokens/s (BASELINE.json configs[1]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--d-model 32]

One "step" = one forward + backward of one ADN-SSD mixer (models/ADNssd.py::Mamba2 of the reference) over a
synthetic batch B=16 of 128x128 token grids (262 144 tokens per GPU), bf16 I/O, random-init weights.
  value  : tokens/s with u / dout already resident in HBM (CUDA events, max over ranks)
  e2e    : the same through the public module call (adnm_unet_b200.Mamba2) with HOST buffers: pinned-host u and
           dout copied H2D and out / du copied D2H inside the timed region, every step
  roofline : dominant kernel of the step, timed per launch with CUDA events inside the library
           (adn_prof_*), algorithmic bytes / time vs MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle port (oracle/adnssd_oracle.py, PyTorch fp32, all host threads) on a bounded sample
N > 1 (torchrun): batch-sharded data parallel, one process per GPU, weak scaling; the mixer's parameter
gradients are all-reduced over NCCL every step (the only exchange the path has).
`--impl reference` times the reference's CPU algorithm (the oracle port; the reference is pure Python and
/root/reference does not exist on the GPU box) on the host cores for the same metric.
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
    """The CPU algorithm of the reference (oracle port), fwd + autograd bwd, fp32."""
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


def time_cpu(d_model, batch, steps, warmup):
    threads = os.cpu_count() or 1
    step = oracle_step_fn(d_model, batch, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch * GRID * GRID / dt, dt, threads


CPU_BATCH = 4        # samples per CPU step (the GPU arm runs B=16 per GPU; tokens/s is per token, so the metric is comparable)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = CPU_BATCH
    steps, warmup = max(1, min(args.steps, 300)), max(1, min(args.warmup, 5))     # <= ~20 s of CPU work
    val, dt, threads = time_cpu(args.d_model, batch, steps, warmup)
    sample = f"oracle port (PyTorch fp32 CPU), B={batch} x {GRID}x{GRID} tokens per step, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ADN-SSD mixer fwd+bwd (BASELINE configs[1]): D={args.d_model}, headdim={HEADDIM}, d_state={D_STATE}, "
                               f"B={B_PER_GPU}/GPU, {GRID}x{GRID} tokens",
                   "note": f"CPU arm: each step is a bounded sample of that workload, B={batch} of the {B_PER_GPU} samples"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import adnm_unet_b200 as A
    from adnm_unet_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its banner there)
        dist.init_process_group("nccl", device_id=dev)
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
        with torch.cuda.graph(graph, pool=pool):
            out = mixer(u_static, GRID, GRID)
            out.backward(g_static)
        # the 18 parameter gradients are views of ONE flat fp32 buffer (adnm_unet_b200/mixer.py): all-reduce it in one call
        flat = torch.empty(0, dtype=torch.float32, device=dev).set_(params[0].grad.untyped_storage())
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

    def step_resident(i):
        if graphs is None:
            return step_eager(i)
        k = i % N_INPUT_SETS
        graph, out, du, flat = graphs[k]
        cur = torch.cuda.current_stream()
        if world > 1:
            cur.wait_event(ev_comm[k])        # the previous all-reduce of this graph's gradient buffer has finished
        graph.replay()
        if world > 1:
            allreduce_async(k, flat, cur)
        return out, du

    def drain_resident():
        if world > 1:
            torch.cuda.current_stream().wait_stream(s_comm)

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
            graph.replay()
            if world > 1:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
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

    for i in range(max(args.warmup, 3)):
        step_resident(i)
        step_e2e(i)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
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

    if rank == 0:
        top_ms = per[top][0] / per[top][1]
        # dominant kernel: achieved = bytes it must move per launch / its mean launch time
        top_bytes = KERNEL_ALG_BYTES.get(top, lambda D, T: None)(D, tokens)
        roof = {"bound": "hbm", "kernel": top, "achieved": (top_bytes / (top_ms * 1e-3) / 1e9) if top_bytes else None,
                "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src, "traffic": load_traffic(top),
                "kernel_ms_per_launch": top_ms, "kernel_share_of_step": per[top][0] / tot,
                "step_achieved": alg_bytes_step / (step_ms * 1e-3) / 1e9, "step_algorithmic_bytes": alg_bytes_step}
        roof["frac"] = (roof["achieved"] / hbm_peak) if roof["achieved"] else None
        roof["step_frac"] = roof["step_achieved"] / hbm_peak
        # bounded CPU sample (~10 s): the oracle port on every host core, CPU_BATCH samples per step
        cpu_val, cpu_dt, cpu_threads = time_cpu(D, CPU_BATCH, 150, 3) if world == 1 and not args.no_cpu else (None, None, None)
        line = {
            "metric": METRIC, "value": world * tokens * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ADN-SSD mixer fwd+bwd (BASELINE configs[1]): D={D}, headdim={HEADDIM}, d_state={D_STATE}, "
                                   f"B={B}/GPU, {GRID}x{GRID} tokens", "global_batch": B * world, "tokens_per_step": tokens * world,
                       "parallelism": f"dp{world}", "grad_allreduce": "NCCL AVG of one flat fp32 buffer per step on a side stream (overlaps the next step)" if world > 1 else "none (1 GPU)", "launch": "cuda_graph_replay" if args.graph else "eager", "l2": f"{N_INPUT_SETS} rotating input sets (> L2) + ~1 GB of intermediates rewritten per step"},
            "e2e": {"value": world * tokens * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": world * 2 * tokens * D * 2, "d2h_bytes_per_step": world * 2 * tokens * D * 2,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kernels,
        }
        if cpu_val is not None:
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                                    "sample": f"oracle port (PyTorch fp32, fwd + autograd bwd), B={CPU_BATCH} x {GRID}x{GRID} tokens per step, "
                                              f"150 steps after 3 warm-ups, {cpu_dt * 1e3:.1f} ms/step"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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


def load_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full summary."""
    p = os.path.join(ROOT, "profiles", "r01_mixer_kernels.json")
    if not os.path.isfile(p):
        return None
    with open(p) as f:
        d = json.load(f)
    e = d.get("kernels", {}).get(kernel)
    return (e["dram_read_bytes"] + e["dram_write_bytes"]) if e else None


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

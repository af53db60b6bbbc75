"""Per-module device time of one ADNM-UNet training step (forward and backward separately), CUDA events on module hooks.
Usage: python profiles/module_times.py [img] [batch] [depth] -> JSON.  Nested modules are reported inclusively: a parent's
time contains its children's (the table lists `depth` levels of the module tree plus every Block / mixer / FFN / WTConv2d)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import refhost  # noqa: E402


def main():
    img = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    variant = sys.argv[4] if len(sys.argv) > 4 else "dropin"
    infer = len(sys.argv) > 5 and sys.argv[5] == "infer"      # eval + no_grad forward only (BASELINE configs[3])
    dev = torch.device("cuda", 0)
    model = refhost.build_adnm_unet(img, dropin=(variant == "dropin"), seed=0).to(dev)
    loss_fn = refhost.reference_loss()
    d = torch.rand(B, 25, 1, img, img).to(dev)
    if infer:
        model.eval()
    interesting = ("Block", "Mamba2", "FeedForward", "WTConv2d", "WTConvLayer", "Attention", "OutProj", "PatchEmbed", "WTLayer",
                   "EncoderToDecoder", "RMSNorm", "StandaloneRMSNorm")
    mods = {n: m for n, m in model.named_modules() if n and (n.count(".") < depth or type(m).__name__ in interesting)}
    ev = {n: {} for n in mods}

    def rec(n, key):
        def f(*a):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev[n].setdefault(key, []).append(e)
        return f

    for n, m in mods.items():
        m.register_forward_pre_hook(rec(n, "f0"))
        m.register_forward_hook(rec(n, "f1"))
        if not infer:
            m.register_full_backward_pre_hook(rec(n, "b0"))
            m.register_full_backward_hook(rec(n, "b1"))

    def step():
        if infer:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                model(d[:, :5])
            return
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(d[:, :5])
        loss = loss_fn(out.float(), d[:, 5:])
        loss.backward()
        model.zero_grad(set_to_none=True)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    for n in ev:
        ev[n].clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    rows = []
    for n, m in mods.items():
        r = ev[n]
        f = sum(a.elapsed_time(b) for a, b in zip(r.get("f0", []), r.get("f1", [])))
        b = sum(a.elapsed_time(b) for a, b in zip(r.get("b0", []), r.get("b1", []))) if len(r.get("b0", [])) == len(r.get("b1", [])) else float("nan")
        rows.append({"module": n, "type": type(m).__name__, "fwd_ms": round(f, 3), "bwd_ms": round(b, 3)})
    rows.sort(key=lambda r: -(r["fwd_ms"] + (r["bwd_ms"] if r["bwd_ms"] == r["bwd_ms"] else 0)))
    by_type = {}
    for r in rows:
        t = by_type.setdefault(r["type"], [0.0, 0.0, 0])
        t[0] += r["fwd_ms"]; t[1] += r["bwd_ms"] if r["bwd_ms"] == r["bwd_ms"] else 0; t[2] += 1
    print(json.dumps({"img": img, "batch": B, "variant": variant, "step_ms_with_hooks": e0.elapsed_time(e1),
                      "by_type": {k: {"fwd_ms": round(v[0], 3), "bwd_ms": round(v[1], 3), "n": v[2]} for k, v in by_type.items()},
                      "rows": rows[:70]}))


if __name__ == "__main__":
    main()

"""GPU probe: parity numbers (not asserts) of the drop-ins inside the reference network + a per-module time breakdown of one
training step.  Usage: python profiles/fullmodel_probe.py [img] [batch]  -> JSON on stdout."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import refhost  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def grads(m):
    return {k: p.grad for k, p in m.named_parameters() if p.requires_grad}


def main():
    img = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    out = {"img": img, "batch": B}
    ref = refhost.build_adnm_unet(img, dropin=False).cuda()
    new = refhost.build_adnm_unet(img, dropin=True).cuda()
    new.load_state_dict(ref.state_dict(), strict=True)
    loss_fn = refhost.reference_loss()
    g = torch.Generator().manual_seed(0)
    data = torch.rand(B, 25, 1, img, img, generator=g)
    imgs, tgt = data[:, :5].cuda(), data[:, 5:].cuda()
    o32 = ref(imgs); l32 = loss_fn(o32, tgt); l32.backward()
    g32 = {k: (None if v is None else v.clone()) for k, v in grads(ref).items()}
    ref.zero_grad(set_to_none=True)
    on = new(imgs); ln = loss_fn(on, tgt); ln.backward()
    gn = grads(new)
    gmax = max(float(v.abs().max()) for v in g32.values() if v is not None)
    e = {k: float((gn[k].double() - v.double()).abs().max()) / max(float(v.abs().max()), 1e-7 * gmax) for k, v in g32.items() if v is not None}
    worst = sorted(e.items(), key=lambda kv: -kv[1])[:8]
    out["fp32"] = {"out": rel(on, o32), "loss": [float(ln), float(l32)], "n_none": sum(v is None for v in g32.values()),
                   "none_equal": {k for k, v in g32.items() if v is None} == {k for k, v in gn.items() if v is None},
                   "worst_param_grads": worst}
    new.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        orr = ref(imgs)
    lr_ = loss_fn(orr.float(), tgt); lr_.backward()
    gr = grads(ref)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        on = new(imgs)
    ln = loss_fn(on.float(), tgt); ln.backward()
    gn = grads(new)
    den = torch.sqrt(sum((v.double() ** 2).sum() for v in g32.values() if v is not None))
    d_new = torch.sqrt(sum(((gn[k].double() - v.double()) ** 2).sum() for k, v in g32.items() if v is not None))
    d_ref = torch.sqrt(sum(((gr[k].double() - v.double()) ** 2).sum() for k, v in g32.items() if v is not None))
    e_new = {k: rel(gn[k], v) for k, v in g32.items() if v is not None}
    e_ref = {k: rel(gr[k], v) for k, v in g32.items() if v is not None}
    out["bf16"] = {"out_new_vs_fp32": rel(on.float(), o32), "out_refbf16_vs_fp32": rel(orr.float(), o32),
                   "loss": [float(ln), float(lr_), float(l32)],
                   "grad_l2_new": float(d_new / den), "grad_l2_refbf16": float(d_ref / den),
                   "n_param_gt_2e-2_new": sum(v > 2e-2 for v in e_new.values()), "n_param_gt_2e-2_refbf16": sum(v > 2e-2 for v in e_ref.values()),
                   "median_new": sorted(e_new.values())[len(e_new) // 2], "median_refbf16": sorted(e_ref.values())[len(e_ref) // 2],
                   "worst_new": sorted(e_new.items(), key=lambda kv: -kv[1])[:6]}
    print(json.dumps(out, default=str))


if __name__ == "__main__":
    main()

"""Diagnostic: per-tensor errors of the drop-in model and of the eager reference against the fp64 reference (128^2, B=2),
fp32 and bf16 autocast.  Usage: python profiles/fullmodel_errs.py [block=1] -> JSON."""
import copy, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from adnm_unet_b200 import refhost
import test_fullmodel_gpu as T

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
block = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ref = refhost.build_adnm_unet(128, dropin=False, seed=0).cuda()
new = refhost.build_adnm_unet(128, dropin=True, seed=0, block=bool(block)).cuda()
new.load_state_dict(ref.state_dict(), strict=True)
loss_fn = refhost.reference_loss()
imgs, tgt = T._train_batch(128, 2)
truth = T._model_run(copy.deepcopy(ref).double(), loss_fn, imgs, tgt, dtype=torch.float64)
fl = T._grad_floor(truth)
out = {"block": block}
for name, ac in (("fp32", False), ("bf16", True)):
    rn, rr = T._model_run(new, loss_fn, imgs, tgt, autocast=ac), T._model_run(ref, loss_fn, imgs, tgt, autocast=ac)
    e_new, e_ref = T._tensor_errors(rn, truth, fl), T._tensor_errors(rr, truth, fl)
    tol = 1e-4 if not ac else 2e-2
    worse = {k: (e_new[k], e_ref[k]) for k in e_new if e_new[k] > 3.0 * e_ref[k] + tol}
    out[name] = {"n_new_gt_tol": sum(v > tol for v in e_new.values()), "n_ref_gt_tol": sum(v > tol for v in e_ref.values()),
                 "out": (e_new["out"], e_ref["out"]), "loss": (e_new["loss"], e_ref["loss"]),
                 "worse_3x": dict(sorted(worse.items(), key=lambda kv: -kv[1][0])),
                 "top_new": sorted(((k, v, e_ref[k]) for k, v in e_new.items()), key=lambda r: -r[1])[:40]}
print(json.dumps(out))

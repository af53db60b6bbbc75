"""Is the error of du (and of the other outputs) of the bf16 paths a zero-mean rounding error or biased?  Scalar gates of the
host network (shift / scale / beta) sum du over every token and channel: a bias of 1e-4 of the typical magnitude outweighs
zero-mean bf16 rounding there.  Compares ours and the eager reference (both bf16 autocast) with the fp64 reference."""
import copy, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import refhost
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ns = refhost.load_reference()
from adnm_unet_b200.mixer import Mamba2 as New
out = []
for D, g, B in ((128, 16, 32), (256, 8, 32), (32, 128, 2)):
    torch.manual_seed(0)
    ref = ns.ref_Mamba2(d_model=D, headdim=4, d_state=16).cuda()
    new = New(d_model=D, headdim=4, d_state=16).cuda()
    new.load_state_dict(ref.state_dict())
    ref64 = copy.deepcopy(ref).double()
    u = torch.randn(B, g * g, D, device="cuda").bfloat16().float()
    dout = torch.randn(B, g * g, D, device="cuda").bfloat16().float()
    def run(m, dt, ac):
        x = u.to(dt).clone().requires_grad_(True)
        if ac:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = m(x, g, g)
            y.float().backward(dout)
        else:
            y = m(x, g, g); y.backward(dout.to(dt))
        return y.detach().double(), x.grad.double()
    yt, dt_ = run(ref64, torch.float64, False)
    rows = {}
    for name, m in (("ours", new), ("ref_bf16", ref)):
        y, d = run(m, torch.float32, True)
        e = d - dt_
        rows[name] = {"du_err_mean": float(e.mean()), "du_err_std": float(e.std()), "du_abs_mean": float(dt_.abs().mean()),
                      "sum_du": float(d.sum()), "sum_du_true": float(dt_.sum()), "bias_over_noise": float(e.mean() / e.std() * e.numel() ** 0.5),
                      "out_err_mean": float((y - yt).mean()), "out_err_std": float((y - yt).std())}
    out.append({"D": D, "grid": g, "B": B, **rows})
    print(json.dumps(out[-1]), flush=True)

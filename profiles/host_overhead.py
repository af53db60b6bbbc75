"""Diagnostic: host-side (Python / ctypes / autograd) cost of one mixer fwd+bwd step vs its GPU time."""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A
torch.manual_seed(0)
m = A.Mamba2(d_model=32, headdim=4, d_state=16).cuda()
params = [p for n, p in m.named_parameters() if n not in ("scale", "shift", "alpha2")]
B = int(os.environ.get("ADN_B", "16"))
u = torch.randn(B, 128 * 128, 32, device="cuda", dtype=torch.bfloat16, requires_grad=True)
go = torch.randn_like(u)
def step():
    u.grad = None
    for p in params: p.grad = None
    out = m(u, 128, 128); out.backward(go)
for _ in range(10): step()
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"B={B}: host issue time {1e6*(t1-t0)/N:.0f} us/step, wall incl. GPU drain {1e6*(t2-t0)/N:.0f} us/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

# the same loop through adnm_unet_b200.graphed_mixer (forward and backward replayed as CUDA graphs inside autograd)
f = A.graphed_mixer(m, u.detach(), 128, 128)
def gstep():
    u.grad = None
    for p in params: p.grad = None
    f(u).backward(go)
for _ in range(10): gstep()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N): gstep()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"graphed_mixer B={B}: host issue time {1e6*(t1-t0)/N:.0f} us/step, wall incl. GPU drain {1e6*(t2-t0)/N:.0f} us/step")

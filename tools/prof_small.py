"""Short single-kernel driver for `ncu --set full` captures at the bench sizes: MSM 2^24 (uniform) and NTT 2^26."""
import importlib, sys
import torch
sys.path.insert(0, ".")
zk = importlib.import_module("scroll-prover_b200")
ctx = zk.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
def rnd(n):
    t = torch.randint(-(2**63), 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g); t[:, 3] &= 0x0FFFFFFFFFFFFFFF; return t
n = 1 << 24
pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
ctx.g1_generator_mul_batch(rnd(n), out=pts)
srs = ctx.srs_register(pts)
sc = rnd(n)
torch.cuda.synchronize()
for _ in range(2):
    srs.msm(sc)
a = rnd(1 << 26)
w = zk.fr_from_int(pow(zk._ROOT_OF_UNITY, 1 << 2, zk.R_MOD))
torch.cuda.synchronize()
for _ in range(2):
    ctx.best_fft(a, w, 26)
ctx.synchronize()
print("done")

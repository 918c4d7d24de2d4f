"""GPU-box experiment: error of each fp32 intermediate of the engine against an fp64 torch recomputation."""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
import torch
from oracle import dsvgp_oracle as O
from dsvgp_b200 import engine, ops
F32, F64 = torch.float32, torch.float64
def rel(a, b):
    a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1)
    return float((a - b).abs().max() / b.abs().max())
def relf(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm())
variant, n, d, M, p = "dsvgp", 1024, 10, 1024, 2
if len(sys.argv) > 1: n, d, M, p = map(int, sys.argv[1:5])
P, x, Vx, y, nd = O.make_problem(n, d, M, p, F32, seed=1, variant=variant, N=100 * n)
Pg = types.SimpleNamespace(**{k: v.cuda().contiguous() for k, v in P.tensors().items()})
E = engine.ENGINE
elbo, g, mu, var = E.elbo_step(Pg, x.cuda(), Vx.cuda(), y.cuda(), nd, p, p, want_grads=True)
ws = E.workspace(x.cuda().device, F32, n, d, M, p, p); f = E.factor(x.cuda().device, F32, d, M, p)
Mq, nq = ws.Mq, ws.nq
# fp64 recomputation on the GPU with torch
dev = "cuda"
P64 = types.SimpleNamespace(**{k: v.double().to(dev) for k, v in P.tensors().items()})
ell = torch.nn.functional.softplus(P64.raw_ell).reshape(()); osc = torch.nn.functional.softplus(P64.raw_os).reshape(())
s2 = torch.nn.functional.softplus(P64.raw_noise).reshape(()) + 1e-4
import oracle.dsvgp_oracle as OO
Kzz = osc * OO.kernel_closed_form(P64.Z, P64.Z, P64.Vz, P64.Vz, ell) + 1e-3 * torch.eye(Mq, dtype=F64, device=dev)
Kzx = osc * OO.kernel_closed_form(P64.Z, x.double().to(dev), P64.Vz, Vx.double().to(dev), ell)
L = torch.linalg.cholesky(Kzz); W = torch.linalg.inv(L)
print("cond(Kzz) %.2e  |W|max %.2e" % (float(torch.linalg.cond(Kzz)), float(W.abs().max())))
A = W @ Kzx
Ls = P64.Ls_raw.tril(); Em = Ls - torch.eye(Mq, dtype=F64, device=dev)
Bp = Em.T @ A; B = A + Bp; C = Em @ B + Bp
mean = A.T @ P64.m + P64.c; varr = torch.cat([osc.reshape(1), (osc / ell ** 2).expand(p)]).repeat(n) + 1e-4 + (A * C).sum(0) + s2
w = 1.0 / nq
gmu = w * (y.double().to(dev) - mean) / s2; gvar = torch.full_like(gmu, float(-0.5 * w / s2))
t = A @ gmu
dA = P64.m[:, None] * gmu[None] + 2 * C * gvar[None]
dKzx = W.T @ dA
G = (A * gvar[None]) @ A.T
print("Kzz  (after chol: L)  L rel %.1e  W rel %.1e  Wt(fp32) rel %.1e" % (rel(f.L[:Mq, :Mq].tril(), L), rel(f.W[:Mq, :Mq].tril(), W), rel(f.Wt, W)))
# note: ws.Kzx now holds dKzx, ws.B holds Ag, ws.C holds dA
print("A    max-rel %.1e  fro-rel %.1e   amplification |W||K|/|A| (fro) %.1f" % (rel(ws.A[:, :nq], A), relf(ws.A[:, :nq], A), float((W.abs() @ Kzx.abs()).norm() / A.norm())))
print("Bp   max-rel %.1e  fro-rel %.1e" % (rel(ws.Bp[:, :nq], Bp), relf(ws.Bp[:, :nq], Bp)))
print("mean max-rel %.1e   var max-rel %.1e" % (rel(mu, mean), rel(var, varr)))
print("t    max-rel %.1e  fro-rel %.1e" % (rel(ws.t, t), relf(ws.t, t)))
print("dA   max-rel %.1e  fro-rel %.1e" % (rel(ws.C[:, :nq], dA), relf(ws.C[:, :nq], dA)))
print("dKzx max-rel %.1e  fro-rel %.1e" % (rel(ws.Kzx[:, :nq], dKzx), relf(ws.Kzx[:, :nq], dKzx)))
print("G    max-rel %.1e  fro-rel %.1e" % (rel(ws.G, G), relf(ws.G, G)))
X = torch.outer(P64.m, t) + 2 * (Em @ (G + Em.T @ G) + Em.T @ G)
print("X    max-rel %.1e  fro-rel %.1e" % (rel(ws.X, X), relf(ws.X, X)))
dL = -(W.T @ X).tril(); Y = (L.T @ dL).tril(); Phi = Y.clone(); Phi.diagonal().mul_(0.5)
S = W.T @ (0.5 * (Phi + Phi.T)) @ W
print("dKzz max-rel %.1e  fro-rel %.1e" % (rel(ws.S, S), relf(ws.S, S)))
# the same with the exact t / exact X fed through the fp64 tail (how much of dKzz error is input error?)
X32 = ws.X.double(); dL2 = -(W.T @ X32).tril(); Y2 = (L.T @ dL2).tril(); Phi2 = Y2.clone(); Phi2.diagonal().mul_(0.5)
S2 = W.T @ (0.5 * (Phi2 + Phi2.T)) @ W
print("dKzz from engine X through exact tail: max-rel %.1e" % rel(S2, S))
# Z gradient split by path (fp64 autograd)
Zr = P64.Z.clone().requires_grad_(True)
Kzx_r = osc * OO.kernel_closed_form(Zr, x.double().to(dev), P64.Vz, Vx.double().to(dev), ell)
gz_x, = torch.autograd.grad((Kzx_r * dKzx).sum(), Zr)
Zr2 = P64.Z.clone().requires_grad_(True)
Kzz_r = osc * OO.kernel_closed_form(Zr2, Zr2, P64.Vz, P64.Vz, ell)
gz_z, = torch.autograd.grad((Kzz_r * S).sum(), Zr2)
print("dZ via Kzx |.|max %.3e   via Kzz %.3e   total %.3e   engine total err %.1e" % (float(gz_x.abs().max()), float(gz_z.abs().max()), float((gz_x + gz_z).abs().max()), rel(g["Z"], gz_x + gz_z)))

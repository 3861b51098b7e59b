import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200")):
    sys.path.insert(0, p)
import torch
from sqlx import sql as S
torch.manual_seed(0)
B, h, w, Q, D = 1, 8, 16, 12, 16
x = torch.randn(B, 32, h, w, device="cuda"); q = 0.4 * torch.randn(B, Q, 32, device="cuda")
Wp = 0.3 * torch.randn(D, Q, device="cuda"); bp = 0.1 * torch.randn(D, device="cuda")
cen = torch.sort(torch.rand(B, D, device="cuda") * 80, dim=1).values
g = torch.randn(B, 1, h, w, device="cuda")
summ, m, l, _ = S.summary_fwd(x, q)
ds = torch.randn_like(summ)
print("calling dx", flush=True)
dx, dq = S.bwd_dx(x, q, Wp, bp, cen, g, summ, m, l, ds)
torch.cuda.synchronize()
print("done", float(dx.abs().max()), float(dq.abs().max()))

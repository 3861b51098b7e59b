"""Per-op device timing of the SQL kernels (CUDA events, 20 iterations): python tools/time_sql.py [B h w Q D]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200")):
    sys.path.insert(0, p)
import torch
from sqlx import sql as S

B, h, w, Q, D = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (12, 96, 320, 64, 64)
torch.manual_seed(0)
x = torch.randn(B, 32, h, w, device="cuda"); q = 0.4 * torch.randn(B, Q, 32, device="cuda")
Wp = 0.3 * torch.randn(D, Q, device="cuda"); bp = 0.1 * torch.randn(D, device="cuda")
cen = torch.sort(torch.rand(B, D, device="cuda") * 80, dim=1).values.contiguous()
g = torch.randn(B, 1, h, w, device="cuda")
summ, m, l, _ = S.summary_fwd(x, q)
ds = torch.randn_like(summ)
Mx = torch.matmul(Wp, q)
dx0 = torch.zeros_like(x)

def t(name, fn, it=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    print("%-28s %8.1f us" % (name, a.elapsed_time(b) / it * 1e3))

print("B=%d n=%d Q=%d D=%d" % (B, h * w, Q, D))
t("summary_fwd", lambda: S.summary_fwd(x, q))
t("pred_mix_fwd v1", lambda: S.pred_mix_fwd(x, Mx, bp, cen, version=1))
t("pred_mix_fwd", lambda: S.pred_mix_fwd(x, Mx, bp, cen))
pred, stats = S.pred_mix_fwd(x, Mx, bp, cen)
t("bwd_pred_mix v1", lambda: S.bwd_pred_mix(x, Mx, bp, cen, g))
t("bwd_pred_mix", lambda: S.bwd_pred_mix(x, Mx, bp, cen, g, pred, stats))
t("bwd_summary v1 (write)", lambda: S.bwd_summary(x, q, summ, m, l, ds, version=1))
t("bwd_summary (write)", lambda: S.bwd_summary(x, q, summ, m, l, ds))
t("bwd_summary (accumulate)", lambda: S.bwd_summary(x, q, summ, m, l, ds, d_x=dx0))
if Q <= 64 and D <= 64:
    t("old pred_fwd", lambda: S.pred_fwd(x, q, Wp, bp, cen))
    t("old bwd_reduce", lambda: S.bwd_reduce(x, q, Wp, bp, cen, g))
    t("old bwd_dx", lambda: S.bwd_dx(x, q, Wp, bp, cen, g, summ, m, l, ds))

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200")):
    sys.path.insert(0, p)
import torch
from sqlx import sql as S
torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)
B, h, w, Q = 1, 8, 16, 32
n = h * w
x = (torch.arange(32).float()[:, None] * 1000 + torch.arange(n).float()[None, :]).reshape(1, 32, h, w)  # e*1000 + p
q = (torch.arange(Q).float()[:, None] * 100 + torch.arange(32).float()[None, :]).reshape(1, Q, 32)        # q*100 + e
en = S.energy_tc(x.cuda(), q.cuda()).cpu().reshape(-1)
mode = os.environ.get("SQLX_TC_DEBUG")
print("mode", mode)
if mode == "1":
    t = en[:4096].reshape(4, 32, 32)
    print("block0 rows 0..9 (raw smem order):\n", t[0, :10])
    print("block1 row 0:", t[1, 0])
elif mode == "2":
    t = en[:Q * 32].reshape(Q, 32)
    print(t[:10])

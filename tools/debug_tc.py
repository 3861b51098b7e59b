import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200")):
    sys.path.insert(0, p)
import torch
from sqlx import sql as S
torch.set_printoptions(linewidth=200, precision=2, sci_mode=False)
B, h, w, Q = 1, 8, 16, 16   # one tile of 128 px
n = h * w
def run(x, q, name):
    en = S.energy_tc(x.cuda(), q.cuda()).cpu().reshape(Q, n)
    ref = torch.einsum("ep,qe->qp", x.reshape(32, n), q.reshape(Q, 32))
    print("==", name, "max err", float((en - ref).abs().max()))
    print("got  rows q0..3, px 0..15:\n", en[:4, :16])
    print("got  q0, px 32..40:", en[0, 32:40], " q0 px 64..68", en[0, 64:68], "q0 px 96..100", en[0, 96:100])
    print("got  q 0..15 at px 0:", en[:, 0])
    print("ref  rows q0..3, px 0..15:\n", ref[:4, :16])
x = torch.ones(1, 32, h, w); q = torch.ones(1, Q, 32)
run(x, q, "all ones (expect 32)")
for e0 in (0, 1, 4, 9, 31):
    x = torch.ones(1, 32, h, w); q = torch.zeros(1, Q, 32); q[0, :, e0] = torch.arange(1, Q + 1).float()
    run(x, q, "K[q,e0=%d]=q+1 (expect y=q+1)" % e0)
for e0 in (0, 1, 9):
    x = torch.zeros(1, 32, h, w); x[0, e0] = torch.arange(1, n + 1).float().reshape(h, w); q = torch.ones(1, Q, 32)
    run(x, q, "x[e0=%d,p]=p+1 (expect y=p+1)" % e0)

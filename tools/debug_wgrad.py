import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "omni-pq_b200"))
import _pn2 as K
dev = "cuda"
torch.manual_seed(0)
for rows, n, k in [(64, 128, 128), (256, 128, 128), (4096, 256, 132)]:
    a = torch.randn(rows, k, device=dev)
    dyv = torch.randn(rows, n, device=dev)
    y = torch.zeros(rows, n, device=dev)
    one, zero = torch.ones(n, device=dev), torch.zeros(n, device=dev)
    dy = K.rows_dy(y, dyv, rows, n, n, one, zero, zero)      # dy = 1*dz + 0 + 0*y
    src = K.rows_plain(a, rows, k, k)
    dw = K.mlp_wgrad(dy, src, n, k, 0, 0, dev)
    want = dyv.t().double() @ a.double()
    err = (dw.double() - want).abs().max() / want.abs().max()
    print(rows, n, k, "rel err", float(err), " ours[0,:4]", dw[0, :4].tolist(), " want[0,:4]", want[0, :4].float().tolist())
    nz = (dw != 0).float().mean().item()
    print("   nonzero fraction", nz, " ours absmax", float(dw.abs().max()), " want absmax", float(want.abs().max()))

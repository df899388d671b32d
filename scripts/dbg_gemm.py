import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gnnb200
from gnnb200 import ops
from gnnb200.nn import Linear
DEV='cuda'
g = torch.Generator().manual_seed(4)
ref = torch.nn.Linear(256, 512)
lin = Linear(256, 512)
lin.precision = 'tf32'
lin.load_state_dict(ref.state_dict())
lin = lin.to(DEV)
x = torch.randn(9000, 256, generator=g)
go = torch.randn(9000, 512, generator=g)
xg = x.to(DEV).requires_grad_(True)
y = lin(xg)
gg = go.to(DEV)
print('g', gg.shape, gg.stride(), gg.data_ptr() % 16, 'w', lin.weight.data_ptr() % 16, 'x', xg.data_ptr() % 16)
import traceback
try:
    y.backward(gg)
    print('backward ok')
except Exception:
    traceback.print_exc()
    for prec in (1, 2):
        try:
            out = ops.gemm(gg, False, lin.weight.detach(), False, None, False, prec); print(prec, 'direct ok')
        except Exception as e:
            print(prec, 'direct FAIL', e)

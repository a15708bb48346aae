"""One unmasked dictionary-learning epoch at C4 scale for an ncu launch list (complex128, f=2048, k=512, minibatch 8192)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from decomp_b200 import dictionary_learning as dl
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(0)
def crandn(*s): return torch.complex(torch.randn(s, dtype=torch.float64, device=dev, generator=g), torch.randn(s, dtype=torch.float64, device=dev, generator=g))
n, mb, f, k = 4 * 8192, 8192, 2048, 512
D = crandn(k, f); D = D / D.abs().pow(2).sum(-1, keepdim=True).sqrt()
x = crandn(n, k) * torch.rand((n, k), dtype=torch.float64, device=dev, generator=g).round()
y = x @ D + 0.1 * crandn(n, f)
masked = len(sys.argv) > 1 and sys.argv[1] == 'masked'
mask = (torch.rand((n, f), dtype=torch.float64, device=dev, generator=g) > 0.1).double() if masked else None
it, Dn, xn = dl.solve(y, D, 0.1, tol=0.0, minibatch=mb, maxiter=2, lasso_method='fista', lasso_iter=10, mask=mask, random_seed=0)
torch.cuda.synchronize()
print('ok', it)

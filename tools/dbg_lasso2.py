import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import golden_cases as gc
from decomp_b200 import ops
from decomp_b200._device import to_device2d, empty2d
from decomp_b200._lib import rview
case = gc.lasso_cases()['fix_ista_nomask']
y, A, alpha = case['y'], case['A'], case['alpha']
B, f = y.shape; k = A.shape[0]
def rel(a, b): return np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-300)
yd, Ad = to_device2d(y), to_device2d(A)
s = ops.row_norms(Ad, False)
s_ref = np.sqrt((A*A).sum(-1)); print('s', rel(s.cpu().numpy(), s_ref))
An = empty2d(k, f); ops.scale(Ad, An, rowscale=s, invert_row=True)
An_ref = A / s_ref[:, None]; print('An', rel(An.cpu().numpy(), An_ref))
alpha_vec, tol_vec = ops.lasso_vectors(s, alpha, 0.0, mult=float(f))
print('alpha', rel(alpha_vec.cpu().numpy(), alpha / s_ref * f))
G = empty2d(k, k); ops.gemm_nt(An, An, ops.epilogue(ops.EPI_STORE, G))
G_ref = An_ref.dot(An_ref.T); print('G', rel(G.cpu().numpy(), G_ref))
step = torch.empty(1, dtype=torch.float64, device='cuda'); ops.gershgorin_step(G, False, step)
step_ref = 1.0/np.max(np.sum(np.abs(G_ref), axis=0)); print('step', step.item(), step_ref)
yAh = empty2d(B, k); ops.gemm_nt(yd, An, ops.epilogue(ops.EPI_STORE, yAh))
yAh_ref = y.dot(An_ref.T); print('yAh', rel(yAh.cpu().numpy(), yAh_ref))
G_rhs = ops.make_rhs(G, False, False); print('G_rhs', rel(G_rhs.cpu().numpy(), G_ref.T))
X = empty2d(B, k); X.zero_()
W = [empty2d(B, k), empty2d(B, k)]; W[0].copy_(X)
x_ref = np.zeros((B, k)); 
for i in range(5):
    epi = ops.epilogue(ops.EPI_PROX, X, out2=W[(i+1)%2], x=W[i%2], other=yAh, prev=X, colvec=alpha_vec, colvec2=tol_vec,
                       step=step, momentum=0.0, shrink=ops.SHRINK_REAL)
    ops.gemm_nt(W[i%2], G_rhs, epi)
    z = x_ref + step_ref*(yAh_ref - x_ref.dot(G_ref))
    x_ref = np.maximum(np.abs(z) - step_ref*(alpha/s_ref*f), 0)*np.sign(z)
    torch.cuda.synchronize()
    print('iter', i, 'X', rel(X.cpu().numpy(), x_ref), 'W', rel(W[(i+1)%2].cpu().numpy(), x_ref), 'pitches', X.stride(0), G_rhs.stride(0))

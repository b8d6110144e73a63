"""Driver for ncu: cht_matrix on 2^20 angles, N = 16 (the row-block kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray
pol = torch.from_numpy(np.linspace(0, 2 * np.pi, 1 << 20, endpoint=False)).cuda()
for _ in range(3):
    micarray.modal.angular.cht_matrix(16, pol)
torch.cuda.synchronize()
print("ok")

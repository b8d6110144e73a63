import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray
P = 1 << 17
i = np.arange(P)
colat = np.arccos(1 - (2 * i + 1) / P)
azi = np.mod(np.pi * (1 + 5 ** 0.5) * i, 2 * np.pi)
d = [torch.from_numpy(a).cuda() for a in (azi, colat)]
for _ in range(3):
    Y = micarray.modal.angular.sht_matrix(32, d[0], d[1])
torch.cuda.synchronize()
print("ok", tuple(Y.shape))

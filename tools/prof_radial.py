import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray
R = micarray.modal.radial
F = 1 << 22
k = torch.from_numpy(np.geomspace(0.01, 50, F) / 0.1).cuda()
for _ in range(3):
    R.radial_filters(16, k, 0.1, "card", 100.0, "Tikh", check="defer")
torch.cuda.synchronize()
R.check_deferred()
print("ok")

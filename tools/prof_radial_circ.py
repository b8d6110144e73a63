"""Driver for ncu: the fused circular radial filter (circular_pw, rigid, Tikhonov) on 2^20 wavenumbers."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray
R = micarray.modal.radial
F = 1 << 20
k = torch.from_numpy(np.geomspace(0.01, 50, F) / 0.1).cuda()
for _ in range(3):
    R.radial_filters(16, k, 0.1, "rigid", 100.0, "Tikh", check="defer", kind="circular_pw")
torch.cuda.synchronize()
R.check_deferred()
print("ok")

"""Where does a headline step go?  CUDA events between the phases of eager steps (back to back, no sync inside)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfa_numpy_b200 as micarray
from sfa_numpy_b200.pipeline import ModalBeamformer
from oracle import sfa_oracle as O
N, F, L, T = 8, 1024, 2562, 938
azi, colat, w = O.grid_equal_angle(3); azl, col, _ = O.fibonacci_grid(L); k = O.wavenumbers(F)
bf = ModalBeamformer(N, (azi, colat, w), (azl, col), k, 0.042, "rigid", 100.0, "Tikh", T, precision="c64_fast")
bf.X.copy_(torch.from_numpy(O.synth_spectra(F, 64, T, seed=0)))
S = micarray.modal.steering
n = 12
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n)]
for _ in range(3):
    bf.run_eager()
torch.cuda.synchronize()
for i in range(n):
    ev[i][0].record()
    Y_p, Y_q, dn = bf._stage12(check="none")
    ev[i][1].record()
    p = bf.plan
    p.Y_look, p.Y_mic, p.dn = Y_q, Y_p, dn
    p.prepare()
    ev[i][2].record()
    p(bf.X, out=bf.out)
    ev[i][3].record()
torch.cuda.synchronize()
a = sum(ev[i][0].elapsed_time(ev[i][1]) for i in range(2, n)) / (n - 2)
b = sum(ev[i][1].elapsed_time(ev[i][2]) for i in range(2, n)) / (n - 2)
c = sum(ev[i][2].elapsed_time(ev[i][3]) for i in range(2, n)) / (n - 2)
tot = ev[2][0].elapsed_time(ev[n - 1][3]) / (n - 2)
print("eager: stage 1-2 %.3f ms  prepare %.3f ms  stage 3 call %.3f ms  | step (wall of the loop) %.3f ms" % (a, b, c, tot))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("graph replay", bf.run), ("stage-3 call only", lambda: bf.plan(bf.X, out=bf.out))):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-20s %.3f ms per call" % (name, e0.elapsed_time(e1) / 10))

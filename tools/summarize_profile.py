"""Write a text summary (key metrics + hottest SASS lines) of an .ncu-rep into profiles/."""
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
m = subprocess.run([sys.executable, os.path.join(here, "ncu_metrics.py"), rep], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tmp = out + ".src.csv"
open(tmp, "w").write(src)
h = subprocess.run([sys.executable, os.path.join(here, "ncu_hot.py"), tmp, "30"], capture_output=True, text=True).stdout
os.remove(tmp)
note = sys.argv[3] if len(sys.argv) > 3 else ""
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on  (%s)\n# %s\n\n" % (os.path.basename(rep), note))
    f.write(m + "\n# hottest SASS instructions (warp-state samples, top stall reasons)\n" + h)
print("wrote", out)

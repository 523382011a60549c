"""Print the worst cells of one adversarial case: GPU vs oracle, field by field."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import oracle as orc
import test_gpu_adversarial as adv
from helpers import run_gpu, run_oracle, CHECKED
geom, kind, teq, sweeps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
cfg = adv.adversarial(adv.base_config(geom, find_Teq=bool(teq)), kind)
g = run_gpu(cfg, max_sweeps=sweeps); o = run_oracle(orc, cfg, max_sweeps=sweeps)
print(geom, kind, teq, "iters", g["iters"], o["iters"])
for n in CHECKED:
    a, b = np.asarray(g[n]), np.asarray(o[n])
    with np.errstate(all="ignore"):
        r = np.where(a == b, 0.0, np.abs(a - b) / np.abs(b))
    i = int(np.nanargmax(r))
    print("%-9s worst cell %3d rel %.2e gpu %.17g orc %.17g | nH %.3e T %.6g xH1 %.3e xHe1 %.3e xHe2 %.3e xHe3 %.3e H1i %.3e He2i %.3e"
          % (n, i, r[i], a[i], b[i], cfg["nH"][i], o["T"][i], o["xH1"][i], o["xHe1"][i], o["xHe2"][i], o["xHe3"][i], o["H1i_src"][i], o["He2i_src"][i]))

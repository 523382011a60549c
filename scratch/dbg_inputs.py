import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
Nl, Nnu = 96, 128
idx = np.array([0, 17, 255, 256 * 7 + 3, 256 * 31 + 255, 4097, 8191])
host = Plan(0, idx.size, Nl, Nnu, 8)
a = configs.c5_batch_arrays(idx, Nl, Nnu)
host.set_batch(0, a["Edges"], a["nH"], a["nHe"], a["T"], a["E"], a["shape"], a["spec_index"], a["z"])
host.upload()
dev = Plan(0, idx.size, Nl, Nnu, 8)
nH0, R_kpc, z, _ = configs.c5_params(idx)
E = dev.set_sphere_family(nH0, R_kpc * configs.KPC_CM, z)
dev.upload()
print("E maxrel", np.max(np.abs(E / a["E"][0] - 1)))
for b in range(idx.size):
    h, d = host.fetch_inputs(b), dev.fetch_inputs(b)
    with np.errstate(all="ignore"):
        r = np.where(d["spec"] == h["spec"], 0.0, np.abs(d["spec"] - h["spec"]) / np.abs(h["spec"]))
    k, c = np.unravel_index(np.argmax(r), r.shape)
    print("b", b, "z", z[b], "max rel", r.max(), "at k", k, "col", c, "E", E[k], "host", h["spec"][k, c], "dev", d["spec"][k, c],
          "cols maxima", [float(r[:, q].max()) for q in range(9)])

"""Time the full drop-in path: reference PlumedMain (oracle/_ref) + LOAD plugin, driven through plumed_cmd like an
MD engine, vs the same input on the built-in CPU action.  python scripts/plugin_e2e.py [natoms] [steps] [cpu_atoms]"""
import os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
os.environ.setdefault("PLUMED_NUM_THREADS", str(os.cpu_count() or 1))
os.environ["PLUMED_IGNORE_NL_MEMORY_ERROR"] = "1"
import numpy as np
from helpers import water_box
from oracle import refplumed as RP
from plumed2_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
PLUGIN = os.path.join(os.path.dirname(capi.LIB_PATH), "libb200coord_plumed.so")
pos, box = water_box(n, 100.0)
rng = np.random.default_rng(0)
frames = [pos + 0.002 * rng.standard_normal(pos.shape) for _ in range(4)]
body = "GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} %s NL_CUTOFF=1.0 NL_STRIDE=10"
for pin in ("0", "1"):
    os.environ["B200COORD_PIN_HOST"] = pin
    p = RP.Plumed(n, ["LOAD FILE=" + PLUGIN, "c: COORDINATION " + body % (n, "NLIST"), "RESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1"])
    for s in range(10):
        p.calc(s, frames[s % 4], box)
    t0 = time.perf_counter()
    for s in range(10, 10 + steps):
        r = p.calc(s, frames[s % 4], box)
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps({"arm": "plugin through plumed_cmd", "pin_host": pin, "natoms": n, "ms_per_step": 1e3 * dt, "bias": r["bias"]}))
    p.close()

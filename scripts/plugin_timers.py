"""PLUMED's own per-phase timers (DEBUG DETAILED_TIMERS) for the plugin at N atoms: where the host time goes"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
os.environ.setdefault("PLUMED_NUM_THREADS", str(os.cpu_count() or 1))
os.environ["B200COORD_PIN_HOST"] = "1"
import numpy as np
from helpers import water_box
from oracle import refplumed as RP
from plumed2_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
PLUGIN = os.path.join(os.path.dirname(capi.LIB_PATH), "libb200coord_plumed.so")
pos, box = water_box(n, 100.0)
frames = [pos + 0.002 * np.random.default_rng(i).standard_normal(pos.shape) for i in range(2)]
p = RP.Plumed(n, ["DEBUG DETAILED_TIMERS", "LOAD FILE=" + PLUGIN,
                  "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10" % n,
                  "RESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1"], log="/tmp/plumed_timers.log")
for s in range(30):
    p.calc(s, frames[s % 2], box)
p.close()
txt = open("/tmp/plumed_timers.log").read()
i = txt.find("Cycles        Total")
print(txt[i - 60:])

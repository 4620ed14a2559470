"""Run the BASELINE.json configurations at (or near) full size on one GPU and print timings + sanity checks.
   python scripts/configs_at_scale.py [c1 c2 c3 c4 c5]"""
import os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import numpy as np
import plumed2_b200 as P
from helpers import water_box, rel_err

def timed(c, frames, box, nsteps, label, pairs_from="nl_size"):
    t_first = None
    for s in range(nsteps):
        c.prepare(s)
        t0 = time.perf_counter(); c.calculate(frames[s % len(frames)], box); dt = time.perf_counter() - t0
        if s == 0: t_first = dt
    t0 = time.perf_counter()
    for s in range(nsteps, 2 * nsteps):
        c.prepare(s); c.calculate(frames[s % len(frames)], box)
    dt = (time.perf_counter() - t0) / nsteps
    st = c.stats()
    out = dict(config=label, ms_per_step_e2e_pageable=1e3 * dt, first_step_ms=1e3 * t_first, nl_size=st["nl_size"],
               pair_evals_per_s=st["nl_size"] / dt, sweep_ms=st["last_sweep_ms"], build_ms=st["last_build_ms"],
               value=c.value, ncells=st["ncells"], sum_deriv=float(np.abs(c.derivatives.sum(axis=0)).max()),
               virial_asym=float(np.abs(c.virial - c.virial.T).max()))
    print(json.dumps(out)); sys.stdout.flush()
    return out

def frames_of(pos, k=4, amp=0.002, seed=1):
    rng = np.random.default_rng(seed)
    return [pos + amp * rng.standard_normal(pos.shape) for _ in range(k)]

which = sys.argv[1:] or ["c1", "c2", "c3", "c5"]
if "c1" in which:  # 1k atoms, no NL
    pos, box = water_box(1000, 100.0)
    c = P.Coordination.from_input("c: COORDINATION GROUPA=1-1000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12}")
    timed(c, frames_of(pos), box, 50, "configs[0] 1k atoms, no NL"); c.close()
if "c2" in which:  # 100k atoms NLIST
    pos, box = water_box(100000, 100.0)
    c = P.Coordination.from_input("c: COORDINATION GROUPA=1-100000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10")
    timed(c, frames_of(pos), box, 50, "configs[1] 100k atoms NLIST 1.0/10"); c.close()
if "c3" in which:  # 10k solute vs 1M solvent, triclinic, EXP, rebuild every step
    na, nb = 10000, 1000000
    pos, box = water_box(na + nb, 100.0, seed=3, triclinic=True)
    for mode in ("NLIST", "NLISTCELLS"):
        c = P.Coordination.from_input("c: COORDINATION GROUPA=1-%d GROUPB=%d-%d SWITCH={EXP R_0=0.2 D_MAX=0.9} %s NL_CUTOFF=1.0 NL_STRIDE=1" % (na, na + 1, na + nb, mode))
        timed(c, frames_of(pos), box, 10, "configs[2] 10k x 1M triclinic EXP %s stride 1" % mode); c.close()
if "c5" in which:  # 4M atoms at 33.4/nm^3, single GPU
    pos, box = water_box(4000000, 33.4, seed=5)
    c = P.Coordination.from_input("c: COORDINATION GROUPA=1-4000000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10")
    timed(c, frames_of(pos, k=2), box, 10, "configs[4] 4M atoms 33.4/nm^3 NLIST 1.0/10 (1 GPU)"); c.close()

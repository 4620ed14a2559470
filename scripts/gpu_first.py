import sys, time, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,R+'/tests')
import plumed2_b200 as P
from oracle import oracle as O
from helpers import water_box, oracle_eval, rel_err
def run(line, n, triclinic=False, density=100.0, nb=0):
    pos, box = water_box(n+nb, density, triclinic=triclinic)
    c = P.Coordination.from_input(line)
    c.prepare(0)
    t=time.time(); v = c.calculate(pos, box); t1=time.time()-t
    t=time.time(); v = c.calculate(pos, box); t2=time.time()-t
    st = c.stats()
    # oracle
    import re
    lab, act, kv, flags = P.coordination.split_input_line(line)
    sw = O.make_switch(kv["SWITCH"]) if "SWITCH" in kv else O.make_switch(nn=int(kv.get("NN",6)),mm=int(kv.get("MM",0)),r0=float(kv["R_0"]),d0=float(kv.get("D_0",0)))
    mode = "classic" if "NLIST" in flags else ("cells" if "NLISTCELLS" in flags else "none")
    style = "single" if nb==0 else ("pair" if "PAIR" in flags else "two")
    ov, od, ovir, pairs, npairs = oracle_eval(pos, box, style, n, nb, sw, do_pbc="NOPBC" not in flags, nl_mode=mode, cutoff=float(kv.get("NL_CUTOFF",1e30)), stride=int(kv.get("NL_STRIDE",0)), fast_list=True)
    print(line)
    print("  value gpu %.15g oracle %.15g rel %.2e | deriv rel %.2e | virial rel %.2e | t1 %.1f ms t2 %.1f ms sweep %.3f ms build %.3f ms nl %d evals %d"%(v, ov, abs(v-ov)/abs(ov), rel_err(c.derivatives, od), rel_err(c.virial, ovir), t1*1e3, t2*1e3, st['last_sweep_ms'], st['last_build_ms'], st['nl_size'], st['pair_evals']))
    if pairs is not None and mode=="classic":
        gp = c.neighbor_pairs()
        ps = pairs[np.lexsort((pairs[:,1],pairs[:,0]))]
        print("  pair sets identical:", gp.shape==ps.shape and bool((gp==ps).all()), gp.shape, ps.shape)
    c.close()
run("c: COORDINATION GROUPA=1-1000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12}", 1000)
run("c: COORDINATION GROUPA=1-1000 R_0=0.3", 1000)
run("c: COORDINATION GROUPA=1-5000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10", 5000)
run("c: COORDINATION GROUPA=1-5000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=10", 5000)
run("c: COORDINATION GROUPA=1-5000 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLIST NL_CUTOFF=1.0 NL_STRIDE=1", 5000, triclinic=True)
run("c: COORDINATION GROUPA=1-500 GROUPB=501-5000 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLIST NL_CUTOFF=1.0 NL_STRIDE=1", 500, triclinic=True, nb=4500)
run("c: COORDINATION GROUPA=1-500 GROUPB=501-5000 SWITCH={GAUSSIAN R_0=0.2 D_MAX=0.9} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=1", 500, triclinic=True, nb=4500)
run("c: COORDINATION GROUPA=1-100000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10", 100000)

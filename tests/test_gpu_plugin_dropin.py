"""Drop-in test on the GPU box: the UNMODIFIED reference PlumedMain (oracle/_ref, driven through plumed_cmd
like an MD engine) runs the same plumed.dat twice -- once with the built-in CPU COORDINATION, once with
`LOAD FILE=libb200coord_plumed.so` in front, which makes our action answer to the key COORDINATION -- and the
bias, the forces on all atoms and the virial coming back through setForces/setVirial must agree.
Pattern follows plugins/cudaCoord/regtest (CPU and GPU action in one comparison)."""
import os

import numpy as np
import pytest

from helpers import rel_err, water_box
from oracle import refplumed as R
from plumed2_b200 import capi

pytestmark = [pytest.mark.gpu, pytest.mark.ref]

PLUGIN = os.path.join(os.path.dirname(capi.LIB_PATH), "libb200coord_plumed.so")


def _need():
    if not R.available():
        pytest.skip("oracle/_ref (the reference build) is not present on this box")
    if not os.path.exists(PLUGIN):
        pytest.skip("plugin .so not built")


def run_both(natoms, lines, frames, box, watch=("c",), charges=None, key="COORDINATION"):
    os.environ["PLUMED_IGNORE_NL_MEMORY_ERROR"] = "1"
    outs = []
    for load in (False, True):
        pre = ["LOAD FILE=" + PLUGIN] if load else []
        p = R.Plumed(natoms, pre + lines, watch=watch, log="/tmp/plumed_%s.log" % ("gpu" if load else "cpu"))
        if charges is not None:
            p.charges[:] = charges
        res = []
        for step, pos in enumerate(frames):
            r = p.calc(step, pos, box)
            r["values"] = {w: p.value(w) for w in watch}
            res.append(r)
        p.close()
        outs.append(res)
    if "B200-native " + key not in open("/tmp/plumed_gpu.log").read():
        raise AssertionError("the loaded plugin did not take over the %s key" % key)
    assert "B200-native " + key not in open("/tmp/plumed_cpu.log").read()
    return outs


def compare(cpu, gpu, tol=1e-10):
    for step, (a, b) in enumerate(zip(cpu, gpu)):
        for k in a["values"]:
            assert abs(a["values"][k] - b["values"][k]) <= tol * max(abs(a["values"][k]), 1e-300), (step, k)
        assert abs(a["bias"] - b["bias"]) <= tol * max(abs(a["bias"]), 1e-300), step
        assert rel_err(b["forces"], a["forces"]) <= tol, (step, rel_err(b["forces"], a["forces"]))
        assert rel_err(b["virial"], a["virial"]) <= tol, step


def trajectory(n, nframes, seed, triclinic=False, step=0.01):
    pos, box = water_box(n, 100.0, seed=seed, triclinic=triclinic)
    rng = np.random.default_rng(seed)
    frames = []
    for _ in range(nframes):
        pos = pos + step * rng.standard_normal(pos.shape)
        frames.append(pos.copy())
    return frames, box


def test_fp32_mode_through_the_plugin(monkeypatch):
    """the opt-in FP32 sweep, switched on by the environment (same plumed.dat for both arms) and by the plugin's
    additive GPU_FP32 flag; 1e-5 against the reference's CPU action"""
    _need()
    frames, box = trajectory(2000, 5, seed=14)
    body = "GROUPA=1-2000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3"
    lines = ["c: COORDINATION " + body, "RESTRAINT ARG=c AT=100 KAPPA=0.01 SLOPE=0.5"]
    monkeypatch.setenv("B200COORD_FP32", "1")
    cpu, gpu = run_both(2000, lines, frames, box)
    assert "FP32 pair arithmetic" in open("/tmp/plumed_gpu.log").read()
    compare(cpu, gpu, tol=1e-5)
    assert any(np.any(a["forces"] != b["forces"]) for a, b in zip(cpu, gpu))  # a different arithmetic did run
    monkeypatch.delenv("B200COORD_FP32")
    p = R.Plumed(2000, ["LOAD FILE=" + PLUGIN, lines[0] + " GPU_FP32", lines[1]], watch=("c",), log="/tmp/plumed_gpu32.log")
    flagged = []
    for step, pos in enumerate(frames):
        r = p.calc(step, pos, box)
        r["values"] = {"c": p.value("c")}
        flagged.append(r)
    p.close()
    assert "FP32 pair arithmetic" in open("/tmp/plumed_gpu32.log").read()
    compare(cpu, flagged, tol=1e-5)
    for a, b in zip(gpu, flagged):  # same mode either way: identical results
        assert a["values"]["c"] == b["values"]["c"] and np.array_equal(a["forces"], b["forces"])


@pytest.mark.parametrize("body", [
    "GROUPA=1-2000 R_0=0.3",
    "GROUPA=1-2000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3",
    "GROUPA=1-2000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=3",
    "GROUPA=1-300 GROUPB=301-2000 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=1",
    "GROUPA=1-1000 GROUPB=1001-2000 R_0=0.4 PAIR",
    "GROUPA=1-2000 SWITCH={GAUSSIAN R_0=0.25 D_MAX=0.9} NOPBC",
])
@pytest.mark.parametrize("tri", [False, True])
def test_driver_style_runs_match(body, tri):
    _need()
    frames, box = trajectory(2000, 7, seed=12, triclinic=tri)
    lines = ["c: COORDINATION " + body, "RESTRAINT ARG=c AT=100 KAPPA=0.01 SLOPE=0.5"]
    cpu, gpu = run_both(2000, lines, frames, box)
    compare(cpu, gpu)


@pytest.mark.parametrize("body", [
    "GROUPA=1-300 GROUPB=301-2000 I=0.1 EPSILON=80.0 TEMP=300",
    "GROUPA=1-2000 I=0.2 EPSILON=78.4 NLIST NL_CUTOFF=1.2 NL_STRIDE=3",
    "GROUPA=1-300 GROUPB=301-2000 I=0.0 EPSILON=80.0 NLISTCELLS NL_CUTOFF=1.2 NL_STRIDE=2",
])
def test_dhenergy_sibling_action(body):
    """DHENERGY (src/colvar/DHEnergy.cpp) answered by the plugin: charges come from the MD engine (setCharges)"""
    _need()
    frames, box = trajectory(2000, 5, seed=17, triclinic=True)
    q = np.random.default_rng(3).standard_normal(2000)
    lines = ["c: DHENERGY " + body, "RESTRAINT ARG=c AT=1 KAPPA=0.01 SLOPE=0.5"]
    cpu, gpu = run_both(2000, lines, frames, box, charges=q, key="DHENERGY")
    compare(cpu, gpu)


@pytest.mark.parametrize("body", [
    "GROUPA=1-300 GROUPB=301-2000 D_0=0.2 D_MAX=0.5 C=0.8",
    "GROUPA=1-2000 D_0=0.15 D_MAX=0.5 C=0.6 NLIST NL_CUTOFF=0.7 NL_STRIDE=3 ENERGY_UNITS=kcal/mol",
    "GROUPA=1-1000 GROUPB=1001-2000 D_0=0.2 D_MAX=0.9 C=0.4 PAIR",
])
def test_ghbfix_sibling_action(body, tmp_path):
    """GHBFIX (src/colvar/GHBFIX.cpp) answered by the plugin: the TYPES / PARAMS files are read by the action itself"""
    _need()
    frames, box = trajectory(2000, 5, seed=19, triclinic=True)
    rng = np.random.default_rng(6)
    names = ["a", "b", "c"]
    (tmp_path / "types.dat").write_text("#! FIELDS itype\n" + "\n".join(names[t] for t in rng.integers(0, 3, 2000)) + "\n")
    (tmp_path / "params.dat").write_text("#! FIELDS itype jtype eta\n" + "\n".join(
        "%s %s %r" % (i, j, float(rng.standard_normal())) for i in names for j in names if (i, j) != ("c", "c")) + "\nzz a 0.5\n")
    lines = ["c: GHBFIX %s TYPES=%s PARAMS=%s" % (body, tmp_path / "types.dat", tmp_path / "params.dat"),
             "RESTRAINT ARG=c AT=1 KAPPA=0.01 SLOPE=0.5"]
    cpu, gpu = run_both(2000, lines, frames, box, key="GHBFIX")
    compare(cpu, gpu)


def test_metad_with_grid_bias():
    """configs[3] in miniature: COORDINATION driving METAD with a GRID bias through plumed_cmd; forces and
    virial returned every step (hills are deposited, so the bias history must stay identical too)"""
    _need()
    n = 3000
    frames, box = trajectory(n, 12, seed=44, step=0.02)
    outs = []
    for load in (False, True):
        hills = "/tmp/HILLS_%s" % ("gpu" if load else "cpu")
        if os.path.exists(hills):
            os.remove(hills)
        lines = ["c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=4" % n,
                 "md: METAD ARG=c SIGMA=20 HEIGHT=1.2 PACE=2 GRID_MIN=0 GRID_MAX=60000 GRID_BIN=3000 FILE=" + hills]
        pre = ["LOAD FILE=" + PLUGIN] if load else []
        p = R.Plumed(n, pre + lines, watch=("c",))
        res = []
        for step, pos in enumerate(frames):
            r = p.calc(step, pos, box)
            r["values"] = {"c": p.value("c")}
            res.append(r)
        p.close()
        outs.append(res)
    compare(outs[0], outs[1], tol=1e-9)
    assert abs(outs[0][-1]["bias"]) > 0.0


def test_plumed_driver_cli(tmp_path):
    """the same through the `plumed driver` executable and an xyz trajectory (configs[0] workflow)"""
    _need()
    import subprocess
    n = 1000
    frames, box = trajectory(n, 3, seed=8)
    xyz = tmp_path / "traj.xyz"
    with open(xyz, "w") as f:
        for pos in frames:
            f.write("%d\n%.10f %.10f %.10f\n" % (n, box[0, 0], box[1, 1], box[2, 2]))
            for p in pos:
                f.write("Ar %.12f %.12f %.12f\n" % tuple(p))
    dat = tmp_path / "plumed.dat"
    dat.write_text("cpu: COORDINATION GROUPA=1-1000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12}\n"
                   "LOAD FILE=%s\n"
                   "gpu: COORDINATION GROUPA=1-1000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12}\n"
                   "diff: CUSTOM ARG=cpu,gpu FUNC=y-x PERIODIC=NO\n"
                   "PRINT ARG=cpu,gpu,diff FILE=%s FMT=%%.14g\n"
                   "DUMPDERIVATIVES ARG=cpu,gpu FILE=%s FMT=%%.14g\n" % (PLUGIN, tmp_path / "COLVAR", tmp_path / "DERIV"))
    env = dict(os.environ, PLUMED_NUM_THREADS="1")
    r = subprocess.run([R.PLUMED_BIN, "driver", "--plumed", str(dat), "--ixyz", str(xyz), "--length-units", "nm"],
                       cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    rows = np.loadtxt(tmp_path / "COLVAR", comments="#")
    assert rows.shape == (3, 4)
    assert np.all(np.abs(rows[:, 3]) <= 1e-10 * np.abs(rows[:, 1]))
    der = np.loadtxt(tmp_path / "DERIV", comments="#")
    assert rel_err(der[:, 3], der[:, 2]) <= 1e-10


def test_top_level_d_max_keeps_every_digit():
    """the plugin's additive top-level D_MAX (cudaCoord semantics) builds `RATIONAL R_0= D_0= NN= MM= D_MAX=` from the
    parsed doubles: values that need more than six decimals must arrive intact (the same line through the CPU action's
    SWITCH={...} form is the reference)"""
    _need()
    n = 1500
    frames, box = trajectory(n, 4, seed=23)
    outs = []
    for load in (False, True):
        body = ("GROUPA=1-%d R_0=0.2512345678 D_0=0.0000004321 NN=6 MM=12 D_MAX=0.7654321987" % n) if load else \
               ("GROUPA=1-%d SWITCH={RATIONAL R_0=0.2512345678 D_0=0.0000004321 NN=6 MM=12 D_MAX=0.7654321987}" % n)
        pre = ["LOAD FILE=" + PLUGIN] if load else []
        p = R.Plumed(n, pre + ["c: COORDINATION " + body, "RESTRAINT ARG=c AT=100 KAPPA=0.01 SLOPE=0.5"], watch=("c",),
                     log="/tmp/plumed_dmax_%d.log" % load)
        res = []
        for step, pos in enumerate(frames):
            r = p.calc(step, pos, box)
            r["values"] = {"c": p.value("c")}
            res.append(r)
        p.close()
        outs.append(res)
    compare(outs[0], outs[1])
    assert "on CUDA device" in open("/tmp/plumed_dmax_1.log").read() or "on the current CUDA device" in open("/tmp/plumed_dmax_1.log").read()


def test_gpu_devices_keyword_shards_one_process_over_two_gpus():
    """GPU_DEVICES=0,1 on the input line: `plumed driver`-style single process, the plugin shards the i-atoms over two
    devices (b200coord_group_*).  Same numbers as the reference's CPU action, forces and virial included."""
    _need()
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = 6000
    frames, box = trajectory(n, 6, seed=29, triclinic=True)
    outs = []
    for load in (False, True):
        body = "GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3" % n
        pre = ["LOAD FILE=" + PLUGIN] if load else []
        p = R.Plumed(n, pre + ["c: COORDINATION " + body + (" GPU_DEVICES=0,1" if load else ""),
                               "RESTRAINT ARG=c AT=100 KAPPA=0.01 SLOPE=0.5"], watch=("c",), log="/tmp/plumed_devs_%d.log" % load)
        res = []
        for step, pos in enumerate(frames):
            r = p.calc(step, pos, box)
            r["values"] = {"c": p.value("c")}
            res.append(r)
        p.close()
        outs.append(res)
    compare(outs[0], outs[1])
    assert "sharded over 2 CUDA devices" in open("/tmp/plumed_devs_1.log").read()


@pytest.mark.parametrize("body", [
    "GROUPA=1-3000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3",
    "GROUPA=1-500 GROUPB=501-3000 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=2",
    "GROUPA=1-600 GROUPB=401-3000 R_0=0.3",  # atoms 401-600 are in both groups: the action must fall back to PLUMED's loops
])
def test_host_fast_path_gathers_and_scatters_itself(body, monkeypatch):
    """large systems: the action gathers its atoms and scatters its forces itself (OpenMP) instead of PLUMED's serial
    retrieveAtoms / setForcesOnAtoms; forced on here for a small system (B200COORD_HOST_FAST=2).  Bias, forces on all atoms
    and virial must be the CPU action's, with RESTRAINT and with a second action (a function of the CV) pushing forces."""
    _need()
    monkeypatch.setenv("B200COORD_HOST_FAST", "2")
    n = 3000
    frames, box = trajectory(n, 6, seed=37, triclinic=True)
    lines = ["c: COORDINATION " + body, "f: CUSTOM ARG=c FUNC=0.001*x*x PERIODIC=NO",
             "RESTRAINT ARG=c,f AT=100,3 KAPPA=0.01,0.5 SLOPE=0.5,0.1"]
    cpu, gpu = run_both(n, lines, frames, box)
    compare(cpu, gpu)
    said = "gathered and forces scattered by the action itself" in open("/tmp/plumed_gpu.log").read()
    assert said == ("GROUPB=401" not in body)


@pytest.mark.parametrize("body", [
    "GROUPA=1-3000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3",
    "GROUPA=2-3000:2 GROUPB=1-2999:2 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=2",
    "GROUPA=1-600 GROUPB=401-2900 R_0=0.3",  # atoms 401-600 own two derivative rows each; atoms 2901-3000 get no force
])
def test_engine_arrays_on_the_device(body):
    """an MD engine that keeps positions and forces on the GPU publishes the two device arrays
    (b200coord_coupling_publish); the action with GPU_COUPLING=<name> reads positions there and adds force-on-CV x
    derivative there.  PLUMED itself is handed zeros as host positions, to show they are not looked at.  The device
    force array, the virial and the bias must be what the CPU action returns through setForces / setVirial."""
    _need()
    import ctypes as C
    L = capi.lib()
    n = 3000
    frames, box = trajectory(n, 6, seed=41, triclinic=True)
    lines = ["c: COORDINATION " + body, "f: CUSTOM ARG=c FUNC=0.001*x*x PERIODIC=NO",
             "RESTRAINT ARG=c,f AT=100,3 KAPPA=0.01,0.5 SLOPE=0.5,0.1"]
    p = R.Plumed(n, lines, watch=("c",), log="/tmp/plumed_cpu.log")
    cpu = []
    for step, pos in enumerate(frames):
        r = p.calc(step, pos, box)
        r["values"] = {"c": p.value("c")}
        cpu.append(r)
    p.close()

    d_pos, d_force = C.c_void_p(), C.c_void_p()
    assert L.b200coord_device_alloc(24 * n, C.byref(d_pos)) == 0
    assert L.b200coord_device_alloc(24 * n, C.byref(d_force)) == 0
    try:
        assert L.b200coord_coupling_publish(b"engine0", 0, d_pos, d_force, n) == 0
        host = np.zeros((n, 3))
        with pytest.raises(R.PlumedError):  # a name nobody published
            R.Plumed(n, ["LOAD FILE=" + PLUGIN, "c: COORDINATION " + body + " GPU_COUPLING=nobody"], log="/tmp/plumed_gpu_bad.log")
        p = R.Plumed(n, ["LOAD FILE=" + PLUGIN, lines[0] + " GPU_COUPLING=engine0"] + lines[1:], watch=("c",),
                     log="/tmp/plumed_gpu.log")
        zeros = np.zeros((n, 3))
        for step, pos in enumerate(frames):
            x = np.ascontiguousarray(pos)
            assert L.b200coord_memcpy_h2d(d_pos, x.ctypes.data_as(C.c_void_p), 24 * n) == 0
            assert L.b200coord_memcpy_h2d(d_force, zeros.ctypes.data_as(C.c_void_p), 24 * n) == 0
            r = p.calc(step, host, box)
            got = np.empty((n, 3))
            assert L.b200coord_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_force, 24 * n) == 0
            a = cpu[step]
            assert abs(p.value("c") - a["values"]["c"]) <= 1e-10 * abs(a["values"]["c"])
            assert abs(r["bias"] - a["bias"]) <= 1e-10 * abs(a["bias"])
            assert not np.any(r["forces"])  # nothing came back through the host array
            assert rel_err(got, a["forces"]) <= 1e-10, (step, rel_err(got, a["forces"]))
            assert rel_err(r["virial"], a["virial"]) <= 1e-10
        p.close()
        assert "GPU_COUPLING" in open("/tmp/plumed_gpu.log").read()
    finally:
        L.b200coord_coupling_withdraw(b"engine0")
        L.b200coord_device_free(d_pos)
        L.b200coord_device_free(d_force)

"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): one process per GPU, i-atoms sharded by rank,
positions and derivative rows exchanged inside the library -- NCCL all-gathers, or (peer-pull) NVLink peer memory: rows
pulled by the rank that returns them, positions pulled by the gather on steps that keep the list -- value/virial all-reduced.
Every rank must end up with the single-GPU result (= the oracle's) for its slice."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from helpers import oracle_from_line, rel_err, water_box

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


LINES = ["c: COORDINATION GROUPA=1-6000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3",
         "c: COORDINATION GROUPA=1-500 GROUPB=501-6000 SWITCH={EXP R_0=0.2 D_MAX=0.9} NLISTCELLS NL_CUTOFF=1.0 NL_STRIDE=1",
         "c: COORDINATION GROUPA=1-3000 GROUPB=3001-6000 R_0=0.5 PAIR"]


def _worker(rank, world, port, out_dir, peer):
    import torch
    import torch.distributed as dist
    import plumed2_b200 as P
    from plumed2_b200 import capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = capi.lib()
        n = 6000
        pos0, box = water_box(n, 100.0, seed=31, triclinic=True)
        rng = np.random.default_rng(1)
        for li, line in enumerate(LINES):
            c = P.Coordination.from_input(line, device=rank, rank=rank, nranks=world)
            ids = [P.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            c.comm_init(ids[0])
            if peer and "PAIR" not in line:
                handles = [None] * world
                dist.all_gather_object(handles, c.peer_export())
                c.peer_attach(handles)
            lo, cnt = P.shard_range(n, rank, world)
            sb, sc = C.c_uint(), C.c_uint()
            capi.check(L.b200coord_my_slice(c._ctx, C.byref(sb), C.byref(sc)))
            assert (sb.value, sc.value) == (lo, cnt)
            pos = pos0.copy()
            for step in range(4):
                pos = pos + 0.01 * rng.standard_normal(pos.shape)
                c.prepare(step)
                c._set_box(box)
                sl = np.ascontiguousarray(pos[lo:lo + cnt])
                der = np.zeros((cnt, 3))
                vir = np.zeros(9)
                val = C.c_double(0)
                capi.check(L.b200coord_calculate_distributed(c._ctx, sl.ctypes.data_as(C.c_void_p), C.byref(val),
                                                             der.ctypes.data_as(C.c_void_p),
                                                             vir.ctypes.data_as(C.POINTER(C.c_double))), c._ctx)
                np.savez(os.path.join(out_dir, "r%d_l%d_s%d.npz" % (rank, li, step)), value=val.value, deriv=der,
                         virial=vir.reshape(3, 3), lo=lo, cnt=cnt, pos=pos)
            c.close()
            dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("peer", [False, True], ids=["nccl-allgather", "peer-pull"])
def test_multi_gpu_distributed_step_matches_oracle(tmp_path, peer, world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), peer), nprocs=world, join=True)
    _, box = water_box(6000, 100.0, seed=31, triclinic=True)
    for li, line in enumerate(LINES):
        list_pos = None
        stride = 3 if "NL_STRIDE=3" in line else 1
        for step in range(4):
            parts = [np.load(tmp_path / ("r%d_l%d_s%d.npz" % (r, li, step))) for r in range(world)]
            pos = parts[0]["pos"]
            if step % stride == 0 or list_pos is None:
                list_pos = pos.copy()
            ref = oracle_from_line(line, pos, box, list_positions=list_pos)
            full = np.zeros((6000, 3))
            for p in parts:
                assert abs(float(p["value"]) - ref["value"]) <= 1e-10 * abs(ref["value"]), (line, step)
                assert rel_err(p["virial"], ref["virial"]) <= 1e-10
                full[int(p["lo"]):int(p["lo"]) + int(p["cnt"])] = p["deriv"]
            assert rel_err(full, ref["deriv"]) <= 1e-10, (line, step)


@pytest.mark.parametrize("ndev", [2, 8])
def test_group_of_devices_in_one_process(ndev):
    """b200coord_group_*: several GPUs inside ONE process (what the plugin's GPU_DEVICES keyword uses): worker thread per
    device, NCCL + peer memory between the contexts of this process.  Rebuild steps, steps that keep the list (positions
    pulled over NVLink), two groups; against the oracle."""
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    import plumed2_b200 as P
    from plumed2_b200 import capi
    L = capi.lib()
    n = 9000
    pos0, box = water_box(n, 100.0, seed=51, triclinic=True)
    for line in ["c: COORDINATION GROUPA=1-9000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=3",
                 "c: COORDINATION GROUPA=1-3000 GROUPB=3001-9000 SWITCH={EXP R_0=0.2 D_MAX=0.8} NLISTCELLS NL_CUTOFF=0.9 NL_STRIDE=2"]:
        single = P.Coordination.from_input(line, device=0)  # borrows the parsed config / switch
        cfg = capi.Config.from_buffer_copy(single._cfg)
        devs = (C.c_int * ndev)(*range(ndev))
        g = C.c_void_p()
        absidx = np.ascontiguousarray(single.atoms)
        capi.check(L.b200coord_group_create(C.byref(cfg), C.byref(single.switch), absidx.ctypes.data_as(C.POINTER(C.c_uint)),
                                            devs, ndev, C.byref(g)))
        assert L.b200coord_group_size(g) == ndev
        rng = np.random.default_rng(4)
        pos, list_pos = pos0.copy(), None
        b9 = np.ascontiguousarray(box.reshape(9))
        for step in range(6):
            pos = pos + 0.006 * rng.standard_normal(pos.shape)
            will = C.c_int(0)
            assert L.b200coord_group_prepare(g, step, 0, C.byref(will)) == 0
            if will.value or list_pos is None:
                list_pos = pos.copy()
            assert L.b200coord_group_set_box(g, b9.ctypes.data_as(C.POINTER(C.c_double))) == 0
            der = np.zeros((n, 3))
            vir = np.zeros(9)
            val = C.c_double(0)
            p = np.ascontiguousarray(pos)
            rc = L.b200coord_group_calculate(g, p.ctypes.data_as(C.c_void_p), C.byref(val), der.ctypes.data_as(C.c_void_p),
                                             vir.ctypes.data_as(C.POINTER(C.c_double)))
            assert rc == 0, L.b200coord_group_last_error(g)
            ref = oracle_from_line(line, pos, box, list_positions=list_pos)
            assert abs(val.value - ref["value"]) <= 1e-10 * abs(ref["value"]), (line, step)
            assert rel_err(der, ref["deriv"]) <= 1e-10 and rel_err(vir.reshape(3, 3), ref["virial"]) <= 1e-10, (line, step)
        L.b200coord_group_destroy(g)
        single.close()


def test_group_member_that_fails_does_not_leave_its_peers_waiting(monkeypatch):
    """a device of an in-process group that fails before a collective (here on purpose: B200COORD_TEST_FAIL=<rank>:<call>)
    used to leave the other workers in their all-reduce forever; now their communicators are aborted, the call returns
    the member's error, and the group keeps reporting it"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import time
    import plumed2_b200 as P
    from plumed2_b200 import capi
    L = capi.lib()
    n = 4000
    pos, box = water_box(n, 100.0, seed=53)
    line = "c: COORDINATION GROUPA=1-4000 SWITCH={RATIONAL R_0=0.3 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=5"
    single = P.Coordination.from_input(line, device=0)
    cfg = capi.Config.from_buffer_copy(single._cfg)
    devs = (C.c_int * 2)(0, 1)
    g = C.c_void_p()
    absidx = np.ascontiguousarray(single.atoms)
    monkeypatch.setenv("B200COORD_TEST_FAIL", "1:1")
    capi.check(L.b200coord_group_create(C.byref(cfg), C.byref(single.switch), absidx.ctypes.data_as(C.POINTER(C.c_uint)),
                                        devs, 2, C.byref(g)))
    monkeypatch.delenv("B200COORD_TEST_FAIL")
    b9 = np.ascontiguousarray(box.reshape(9))
    der, vir, val = np.zeros((n, 3)), np.zeros(9), C.c_double(0)
    rcs = []
    t0 = time.time()
    for step in range(3):
        will = C.c_int(0)
        L.b200coord_group_prepare(g, step, 0, C.byref(will))
        L.b200coord_group_set_box(g, b9.ctypes.data_as(C.POINTER(C.c_double)))
        p = np.ascontiguousarray(pos + 0.001 * step)
        rcs.append(L.b200coord_group_calculate(g, p.ctypes.data_as(C.c_void_p), C.byref(val), der.ctypes.data_as(C.c_void_p),
                                               vir.ctypes.data_as(C.POINTER(C.c_double))))
    assert rcs[0] == 0 and rcs[1] != 0 and rcs[2] != 0, rcs
    assert b"injected" in L.b200coord_group_last_error(g)
    assert time.time() - t0 < 60
    L.b200coord_group_destroy(g)
    single.close()

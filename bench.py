#!/usr/bin/env python
"""bench.py -- COORDINATION pair evaluations / s (value + 3N derivatives + virial) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1] at the north-star headline size): one group of `--natoms-per-gpu` x N atoms
(default 1,000,000 per GPU), uniform water-like box at 100 atoms/nm^3, orthorhombic PBC,
`COORDINATION GROUPA=1-n SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10`.
A step = prepare() + calculate() on one frame (frames: base configuration + a small wiggle, so the frozen list
stays valid between the rebuilds that happen every 10th step).  The i-atoms are sharded over the ranks and the
per-GPU share is fixed -> "scaling": "weak".

Numerator of the metric for BOTH arms: the number of pairs within NL_CUTOFF at the last rebuild (= the size of
the reference's NLIST list); evaluating a pair from both ends on the GPU does not count twice.

  value : inputs resident in HBM (positions of all frames uploaded before the timed region, result left on the
          device), all K steps enqueued on the library's stream and timed with CUDA events on that stream.
  e2e   : the same K steps through the reference-facing C-ABI call with pinned HOST buffers: each rank uploads
          its slice of the positions, downloads its slice of the derivatives + value + virial every step.
  roofline : the pair-sweep kernel; achieved = 68 algorithmic FLOP per listed pair / CUDA-event
          duration of the kernel, peak = FP64 FMA rate measured in this run with a DFMA microbenchmark.
  cpu_baseline : the REAL reference (oracle/_ref) timed on this box's host cores on a bounded sample.

--impl reference times the reference's own CPU COORDINATION (oracle/_ref through plumed_cmd, all host threads)
on a bounded sample of the same workload (same density and keywords, fewer atoms; the reference's cost is linear
in the atom count with NLISTCELLS, the only list flavour it can run above 32768 atoms).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
DENSITY = 100.0
FLOP_PER_PAIR = 68.0  # ortho PBC + rational 6/12, SURVEY.md 8(d)
SWITCH = "RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8"
NL_CUTOFF, NL_STRIDE = 1.0, 10
METRIC = "COORDINATION pair evals/s (value+3N derivs+virial)"
UNIT = "pair_evals/s"


def make_frames(n, nframes, seed=SEED):
    """base configuration + small cumulative wiggle (<=0.01 nm from the base), box edge L"""
    rng = np.random.default_rng(seed)
    L = (n / DENSITY) ** (1.0 / 3.0)
    base = rng.random((n, 3)) * L
    frames = []
    for f in range(nframes):
        d = rng.standard_normal((n, 3))
        frames.append(base + 0.002 * d)
    return frames, np.diag([L, L, L])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None
        self.windows = []  # [t0, t1] of the timed regions (time.time()); samples are stamped on receipt

    def wait_first(self, timeout=3.0):
        """nvidia-smi takes a moment to start: do not enter a timed region before it delivers"""
        t_end = time.time() + timeout
        while not self.rows and time.time() < t_end:
            time.sleep(0.01)

    def open_window(self):
        self.windows.append([time.time(), None])

    def close_window(self):
        self.windows[-1][1] = time.time()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")] + [time.time()])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        inside = [r for r in self.rows if any(w[0] <= r[-1] <= (w[1] or r[-1]) for w in self.windows)]
        which = "timed regions (value + e2e)"
        if not inside:  # regions shorter than the sampling period: fall back to everything sampled under load
            inside, which = self.rows, "whole run (timed regions shorter than the 20 ms sampling period)"
        sm = [float(r[0]) for r in inside if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in inside:
            if len(r) >= 8:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": which}


def pin_to_gpu_numa_node(local):
    """Multi-GPU runs: keep this rank's threads (and therefore its page-locked staging buffers, which are placed by
    first touch) on the NUMA node its GPU hangs off, so that eight ranks do not push their PCIe traffic through one
    socket.  Best effort: returns the node or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def pinned_array(L, shape):
    """numpy view of page-locked host memory from the library's allocator"""
    nbytes = int(np.prod(shape)) * 8
    ptr = C.c_void_p()
    from plumed2_b200 import capi
    capi.check(L.b200coord_host_alloc(nbytes, C.byref(ptr)))
    buf = (C.c_double * (nbytes // 8)).from_address(ptr.value)
    return np.ctypeslib.as_array(buf).reshape(shape), ptr


# ------------------------------------------------------------------------------------------------
def reference_cpu(sample_atoms, steps, warmup, threads):
    """the reference's CPU COORDINATION on a bounded sample; returns dict(value, ms_per_step, ...).

    Both list flavours are timed and the FASTER one is reported: NLIST (the keyword our arm uses; the reference
    can only run it below 32768 atoms and its rebuild is O(N^2), so a 20000-atom sample flatters it) and NLISTCELLS
    (the only flavour it can run at the headline size; cost linear in N, but it sweeps the 27-cell superset)."""
    os.environ["PLUMED_NUM_THREADS"] = str(threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["PLUMED_IGNORE_NL_MEMORY_ERROR"] = "1"
    from oracle import refplumed as R
    from oracle import oracle as O
    if not R.available():
        return None
    n = sample_atoms
    frames, box = make_frames(n, 4)
    # numerator: pairs within NL_CUTOFF at the rebuild frame (what NLIST lists)
    nl = O.NeighborList(O.NL_SINGLELIST, n, 0, cutoff=NL_CUTOFF, stride=NL_STRIDE)
    nl.update(O.make_pbc(box), frames[0], fast=True)
    pairs = int(nl.size())
    best = None
    flavours = {}
    for flavour in (["NLIST", "NLISTCELLS"] if n <= 32768 else ["NLISTCELLS"]):
        line = "c: COORDINATION GROUPA=1-%d SWITCH={%s} %s NL_CUTOFF=%r NL_STRIDE=%d" % (n, SWITCH, flavour, NL_CUTOFF, NL_STRIDE)
        p = R.Plumed(n, [line, "RESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1"])
        for s in range(warmup):
            p.calc(s, frames[s % len(frames)], box)
        nsteps = steps * (3 if flavour == "NLIST" else 1)  # NLIST steps are ~8x cheaper: time three times as many
        t0 = time.perf_counter()
        for s in range(warmup, warmup + nsteps):
            p.calc(s, frames[s % len(frames)], box)
        dt = time.perf_counter() - t0
        p.close()
        r = {"value": pairs * nsteps / dt, "ms_per_step": 1e3 * dt / nsteps, "pairs_per_step": pairs, "atoms": n,
             "keywords": line, "seconds": dt, "flavour": flavour, "steps": nsteps}
        flavours[flavour] = r["value"]
        if best is None or r["value"] > best["value"]:
            best = r
    best["flavours"] = flavours
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    r = reference_cpu(args.ref_sample_atoms, args.steps, args.warmup, threads)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference build) not present on this box"}))
        return
    sample = ("%d-atom box of the same density/keywords, faster of the reference's two list flavours (%s; pair evals/s: %s); "
              "%d warm-up + %d timed steps incl. rebuilds every %d.  The reference cannot run NLIST above 32768 atoms, "
              "NLISTCELLS costs the same per atom at any size" %
              (r["atoms"], r["flavour"], ", ".join("%s %.3g" % kv for kv in r["flavours"].items()), args.warmup, args.steps,
               NL_STRIDE))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_keywords": r["keywords"], "pairs_per_step_sample": r["pairs_per_step"]}
    print(json.dumps(line))


def workload_config(args, world):
    n = args.natoms_per_gpu * world
    return {"workload": "BASELINE configs[1] at headline size: COORDINATION single group, %d atoms (%d per GPU), "
                        "orthorhombic PBC, SWITCH={%s} NLIST NL_CUTOFF=%g NL_STRIDE=%d, 100 atoms/nm^3"
                        % (n, args.natoms_per_gpu, SWITCH, NL_CUTOFF, NL_STRIDE),
            "natoms": n, "natoms_per_gpu": args.natoms_per_gpu, "parallelism": "i-atom shards x%d%s" % (world, "" if world == 1 else (", NCCL all-gather" if args.no_peer else ", in-kernel NVLink peer stores")),
            "pair_count": "pairs within NL_CUTOFF at the last rebuild (NLIST size); both arms",
            "cache": "per-step inputs (neighbour list %.1f GB + positions) exceed the 126 MB L2" %
                     (n / world * 419 * 4 / 1e9)}


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = pin_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import plumed2_b200 as P
    from plumed2_b200 import capi
    L = capi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        capi.check(L.b200coord_device_synchronize())

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = args.natoms_per_gpu * world
    K, W, F = args.steps, args.warmup, args.frames
    frames, box = make_frames(n, F)
    line = "c: COORDINATION GROUPA=1-%d SWITCH={%s} NLIST NL_CUTOFF=%r NL_STRIDE=%d" % (n, SWITCH, NL_CUTOFF, NL_STRIDE)
    c = P.Coordination.from_input(line, device=local, rank=rank, nranks=world,
                                  precision=capi.FP32 if args.fp32 else capi.FP64)
    if world > 1:
        ids = [P.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        c.comm_init(ids[0])
        if not args.no_peer:  # fused sweep + exchange: derivative rows go to the peers from inside the sweep kernel
            handles = [None] * world
            dist.all_gather_object(handles, c.peer_export())
            c.peer_attach(handles)
    c._set_box(box)
    ctx = c._ctx
    sb, sc = C.c_uint(), C.c_uint()
    capi.check(L.b200coord_my_slice(ctx, C.byref(sb), C.byref(sc)))
    lo, cnt = sb.value, sc.value

    peak = C.c_double(0)
    capi.check(L.b200coord_measure_fp64_peak(local, C.byref(peak)))

    # ---------------- value: inputs resident in HBM
    d_frames = []
    for f in frames:
        p = C.c_void_p()
        capi.check(L.b200coord_device_alloc(f.nbytes, C.byref(p)))
        capi.check(L.b200coord_memcpy_h2d(p, np.ascontiguousarray(f).ctypes.data_as(C.c_void_p), f.nbytes))
        d_frames.append(p)
    d_out = C.c_void_p()
    capi.check(L.b200coord_device_alloc((3 * n + 10) * 8, C.byref(d_out)))
    step = 0
    for _ in range(W):
        c.prepare(step)
        capi.check(L.b200coord_enqueue_device(ctx, d_frames[step % F], d_out), ctx)
        step += 1
    barrier()
    st0 = c.stats()
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    barrier()
    sampler.open_window()
    capi.check(L.b200coord_stream_mark(ctx, 0), ctx)
    for _ in range(K):
        c.prepare(step)
        capi.check(L.b200coord_enqueue_device(ctx, d_frames[step % F], d_out), ctx)
        step += 1
    capi.check(L.b200coord_stream_mark(ctx, 1), ctx)
    ms = C.c_float(0)
    capi.check(L.b200coord_stream_elapsed_ms(ctx, C.byref(ms)), ctx)
    sampler.close_window()
    barrier()
    st1 = c.stats()
    value_ms = allmax(float(ms.value))
    pairs_per_step = allsum(float(st1["nl_size"]))
    launches = int(st1["kernel_launches"] - st0["kernel_launches"])
    sweep_ms = st1["sweep_ms_sum"] / max(1, st1["sweep_count"])
    build_ms = st1["build_ms_sum"] / max(1, st1["build_count"])
    tail = np.zeros(10)
    capi.check(L.b200coord_memcpy_d2h(tail.ctypes.data_as(C.c_void_p), C.c_void_p(d_out.value + 3 * n * 8), 80))
    value_cv = float(tail[9])
    for p in d_frames:
        L.b200coord_device_free(p)
    L.b200coord_device_free(d_out)

    # ---------------- e2e: pinned host buffers through the C-ABI call, H2D + D2H inside the timed region
    h_frames = []
    for f in frames:
        a, _ptr = pinned_array(L, (max(cnt, 1), 3))
        a[:cnt] = f[lo:lo + cnt]
        h_frames.append(a)
    h_deriv, _p2 = pinned_array(L, (max(cnt, 1), 3))
    vir = np.zeros(9)
    val = C.c_double(0)

    def e2e_step(s):
        c.prepare(s)
        src = h_frames[s % F]
        if world > 1:
            capi.check(L.b200coord_calculate_distributed(ctx, src.ctypes.data_as(C.c_void_p), C.byref(val),
                                                         h_deriv.ctypes.data_as(C.c_void_p),
                                                         vir.ctypes.data_as(C.POINTER(C.c_double))), ctx)
        else:
            capi.check(L.b200coord_calculate(ctx, src.ctypes.data_as(C.c_void_p), C.byref(val),
                                             h_deriv.ctypes.data_as(C.c_void_p),
                                             vir.ctypes.data_as(C.POINTER(C.c_double))), ctx)

    for _ in range(W):
        e2e_step(step)
        step += 1
    barrier()
    sampler.open_window()
    capi.check(L.b200coord_stream_mark(ctx, 0), ctx)
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step(step)
        step += 1
    capi.check(L.b200coord_stream_mark(ctx, 1), ctx)
    capi.check(L.b200coord_stream_elapsed_ms(ctx, C.byref(ms)), ctx)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.close_window()
    clocks = sampler.stop()
    barrier()
    e2e_ms = allmax(max(float(ms.value), wall_ms))
    st2 = c.stats()
    e2e_pairs = allsum(float(st2["nl_size"]))

    if rank == 0:
        value = pairs_per_step * K / (value_ms * 1e-3)
        e2e_value = e2e_pairs * K / (e2e_ms * 1e-3)
        my_pairs = float(st1["nl_size"])
        achieved = FLOP_PER_PAIR * my_pairs / (sweep_ms * 1e-3) / 1e12 if sweep_ms > 0 else None
        traffic = None
        tf = os.path.join(ROOT, "profiles", "sweep_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        rows_mine = n / world
        # algorithmic HBM bytes of one sweep: 4 B per list entry (2 per pair), the 32 B records once, 24 B of derivatives
        list_bytes = 2.0 * my_pairs * 4 + 32.0 * n + 24.0 * rows_mine
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": value_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32 pair arithmetic, f64 minimum image and accumulation (opt-in, 1e-5)" if args.fp32 else "f64",
               "data": "synthetic", "config": workload_config(args, world),
               "pairs_per_step": pairs_per_step, "cv_value": value_cv,
               "roofline": {"kernel": "k_sweep_list<rationalfix6, orthorhombic, %s>" % ("float" if args.fp32 else "double"),
                            "bound": "fp64",
                            "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                            "frac": (achieved / peak.value) if achieved else None, "traffic": traffic,
                            "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                            "algorithmic_flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": my_pairs,
                            "kernel_ms": sweep_ms, "executed_pair_evals_per_launch": 2.0 * my_pairs,
                            "hbm": {"achieved_gbs": list_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None,
                                    "peak_gbs": hbm_peak, "source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
               "rebuild_ms": build_ms, "rebuilds_in_timed_region": int(st1["build_count"]),
               "rebuild_kinds": {"super_list_builds_total": int(st1["super_builds"]),
                                 "filter_rebuilds_total": int(st1["filter_rebuilds"]), "rebuilds_total": int(st1["rebuilds"])},
               "clocks": clocks, "gpu_launches": launches,
               "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / K,
                       "h2d_bytes_per_step": int(cnt * 24), "d2h_bytes_per_step": int(cnt * 24 + 80),
                       "api": "b200coord_calculate_distributed" if world > 1 else "b200coord_calculate",
                       "numa_node_of_rank0": numa_node}}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r = reference_cpu(args.ref_sample_atoms, NL_STRIDE, 0, threads)
            if r is not None:
                out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference",
                                       "sample": "%d-atom box, same density/keywords, faster of NLIST / NLISTCELLS (%s; %s), "
                                                 "%d steps with a rebuild every %d, %.1f s" %
                                                 (r["atoms"], r["flavour"],
                                                  ", ".join("%s %.3g" % kv for kv in r["flavours"].items()), r["steps"],
                                                  NL_STRIDE, r["seconds"])}
            else:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                                       "sample": "oracle/_ref not present on this box"}
        print(json.dumps(out))
    c.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--natoms-per-gpu", type=int, default=1000000)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--ref-sample-atoms", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fp32", action="store_true",
                    help="opt-in FP32 sweep (B200COORD_FP32); the default and the headline number are FP64")
    ap.add_argument("--no-peer", action="store_true", help="combine with NCCL all-gather instead of in-kernel peer stores")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- COORDINATION pair evaluations / s (value + 3N derivatives + virial) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1] at the north-star headline size): one group of `--natoms-per-gpu` x N atoms
(default 1,000,000 per GPU), uniform water-like box at 100 atoms/nm^3, orthorhombic PBC,
`COORDINATION GROUPA=1-n SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10`.
A step = prepare() + calculate() on one frame.  The i-atoms are sharded over the ranks and the per-GPU share is
fixed -> "scaling": "weak".

Frames: a CUMULATIVE random walk (0.0013 nm rms per coordinate and step, about water at 300 K with a 2 fs step), one
new frame per step, so everything the engine conditions on displacements sees what an MD run shows it: partners that
come back inside D_MAX + skin (far parts of the rows get visited), a super-list that expires.  `regimes` also times
the two ends: "best" = frames that only jitter around one configuration (round-1 bench: both shortcuts always
apply), "worst" = the same drifting frames with both shortcuts switched off (every far part visited, every rebuild a
full cell scan).  The headline `value` is "typical".

Numerator of the metric for BOTH arms: the number of pairs within NL_CUTOFF at the last rebuild (= the size of the
reference's NLIST list); evaluating a pair from both ends on the GPU does not count twice.

  value : inputs resident in HBM (all frames uploaded before the timed region, result left on the device), the K
          steps enqueued on the library's stream and timed with CUDA events on that stream.
  sustained : the same K-step block repeated (frames walked back and forth) until >= 1 s has been timed.
  e2e   : the same K steps through the reference-facing C-ABI call with pinned HOST buffers: each rank uploads its
          slice of the positions, downloads its slice of the derivatives + value + virial every step.
  e2e_plumed : (1 GPU) the same steps through the UNMODIFIED PlumedMain of oracle/_ref, driven with plumed_cmd like
          an MD engine, with `LOAD FILE=libb200coord_plumed.so` -- the whole drop-in path including PLUMED's own
          per-step host work (atom gather, Value stores, apply()).  The reference build only hosts the plugin here.
  parity_check : (N > 1) one untimed step of the distributed engine against a single-GPU context on the same frame.
  roofline : the pair-sweep kernel; achieved = 68 algorithmic FLOP per listed pair / CUDA-event
          duration of the kernel, peak = FP64 FMA rate measured in this run with a DFMA microbenchmark.
  other_configs : BASELINE.json configs[0], [2] (both list flavours), [3] (through plumed_cmd) and [4] (strong
          scaling over the N GPUs of this run), a few steps each.
  cuda_baseline : (1 GPU) the reference's own CUDA prototype plugins/cudaCoord, unchanged, next to our plugin in one
          PLUMED input (100k and 1M atoms): time in calculate() of each action.
  cpu_baseline : the REAL reference (oracle/_ref) timed on this box's host cores on a bounded sample.

--impl reference times the reference's own CPU COORDINATION (oracle/_ref through plumed_cmd, all host threads) at
the literal size of BASELINE configs[1] (100,000 atoms; NLISTCELLS + PLUMED_IGNORE_NL_MEMORY_ERROR, the only list
flavour the reference can run above 32768 atoms) and says so in ITS config.workload.
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
DRIFT = 0.0013  # nm rms per coordinate and step
METRIC = "COORDINATION pair evals/s (value+3N derivs+virial)"
UNIT = "pair_evals/s"


def box_edge(n, density=DENSITY):
    return (n / density) ** (1.0 / 3.0)


def make_frames(n, nframes, seed=SEED, drift=None, density=DENSITY):
    """numpy frames: base configuration + i.i.d. jitter (drift None) or a cumulative random walk"""
    rng = np.random.default_rng(seed)
    L = box_edge(n, density)
    pos = rng.random((n, 3)) * L
    frames = []
    for _ in range(nframes):
        if drift is None:
            frames.append(pos + 0.002 * rng.standard_normal((n, 3)))
        else:
            pos = pos + drift * rng.standard_normal((n, 3))
            frames.append(pos.copy())
    return frames, np.diag([L, L, L])


def make_frames_device(n, nframes, drift, seed=SEED, density=DENSITY, box=None):
    """the same kind of frames generated on the GPU (identical on every rank: same seed, same generator)"""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    L = box_edge(n, density)
    boxm = np.diag([L, L, L]) if box is None else box
    frac = torch.rand((n, 3), generator=g, device="cuda", dtype=torch.float64)
    pos = frac @ torch.tensor(boxm, device="cuda", dtype=torch.float64)
    frames = []
    for _ in range(nframes):
        step = torch.randn((n, 3), generator=g, device="cuda", dtype=torch.float64)
        if drift is None:
            frames.append(pos + 0.002 * step)
        else:
            pos = pos + drift * step
            frames.append(pos.clone())
    return frames, boxm


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
        which = "timed regions (value + sustained + e2e)"
        if not inside:  # regions shorter than the sampling period: fall back to everything sampled under load
            inside, which = self.rows, "whole run (timed regions shorter than the 20 ms sampling period)"
        sm = [float(r[0]) for r in inside if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in inside if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in inside:
            if len(r) >= 8:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm), "window": which}


def pin_to_gpu_numa_node(local):
    """Multi-GPU runs: keep this rank's threads (and therefore its page-locked staging buffers, which are placed by
    first touch) on the NUMA node its GPU hangs off, so that eight ranks do not push their PCIe traffic through one
    socket.  Best effort: returns the node or None (single-node boxes report -1)."""
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
    return _pin_from_topo(local)


def _pin_from_topo(local):
    """sysfs has no NUMA node for the device (containers, VMs): take the CPU / NUMA affinity `nvidia-smi topo -m` prints"""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        lines = [l for l in out.splitlines() if l.strip()]
        hdr = next(l for l in lines if "CPU Affinity" in l)
        cols = [c.strip() for c in hdr.split("\t")]
        row = next(l for l in lines if l.split("\t")[0].strip() == "GPU%d" % local)
        cells = [c.strip() for c in row.split("\t")]
        # the header has one leading empty cell per row label
        off = len(cells) - len(cols)
        cpu_aff = cells[cols.index("CPU Affinity") + off]
        numa = cells[cols.index("NUMA Affinity") + off] if "NUMA Affinity" in cols else None
        cpus = set()
        for part in cpu_aff.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return "numa %s (cpus %s, nvidia-smi topo)" % (numa, cpu_aff)
        return "numa %s (cpus %s, nvidia-smi topo; not narrower than the current affinity)" % (numa, cpu_aff)
    except Exception:
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
# the reference on the host
def reference_cpu(natoms, steps, warmup, threads, flavours=("NLIST", "NLISTCELLS"), drift=DRIFT, steps_cells=None):
    """the reference's CPU COORDINATION through plumed_cmd; returns the faster flavour's dict plus all flavours.

    NLIST is the keyword our arm uses; the reference can only run it below 32768 atoms and its rebuild is O(N^2).
    NLISTCELLS is the only flavour it can run above that (cost linear in N, but it sweeps the 27-cell superset)."""
    os.environ["PLUMED_NUM_THREADS"] = str(threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["PLUMED_IGNORE_NL_MEMORY_ERROR"] = "1"
    from oracle import refplumed as R
    from oracle import oracle as O
    if not R.available():
        return None
    n = natoms
    frames, box = make_frames(n, min(warmup + steps, 2 * NL_STRIDE), drift=drift)
    # numerator: pairs within NL_CUTOFF at the rebuild frame (what NLIST lists)
    nl = O.NeighborList(O.NL_SINGLELIST, n, 0, cutoff=NL_CUTOFF, stride=NL_STRIDE)
    nl.update(O.make_pbc(box), frames[0], fast=True)
    pairs = int(nl.size())
    best, out = None, {}
    for flavour in flavours:
        if flavour == "NLIST" and n > 32768:
            continue
        line = "c: COORDINATION GROUPA=1-%d SWITCH={%s} %s NL_CUTOFF=%r NL_STRIDE=%d" % (n, SWITCH, flavour, NL_CUTOFF, NL_STRIDE)
        p = R.Plumed(n, [line, "RESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1"])
        nst = steps_cells if (flavour == "NLISTCELLS" and steps_cells) else steps  # the superset flavour is ~8x slower
        for s in range(warmup):
            p.calc(s, frames[s % len(frames)], box)
        t0 = time.perf_counter()
        for s in range(warmup, warmup + nst):
            p.calc(s, frames[s % len(frames)], box)
        dt = time.perf_counter() - t0
        p.close()
        r = {"value": pairs * nst / dt, "ms_per_step": 1e3 * dt / nst, "pairs_per_step": pairs, "atoms": n,
             "keywords": line, "seconds": dt, "flavour": flavour, "steps": nst}
        out[flavour] = r
        if best is None or r["value"] > best["value"]:
            best = r
    best = dict(best)
    best["flavours"] = {k: v["value"] for k, v in out.items()}
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_atoms
    r = reference_cpu(n, args.steps, args.warmup, threads, flavours=("NLISTCELLS",) if n > 32768 else ("NLIST", "NLISTCELLS"))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference build) not present on this box"}))
        return
    sample = ("the whole %d-atom box of BASELINE configs[1] (%s; pair evals/s: %s), %d warm-up + %d timed steps incl. a list "
              "rebuild every %d.  The reference cannot run NLIST above 32768 atoms (SURVEY 9.5); NLISTCELLS costs the same "
              "per atom at any size, so this is also its rate at the 1 M atoms of the GPU arm" %
              (r["atoms"], r["flavour"], ", ".join("%s %.3g" % kv for kv in r["flavours"].items()), args.warmup, args.steps,
               NL_STRIDE))
    cfg = {"workload": "BASELINE configs[1] at its literal size: COORDINATION single group, %d atoms, orthorhombic PBC, "
                       "SWITCH={%s} %s NL_CUTOFF=%g NL_STRIDE=%d, 100 atoms/nm^3, drifting frames; reference CPU code on "
                       "%d host threads" % (r["atoms"], SWITCH, r["flavour"], NL_CUTOFF, NL_STRIDE, threads),
           "natoms": r["atoms"], "gpu_arm_natoms_per_gpu": args.natoms_per_gpu,
           "pair_count": "pairs within NL_CUTOFF at the last rebuild (NLIST size); both arms"}
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_keywords": r["keywords"], "pairs_per_step_sample": r["pairs_per_step"]}
    print(json.dumps(line))


def workload_config(args, world):
    n = args.natoms_per_gpu * world
    return {"workload": "BASELINE configs[1] at headline size: COORDINATION single group, %d atoms (%d per GPU), "
                        "orthorhombic PBC, SWITCH={%s} NLIST NL_CUTOFF=%g NL_STRIDE=%d, 100 atoms/nm^3"
                        % (n, args.natoms_per_gpu, SWITCH, NL_CUTOFF, NL_STRIDE),
            "natoms": n, "natoms_per_gpu": args.natoms_per_gpu,
            "parallelism": "i-atom shards x%d%s" % (world, "" if world == 1 else (", NCCL all-gather" if args.no_peer else ", NVLink peer memory")),
            "frames": "cumulative random walk, %.4f nm rms per coordinate and step, one new frame per step" % DRIFT,
            "pair_count": "pairs within NL_CUTOFF at the last rebuild (NLIST size); both arms",
            "cache": "per-step inputs (neighbour list %.1f GB + positions) exceed the 126 MB L2" %
                     (n / world * 419 * 4 / 1e9)}


# ------------------------------------------------------------------------------------------------
class Engine:
    """one context per rank + the plumbing the legs share"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, self.world))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
        torch.cuda.set_device(self.local)
        self.numa_node = pin_to_gpu_numa_node(self.local) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        import plumed2_b200 as P
        from plumed2_b200 import capi
        self.P, self.capi, self.L = P, capi, capi.lib()
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        self.capi.check(self.L.b200coord_device_synchronize())

    def allmax(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def context(self, line, env=None, sharded=True, **kw):
        """a Coordination context of this rank; env = A/B switches read at creation"""
        env = env or {}
        os.environ.update(env)
        try:
            c = self.P.Coordination.from_input(line, device=self.local, rank=self.rank if sharded else 0,
                                               nranks=self.world if sharded else 1, **kw)
        finally:
            for k in env:
                os.environ.pop(k)
        c.sharded = bool(sharded and self.world > 1)
        if sharded and self.world > 1:
            ids = [self.P.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(ids, src=0)
            c.comm_init(ids[0])
            if not self.args.no_peer:
                handles = [None] * self.world
                self.dist.all_gather_object(handles, c.peer_export())
                c.peer_attach(handles)
        return c

    def device_steps(self, c, frames, d_out, step, count, order=None):
        """enqueue `count` steps on the context's stream; frames[i] are device tensors; returns the next step number"""
        L, capi, ctx = self.L, self.capi, c._ctx
        F = len(frames)
        # several ranks: every rank keeps the derivatives of its own slice of the atoms (like the e2e call)
        call = L.b200coord_enqueue_device_distributed if (self.world > 1 and c.sharded) else L.b200coord_enqueue_device
        for i in range(count):
            c.prepare(step)
            f = frames[order[i] if order is not None else step % F]
            capi.check(call(ctx, C.c_void_p(f.data_ptr()), C.c_void_p(d_out.data_ptr())), ctx)
            step += 1
        return step

    def out_buffer(self, c, n):
        """device buffer for the result of a device-resident step: [derivatives | virial 9 | value]"""
        if self.world > 1 and c.sharded:
            sb, sc = C.c_uint(), C.c_uint()
            self.capi.check(self.L.b200coord_my_slice(c._ctx, C.byref(sb), C.byref(sc)))
            return self.torch.empty(3 * sc.value + 10, dtype=self.torch.float64, device="cuda")
        return self.torch.empty(3 * n + 10, dtype=self.torch.float64, device="cuda")

    def timed_device_block(self, c, frames, d_out, step, count, order=None):
        L, capi, ctx = self.L, self.capi, c._ctx
        capi.check(L.b200coord_stream_mark(ctx, 0), ctx)
        step = self.device_steps(c, frames, d_out, step, count, order)
        capi.check(L.b200coord_stream_mark(ctx, 1), ctx)
        ms = C.c_float(0)
        capi.check(L.b200coord_stream_elapsed_ms(ctx, C.byref(ms)), ctx)
        return step, float(ms.value)


def regime_run(E, line, box, frames, W, K, env=None, sampler=None, sustained=False):
    """W warm-up + K timed device-resident steps of a fresh context; dict with times and what the engine did"""
    torch = E.torch
    n = frames[0].shape[0]
    c = E.context(line, env=env)
    c._set_box(box)
    d_out = E.out_buffer(c, n)
    step = E.device_steps(c, frames, d_out, 0, W)
    E.barrier()
    st0 = c.stats()
    if sampler:
        sampler.open_window()
    step, ms = E.timed_device_block(c, frames, d_out, step, K)
    if sampler:
        sampler.close_window()
    E.barrier()
    st1 = c.stats()
    ms = E.allmax(ms)
    pairs = E.allsum(float(st1["nl_size"]))
    out = {"ms_per_step": ms / K, "value": pairs * K / (ms * 1e-3), "pairs_per_step": pairs,
           "sweep_ms": st1["sweep_ms_sum"] / max(1, st1["sweep_count"]),
           "rebuild_ms": st1["build_ms_sum"] / max(1, st1["build_count"]), "rebuilds": int(st1["build_count"]),
           "rebuild_ms_max": st1["build_ms_max"],
           "super_list_builds_total": int(st1["super_builds"]), "filter_rebuilds_total": int(st1["filter_rebuilds"]),
           "rebuilds_total": int(st1["rebuilds"]),
           "entries_evaluated_last_step": int(st1["pair_evals"]), "entries_listed": 2 * int(st1["nl_size"]),
           "gpu_launches": int(st1["kernel_launches"] - st0["kernel_launches"]), "my_pairs": float(st1["nl_size"])}
    if sustained:
        # the K-step block again and again (frames walked back and forth, so that consecutive frames stay one step
        # apart) until at least a second has been timed
        F = len(frames)
        fwd = [(W + i) % F for i in range(K)]
        total_ms, total_steps, flip = 0.0, 0, False
        if sampler:
            sampler.open_window()
        while total_ms < 1000.0 and total_steps < 200 * K:
            order = fwd[::-1] if not flip else fwd
            flip = not flip
            step, ms = E.timed_device_block(c, frames, d_out, step, K, order)
            total_ms += E.allmax(ms)
            total_steps += K
        if sampler:
            sampler.close_window()
        st2 = c.stats()
        out["sustained"] = {"ms_per_step": total_ms / total_steps, "value": pairs * total_steps / (total_ms * 1e-3),
                            "steps": total_steps, "seconds": total_ms * 1e-3,
                            "sweep_ms": st2["sweep_ms_sum"] / max(1, st2["sweep_count"])}
    out["cv_value"] = float(d_out[-1].item())
    return c, d_out, step, out


def e2e_run(E, c, frames, step, W, K, sampler):
    """the same steps through b200coord_calculate[_distributed] with pinned host buffers"""
    L, capi, ctx = E.L, E.capi, c._ctx
    sb, sc = C.c_uint(), C.c_uint()
    capi.check(L.b200coord_my_slice(ctx, C.byref(sb), C.byref(sc)))
    lo, cnt = sb.value, sc.value
    h_frames = []
    for f in frames:
        a, _ptr = pinned_array(L, (max(cnt, 1), 3))
        a[:cnt] = f[lo:lo + cnt].cpu().numpy()
        h_frames.append(a)
    h_deriv, _p2 = pinned_array(L, (max(cnt, 1), 3))
    vir = np.zeros(9)
    val = C.c_double(0)
    F = len(frames)

    def one(s):
        c.prepare(s)
        src = h_frames[s % F]
        if E.world > 1:
            capi.check(L.b200coord_calculate_distributed(ctx, src.ctypes.data_as(C.c_void_p), C.byref(val),
                                                         h_deriv.ctypes.data_as(C.c_void_p),
                                                         vir.ctypes.data_as(C.POINTER(C.c_double))), ctx)
        else:
            capi.check(L.b200coord_calculate(ctx, src.ctypes.data_as(C.c_void_p), C.byref(val),
                                             h_deriv.ctypes.data_as(C.c_void_p),
                                             vir.ctypes.data_as(C.POINTER(C.c_double))), ctx)

    for _ in range(W):
        one(step)
        step += 1
    E.barrier()
    sampler.open_window()
    ms = C.c_float(0)
    capi.check(L.b200coord_stream_mark(ctx, 0), ctx)
    t0 = time.perf_counter()
    for _ in range(K):
        one(step)
        step += 1
    capi.check(L.b200coord_stream_mark(ctx, 1), ctx)
    capi.check(L.b200coord_stream_elapsed_ms(ctx, C.byref(ms)), ctx)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.close_window()
    E.barrier()
    e2e_ms = E.allmax(max(float(ms.value), wall_ms))
    pairs = E.allsum(float(c.stats()["nl_size"]))
    # what the copies alone cost on this box: the same H2D + D2H per step on every rank at once, no kernels in between
    scratch = C.c_void_p()
    capi.check(L.b200coord_device_alloc(max(cnt, 1) * 24, C.byref(scratch)))
    E.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        capi.check(L.b200coord_memcpy_h2d(scratch, h_frames[i % F].ctypes.data_as(C.c_void_p), cnt * 24))
        capi.check(L.b200coord_memcpy_d2h(h_deriv.ctypes.data_as(C.c_void_p), scratch, cnt * 24))
    copy_ms = E.allmax(1e3 * (time.perf_counter() - t0)) / K
    E.barrier()
    L.b200coord_device_free(scratch)
    # frames known in advance (trajectory post-processing): b200coord_submit keeps two steps in flight, so the upload of
    # the next frame and the download of the previous result run under the sweep.  Reported beside e2e, never as e2e:
    # an MD engine cannot hand over frame k+1 before it has the forces of frame k.
    ahead = None
    if E.world == 1:
        h_d2, _p3 = pinned_array(L, (max(cnt, 1), 3))
        outs = [(C.c_double(0), h_deriv, np.zeros(9)), (C.c_double(0), h_d2, np.zeros(9))]

        def sub(s):
            c.prepare(s)
            v, d, w = outs[s & 1]
            capi.check(L.b200coord_submit(ctx, h_frames[s % F].ctypes.data_as(C.c_void_p), C.byref(v),
                                          d.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p)), ctx)

        for _ in range(W):
            sub(step)
            step += 1
        capi.check(L.b200coord_collect(ctx), ctx)
        t0 = time.perf_counter()
        for _ in range(K):
            sub(step)
            step += 1
        capi.check(L.b200coord_collect(ctx), ctx)
        ahead_ms = 1e3 * (time.perf_counter() - t0)
        ahead = {"ms_per_step": ahead_ms / K, "value": pairs * K / (ahead_ms * 1e-3), "unit": UNIT,
                 "api": "b200coord_submit / b200coord_collect (two steps in flight; same bytes per step as e2e)",
                 "cv_value": outs[(step - 1) & 1][0].value}
    res = {"value": pairs * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
           "frames_submitted_ahead": ahead,
           "copies_alone_ms_per_step": copy_ms,
           "copies_alone_note": "the same %d-byte upload + download per rank and step with nothing in between, all ranks at "
                                "once: the share of e2e that is the host <-> device link of this box" % (cnt * 24),
           "h2d_bytes_per_step": int(cnt * 24), "d2h_bytes_per_step": int(cnt * 24 + 80),
           "api": "b200coord_calculate_distributed" if E.world > 1 else "b200coord_calculate",
           "numa_node_of_rank0": E.numa_node}
    return step, res, (lo, cnt, h_frames, h_deriv, float(val.value), vir.copy())


def parity_check(E, c, line, box, frame, step, e2e_state):
    """N > 1: one more (untimed) distributed step against a single-GPU context on the same frame.  The frozen list of
    the distributed engine and the fresh list of the checker both hold every pair that can be inside D_MAX (the walk
    moves atoms by ~0.01 nm per list interval, the list reaches 0.2 nm beyond D_MAX), so the numbers must agree to
    rounding."""
    torch, dist, L, capi = E.torch, E.dist, E.L, E.capi
    lo, cnt, _h, h_deriv, _v, _w = e2e_state
    n = frame.shape[0]
    src, _p = pinned_array(L, (max(cnt, 1), 3))
    src[:cnt] = frame[lo:lo + cnt].cpu().numpy()
    vir = np.zeros(9)
    val = C.c_double(0)
    c.prepare(step * NL_STRIDE + 1)  # never a rebuild step
    capi.check(L.b200coord_calculate_distributed(c._ctx, src.ctypes.data_as(C.c_void_p), C.byref(val),
                                                 h_deriv.ctypes.data_as(C.c_void_p),
                                                 vir.ctypes.data_as(C.POINTER(C.c_double))), c._ctx)
    full = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    ref_tail = torch.zeros(10, dtype=torch.float64, device="cuda")
    if E.rank == 0:
        c1 = E.context(line, sharded=False)
        c1.prepare(0)
        c1.calculate(frame.cpu().numpy(), box)
        full.copy_(torch.from_numpy(np.ascontiguousarray(c1.derivatives)))
        ref_tail[:9] = torch.from_numpy(np.asarray(c1.virial).reshape(9))
        ref_tail[9] = c1.value
        c1.close()
    dist.broadcast(full, src=0)
    dist.broadcast(ref_tail, src=0)
    mine = torch.from_numpy(h_deriv[:cnt].copy()).cuda()
    scale = float(full.abs().max().item())
    err = float((mine - full[lo:lo + cnt]).abs().max().item()) / max(scale, 1e-300) if cnt else 0.0
    err = E.allmax(err)
    rt = ref_tail.cpu().numpy()
    return {"max_rel_derivatives": err, "rel_value": abs(val.value - rt[9]) / abs(rt[9]),
            "max_rel_virial": float(np.abs(vir - rt[:9]).max() / np.abs(rt[:9]).max()),
            "against": "single-GPU context of the same library on the same frame (rank 0), every rank compares its own slice",
            "tolerance": 1e-10}


def plumed_e2e(E, line, box, frames, W, K):
    """the same steps through the unmodified PlumedMain + LOAD plugin (1 GPU); wall clock around plumed_cmd("calc")"""
    import torch
    from oracle import refplumed as R
    plugin = os.path.join(os.path.dirname(E.capi.LIB_PATH), "libb200coord_plumed.so")
    if not (R.available() and os.path.exists(plugin)):
        return {"unavailable": "oracle/_ref or the plugin .so is not present on this box"}
    os.environ["PLUMED_NUM_THREADS"] = str(os.cpu_count() or 1)
    os.environ["PLUMED_IGNORE_NL_MEMORY_ERROR"] = "1"
    os.environ["B200COORD_DEVICE"] = str(E.local)
    os.environ["B200COORD_PIN_HOST"] = "1"
    n = frames[0].shape[0]
    host = [f.cpu().numpy() for f in frames]
    p = R.Plumed(n, ["LOAD FILE=" + plugin, "DEBUG DETAILED_TIMERS", line, "RESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1"],
                 log="/tmp/bench_plumed_e2e.log")
    F = len(host)
    # what an MD engine does every step: hand over its arrays (forces are ADDED to, so nothing is cleared) and call calc
    forces = np.zeros((n, 3))
    virial = np.zeros((3, 3))
    boxm = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(3, 3))

    def one(s):
        p.cmd("setStep", C.c_int(int(s)))
        p.cmd("setPositions", host[s % F])
        p.cmd("setMasses", p.masses)
        p.cmd("setCharges", p.charges)
        p.cmd("setBox", boxm)
        p.cmd("setForces", forces)
        p.cmd("setVirial", virial)
        p.cmd("calc", None)
        p._keep = p._keep[-64:]

    def timed(step_fn):
        for s in range(W):
            step_fn(s)
        t0 = time.perf_counter()
        for s in range(W, W + K):
            step_fn(s)
        return time.perf_counter() - t0

    def read_timers(logfile):
        timers = []
        try:
            for ln in open(logfile):
                if "PLUMED:" in ln and any(k in ln for k in ("Prepare", "Sharing", "Waiting", "Calculating", "Applying", "4A ", "5A ", "Update")):
                    timers.append(" ".join(ln.split()[1:]))
        except Exception:
            pass
        return timers[:40]

    dt = timed(one)
    p.close()
    out = {"ms_per_step": 1e3 * dt / K, "threads": os.cpu_count() or 1, "plumed_timers": read_timers("/tmp/bench_plumed_e2e.log"),
           "path": "plumed_cmd(setPositions..calc) -> PlumedMain -> CoordinationB200 (LOAD) -> libb200coord; forces and "
                   "virial returned to the caller every step; the action gathers positions / adds forces itself over OpenMP "
                   "threads (B200COORD_HOST_FAST) through page-locked buffers"}

    # the engine's positions and forces stay on the device (GPU_COUPLING); PLUMED is still handed host
    # arrays (zeros: nobody reads them) and does its own per-atom bookkeeping on them
    try:
        L = E.capi.lib()
        d_pos = torch.empty((n, 3), dtype=torch.float64, device=frames[0].device)
        d_force = torch.zeros((n, 3), dtype=torch.float64, device=frames[0].device)
        key = b"bench_engine"
        if L.b200coord_coupling_publish(key, E.local, C.c_void_p(d_pos.data_ptr()), C.c_void_p(d_force.data_ptr()), n) != 0:
            raise RuntimeError("coupling_publish failed")
        p = R.Plumed(n, ["LOAD FILE=" + plugin, "DEBUG DETAILED_TIMERS", line + " GPU_COUPLING=bench_engine",
                         "RESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1"], log="/tmp/bench_plumed_e2e_dev.log")
        hostpos = np.zeros((n, 3))

        def one_dev(s):
            d_pos.copy_(frames[s % F])  # the engine's integrator wrote new positions
            torch.cuda.current_stream().synchronize()
            p.cmd("setStep", C.c_int(int(s)))
            p.cmd("setPositions", hostpos)
            p.cmd("setMasses", p.masses)
            p.cmd("setCharges", p.charges)
            p.cmd("setBox", boxm)
            p.cmd("setForces", forces)
            p.cmd("setVirial", virial)
            p.cmd("calc", None)
            p._keep = p._keep[-64:]

        dt = timed(one_dev)
        p.close()
        L.b200coord_coupling_withdraw(key)
        out["device_coupled"] = {
            "ms_per_step": 1e3 * dt / K, "plumed_timers": read_timers("/tmp/bench_plumed_e2e_dev.log"),
            "force_checksum": float(d_force.abs().sum().item()),
            "path": "as above with GPU_COUPLING: positions read from and forces added to device arrays the caller "
                    "published (b200coord_coupling_publish); no per-atom data crosses the host link"}
    except Exception as ex:  # noqa: BLE001
        out["device_coupled"] = {"unavailable": repr(ex)[:200]}
    return out


# ------------------------------------------------------------------------------------------------
def _action_timers(logfile, labels):
    """per-action calculate() times from PLUMED's DETAILED_TIMERS table: label -> (cycles, avg s, min s)"""
    out = {}
    try:
        for ln in open(logfile):
            t = ln.split()
            if len(t) >= 9 and t[0] == "PLUMED:" and t[1] == "4A" and t[3] in labels:
                out[t[3]] = (int(t[4]), float(t[6]), float(t[7]))
    except Exception:
        pass
    return out


def cuda_baseline(E):
    """The reference's own CUDA prototype (plugins/cudaCoord, unchanged, compiled for sm_100a by oracle/Makefile) next to
    our plugin in ONE PLUMED input on the cases it supports (orthorhombic PBC, rational switch): both actions are
    driven by the same unmodified PlumedMain on the same frames, and PLUMED's own DETAILED_TIMERS give the time each
    spends in calculate() (host<->device copies included, PLUMED's shared host work excluded)."""
    from oracle import refplumed as R
    plugin = os.path.join(os.path.dirname(E.capi.LIB_PATH), "libb200coord_plumed.so")
    cudacoord = os.path.join(ROOT, "oracle", "_ref", "lib", "CudaCoordination.so")
    if not (R.available() and os.path.exists(plugin) and os.path.exists(cudacoord)):
        return {"unavailable": "oracle/_ref (with CudaCoordination.so) or the plugin .so is not present on this box"}
    os.environ["PLUMED_NUM_THREADS"] = str(os.cpu_count() or 1)
    os.environ["PLUMED_IGNORE_NL_MEMORY_ERROR"] = "1"
    os.environ["B200COORD_DEVICE"] = str(E.local)
    os.environ["B200COORD_PIN_HOST"] = "1"
    out = {"what": "calculate() per step from PLUMED's DETAILED_TIMERS, both actions in one input, %d threads" % (os.cpu_count() or 1)}
    for n in (100000, 1000000):
        W, K = 10, 30
        frames, box = make_frames(n, W + K, seed=SEED + 11, drift=DRIFT)
        log = "/tmp/bench_cuda_baseline_%d.log" % n
        lines = ["LOAD FILE=" + cudacoord, "LOAD FILE=" + plugin, "DEBUG DETAILED_TIMERS",
                 "ref: CUDACOORDINATION GROUPA=1-%d R_0=0.3 NN=6 MM=12 D_MAX=0.8 NL_CUTOFF=%r NL_STRIDE=%d" % (n, NL_CUTOFF, NL_STRIDE),
                 "c: COORDINATION GROUPA=1-%d SWITCH={%s} NLIST NL_CUTOFF=%r NL_STRIDE=%d" % (n, SWITCH, NL_CUTOFF, NL_STRIDE),
                 "RESTRAINT ARG=c,ref AT=0,0 KAPPA=0,0 SLOPE=1,1"]
        try:
            p = R.Plumed(n, lines, log=log, watch=("c", "ref"))
            for s in range(W + K):
                p.calc(s, frames[s], box)
            vc, vr = p.value("c"), p.value("ref")
            p.close()
            tm = _action_timers(log, ("c", "ref"))
            r = {"steps": W + K, "cv_ours": vc, "cv_cudacoord": vr, "rel_diff": abs(vc - vr) / abs(vr) if vr else None}
            if "c" in tm and "ref" in tm:
                r.update({"ours_ms_avg": 1e3 * tm["c"][1], "ours_ms_min": 1e3 * tm["c"][2],
                          "cudacoord_ms_avg": 1e3 * tm["ref"][1], "cudacoord_ms_min": 1e3 * tm["ref"][2],
                          "speedup_avg": tm["ref"][1] / tm["c"][1], "speedup_min": tm["ref"][2] / tm["c"][2]})
            out["%d atoms" % n] = r
        except Exception as e:  # cudaCoord refuses some sizes ("try by reducing the cell dimensions ...")
            out["%d atoms" % n] = {"failed": str(e)[:300]}
        del frames
    return out


def other_configs(E, peak_tflops, args):
    """BASELINE.json configs[0], [2], [3], [4]: a few device-resident steps each (jittering frames: these lines are
    here for coverage of the kernels they run, not for the displacement-conditioned shortcuts)"""
    torch = E.torch
    out = {}

    def run(name, line, n, box, flop, steps=10, warm=3, drift=None, note=None, sharded=False):
        frames, _ = make_frames_device(n, 4 if drift is None else steps + warm, drift, seed=SEED + 7, box=box)
        c = E.context(line, sharded=sharded)
        c._set_box(box)
        d_out = E.out_buffer(c, n)
        step = E.device_steps(c, frames, d_out, 0, warm)
        E.barrier()
        step, ms = E.timed_device_block(c, frames, d_out, step, steps)
        E.barrier()
        st = c.stats()
        ms = E.allmax(ms)
        pairs = E.allsum(float(st["nl_size"]))
        sweep = st["sweep_ms_sum"] / max(1, st["sweep_count"])
        r = {"keywords": line if len(line) < 200 else line[:200] + "...", "natoms": n, "ms_per_step": ms / steps,
             "pairs_per_step": pairs, "value": pairs * steps / (ms * 1e-3), "sweep_ms": sweep,
             "rebuild_ms": st["build_ms_sum"] / max(1, st["build_count"]), "rebuilds": int(st["build_count"]),
             "flop_per_pair": flop,
             "roofline_frac": (flop * float(st["nl_size"]) / (sweep * 1e-3) / 1e12 / peak_tflops) if sweep > 0 else None,
             "cv_value": float(d_out[-1].item())}
        if note:
            r["note"] = note
        c.close()
        del frames, d_out
        torch.cuda.empty_cache()
        out[name] = r

    only = getattr(args, "only_other", None)
    if E.world == 1 and only:
        # one configuration alone (profiling runs): configs[0] | configs[2]-NLIST | configs[2]-NLISTCELLS
        if only == "configs[0]":
            run("configs[0]", "c: COORDINATION GROUPA=1-1000 SWITCH={RATIONAL R_0=0.3 NN=6 MM=12}", 1000,
                np.diag([box_edge(1000)] * 3), 68.0, steps=20)
        else:
            na, nb = 10000, 1000000
            n = na + nb
            L2 = box_edge(n)
            tri = L2 * np.array([[1.0, 0.0, 0.0], [0.2, 1.0, 0.0], [0.1, 0.3, 1.0]])
            flavour = only.split("-")[1]
            run(only, "c: COORDINATION GROUPA=1-%d GROUPB=%d-%d SWITCH={EXP R_0=0.2 D_MAX=0.9} %s NL_CUTOFF=1.0 NL_STRIDE=1"
                % (na, na + 1, n, flavour), n, tri, 102.0)
        return out
    if E.world == 1:
        # configs[0]: plumed driver, 1k atoms, no NL
        n = 1000
        L0 = box_edge(n)
        run("configs[0]", "c: COORDINATION GROUPA=1-%d SWITCH={RATIONAL R_0=0.3 NN=6 MM=12}" % n, n, np.diag([L0] * 3), 68.0,
            steps=20)
        # configs[2]: 10k solute vs 1M solvent, triclinic, EXP, rebuild every step; both list flavours
        na, nb = 10000, 1000000
        n = na + nb
        L2 = box_edge(n)
        tri = L2 * np.array([[1.0, 0.0, 0.0], [0.2, 1.0, 0.0], [0.1, 0.3, 1.0]])
        for flavour in ("NLIST", "NLISTCELLS"):
            run("configs[2] " + flavour,
                "c: COORDINATION GROUPA=1-%d GROUPB=%d-%d SWITCH={EXP R_0=0.2 D_MAX=0.9} %s NL_CUTOFF=1.0 NL_STRIDE=1"
                % (na, na + 1, n, flavour), n, tri, 102.0,
                note="pairs = list size: NLIST pairs within NL_CUTOFF, NLISTCELLS the 27-cell superset the reference iterates")
        out["configs[3]"] = config3_metad(E)
    # configs[4]: 4M atoms, strong scaling over the ranks of this run
    n = 4000000
    L4 = box_edge(n, 33.4)
    frames_note = "4M atoms at 33.4 atoms/nm^3 (O-only water, SURVEY 8(d)); total work fixed, sharded over %d GPU(s)" % E.world
    run("configs[4]", "c: COORDINATION GROUPA=1-%d SWITCH={%s} NLIST NL_CUTOFF=%r NL_STRIDE=%d" % (n, SWITCH, NL_CUTOFF, NL_STRIDE),
        n, np.diag([L4] * 3), 68.0, steps=20, warm=5, note=frames_note, sharded=True)
    # the frames of configs[4] use density 100 positions scaled into the 33.4 box by make_frames_device(box=...)
    return out


def config3_metad(E):
    """configs[3]: COORDINATION (250k atoms) driving METAD with a GRID through plumed_cmd, forces + virial returned"""
    from oracle import refplumed as R
    plugin = os.path.join(os.path.dirname(E.capi.LIB_PATH), "libb200coord_plumed.so")
    if not (R.available() and os.path.exists(plugin)):
        return {"unavailable": "oracle/_ref or the plugin .so is not present on this box"}
    os.environ["PLUMED_NUM_THREADS"] = str(os.cpu_count() or 1)
    os.environ["PLUMED_IGNORE_NL_MEMORY_ERROR"] = "1"
    os.environ["B200COORD_DEVICE"] = str(E.local)
    os.environ["B200COORD_PIN_HOST"] = "1"
    n = 250000
    frames, box = make_frames(n, 25, seed=SEED + 3, drift=DRIFT)
    body = "GROUPA=1-%d SWITCH={%s} NLIST NL_CUTOFF=%r NL_STRIDE=%d" % (n, SWITCH, NL_CUTOFF, NL_STRIDE)
    c = E.context("c: COORDINATION " + body, sharded=False)
    c.prepare(0)
    v0 = c.calculate(frames[0], box)
    c.close()
    lines = ["LOAD FILE=" + plugin, "c: COORDINATION " + body,
             "METAD ARG=c SIGMA=%g HEIGHT=1.2 PACE=5 GRID_MIN=%g GRID_MAX=%g GRID_BIN=2000 FILE=/tmp/bench_HILLS" %
             (2e-4 * v0, 0.9 * v0, 1.1 * v0)]
    p = R.Plumed(n, lines, log="/tmp/bench_config3.log")
    W, K = 5, 20
    for s in range(W):
        p.calc(s, frames[s], box)
    t0 = time.perf_counter()
    for s in range(W, W + K):
        r = p.calc(s, frames[s], box)
    dt = time.perf_counter() - t0
    p.close()
    return {"keywords": lines[1][:120] + " + " + lines[2][:60] + "...", "natoms": n, "ms_per_step": 1e3 * dt / K,
            "path": "plumed_cmd -> PlumedMain -> CoordinationB200 (LOAD) + the reference's own METAD; wall clock",
            "bias_last_step": r["bias"], "max_abs_force_last_step": float(np.abs(r["forces"]).max()), "cv_first_frame": v0}


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    E = Engine(args)
    torch, capi, L = E.torch, E.capi, E.L
    world, rank, local = E.world, E.rank, E.local
    if args.only_other:
        print(json.dumps(other_configs(E, 34.2, args)))
        return
    n = args.natoms_per_gpu * world
    K, W = args.steps, args.warmup
    line = "c: COORDINATION GROUPA=1-%d SWITCH={%s} NLIST NL_CUTOFF=%r NL_STRIDE=%d" % (n, SWITCH, NL_CUTOFF, NL_STRIDE)

    peak = C.c_double(0)
    capi.check(L.b200coord_measure_fp64_peak(local, C.byref(peak)))

    # the first steady-state rebuild (a single pass into fixed-capacity rows) happens at step NL_STRIDE, and the first
    # launch of its kernels in a process costs up to 30 ms on a box whose page cache is cold (measured: the first bench
    # process on a fresh box only).  The untimed warm-up therefore always covers that step, whatever --warmup says.
    W_asked, W = W, max(W, NL_STRIDE + 1)
    frames, box = make_frames_device(n, W + K, DRIFT)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    E.barrier()

    # ---------------- value (typical regime) + sustained
    c, d_out, step, typ = regime_run(E, line, box, frames, W, K, sampler=sampler, sustained=True)
    # ---------------- e2e through the C-ABI call with host buffers, same context, same frames
    step = ((step + NL_STRIDE - 1) // NL_STRIDE) * NL_STRIDE  # the e2e leg starts on a rebuild step like the value leg
    step, e2e, e2e_state = e2e_run(E, c, frames, step, W, K, sampler)
    clocks = sampler.stop()
    parity = None
    if world > 1:
        parity = parity_check(E, c, line, box, frames[-1], step, e2e_state)
    c.close()
    del d_out
    torch.cuda.empty_cache()

    regimes = {"typical": {k: typ[k] for k in ("ms_per_step", "value", "sweep_ms", "rebuild_ms", "rebuilds",
                                                "super_list_builds_total", "filter_rebuilds_total", "rebuilds_total",
                                                "entries_evaluated_last_step", "entries_listed")}}
    regimes["typical"]["what"] = "drifting frames, engine defaults (the headline)"
    if world == 1 and not args.no_regimes:
        jit, _ = make_frames_device(n, 4, None)
        cb, _d, _s, best = regime_run(E, line, box, jit, W, K)
        cb.close()
        del jit, _d
        torch.cuda.empty_cache()
        cw, _d, _s, worst = regime_run(E, line, box, frames, W, K,
                                       env={"B200COORD_NO_FAR_SPLIT": "1", "B200COORD_NO_SUPERLIST": "1"})
        cw.close()
        del _d
        torch.cuda.empty_cache()
        keys = ("ms_per_step", "value", "sweep_ms", "rebuild_ms", "rebuilds", "super_list_builds_total",
                "filter_rebuilds_total", "entries_evaluated_last_step", "entries_listed")
        regimes["best"] = {k: best[k] for k in keys}
        regimes["best"]["what"] = "frames that jitter around one configuration: far parts never visited, every rebuild filters the super-list"
        regimes["worst"] = {k: worst[k] for k in keys}
        regimes["worst"]["what"] = "drifting frames, B200COORD_NO_FAR_SPLIT=1 B200COORD_NO_SUPERLIST=1: every listed pair evaluated every step, every rebuild a full cell scan"

    e2e_pl = None
    others = None
    if world == 1 and not args.no_plumed_e2e:
        e2e_pl = plumed_e2e(E, line, box, frames, W, K)
        if "ms_per_step" in e2e_pl:
            e2e_pl["value"] = typ["pairs_per_step"] / (e2e_pl["ms_per_step"] * 1e-3)
            e2e_pl["unit"] = UNIT
            dc = e2e_pl.get("device_coupled") or {}
            if "ms_per_step" in dc:
                dc["value"] = typ["pairs_per_step"] / (dc["ms_per_step"] * 1e-3)
                dc["unit"] = UNIT
    del frames
    torch.cuda.empty_cache()
    if not args.no_other_configs:
        others = other_configs(E, peak.value, args)
    cuda_base = None
    if world == 1 and not args.no_cuda_baseline:
        cuda_base = cuda_baseline(E)

    if rank == 0:
        sweep_ms = typ["sweep_ms"]
        my_pairs = typ["my_pairs"]
        achieved = FLOP_PER_PAIR * my_pairs / (sweep_ms * 1e-3) / 1e12 if sweep_ms > 0 else None
        traffic, traffic_src = None, None
        tf = os.path.join(ROOT, "profiles", "sweep_traffic.json")
        if os.path.exists(tf):
            try:
                tj = json.load(open(tf))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "recorded, not measured in this run: " + tj.get("source", "profiles/sweep_traffic.json")
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
        out = {"metric": METRIC, "value": typ["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_asked,
               "ms_per_step": typ["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": dict(workload_config(args, world), warmup_steps_run=W,
                              warmup_note="untimed warm-up = max(--warmup, NL_STRIDE + 1) steps: it always includes the "
                                          "first steady-state rebuild (first launch of its kernels in the process)"),
               "pairs_per_step": typ["pairs_per_step"], "cv_value": typ["cv_value"],
               "regimes": regimes, "sustained": typ.get("sustained"),
               "roofline": {"kernel": "k_sweep_img<rationalfix6, double> (image-mode list sweep)", "bound": "fp64",
                            "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                            "frac": (achieved / peak.value) if achieved else None,
                            "traffic": traffic, "traffic_source": traffic_src,
                            "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                            "algorithmic_flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": my_pairs,
                            "kernel_ms": sweep_ms,
                            "entries_evaluated_per_launch": typ["entries_evaluated_last_step"],
                            "entries_listed": typ["entries_listed"],
                            "hbm": {"achieved_gbs": list_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None,
                                    "peak_gbs": hbm_peak, "source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
               "rebuild_ms": typ["rebuild_ms"], "rebuilds_in_timed_region": typ["rebuilds"],
               "rebuild_kinds": {"super_list_builds_total": typ["super_list_builds_total"],
                                 "filter_rebuilds_total": typ["filter_rebuilds_total"], "rebuilds_total": typ["rebuilds_total"]},
               "clocks": clocks, "gpu_launches": typ["gpu_launches"], "e2e": e2e}
        if e2e_pl is not None:
            out["e2e_plumed"] = e2e_pl
        if parity is not None:
            out["parity_check"] = parity
        if others is not None:
            out["other_configs"] = others
        if cuda_base is not None:
            out["cuda_baseline"] = cuda_base
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r = reference_cpu(args.ref_sample_atoms, 12 * NL_STRIDE, 1, threads, steps_cells=NL_STRIDE)
            if r is not None:
                out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference",
                                       "sample": "%d-atom box, same density/keywords/frames, faster of NLIST / NLISTCELLS (%s; %s), "
                                                 "%d steps with a rebuild every %d, %.1f s.  `--impl reference` runs the "
                                                 "literal configs[1] size (100k atoms, NLISTCELLS)" %
                                                 (r["atoms"], r["flavour"],
                                                  ", ".join("%s %.3g" % kv for kv in r["flavours"].items()), r["steps"],
                                                  NL_STRIDE, r["seconds"])}
            else:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                                       "sample": "oracle/_ref not present on this box"}
        print(json.dumps(out))
    if world > 1:
        E.dist.barrier()
        E.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--natoms-per-gpu", type=int, default=1000000)
    ap.add_argument("--ref-atoms", type=int, default=100000, help="atoms of the --impl reference run (configs[1]: 100000)")
    ap.add_argument("--ref-sample-atoms", type=int, default=20000, help="atoms of the in-run cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="skip the best / worst regime runs")
    ap.add_argument("--no-plumed-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-cuda-baseline", action="store_true", help="skip the plugins/cudaCoord comparison")
    ap.add_argument("--only-other", default=None, help="profiling: run only this entry of other_configs and print it")
    ap.add_argument("--quick", action="store_true", help="only value + e2e (profiling runs)")
    ap.add_argument("--no-peer", action="store_true", help="combine with NCCL all-gather instead of NVLink peer memory")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.quick:
        args.no_cpu_baseline = args.no_regimes = args.no_plumed_e2e = args.no_other_configs = args.no_cuda_baseline = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

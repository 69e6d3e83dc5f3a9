#!/usr/bin/env python
"""bench.py -- rays/s of the per-ray trace loop (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this framework (CUDA, C-ABI)
  python bench.py --impl reference [...]                          the CPU arm (see below)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   N > 1, one rank per GPU
  python bench.py --single-process --gpus N                       one process drives N GPUs (ort_init(N))

A "step" is one pass of the hot path over one batch of synthetic rays: the ring-source phase of
BASELINE.json configs[1] (ring of point sources on the bottle surface, clearBottle-large +
planoConvex-f39.9mm + achromaticDoublet-f50.0mm, 785 nm), RAYS_PER_GPU rays per GPU per step
(weak scaling: every rank traces its own contiguous ray-index range of the job, rank 0
receives the ncclReduce'd image).  Rays are generated on the device by the counter-based
sources, so there are no input arrays: the only per-step traffic is the scene (H2D, kernel
parameters) and the 401x401 uint64 image + status histogram (D2H).

  value  = rays/s from CUDA events on the library's stream (image clear + trace kernels + NCCL
           reduce), summed over the K steps, max over ranks
  e2e    = rays/s from the host clock around the same K `ort_trace` calls with HOST buffers
           (scene H2D, image D2H inside), barrier + device synchronize on both sides
  roofline: FP64 ALU.  achieved = sum over final ray statuses of count x algorithmic flops
           (SURVEY.md 8(d) stage table) / device time; peak = DFMA micro-kernel
           measured in this run (MEASURED_PEAKS.json has no FP64 entry).  For the ring loop the
           dominant kernel (ort_ring_cull_kernel) decides 99 % of the rays in integer + fp32
           arithmetic and issues no fp64 at all, so the algorithmic-fp64 fraction says how much
           of the reference's fp64 work per second is being REPLACED, not how busy the FP64 pipe
           is; `roofline.issue` adds the bound that kernel actually runs against (warp
           instructions issued / issue slots, instruction count from the committed ncu profile).
  extra:   the same measurement (device-timed, own warm-up, NOT part of `value`) for the other
           loop and the other BASELINE.json configs: the point loop of config 2, config 1 both
           loops, config 3 (15 scenes in one call) both loops, config 4 (scatter) faithful and with
           the outer-ellipse fix, the fp32 variant, the ring loop with the fp32 filter switched off
           (all rays past L2's aperture in fp64), and config 5 as written -- 1e11 ring rays in
           total, split over the ranks (strong scaling).
  cpu_baseline / --impl reference: the reference is Fortran and no Fortran compiler exists in
           this image, so the CPU arm is the C++/OpenMP oracle built with the reference's own
           optimisation flags (oracle/Makefile: -O2 -march=native -flto -mavx -fopenmp; kind "port")
           on all host cores.  If oracle/_ref/raytrace exists (a box with gfortran: `make -C oracle
           ref`) the reference binary itself is timed instead (kind "reference").

No torch anywhere: ranks started by torchrun find each other through files in /tmp (one node).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# BASELINE.json configs (index = position in its `configs` list + 1); config 2 / ring is the
# headline (configs[1]); the others are reported under `extra` and exercised by the parity tests
CONFIGS = {
    1: dict(bottles=["clearBottle-small.params"], l2="planoConvex.params", l3="achromaticDoublet.params",
            text="clearBottle-small -> planoConvex -> achromaticDoublet"),
    2: dict(bottles=["clearBottle-large.params"], l2="planoConvex-f39.9mm.params",
            l3="achromaticDoublet-f50.0mm.params",
            text="clearBottle-large -> planoConvex-f39.9mm -> achromaticDoublet-f50.0mm"),
    3: dict(bottles=["clearBottle-large_%dmm.params" % mm for mm in range(-14, 15, 2)],
            l2="planoConvex-f39.9mm.params", l3="achromaticDoublet-f50.0mm.params",
            text="15 clearBottle-large offset files (-14..+14 mm) in one ort_trace call -> "
                 "planoConvex-f39.9mm -> achromaticDoublet-f50.0mm"),
    4: dict(bottles=["scatterBottle-ellipse-long.params"], l2="planoConvex-f39.9mm.params",
            l3="achromaticDoublet-f50.0mm.params",
            text="scatterBottle-ellipse-long (mu_a 1, mu_s 30 1/m contents: tauint + stokes) -> "
                 "planoConvex-f39.9mm -> achromaticDoublet-f50.0mm"),
}
PHASES = {"ring": 1, "point": 2}


def workload_text(cfg, phase, flat=False, fix=False):
    src = ("ring source on the bottle surface, 785 nm" if phase == "ring"
           else "point source at the bottle centre, lenses at 843 nm")
    return "config%d-%s: %s; %s -> 401x401 detector%s%s" % (
        cfg, phase, src, CONFIGS[cfg]["text"], " [fixed outer ellipse]" if fix else "",
        " [no compaction]" if flat else "")


WORKLOAD = workload_text(2, "ring")
FILES = (CONFIGS[2]["bottles"][0], CONFIGS[2]["l2"], CONFIGS[2]["l3"])
RAYS_PER_GPU = 1 << 34
CPU_SAMPLE = 30_000_000
PROFILE_TAG = "r02"     # profiles/<tag>_{ring,point}_full.txt: the committed ncu summaries


# Algorithmic fp64 flops by final status (SURVEY.md 8(d) convention: + - * / sqrt and libm calls
# count 1, as written in the reference).  Stage table: ring 51 | point 11, bottle wall 98 each,
# L2 flat: 14 to the aperture test + 56, L2 curved 104 (30 of it the intersection), L3 s1 108,
# s2 104, s3 104, iris 14, plane move 9, makeImage 37; one scatter event = tauint 28 + stokes 60.
SCATTER_EVENT_FLOPS = 88.0


def flops_by_status(phase, use_bottle=True, iris_before=False, iris_after=False):
    f = np.zeros(32)
    src = 51.0 if phase == 1 else 11.0
    b_in = b_out = 0.0
    if phase == 2 and use_bottle:
        f[1] = src + 24                 # inner wall miss: intersection only
        f[2] = f[3] = f[24] = src + 24  # scatter loop endings (+ SCATTER_EVENT_FLOPS per counted event)
        f[6] = f[7] = src + 98 + 24
        f[4] = src + 98                 # reflected at the inner wall
        f[5] = src + 98 + 24
        f[8] = src + 196
        b_in, b_out = 98.0, 98.0
    s0 = src + b_in + b_out
    f[9] = s0 + 14
    f[10] = s0 + 14 + 56 + 30
    f[11] = s0 + 14 + 56 + 104
    l2 = s0 + 174
    ib = 14.0 if iris_before else 0.0
    f[12] = l2 + ib
    f[13] = l2 + ib + 30
    f[14] = l2 + ib + 40
    f[15] = l2 + ib + 108
    f[16] = l2 + ib + 108 + 30
    f[17] = l2 + ib + 212
    f[18] = l2 + ib + 212 + 30
    f[19] = l2 + ib + 316
    l3 = l2 + ib + 316
    ia = 14.0 if iris_after else 0.0
    f[20] = l3 + ia
    f[21] = l3 + ia + 9 + 20
    f[22] = f[23] = f[0] = l3 + ia + 9 + 37
    return f


def _profile(phase_name):
    for tag in (PROFILE_TAG, "r01"):
        path = os.path.join(ROOT, "profiles", "%s_%s_full.txt" % (tag, phase_name))
        if os.path.exists(path):
            return path
    return None


def ncu_inst_per_ray(phase_name):
    """warp instructions per ray of the dominant kernel, from the committed ncu summary:
    smsp__inst_executed.sum / rays of the profiled launch (tools/gpu_profile.sh uses 2^27)."""
    path = _profile(phase_name)
    if path:
        for line in open(path):
            f = line.split()
            if len(f) >= 2 and f[0] == "smsp__inst_executed.sum":
                return float(f[1]) / float(1 << 27), os.path.relpath(path, ROOT)
    return None, None


def ncu_traffic(phase_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one trace-kernel launch, from the committed
    ncu --set full summary; None when absent."""
    path = _profile(phase_name)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, False
    if path:
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[1]) * unit.get(f[2], 1.0)
                seen = True
    return total if seen else None


def issue_bound(phase_name, total_rays, dev_s, ngpus, clocks):
    """The bound the loops actually run against: warp instructions issued per second over the
    issue slots available (148 SMs x 4 schedulers x SM clock under load)."""
    ipr, src = ncu_inst_per_ray(phase_name)
    mhz = (clocks or {}).get("sm_mhz")
    if not ipr or not mhz:
        return None
    achieved = total_rays * ipr / dev_s
    peak = 148 * 4 * mhz * 1e6 * ngpus
    return {"warp_inst_per_ray": ipr, "achieved_inst_per_s": achieved, "peak_inst_per_s": peak,
            "frac": achieved / peak, "source": "smsp__inst_executed.sum of %s / 2^27 rays" % src}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def stop(self, window=None):
        """`window` = (t0, t1) of the timed region (perf_counter): only the samples taken inside it count
        (the sampler is started before the warm-up steps, because nvidia-smi needs ~0.3 s to deliver its
        first line and a timed region of K = 5 steps is only half a second long)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r[1:] for r in self.rows if window is None or window[0] <= r[0] <= window[1] + 0.1]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class Rendezvous:
    """Barrier / broadcast / all-gather between the ranks torchrun started on this node, through files
    in a directory named after the launch (MASTER_PORT + the launching agent's pid).  Replaces
    torch.distributed, which this benchmark no longer imports (PyTorch is not part of the product)."""

    def __init__(self, rank, world):
        self.rank, self.world, self.seq = rank, world, 0
        self.dir = None
        if world > 1:
            key = "%s_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "x"),
                                os.getppid())
            self.dir = os.path.join(tempfile.gettempdir(), "ort_bench_" + key)
            os.makedirs(self.dir, exist_ok=True)

    def _put(self, name, data):
        tmp = os.path.join(self.dir, name + ".tmp%d" % self.rank)
        with open(tmp, "wb") as f:
            f.write(data)
        os.rename(tmp, os.path.join(self.dir, name))

    def _get(self, name, timeout=120.0):
        path, t0 = os.path.join(self.dir, name), time.time()
        while not os.path.exists(path):
            if time.time() - t0 > timeout:
                raise SystemExit("bench.py: rank %d timed out waiting for %s" % (self.rank, name))
            time.sleep(0.0005)
        with open(path, "rb") as f:
            return f.read()

    def allgather(self, obj):
        """every rank's JSON-able `obj`, in rank order (also the barrier)"""
        if self.world == 1:
            return [obj]
        self.seq += 1
        self._put("g%d.%d" % (self.seq, self.rank), json.dumps(obj).encode())
        return [json.loads(self._get("g%d.%d" % (self.seq, r))) for r in range(self.world)]

    def barrier(self):
        self.allgather(0)

    def bcast(self, data):
        """bytes from rank 0 to everyone"""
        if self.world == 1:
            return data
        self.seq += 1
        if self.rank == 0:
            self._put("b%d" % self.seq, data)
        return self._get("b%d" % self.seq)

    def close(self):
        """last barrier, then rank 0 removes the directory -- only after every other rank has said that it
        is done reading"""
        if self.dir:
            self.barrier()
            if self.rank != 0:
                self._put("bye.%d" % self.rank, b"")
            else:
                for r in range(1, self.world):
                    self._get("bye.%d" % r, timeout=60.0)
                shutil.rmtree(self.dir, ignore_errors=True)


# ---------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------
CPU_FLAGS = "g++ -O2 -march=native -flto -mavx -fopenmp (the reference's src/Makefile flags)"


def cpu_arm(nrays, threads, first_ray=0):
    """Times the CPU oracle (C++/OpenMP restatement of the Fortran path, built with the reference's
    own optimisation flags) on `nrays` rays of the headline workload."""
    from opticalraytrace_b200 import abi
    from tests import oracle_lib as O
    scene = O.make_scene(*FILES)
    job = abi.default_job(abi.PHASE_RING, nrays, first_ray=first_ray)
    t0 = time.perf_counter()
    O.trace(job, scene, nthreads=threads, fast=True)
    return nrays / (time.perf_counter() - t0)


def reference_binary_arm(nphotons, threads):
    """oracle/_ref/raytrace (the UNMODIFIED reference, only where a Fortran compiler built it: `make -C
    oracle ref`): one run of its two ray loops on the headline configuration from a scratch copy of the
    drop-in directory layout.  -> rays/s over both loops, or None when the binary does not exist."""
    exe = os.path.join(ROOT, "oracle", "_ref", "raytrace")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        for d in ("bin", "data"):
            os.makedirs(os.path.join(tmp, d))
        shutil.copytree(os.path.join(ROOT, "res"), os.path.join(tmp, "res"))
        lines = ["0.5d-3", "785d-9", str(int(nphotons)), "5.0", "1.45", ".true.", ".false.", ".false.", "1.d-2",
                 "0.0", "point", "none", "1.0", FILES[0], FILES[1], FILES[2], "bessel-normal.dat", "bench",
                 "1.5d-3", "1.d-3"]
        open(os.path.join(tmp, "res", "bench.params"), "w").write("\n".join(lines) + "\n")
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        t0 = time.perf_counter()
        p = subprocess.run([exe, "bench.params"], cwd=os.path.join(tmp, "bin"), env=env, capture_output=True)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            return None
    return 2.0 * nphotons / dt


def cpu_baseline(sample, cores):
    ref = reference_binary_arm(min(sample, 2_000_000_000), cores)
    if ref:
        return {"value": ref, "unit": "rays/s", "cores": cores, "kind": "reference",
                "sample": "oracle/_ref/raytrace (gfortran build of the unmodified reference), %d rays per loop, both "
                          "loops, OMP_NUM_THREADS=%d" % (sample, cores)}
    cpu_arm(sample // 10, cores)
    v = cpu_arm(sample, cores, first_ray=1 << 40)
    return {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d rays of the same workload, C++/OpenMP restatement of the Fortran path (oracle/), %s, "
                      "%d threads; no Fortran compiler in this image (command -v gfortran fails)" % (sample, CPU_FLAGS, cores)}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = CPU_SAMPLE
    ref = reference_binary_arm(sample, cores)
    kind = "reference" if ref else "port"
    if ref:
        for _ in range(args.warmup):
            reference_binary_arm(sample // 10, cores)
        t0 = time.perf_counter()
        for k in range(args.steps):
            reference_binary_arm(sample, cores)
        dt = time.perf_counter() - t0
        v = args.steps * 2.0 * sample / dt
        what = ("oracle/_ref/raytrace (gfortran build of the unmodified reference), both loops, %d rays per loop and "
                "step, OMP_NUM_THREADS=%d" % (sample, cores))
    else:
        from tests import oracle_lib as O
        O.fast_lib()
        for _ in range(args.warmup):
            cpu_arm(sample // 10, cores)
        t0 = time.perf_counter()
        for k in range(args.steps):
            cpu_arm(sample, cores, first_ray=k * sample)
        dt = time.perf_counter() - t0
        v = args.steps * sample / dt
        what = ("%d rays per step, C++/OpenMP restatement of the Fortran path, %s, %d threads (no Fortran compiler "
                "in this image)" % (sample, CPU_FLAGS, cores))
    line = {
        "impl": "reference", "metric": "rays/sec", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": sample,
                   "note": "a throughput metric: the CPU arm runs a bounded sample of the workload per step "
                           "(3e7 rays against 2^34 per GPU in the b200 arm)"},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "sample": what},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args, rank, world, local):
        from opticalraytrace_b200 import abi, lib
        self.abi, self.lib, self.args = abi, lib, args
        self.rank, self.world = rank, world
        self.rdv = Rendezvous(rank, world)
        if lib.device_count() <= 0:
            raise SystemExit("bench.py: no CUDA device -- the trace loop has no CPU fallback")
        # stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...", printed to stdout when the
        # environment sets NCCL_DEBUG) goes to stderr with everything else a library may say while it starts
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            if args.single_process:
                self.ngpus = lib.init(args.gpus)
            else:
                nccl_id = None
                if world > 1:
                    nccl_id = self.rdv.bcast(lib.nccl_unique_id() if rank == 0 else b"")
                lib.init_rank(local, rank, world, nccl_id)
                self.ngpus = world
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        self.local = local

    def sync(self):
        self.rdv.barrier()
        self.lib.synchronize()

    def scenes(self, cfgid, phase):
        cfg, lib = CONFIGS[cfgid], self.lib
        out = []
        for bottle in cfg["bottles"]:
            st = lib.make_settings(bottle, cfg["l2"], cfg["l3"], nphotons=1000)
            # the point loop runs with the lenses re-built at 843 nm (reference src/main.f90:113-117)
            out.append(lib.build_scene(st, os.path.join(ROOT, "res"), None if phase == 1 else 843e-9)[0])
        return st, out

    def measure(self, cfgid, phase_name, rays, steps, warmup, *, precision=64, flags=0, strong=False, sampler=None):
        """`steps` timed ort_trace calls of `rays` rays per process and step (strong: `rays` in total, split
        over the ranks).  -> dict on rank 0, None elsewhere"""
        abi, lib = self.abi, self.lib
        phase = PHASES[phase_name]
        st, scenes = self.scenes(cfgid, phase)
        nsc = len(scenes)
        procs = self.world
        devs = self.ngpus if self.args.single_process else 1   # devices this process drives
        if strong:      # `rays` per step in total: this process takes its slice of the ray-index range
            lo, hi = rays * self.rank // procs, rays * (self.rank + 1) // procs
            nr = (hi - lo) // nsc
        else:           # `rays` per GPU and step (ort_init(N) splits a job's range over its N devices itself)
            lo, nr = 0, (rays // nsc) * devs
        mine = nr * nsc

        def step(k):
            job = lib.job_from_settings(st, phase)
            job.first_ray = (lo + k * rays) if strong else (k * procs + self.rank) * nr
            job.nrays = nr
            job.precision = precision
            job.flags |= flags
            return lib.trace(job, scenes, want_image=True)

        if sampler:
            sampler.start()
        for k in range(warmup):
            step(k)
        self.sync()
        t0 = time.perf_counter()
        dev_s = red_s = d2h_s = 0.0
        launches = h2d = d2h = 0
        hist = np.zeros(32, dtype=np.int64)
        for k in range(steps):
            _, _, h, tm = step(warmup + k)
            dev_s += tm.trace_seconds + tm.reduce_seconds
            red_s += tm.reduce_seconds
            d2h_s += tm.d2h_seconds
            launches += tm.kernel_launches
            h2d, d2h = tm.h2d_bytes, tm.d2h_bytes
            hist += h.sum(axis=0)
        self.sync()
        wall_s = time.perf_counter() - t0
        clocks = sampler.stop((t0, t0 + wall_s)) if sampler else None
        every = self.rdv.allgather([dev_s, wall_s, red_s, int(launches), int(mine)])
        if self.rank != 0:
            return None
        dev_s, wall_s, red_s = (max(e[i] for e in every) for i in range(3))
        launches = sum(e[3] for e in every)
        per_step = sum(e[4] for e in every)
        total = float(per_step) * steps
        # rank 0 / device 0 is the reduce root: its histogram holds the counts of every rank
        events = float(hist[abi.SCATTER_EVENTS_SLOT]) if hasattr(abi, "SCATTER_EVENTS_SLOT") else 0.0
        flops = float((flops_by_status(phase) * hist).sum()) + SCATTER_EVENT_FLOPS * events
        return dict(total=total, dev_s=dev_s, wall_s=wall_s, red_s=red_s, d2h_s=d2h_s, launches=launches, h2d=int(h2d),
                    d2h=int(d2h), hist=hist, flops=flops, clocks=clocks, nsc=nsc, rays_per_step=per_step,
                    scatter_events=events)

    def extra(self, name, cfgid, phase_name, rays, steps=2, **kw):
        r = self.measure(cfgid, phase_name, rays, steps, 1, **kw)
        if r is None:
            return None
        tf = r["flops"] / r["dev_s"] * 1e-12
        out = {"workload": workload_text(cfgid, phase_name, fix=bool(kw.get("flags", 0) & self.abi.FLAG_FIX_OUTER_ELLIPSE)),
               "value": r["total"] / r["dev_s"], "unit": "rays/s", "rays_per_step": r["rays_per_step"], "steps": steps,
               "ms_per_step": r["dev_s"] / steps * 1e3, "e2e": r["total"] / r["wall_s"],
               "gpu_launches": r["launches"], "dtype": "f%d" % kw.get("precision", 64),
               "roofline": {"achieved": tf, "unit": "TFLOP/s", "frac": tf / (self.peak_tf * self.ngpus),
                            "flops_per_launched_ray": r["flops"] / r["total"]},
               "binned_fraction": float(r["hist"][0]) / r["total"]}
        if r["scatter_events"]:
            out["scatter_events_per_ray"] = r["scatter_events"] / r["total"]
        if kw.get("strong"):
            out["scaling"] = "strong"
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU, help="rays per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other loops / configs (`extra`)")
    ap.add_argument("--phase", default="ring", choices=["ring", "point"],
                    help="ring = BASELINE.json configs[1] (the headline); point = the other loop")
    ap.add_argument("--flat", action="store_true", help="diagnostic: kernel without compaction")
    ap.add_argument("--no-filter", action="store_true", help="ring loop without the fp32 culling filter")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4],
                    help="BASELINE.json config (2 = the headline)")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32],
                    help="64 = the reference's arithmetic (headline); 32 = the fp32 variant (1e-5 parity)")
    ap.add_argument("--fix-ellipse", action="store_true",
                    help="config 4: opt-in outer-ellipse fix instead of the reference's half radii")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --total-rays per step in total, split over the ranks (BASELINE.json config 5)")
    ap.add_argument("--total-rays", type=float, default=1e11)
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives --gpus devices (ort_init(N) + ncclCommInitAll: what install.sh -n N uses)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.single_process and world > 1:
        raise SystemExit("bench.py: --single-process is one process; do not start it under torchrun")

    B = Bench(args, rank, world, local)
    abi, lib = B.abi, B.lib
    flags = 0
    if args.flat:
        flags |= abi.FLAG_NO_COMPACTION
    if args.fix_ellipse:
        flags |= abi.FLAG_FIX_OUTER_ELLIPSE
    if args.no_filter:
        flags |= abi.FLAG_NO_FILTER
    headline = (args.config == 2 and args.phase == "ring" and args.precision == 64 and not flags)
    strong = args.scaling == "strong"

    B.peak_tf, peak_mhz = lib.measure_fp64_peak()
    sampler = ClockSampler(local) if rank == 0 else None
    rays = int(args.total_rays) if strong else args.rays
    r = B.measure(args.config, args.phase, rays, args.steps, args.warmup, precision=args.precision, flags=flags,
                  strong=strong, sampler=sampler)

    extra = {}
    if headline and not strong and not args.no_extra and args.rays == RAYS_PER_GPU:
        todo = [
            ("config2_point", 2, "point", 1 << 32, {}),
            ("config1_ring", 1, "ring", 1 << 33, {}),
            ("config1_point", 1, "point", 1 << 32, {}),
            ("config3_ring", 3, "ring", 15 << 29, {}),
            ("config3_point", 3, "point", 15 << 28, {}),
            ("config4_point", 4, "point", 1 << 30, {}),
            ("config4_point_fixed_outer_ellipse", 4, "point", 1 << 30, dict(flags=abi.FLAG_FIX_OUTER_ELLIPSE)),
            ("config2_ring_fp32", 2, "ring", 1 << 33, dict(precision=32)),
            ("config2_point_fp32", 2, "point", 1 << 32, dict(precision=32)),
            ("config2_ring_fp64_only", 2, "ring", 1 << 33, dict(flags=abi.FLAG_NO_FILTER)),
            ("config5_strong_1e11", 2, "ring", int(1e11), dict(strong=True)),
        ]
        for name, cfgid, ph, n, kw in todo:
            steps = 1 if name.startswith("config5") else 2
            e = B.extra(name, cfgid, ph, n, steps=steps, **kw)
            if e is not None:
                extra[name] = e

    if rank == 0:
        ng = B.ngpus
        achieved = r["flops"] / r["dev_s"] * 1e-12
        cpu = None
        if ng == 1 and not args.no_cpu and headline:
            cpu = cpu_baseline(CPU_SAMPLE * 8, os.cpu_count() or 1)
        profiled = args.config == 2 and args.precision == 64 and not flags
        line = {
            "metric": "rays/sec", "value": r["total"] / r["dev_s"], "unit": "rays/s",
            "n_gpus": ng, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["dev_s"] / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f%d" % args.precision, "data": "synthetic",
            "config": {"workload": workload_text(args.config, args.phase, args.flat, args.fix_ellipse)
                                   + (" [fp32 filter off]" if args.no_filter else ""),
                       "rays_per_gpu_per_step": r["rays_per_step"] // ng, "scenes": r["nsc"],
                       "rays_per_step": r["rays_per_step"],
                       "parallelism": ("one process, ray-range x%d devices" if args.single_process
                                       else "one process per GPU, ray-range x%d") % ng,
                       "l2": "no input arrays (rays are generated on the device); the 2.6 MB "
                             "image buffer is re-zeroed every step",
                       "seed": 123456789},
            "e2e": {"value": r["total"] / r["wall_s"], "unit": "rays/s",
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": int(r["launches"]),
            "reduce_ms_per_step": r["red_s"] / args.steps * 1e3,
            # there is no flush kernel (hits go to the L2-resident image with RED): the only image
            # traffic is this device-to-host read-back of image + histogram on rank 0
            "image_readback": {"bytes_per_step": r["d2h"], "ms_per_step": r["d2h_s"] / args.steps * 1e3,
                               "GB/s": (r["d2h"] * args.steps / r["d2h_s"] * 1e-9) if r["d2h_s"] > 0 else None},
            "clocks": r["clocks"],
            "roofline": {"bound": "alu_fp64", "achieved": achieved, "peak": B.peak_tf * ng,
                         "unit": "TFLOP/s", "frac": achieved / (B.peak_tf * ng) if B.peak_tf else None,
                         "traffic": ncu_traffic(args.phase) if profiled else None,
                         "traffic_note": "DRAM bytes of one ncu-profiled launch of 2^27 rays "
                                         "(profiles/<round>_<loop>_full.txt): the loops have no per-ray memory "
                                         "traffic, the image stays in L2",
                         "peak_source": "DFMA micro-kernel measured in this run on rank 0 at %.0f MHz, "
                                        "x n_gpus (MEASURED_PEAKS.json holds no FP64 figure)" % peak_mhz,
                         "flops_per_launched_ray": r["flops"] / r["total"],
                         "issue": issue_bound(args.phase, r["total"], r["dev_s"], ng, r["clocks"]) if profiled else None},
            "cpu_baseline": cpu,
            "status_fractions": {abi.STATUS_NAMES[i]: r["hist"][i] / r["total"]
                                 for i in range(26) if r["hist"][i]},
        }
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    B.rdv.close()
    lib.finalize()


if __name__ == "__main__":
    main()

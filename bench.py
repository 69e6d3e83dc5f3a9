#!/usr/bin/env python
"""bench.py -- rays/s of the per-ray trace loop (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this framework (CUDA, C-ABI)
  python bench.py --impl reference [...]                          the CPU arm (see below)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   N > 1

A "step" is one pass of the hot path over one batch of synthetic rays: the ring-source phase of
BASELINE.json configs[1] (ring of point sources on the bottle surface, clearBottle-large +
planoConvex-f39.9mm + achromaticDoublet-f50.0mm, 785 nm), RAYS_PER_GPU rays per GPU per step
(weak scaling: every rank traces its own contiguous ray-index range of the job, rank 0
receives the ncclReduce'd image).  Rays are generated on the device by the counter-based
sources, so there are no input arrays: the only per-step traffic is the scene (H2D) and the
401x401 uint64 image + status histogram (D2H).

  value  = rays/s from CUDA events on the library's stream (image clear + trace kernel + NCCL
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
  cpu_baseline / --impl reference: the reference is Fortran and no Fortran compiler exists in
           this image, so the CPU arm is the C++/OpenMP oracle (kind "port": statement-by-
           statement restatement of the Fortran path) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# BASELINE.json configs (index = position in its `configs` list + 1); config 2 / ring is the
# headline (configs[1]); the others are reported in BASELINE.md and exercised by the parity tests
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

# Algorithmic fp64 flops by final status (SURVEY.md 8(d) convention: + - * / sqrt and libm calls
# count 1, as written in the reference).  Stage table: ring 51 | point 11, bottle wall 98 each,
# L2 flat: 14 to the aperture test + 56, L2 curved 104 (30 of it the intersection), L3 s1 108,
# s2 104, s3 104, iris 14, plane move 9, makeImage 37.
def flops_by_status(phase, use_bottle=True, iris_before=False, iris_after=False):
    f = np.zeros(32)
    src = 51.0 if phase == 1 else 11.0
    b_in = b_out = 0.0
    if phase == 2 and use_bottle:
        f[1] = src + 24                 # inner wall miss: intersection only
        f[2] = f[3] = f[24] = src + 24  # scatter loop endings: lower bound (tauint ~28 and stokes
        f[6] = f[7] = src + 98 + 24     #   ~60 flops per scatter event are not counted)
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


def ncu_inst_per_ray(phase_name):
    """warp instructions per ray of the dominant kernel, from the committed ncu summary:
    smsp__inst_executed.sum / rays of the profiled launch (tools/gpu_profile.sh uses 2^27)."""
    path = os.path.join(ROOT, "profiles", "r01_%s_full.txt" % phase_name)
    try:
        for line in open(path):
            f = line.split()
            if len(f) >= 2 and f[0] == "smsp__inst_executed.sum":
                return float(f[1]) / float(1 << 27)
    except OSError:
        pass
    return None


def ncu_traffic(phase_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one trace-kernel launch, from the committed
    ncu --set full summary (profiles/r01_<loop>_full.txt); None when absent."""
    path = os.path.join(ROOT, "profiles", "r01_%s_full.txt" % phase_name)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, False
    try:
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[1]) * unit.get(f[2], 1.0)
                seen = True
    except OSError:
        return None
    return total if seen else None


def issue_bound(args, total_rays, dev_s, world, clocks):
    """The bound the loops actually run against: warp instructions issued per second over the
    issue slots available (148 SMs x 4 schedulers x SM clock under load)."""
    ipr = ncu_inst_per_ray(args.phase) if (args.config == 2 and args.precision == 64 and not args.flat) else None
    mhz = (clocks or {}).get("sm_mhz")
    if not ipr or not mhz:
        return None
    achieved = total_rays * ipr / dev_s
    peak = 148 * 4 * mhz * 1e6 * world
    return {"warp_inst_per_ray": ipr, "achieved_inst_per_s": achieved, "peak_inst_per_s": peak,
            "frac": achieved / peak,
            "source": "smsp__inst_executed.sum of profiles/r01_%s_full.txt / 2^27 rays" % args.phase}


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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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


def cpu_arm(nrays, threads, first_ray=0):
    """Times the CPU oracle (C++/OpenMP restatement of the Fortran path) on `nrays` rays."""
    from opticalraytrace_b200 import abi
    from tests import oracle_lib as O
    scene = O.make_scene(*FILES)
    job = abi.default_job(abi.PHASE_RING, nrays, first_ray=first_ray)
    t0 = time.perf_counter()
    O.trace(job, scene, nthreads=threads)
    return nrays / (time.perf_counter() - t0)


def run_reference(args, rank):
    if rank != 0:
        return
    from tests import oracle_lib as O
    O.lib()
    cores = os.cpu_count() or 1
    sample = CPU_SAMPLE
    for _ in range(args.warmup):
        cpu_arm(sample // 10, cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_arm(sample, cores, first_ray=k * sample)
    dt = time.perf_counter() - t0
    v = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": "rays/sec", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": sample},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%d rays per step, C++/OpenMP restatement of the Fortran path "
                                   "(no Fortran compiler in this image)" % sample},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU, help="rays per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--phase", default="ring", choices=["ring", "point"],
                    help="ring = BASELINE.json configs[1] (the headline); point = the other loop")
    ap.add_argument("--flat", action="store_true", help="diagnostic: kernel without compaction")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4],
                    help="BASELINE.json config (2 = the headline)")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32],
                    help="64 = the reference's arithmetic (headline); 32 = the fp32 variant (1e-5 parity)")
    ap.add_argument("--fix-ellipse", action="store_true",
                    help="config 4: opt-in outer-ellipse fix instead of the reference's half radii")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from opticalraytrace_b200 import abi, lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the trace loop has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        box = [lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    lib.init_rank(local, rank, world, nccl_id)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = CONFIGS[args.config]
    phase = abi.PHASE_RING if args.phase == "ring" else abi.PHASE_POINT
    nsc = len(cfg["bottles"])
    n = args.rays // nsc          # rays per scene per GPU per step
    scene = []
    for bottle in cfg["bottles"]:
        st = lib.make_settings(bottle, cfg["l2"], cfg["l3"], nphotons=n)
        # the point loop runs with the lenses re-built at 843 nm (reference src/main.f90:113-117)
        scene.append(lib.build_scene(st, os.path.join(ROOT, "res"), None if phase == 1 else 843e-9)[0])

    def step(k, want_image=True, ph=phase, sc=scene, nr=n):
        job = lib.job_from_settings(st, ph)
        job.first_ray = (k * world + rank) * nr
        job.nrays = nr
        job.precision = args.precision
        if args.flat:
            job.flags |= abi.FLAG_NO_COMPACTION
        if args.fix_ellipse:
            job.flags |= abi.FLAG_FIX_OUTER_ELLIPSE
        return lib.trace(job, sc, want_image=want_image)

    for k in range(args.warmup):
        step(k)
    peak_tf, peak_mhz = lib.measure_fp64_peak()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_s, red_s, launches, h2d, d2h, d2h_s = 0.0, 0.0, 0, 0, 0, 0.0
    hist = np.zeros(32, dtype=np.int64)
    for k in range(args.steps):
        _, _, h, tm = step(args.warmup + k)
        dev_s += tm.trace_seconds + tm.reduce_seconds
        red_s += tm.reduce_seconds
        launches += tm.kernel_launches
        h2d, d2h = tm.h2d_bytes, tm.d2h_bytes
        d2h_s += tm.d2h_seconds
        if rank == 0:
            hist += h.sum(axis=0)
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_s, wall_s, red_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s, red_s = (float(x) for x in t.tolist())
    total_rays = float(n) * nsc * world * args.steps

    if rank == 0:
        # rank 0's histogram holds the reduced counts of all ranks (it is the reduce root)
        flops = float((flops_by_status(phase) * hist).sum())
        achieved = flops / dev_s * 1e-12
        cpu = None
        if world == 1 and not args.no_cpu and args.config == 2 and args.phase == "ring":
            cores = os.cpu_count() or 1
            cpu_arm(CPU_SAMPLE // 10, cores)
            v = cpu_arm(CPU_SAMPLE * 4, cores, first_ray=1 << 40)
            cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                   "sample": "%d rays of the same workload, C++/OpenMP restatement of the Fortran "
                             "path (oracle/), all host cores" % (CPU_SAMPLE * 4)}
        line = {
            "metric": "rays/sec", "value": total_rays / dev_s, "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f%d" % args.precision, "data": "synthetic",
            "config": {"workload": workload_text(args.config, args.phase, args.flat, args.fix_ellipse),
                       "rays_per_gpu_per_step": n * nsc, "scenes": nsc,
                       "rays_per_step": n * nsc * world, "parallelism": "ray-range x%d" % world,
                       "l2": "no input arrays (rays are generated on the device); the 2.6 MB "
                             "image buffer is re-zeroed every step",
                       "seed": 123456789},
            "e2e": {"value": total_rays / wall_s, "unit": "rays/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "reduce_ms_per_step": red_s / args.steps * 1e3,
            # there is no flush kernel (hits go to the L2-resident image with RED): the only image
            # traffic is this device-to-host read-back of image + histogram on rank 0
            "image_readback": {"bytes_per_step": int(d2h), "ms_per_step": d2h_s / args.steps * 1e3,
                               "GB/s": (d2h * args.steps / d2h_s * 1e-9) if d2h_s > 0 else None},
            "clocks": clocks,
            "roofline": {"bound": "alu_fp64", "achieved": achieved, "peak": peak_tf * world,
                         "unit": "TFLOP/s", "frac": achieved / (peak_tf * world) if peak_tf else None,
                         "traffic": ncu_traffic(args.phase) if (args.config == 2 and args.precision == 64
                                                                and not args.flat) else None,
                         "traffic_note": "DRAM bytes of one ncu-profiled launch of 2^27 rays "
                                         "(profiles/r01_<loop>_full.txt): the loops have no per-ray memory "
                                         "traffic, the image stays in L2",
                         "peak_source": "DFMA micro-kernel measured in this run on rank 0 at %.0f MHz, "
                                        "x n_gpus (MEASURED_PEAKS.json holds no FP64 figure)" % peak_mhz,
                         "flops_per_launched_ray": flops / (total_rays),
                         "issue": issue_bound(args, total_rays, dev_s, world, clocks)},
            "cpu_baseline": cpu,
            "status_fractions": {abi.STATUS_NAMES[i]: hist[i] / total_rays
                                 for i in range(26) if hist[i]},
        }
        print(json.dumps(line), flush=True)
    lib.finalize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

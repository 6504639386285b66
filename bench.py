#!/usr/bin/env python
"""Benchmark of the bulk particle loop (BASELINE.json: particle-steps/s, Si bulk EMC; HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (configs[1] of BASELINE.json): Si electrons, the shipped bulkSimulation mechanism set
(acoustic + zero- and first-order intervalley, 1000 energy levels of 1 meV), 300 K, 10 kV/cm along
-x, dt = 1e-16 s; 1e8 particles PER GPU (weak scaling), generated on the device from the
reference's initial distributions and advanced out of the initial transient before timing.
One "step" = one time step of every particle of the shard including the per-valley observables of
that step (the reference's moveParticles(dt) + the three getAvg* passes, bulkSimulation.cpp:150-157).
The headline runs the flight + event kernel pair (K1d, bulkFlightKernel / bulkEventKernel): up to SPL = 24
consecutive time steps per launch pair, the particle state crosses HBM once per pair, the per-step
observables of all steps are delivered.  The one-step-per-launch streaming kernel (K1a, bulkTmaKernel,
HBM-bound) is timed next to it ("one_step_per_launch"), and so is the reference's own driver loop
(moveParticles(dt) + getAvg* per step) through the drop-in C++ handler ("dropin_loop").  The rate tables are built on the host by the drop-in
C++ API (libemchost) -- not by the oracle.

N > 1: one process per GPU (torchrun), the ensemble is block-partitioned, no per-step communication;
the [K x valleys x 3] observable series is all-reduced (NCCL) once, inside the timed region.

--impl reference: the reference's OWN OpenMP implementation of the same path
(oracle/_ref/ref_bulk_bench: unmodified basicBulkParticleHandler::moveParticles + observable passes)
on the host cores, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_s"
UNIT = "particle-steps/s"
BYTES_PER_PARTICLE_STEP = 136.0  # 8 fp64 + 1 u32 read and written per particle-step
DT = 1e-16
FIELD = 1e6
DOPING = 1e23
SEED = 12345
SPL = 24  # time steps per launch pair of the headline kernels (K1d)
# FP64 instructions of the flight kernel per particle-step for a field along a coordinate axis (SASS of
# bulkFlightKernel<4, AXIS = 0>, hot loop: 11 DFMA + 9 DMUL + 4 DADD since the transverse components of k are known not to
# change -- flightCoreAxis; 28 before, and 28 for a general field direction; cross-checked against ncu's executed-opcode
# counts, profiles/r2_*flight*.txt)
FLIGHT_FP64_PER_PARTICLE_STEP = 24.0
FLIGHT_FLOP_PER_PARTICLE_STEP = 11 * 2 + 9 + 4
FP64_LANES_PER_SM = 64


def workload_config(particles_per_gpu, n_gpus, extra=None):
    cfg = {
        "workload": "Si bulk EMC (BASELINE configs[1]): 6 X valleys non-parabolic, acoustic + zero/first-order "
                    "intervalley phonons (shipped bulkSimulation set), 300 K, 10 kV/cm, dt=1e-16 s",
        "particles_per_gpu": int(particles_per_gpu),
        "particles_total": int(particles_per_gpu) * n_gpus,
        "steps_per_launch": SPL,
        "kernels": "bulkFlightKernel + bulkEventKernel (K1d), one launch pair per steps_per_launch time steps",
        "observables": "per-step per-valley <E>, <v.E>, occupation of EVERY time step, fused into the step kernel",
        "l2": "no flush needed: state per GPU (%.1f GB) is far larger than L2" % (particles_per_gpu * 68 / 1e9),
        "parallelism": f"particles block-partitioned over {n_gpus} GPU(s), one all-reduce of the observable series",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref, built from the unmodified reference headers)
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bulk_bench")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_ref_bench(particles, steps, warmup, threads):
    out = subprocess.run([REF_BENCH, "--particles", str(int(particles)), "--steps", str(int(steps)), "--warmup",
                          str(int(warmup)), "--threads", str(threads), "--field", str(FIELD), "--dt", str(DT), "--seed",
                          str(SEED)], capture_output=True, text=True, check=True,
                         env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_port_bench(particles, steps):
    """fallback when the reference binary is not there: the single-threaded oracle port"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import pyoracle as po
    from scenarios import build_si
    m = build_si()
    box = [(particles / DOPING) ** (1 / 3)] * 3
    ens, _ = m.generate_initial(box, [5, 5, 5], DOPING, po.mt_state(SEED))
    t0 = time.perf_counter()
    m.bulk_steps(ens, box, [-1, 0, 0], FIELD, DT, steps, po.rng_philox(SEED), first_step=1)
    dt = time.perf_counter() - t0
    return {"particles": ens.n, "steps": steps, "threads": 1, "psteps_per_s": ens.n * steps / dt,
            "move_s": dt, "obs_s": 0.0}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference(steps, warmup, budget_s):
    """Time the reference on a bounded sample: calibrate briefly, then size the ensemble so that
    steps+warmup time steps take about budget_s."""
    cores = host_cores()
    one_thread = None
    if os.path.exists(REF_BENCH):
        cal = run_ref_bench(100000, 20, 5, cores)
        rate = cal["psteps_per_s"]
        particles = max(20000, min(5_000_000, int(rate * budget_s / max(1, steps + warmup))))
        res = run_ref_bench(particles, steps, warmup, cores)
        kind = "reference"
        try:  # also one thread (small sample, a few seconds)
            one = run_ref_bench(100000, 20, 3, 1)
            one_thread = {"value": one["psteps_per_s"], "move_only_value": one.get("psteps_per_s_move_only"),
                          "sample": f"{one['particles']} particles x 20 time steps, 1 thread"}
        except Exception:
            one_thread = None
    else:
        particles = 2000
        res = run_port_bench(particles, min(steps, 20))
        kind, cores = "port", 1
        steps = res["steps"]
    sample = (f"{res['particles']} particles x {steps} time steps of the same workload "
              f"(moveParticles + 3 observable passes per step), {res['threads']} OpenMP thread(s)")
    return {"value": res["psteps_per_s"], "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
            "move_only_value": res.get("psteps_per_s_move_only"), "one_thread": one_thread, "cpu": cpu_model(),
            "build": "g++ -O3 -march=x86-64-v3 -fopenmp (the reference's Release flags with -march=native replaced so "
                     "that the binary built in the dev container runs on the GPU box's host CPU)",
            "seconds": res["move_s"] + res["obs_s"]}, res


CPU_KEYS = ("value", "unit", "cores", "kind", "sample", "move_only_value", "one_thread", "cpu", "build")


def dropin_loop(particles, steps, lookahead, device_index):
    """The reference's own driver loop -- handler.moveParticles(dt) followed by getAvgEnergy / getAvgDriftVelocity /
    getValleyOccupationProbability, once per time step (bulkSimulation.cpp:150-157) -- through the drop-in C++ handler
    (viennaemc_b200/bin/bulkSimulation, steps-per-launch 1 = that loop verbatim).  The handler runs `lookahead` steps per
    launch pair ahead and serves the following calls from the series."""
    import re
    path = os.path.join(ROOT, "viennaemc_b200", "bin", "bulkSimulation")
    if not os.path.exists(path):
        return {"failed": "driver binary missing"}
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([path, "--particles", str(int(particles)), "--steps", str(int(steps)), "--seed", str(SEED),
                            "--lookahead", str(int(lookahead)), "--dt", str(DT), "--field", str(FIELD)], cwd=tmp,
                           capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, EMCGPU_DEVICE=str(device_index)))
    m = re.search(r"Wall time: ([0-9.eE+-]+) s  \(([0-9.eE+-]+) particle-steps/s\)", r.stdout)
    if r.returncode != 0 or not m:
        return {"failed": (r.stdout + r.stderr)[-300:]}
    return {"value": float(m.group(2)), "unit": UNIT, "seconds": float(m.group(1)), "particles": int(particles),
            "steps": int(steps), "lookahead": int(lookahead),
            "what": "host wall clock around the reference's driver loop (one moveParticles(dt) + three getAvg* calls per time "
                    "step, host-generated ensemble uploaded before the loop) through basicBulkParticleHandler on the GPU path"}


def device_run_numbers(budget_s=60.0):
    """Configs 3 and 4 (parity-test cases, not the bench metric): Monte Carlo loop of the self-consistent device runs through
    the drop-in C++ API (viennaemc_b200/bin example drivers), reported next to the headline line.  Untimed region of bench.py."""
    import re
    import tempfile
    out = {}
    # run lengths: >= 1 s of Monte Carlo loop each, so that the clock ramp of an idle GPU does not weigh
    runs = (("resistor2D", "reference_order", ["--steps", "6000", "--transient", "2000", "--avg", "2000", "--red-black", "0"]),
            ("resistor2D", "red_black", ["--steps", "10000", "--transient", "3000", "--avg", "3000", "--red-black", "1"]),
            ("mosfet2D", "reference_order", ["--steps", "300", "--transient", "100", "--avg", "100", "--red-black", "0"]),
            ("mosfet2D", "red_black", ["--steps", "3000", "--transient", "1000", "--avg", "1000", "--red-black", "1"]))
    t_end = time.time() + budget_s
    for exe, order, extra in runs:
        path = os.path.join(ROOT, "viennaemc_b200", "bin", exe)
        if not os.path.exists(path) or time.time() > t_end:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            try:
                r = subprocess.run([path, "--seed", "5", "--progress", "100000", *extra], cwd=tmp, capture_output=True, text=True,
                                   timeout=max(5.0, t_end - time.time()))
            except subprocess.TimeoutExpired:
                continue
        m = re.search(r"(\d+) steps, (\d+) particles at the end.*Monte Carlo loop alone: ([0-9.eE+-]+) s, ([0-9.eE+-]+) "
                      r"particle-steps/s\), ([0-9.eE+-]+) SOR sweeps", r.stdout)
        if r.returncode == 0 and m:
            steps, n, secs, rate, sweeps = int(m.group(1)), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
            out.setdefault(exe, {})[order] = {"particle_steps_per_s": rate, "us_per_step": secs / steps * 1e6, "particles": n,
                                              "steps": steps, "sor_sweeps_per_step": sweeps}
    # the reference's own CPU run of the same devices on this box: the UNMODIFIED example mains with only
    # their run length shortened (oracle/Makefile: _ref/ref_resistor2D_short 3000 steps, _ref/ref_mosfet2D_short 300 steps),
    # 4 OpenMP threads as the examples hard-code; "CPU time" is the examples' own clock around simulation.execute()
    # (whole seconds, includes the initial equilibrium solve and particle creation)
    for exe, steps in (("resistor2D", 3000), ("mosfet2D", 300)):
        path = os.path.join(ROOT, "oracle", "_ref", f"ref_{exe}_short")
        if exe not in out or not os.path.exists(path) or time.time() > t_end + 60.0:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            try:
                r = subprocess.run([path], cwd=tmp, capture_output=True, text=True, timeout=180.0)
            except subprocess.TimeoutExpired:
                continue
        m = re.search(r"CPU time: (\d+) s", r.stdout)
        if r.returncode == 0 and m and int(m.group(1)) > 0:
            secs = float(m.group(1))
            n = next(iter(out[exe].values()))["particles"]
            out[exe]["cpu_reference"] = {"particle_steps_per_s": n * steps / secs, "us_per_step": secs / steps * 1e6, "steps": steps,
                                         "threads": 4, "seconds": secs, "cores": host_cores(), "kind": "reference",
                                         "what": f"unmodified reference examples/{exe} main(), run length shortened to {steps} "
                                                 "steps, on this box's host cores (the example fixes 4 OpenMP threads)"}
    return out


def sharded_device_runs(rank, world, local_rank, id_dir):
    """Configs 3 and 4 on ALL GPUs of the run: the drop-in C++ drivers started once per GPU (this process
    starts the one of its rank), particles split over the ranks, grids replicated, the library's per-step exchange
    (reservoir share table, carriers-per-grid-point grid) as ncclAllReduce over NVLink through libemcnccl.  emcSimulation
    itself asserts that the potential and the averaged concentration are bit-identical on all ranks.  Rank 0 reports."""
    import re
    out = {}
    runs = (("resistor2D", ["--steps", "10000", "--transient", "3000", "--avg", "3000"]),
            ("mosfet2D", ["--steps", "3000", "--transient", "1000", "--avg", "1000"]))
    for exe, extra in runs:
        path = os.path.join(ROOT, "viennaemc_b200", "bin", exe)
        if not os.path.exists(path):
            continue
        env = dict(os.environ, EMCGPU_SHARD="1", RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local_rank),
                   EMCNCCL_ID_FILE=os.path.join(id_dir, exe + ".id"))
        env.pop("EMCGPU_DEVICE", None)
        with tempfile.TemporaryDirectory() as tmp:
            try:
                r = subprocess.run([path, "--seed", "5", "--progress", "100000", *extra], cwd=tmp, capture_output=True, text=True,
                                   timeout=300, env=env)
            except subprocess.TimeoutExpired:
                out[exe] = {"failed": "timeout"}
                continue
        m = re.search(r"(\d+) steps, (\d+) particles at the end.*Monte Carlo loop alone: ([0-9.eE+-]+) s, ([0-9.eE+-]+) "
                      r"particle-steps/s\), ([0-9.eE+-]+) SOR sweeps", r.stdout)
        same = "identical on all ranks" in r.stdout
        calls = re.search(r"(\d+) all-reduces \((\d+) bytes\)", r.stdout)
        if r.returncode == 0 and m:
            steps, n, secs = int(m.group(1)), int(m.group(2)), float(m.group(3))
            out[exe] = {"particle_steps_per_s": float(m.group(4)), "us_per_step": secs / steps * 1e6, "particles_total": n,
                        "steps": steps, "sor_sweeps_per_step": float(m.group(5)), "ranks": world,
                        "grids_identical_on_all_ranks": same,
                        "allreduces_per_step": (int(calls.group(1)) / steps) if calls else None,
                        "allreduce_bytes_per_step": (int(calls.group(2)) / steps) if calls else None}
        else:
            out[exe] = {"failed": (r.stdout + r.stderr)[-300:]}
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base, res = cpu_reference(args.steps, args.warmup, budget_s=90.0)
    ms = 1e3 * base["seconds"] / max(1, res["steps"])
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.particles, args.gpus,
                                      {"reference_sample": base["sample"],
                                       "note": "reference CPU path: throughput does not depend on the GPU count"}),
            "cpu_baseline": {k: base[k] for k in CPU_KEYS},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from viennaemc_b200 import capi, hostapi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:  # convenience: spawn the ranks ourselves
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), *sys.argv]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_local = int(args.particles)
    n_total = n_local * world
    box = [(n_total / DOPING) ** (1.0 / 3.0)] * 3
    ctx = capi.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    hostapi.si_upload(ctx, hostapi.si_spec(box=box, spacing=[b / 5 for b in box], doping=DOPING))
    first_id, last_id = sharding.shard_range(n_total, rank, world)
    assert last_id - first_id == n_local
    ctx.generate_bulk_ensemble(n_local, box, 300.0, 0, seed=SEED, particle_id_base=first_id)
    ctx.rng_philox(SEED)
    ctx.bulk_configure(box, [-1, 0, 0], FIELD, math_mode=capi.MATH_FAST)
    ctx.set_step_index(1)
    n_v = 1
    K, W = args.steps, args.warmup

    # leave the initial transient (untimed set-up; fused launches)
    if args.settle > 0:
        scratch = torch.zeros(args.settle * n_v * 3, dtype=torch.float64, device="cuda")
        ctx.bulk_step_device(DT, args.settle, 16, scratch.data_ptr())
        del scratch
    obs = torch.zeros(K * n_v * 3, dtype=torch.float64, device="cuda")
    warm = torch.zeros(max(1, W) * n_v * 3, dtype=torch.float64, device="cuda")

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- headline: K time steps, SPL per launch pair (K1d), inputs resident in HBM ---------------
    if W > 0:
        ctx.bulk_step_device(DT, W, SPL, warm.data_ptr())
    ctx.set_option("kernel_timing", 1)  # cudaEvents around every kernel launch, on the launching stream
    ctx.kernel_times(reset=True)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def run_steps():
        ctx.bulk_step_device(DT, K, SPL, obs.data_ptr())
        sharding.allreduce_observables(obs)  # the only communication of a bulk run (no-op for one rank)

    ms = timed(run_steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    kernel_ms, kernel_launches = ctx.kernel_times(reset=True)  # [flight, event, other] inside the timed region
    ctx.set_option("kernel_timing", 0)
    value = n_total * K / (ms * 1e-3)
    series = obs.view(K, n_v, 3).cpu().numpy()
    assert np.all(series[:, :, 2].sum(axis=1) == n_total), "particle count not conserved"
    mean_e = float((series[:, 0, 0] / series[:, 0, 2]).mean())
    mean_v = float((series[:, 0, 1] / series[:, 0, 2]).mean())

    # ---- the same K steps with ONE time step per launch (K1a, the HBM-bound streaming kernel) ---
    ctx.bulk_step_device(DT, max(3, W), 1, warm.data_ptr() if W >= 3 else obs.data_ptr())
    one_ms = timed(lambda: ctx.bulk_step_device(DT, K, 1, obs.data_ptr()))
    one_value = n_total * K / (one_ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(n_local, dtype=torch.float64, pin_memory=True) for _ in range(capi.N_STREAMS)]
        host_packed = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
        streams = [h.numpy() for h in host]
        packed = host_packed.numpy().view(np.uint32)
        ctx.get_ensemble_into(streams, packed)  # the job's input now lives in (pinned) host memory
        obs_host = np.zeros((K, n_v, 3))
        base_id = first_id

        def run_e2e():
            # ONE call on host buffers: the ensemble is advanced in place slice by slice, H2D of slice i+1 and D2H of
            # slice i-1 overlap the K time steps of slice i (68 B per particle each way, all inside the timed region)
            o = ctx.bulk_run_host(streams, packed, DT, K, SPL, 0, particle_id_base=base_id)
            obs_host[:] = o

        def run_e2e_resident():
            ctx.set_ensemble_from(streams, packed, base_id)  # H2D, 68 B per particle
            for s in range(0, K, SPL):  # per launch: SPL time steps + D2H of their observables (24 B per valley and step)
                ctx.L.emcgpu_bulk_step(ctx.h, DT, min(SPL, K - s), SPL, obs_host[s].ctypes.data_as(capi._DP))
            ctx.get_ensemble_into(streams, packed)  # D2H, 68 B per particle

        e2e_res_ms = timed(run_e2e_resident)
        assert np.all(obs_host[:, :, 2].sum(axis=1) == n_local)
        unpipelined = {"value": n_total * K / (e2e_res_ms * 1e-3), "ms_total": e2e_res_ms,
                       "what": f"emcgpu_set_ensemble from pinned host arrays + K/{SPL} x emcgpu_bulk_step({SPL} steps, host "
                               "observables) + emcgpu_get_ensemble to pinned host arrays (copies not overlapped), all "
                               "inside the timed region (per rank)"}
        e2e_bytes = {"h2d_bytes_per_step": 68.0 * n_local / K, "d2h_bytes_per_step": 68.0 * n_local / K + 24.0 * n_v}
        try:
            ctx.bulk_run_host(streams, packed, DT, SPL, SPL, 0, particle_id_base=base_id, want_obs=False)  # untimed warm-up
            e2e_ms = timed(run_e2e)
            if not np.all(obs_host[:, :, 2].sum(axis=1) == n_local):
                raise RuntimeError("particle count not conserved in emcgpu_bulk_run_host")
            e2e = {"value": n_total * K / (e2e_ms * 1e-3), "unit": UNIT, **e2e_bytes, "ms_total": e2e_ms,
                   "what": f"emcgpu_bulk_run_host on pinned host arrays ({K} time steps, {SPL} per launch, host "
                           "observables): the ensemble is cut into ~8 slices (~16 for K < 128), each slice is copied in, advanced K steps "
                           "and copied back with the copies of the neighbouring slices overlapping its kernels; all "
                           "inside the timed region (per rank)",
                   "unpipelined": unpipelined}
        except Exception as exc:  # the streamed call failed: report the unpipelined calls and say so (no silent switch)
            sys.stderr.write(f"emcgpu_bulk_run_host failed ({exc}); e2e is the unpipelined set/step/get sequence\n")
            e2e = {"value": unpipelined["value"], "unit": UNIT, **e2e_bytes, "ms_total": e2e_res_ms,
                   "what": unpipelined["what"], "bulk_run_host_error": str(exc)}
        # what the host link allows: the same bytes copied in and copied out with no compute at all.  With both PCIe
        # directions perfectly overlapped a job on host buffers cannot finish before max(t_in, t_out).
        t_in = timed(lambda: ctx.set_ensemble_from(streams, packed, base_id))
        t_out = timed(lambda: ctx.get_ensemble_into(streams, packed))
        ceiling = n_total * K / (max(t_in, t_out) * 1e-3)
        e2e["copy_ceiling"] = {"value": ceiling, "unit": UNIT, "h2d_ms": t_in, "d2h_ms": t_out,
                               "h2d_gbs_per_rank": 68.0 * n_local / (t_in * 1e-3) / 1e9,
                               "d2h_gbs_per_rank": 68.0 * n_local / (t_out * 1e-3) / 1e9,
                               "frac": e2e["value"] / ceiling,
                               "what": f"68 B per particle each way over the host link and nothing else (max over ranks, all {world} "
                                       f"rank(s) copying at once), expressed in the metric for K = {K}: the bound of any run on host "
                                       "buffers at this K; 'frac' = e2e / this"}
        if K < 200 and not args.no_e2e_long:
            # the same call on a run long enough to amortise the two copies (K = 1000 time steps)
            KL = 1000
            long_sampler = ClockSampler(local_rank)
            if rank == 0:
                long_sampler.start()
            long_ms = timed(lambda: ctx.bulk_run_host(streams, packed, DT, KL, SPL, 0, particle_id_base=base_id, want_obs=False))
            e2e["long_run"] = {"steps": KL, "value": n_total * KL / (long_ms * 1e-3), "unit": UNIT, "ms_total": long_ms,
                               "h2d_bytes_per_step": 68.0 * n_local / KL, "d2h_bytes_per_step": 68.0 * n_local / KL}
            if rank == 0:  # a run this long is where a power or thermal cap would show
                e2e["long_run"]["clocks"] = long_sampler.stop()
        # both directions AT ONCE (what a pipelined run actually gets from the link): one half of every array goes in while the
        # other half comes back, on two streams; twice that time = 68 B per particle each way in full duplex
        try:
            dev = [torch.empty(n_local, dtype=torch.float64, device="cuda") for _ in range(capi.N_STREAMS)]
            dev_packed = torch.empty(n_local, dtype=torch.int32, device="cuda")
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            half = n_local // 2

            def duplex():
                cur = torch.cuda.current_stream()
                s_in.wait_stream(cur)
                s_out.wait_stream(cur)
                with torch.cuda.stream(s_in):
                    for a, d in zip(host, dev):
                        d[:half].copy_(a[:half], non_blocking=True)
                    dev_packed[:half].copy_(host_packed[:half], non_blocking=True)
                with torch.cuda.stream(s_out):
                    for a, d in zip(host, dev):
                        a[half:].copy_(d[half:], non_blocking=True)
                    host_packed[half:].copy_(dev_packed[half:], non_blocking=True)
                cur.wait_stream(s_in)
                cur.wait_stream(s_out)

            duplex()
            t_duplex = 2.0 * timed(duplex)
            cc = e2e["copy_ceiling"]
            cc["duplex_ms"] = t_duplex
            cc["duplex_value"] = n_total * K / (t_duplex * 1e-3)
            cc["frac_of_duplex"] = e2e["value"] / cc["duplex_value"]
            cc["what"] += ("; duplex_ms = the same bytes with both directions busy at the same time (two streams, half of every "
                           "array each way, time doubled): the bound of a PIPELINED run, frac_of_duplex = e2e / that")
            del dev, dev_packed
        except Exception as exc:  # reported, never required
            e2e["copy_ceiling"]["duplex_failed"] = str(exc)
        del host, host_packed

    # ---- roofline of the step kernel -------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic_flight = traffic_event = traffic_one = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if int(t.get("particles", 0)) == n_local:
            traffic_flight = t.get("bulkFlightKernel", {}).get("dram_bytes_per_launch")
            traffic_event = t.get("bulkEventKernel", {}).get("dram_bytes_per_launch")
            traffic_one = t.get("bulkTmaKernel", {}).get("dram_bytes_per_launch")
    n_pairs = max(1, int(kernel_launches[0]))
    flight_ms = float(kernel_ms[0]) / n_pairs          # average duration of a flight launch (CUDA events, this run)
    event_ms = float(kernel_ms[1]) / max(1, int(kernel_launches[1]))
    steps_per_pair = K / n_pairs
    psteps_per_launch = n_local * steps_per_pair
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_mhz = float((clocks or {}).get("sm_mhz") or 0.0) or float((clocks or {}).get("sm_max_mhz") or 1965.0)
    fp64_peak = sm_count * FP64_LANES_PER_SM * sm_mhz * 1e6 / 1e12          # T lane-instructions/s of the FP64 pipe
    fp64_ach = FLIGHT_FP64_PER_PARTICLE_STEP * psteps_per_launch / (flight_ms * 1e-3) / 1e12
    algorithmic = BYTES_PER_PARTICLE_STEP * n_local * K / (ms * 1e-3) / 1e9
    moved = ((traffic_flight or 0) + (traffic_event or 0)) / ((flight_ms + event_ms) * 1e-3) / 1e9 if traffic_flight else None
    roofline = {"bound": "fp64",
                "kernel": f"bulkFlightKernel<4, axis> (dominant kernel: {kernel_ms[0] / max(1e-9, kernel_ms[0] + kernel_ms[1]):.0%} "
                          f"of the device time of a launch pair; {steps_per_pair:g} time steps per launch)",
                "achieved": fp64_ach, "peak": fp64_peak, "unit": "T fp64 lane-instructions/s (FP64 pipe: 64 lanes/clk/SM)",
                "frac": fp64_ach / fp64_peak, "traffic": traffic_flight,
                "tflops": FLIGHT_FLOP_PER_PARTICLE_STEP * psteps_per_launch / (flight_ms * 1e-3) / 1e12,
                "peak_source": f"{sm_count} SMs x {FP64_LANES_PER_SM} FP64 lanes x {sm_mhz:.0f} MHz (median SM clock sampled "
                               "during the timed region)",
                "fp64_instructions_per_particle_step": FLIGHT_FP64_PER_PARTICLE_STEP,
                "avg_launch_ms": flight_ms, "particle_steps_per_launch": psteps_per_launch,
                "flight_only_particle_steps_per_s": psteps_per_launch / (flight_ms * 1e-3),
                "event_kernel": {"avg_launch_ms": event_ms, "traffic": traffic_event,
                                 "share_of_pair": kernel_ms[1] / max(1e-9, kernel_ms[0] + kernel_ms[1]),
                                 "note": "latency-bound (gathers of the frozen particles, table rows through L2, dependent "
                                         "fp64 chains of the samplers): FP64 pipe 22 %, issue slots 35 % (profiles/r2_*event*.txt)"},
                "hbm": {"algorithmic_gbs": algorithmic, "algorithmic_frac": algorithmic / peak,
                        "moved_gbs": moved, "moved_frac": (moved / peak) if moved else None, "peak": peak,
                        "peak_source": peak_src,
                        "note": "algorithmic = 136 B x particle-steps (SURVEY 8d) over the whole timed region; moved = ncu DRAM "
                                "bytes of one flight + one event launch (profiles/traffic.json) over their measured "
                                "durations: the state crosses HBM once per launch pair, so the step is bound by the FP64 "
                                "pipe, not by HBM; the HBM-bound one-step kernel is in 'one_step_per_launch'"},
                "note": "per GPU. achieved = 24 FP64 instructions per particle-step (SASS of the flight loop) x particle-steps of "
                        "a launch / average flight-launch duration measured with CUDA events in this run; frac = share of the "
                        "FP64 pipe's issue rate (the binding resource, ncu: sm__pipe_fp64_cycles_active)"}
    one_achieved = BYTES_PER_PARTICLE_STEP * n_local * K / (one_ms * 1e-3) / 1e9
    one_step = {"kernel": "bulkTmaKernel<FAST, PHILOX> (one time step per launch, TMA pipeline)", "value": one_value,
                "unit": UNIT, "ms_per_step": one_ms / K,
                "roofline": {"bound": "hbm", "achieved": one_achieved, "peak": peak, "unit": "GB/s",
                             "frac": one_achieved / peak, "traffic": traffic_one,
                             "algorithmic_bytes_per_launch": BYTES_PER_PARTICLE_STEP * n_local,
                             "avg_launch_ms": one_ms / K}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(n_local, world, {"settle_steps": args.settle, "math": "FAST (FMA, hoisted constants)",
                                                       "rng": "Philox4x32-10 keyed by global particle id and step"}),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "one_step_per_launch": one_step,
            "observables": {"mean_energy_eV": mean_e, "mean_drift_velocity_m_s": mean_v}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                base, _ = cpu_reference(steps=50, warmup=5, budget_s=15.0)
                line["cpu_baseline"] = {k: base[k] for k in CPU_KEYS}
            except Exception as exc:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference",
                                        "sample": f"failed: {exc}"}
        if world == 1 and not args.no_dropin_loop:
            try:
                line["dropin_loop"] = dropin_loop(n_local, 40 * SPL, SPL, local_rank)  # 960 steps: ~0.4 s, start-up of the process amortised
                line["dropin_loop"]["frac_of_value"] = line["dropin_loop"].get("value", 0.0) / value
            except Exception as exc:
                line["dropin_loop"] = {"failed": str(exc)}
        if world == 1 and not args.no_device_runs:
            try:
                line["device_runs"] = device_run_numbers()
            except Exception as exc:
                line["device_runs"] = {"failed": str(exc)}
    ctx.close()
    if world > 1:
        if not args.no_device_runs:
            # configs 3 / 4 sharded over all GPUs of the run: every rank starts the C++ driver of its rank
            box = [tempfile.mkdtemp(prefix="emcnccl") if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            sharded = sharded_device_runs(rank, world, local_rank, box[0])
            dist.barrier()
            if rank == 0:
                line["device_runs_sharded"] = sharded
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=1e8, help="particles per GPU")
    ap.add_argument("--settle", type=int, default=2000, help="untimed time steps before the measurement")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-long", action="store_true", help="skip the 1000-step end-to-end run")
    ap.add_argument("--no-dropin-loop", action="store_true", help="skip the reference-driver-loop leg (C++ handler)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-device-runs", action="store_true", help="skip the (untimed) device-run numbers of configs 3 / 4")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    sys.exit(main_reference(args) if args.impl == "reference" else main_ours(args))


if __name__ == "__main__":
    main()

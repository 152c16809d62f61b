#!/usr/bin/env python
"""bench.py -- skeleton events/s (+ ESS/s) of the grid-based Poisson-thinning hot path on B200.

Workload (BASELINE.json configs[1], "C2"): ZigZag, banana potential d=50 with the manual gradient,
grid_size=0 (constant bound via Brent), 4096 chains per GPU, Philox draws, full PDMPHistory columns stored.
One "step" = every chain advanced by --events accepted events (one launch of the skeleton kernel), followed
by the closed-form moment kernel and (N > 1) one NCCL all-reduce of the moment sums.

  value  events/s, whole job, state and outputs resident in HBM (CUDA events, max over ranks)
  e2e    events/s through pdmpflux_sample_skeleton with HOST (pinned) buffers: H2D of the initial states and
         D2H of the full history inside the timed region
  roofline  algorithmic bytes (16 d + 76 per event: X, V, t, horizon, ar, error_value_ar, 3 int32 counters)
         / skeleton-kernel time vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the C restatement of the reference (oracle/, OpenMP, one chain per thread) on the host cores
`--impl reference` times that CPU restatement alone (Julia, hence the reference itself, is not installed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
ALL_CPUS = os.sched_getaffinity(0)
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (description, sampler ctor, d, vinit unit-norm?, x0 value, oracle (sampler, potential, params, kwargs))
    "c1": dict(desc="ZigZagAD std Gaussian d=10, grid_size=10 (README example)", d=10, unit_v=False, x0=0.0,
               oracle=(0, 0, None, dict())),
    "c2": dict(desc="ZigZag banana d=50, manual gradient, grid_size=0 (Brent constant bounds)", d=50, unit_v=False,
               x0=1.0, oracle=(0, 3, None, dict(grid_size=0))),
    "c3": dict(desc="BPS slanted (equicorrelated rho=0.9) Gaussian d=100, grid_size=10, refresh 0.1", d=100,
               unit_v=True, x0=0.0, oracle=(1, 2, [0.9], dict(tmax=1.0, refresh_rate=0.1))),
    "c4": dict(desc="ZigZag Bayesian logistic regression d=100, n=1e5 synthetic rows, grid_size=10 (FP64 DMMA gradient)",
               d=100, unit_v=False, x0=0.0, oracle=(0, 5, "logreg:100000", dict())),
    "c5f": dict(desc="ForwardECMC std Gaussian d=1000, grid_size=10", d=1000, unit_v=True, x0=0.0,
                oracle=(2, 0, None, dict())),
    "c5b": dict(desc="Boomerang std Gaussian d=1000, grid_size=10, refresh 0.1", d=1000, unit_v=False, x0=0.0,
                oracle=(3, 0, None, dict(tmax=1.0, refresh_rate=0.1, deriv_mode=1))),
}
DEFAULT_CHAINS = {"c1": 65536, "c2": 4096, "c3": 16384, "c4": 4096, "c5f": 8192, "c5b": 8192}
DEFAULT_EVENTS = {"c1": 500, "c2": 1000, "c3": 300, "c4": 8, "c5f": 40, "c5b": 40}


def logreg_data(n, d, seed=2024, sigma0=10.0):
    """Synthetic design of SURVEY.md 8d (C4): rows ~ N(0, I/d), theta* ~ N(0, I), y ~ Bernoulli(sigma(x.theta*))."""
    import numpy as np
    g = np.random.default_rng([seed, n, d])
    X = g.standard_normal((n, d)) / np.sqrt(d)
    theta = g.standard_normal(d)
    y = (g.random(n) < 1.0 / (1.0 + np.exp(-X @ theta))).astype(np.float64)
    return X, y, sigma0


def make_sampler(p, name):
    if name == "c1":
        return p.ZigZagAD(10, p.GaussStd())
    if name == "c2":
        return p.ZigZag(50, p.Banana(), grid_size=0)
    if name == "c3":
        return p.BPS(100, p.GaussEquicorr(0.9), refresh_rate=0.1)
    if name == "c4":
        X, y, s0 = logreg_data(100000, 100)
        return p.ZigZagAD(100, p.LogReg(X, y, s0))
    if name == "c5f":
        return p.ForwardECMC(1000, p.GaussStd())
    if name == "c5b":
        return p.Boomerang(1000, p.GaussStd())
    raise SystemExit(f"unknown config {name}")


FP64_PEAK_TFLOPS = 37.0  # measured FP64 tensor (DMMA) throughput, scratch/dmma_peak.cu; plain DFMA measures 33.4 (scratch/fp64_peak.cu)


def pin_to_gpu_numa_node(index):
    """Run (and allocate pinned host memory) on the CPUs local to the GPU's PCIe root: host buffers on the remote
    NUMA node make the D2H leg of the end-to-end number vary by 2-3x from run to run."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        cpus = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except (OSError, ValueError, subprocess.SubprocessError):
        pass
    return None


def bytes_per_event(d):
    return 16 * d + 76


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(name, target_seconds=12.0, threads=None):
    """C restatement of the reference (oracle/) on the host cores, one chain per OpenMP thread, Philox draws:
    a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_c as oc
    cfgd = CONFIGS[name]
    sampler, pot, pp, kw = cfgd["oracle"]
    d = cfgd["d"]
    threads = threads or os.cpu_count() or 1
    if isinstance(pp, str) and pp.startswith("logreg:"):
        import numpy as np
        X, y, s0 = logreg_data(int(pp.split(":")[1]), d)
        pp = np.concatenate([[float(X.shape[0]), s0], X.ravel(), y])
    cfg = oc.make_cfg(sampler, pot, d, pp, **kw)
    nch = threads * 4
    x0 = np.full((nch, d), cfgd["x0"]); v0 = np.ones((nch, d)) / (d ** 0.5 if cfgd["unit_v"] else 1.0)
    n_ev = 50 if name != "c4" else 1
    if name == "c4":
        nch = threads
        x0, v0 = x0[:nch], v0[:nch]
    t0 = time.perf_counter()
    oc.sample_skeleton(cfg, n_ev + 1, x0, v0, seed=2024, nthreads=threads)
    dt = time.perf_counter() - t0
    want = n_ev * target_seconds / max(dt, 1e-4)          # events per chain for the target time at nch chains
    n_ev = int(max(50 if name != "c4" else 1, min(20000, want)))   # cap the stored history (16 d bytes per event per chain)
    if want > n_ev:
        nch = int(min(threads * 64, max(nch, threads * round(nch * want / n_ev / threads))))
        x0 = np.full((nch, d), cfgd["x0"]); v0 = np.ones((nch, d)) / (d ** 0.5 if cfgd["unit_v"] else 1.0)
    t0 = time.perf_counter()
    r = oc.sample_skeleton(cfg, n_ev + 1, x0, v0, seed=2024, nthreads=threads)
    dt = time.perf_counter() - t0
    assert (r.status == 0).all()
    return {"value": nch * n_ev / dt, "unit": "events/s", "cores": threads, "kind": "port",
            "sample": f"{nch} chains x {n_ev} events, {dt:.1f} s, C restatement of the reference (oracle/pdmp_oracle.c) "
                      f"with OpenMP; Julia (the reference runtime) is not installed"}, dt, nch * n_ev


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.config
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(name, target_seconds=2.0)
    total_t, total_ev = 0.0, 0
    for _ in range(args.steps):
        cb, dt, ev = cpu_baseline(name, target_seconds=max(2.0, 60.0 / max(args.steps, 1)))
        vals.append(cb); total_t += dt; total_ev += ev
    v = total_ev / total_t
    cb = dict(vals[-1]); cb["value"] = v
    line = {"metric": "skeleton events/sec", "value": v, "unit": "events/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {CONFIGS[name]['desc']}, CPU port of the reference, one chain per thread"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the BASELINE config's)")
    ap.add_argument("--events", type=int, default=0, help="events per chain per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import pdmpflux_b200 as p

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    numa_cpus = pin_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    p.lib().pdmpflux_set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    name = args.config
    cfgd = CONFIGS[name]
    d = cfgd["d"]
    nch = args.chains or DEFAULT_CHAINS[name]
    n_ev = args.events or DEFAULT_EVENTS[name]
    sampler = make_sampler(p, name)
    f64 = torch.float64
    x0 = torch.full((nch, d), cfgd["x0"], dtype=f64, device=dev)
    v0 = torch.ones((nch, d), dtype=f64, device=dev) / (d ** 0.5 if cfgd["unit_v"] else 1.0)
    chain_offset, _ = p.dist.shard(nch * world, rank, world)      # weak scaling: nch chains on every rank
    chains = p.DeviceChains(sampler, x0, v0, seed=2024, chain_offset=chain_offset)
    bufs = dict(X=torch.empty((nch, n_ev, d), dtype=f64, device=dev), V=torch.empty((nch, n_ev, d), dtype=f64, device=dev),
                t=torch.empty((nch, n_ev), dtype=f64, device=dev), horizon=torch.empty((nch, n_ev), dtype=f64, device=dev),
                ar=torch.empty((nch, n_ev), dtype=f64, device=dev),
                error_value_ar=torch.empty((nch, n_ev, 5), dtype=f64, device=dev),
                errored_bound=torch.empty((nch, n_ev), dtype=torch.int32, device=dev),
                rejected=torch.empty((nch, n_ev), dtype=torch.int32, device=dev),
                hitting_horizon=torch.empty((nch, n_ev), dtype=torch.int32, device=dev))
    view = p.device_history_view(n_ev, **bufs)
    m1 = torch.empty((nch, d), dtype=f64, device=dev); m2 = torch.empty((nch, d), dtype=f64, device=dev)
    Tl = torch.empty((nch,), dtype=f64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    lib = p.lib()
    from pdmpflux_b200 import _lib as L

    kern_ms = []

    def step(timed):
        k0 = torch.cuda.Event(enable_timing=True); k1 = torch.cuda.Event(enable_timing=True)
        k0.record()
        chains.advance(n_ev, view, 0, stream)                       # skeleton kernel: n_ev events per chain
        k1.record()
        L.check(lib.pdmpflux_skeleton_moments(sampler.flow_kind, d, n_ev, nch, 0, bufs["X"].data_ptr(),
                                              bufs["V"].data_ptr(), bufs["t"].data_ptr(), m1.data_ptr(), m2.data_ptr(),
                                              Tl.data_ptr(), 1, stream))
        sums = p.dist.moment_sums(m1 / Tl[:, None], m2 / Tl[:, None])   # 4 x d sufficient statistics
        p.dist.all_reduce_sums(sums)                                # the only collective: final moment reduction
        if timed:
            kern_ms.append((k0, k1))
        return sums

    for _ in range(max(args.warmup, 3)):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.pdmpflux_launch_count()
    clocks = ClockSampler(local); clocks.start()
    time.sleep(0.25)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        sums = step(True)
    e1.record()
    torch.cuda.synchronize()
    w1 = time.time()
    if world > 1:
        dist.barrier()
    clk = clocks.stop(w0, w1)
    launches = lib.pdmpflux_launch_count() - launches0
    elapsed = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=f64, device=dev)
    kern = torch.tensor([sum(a.elapsed_time(b) for a, b in kern_ms) * 1e-3 / len(kern_ms)], dtype=f64, device=dev)
    per_rank_kernel_ms = [float(kern) * 1e3]
    if world > 1:
        gathered = [torch.empty_like(kern) for _ in range(world)]
        dist.all_gather(gathered, kern)
        per_rank_kernel_ms = [float(g) * 1e3 for g in gathered]
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed); kern = float(kern)
    chains.status()  # raises if any chain stopped
    total_chains = nch * world
    events_per_step = total_chains * n_ev
    value = events_per_step * args.steps / elapsed

    # ESS/s from the last step's cross-chain moment sums (definition: SURVEY.md 8d / sample.ess_from_chain_means)
    stats = p.dist.ess_from_sums(sums)
    ess_per_s = float(stats["ess"].min() / (elapsed / args.steps))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = nch * n_ev * bytes_per_event(d) / kern / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "skeleton_kernel", "kernel_ms": kern * 1e3,
                "bytes_per_event": bytes_per_event(d),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6.65 TB/s"}

    line = {"metric": "skeleton events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {cfgd['desc']}", "chains_per_gpu": nch, "events_per_chain_per_step": n_ev,
                       "draws": "philox4x32-10 keyed (seed=2024, chain, event)", "stored": "full PDMPHistory row",
                       "l2": "outputs per step (%.2f GB) exceed the 126 MB L2" % (nch * n_ev * bytes_per_event(d) / 1e9)},
            "ess_per_s": ess_per_s, "ess_definition": "min over coordinates of C * Var_pi(x_i) / Var_c(chain time-average of x_i), per step window",
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "host_cpus": numa_cpus,
            "per_rank_kernel_ms": per_rank_kernel_ms}

    chains.close()
    del bufs, view
    torch.cuda.empty_cache()
    if not args.no_e2e:
        # every rank runs its shard through the host-buffer call at the same time (they share the host's memory
        # system and CPUs); the slowest rank's time counts
        if world > 1:
            # CPUs this rank may use for rebuilding the V rows: an equal share of the CPUs local to its GPU
            local = sorted(os.sched_getaffinity(0))
            sharing = [None] * world
            dist.all_gather_object(sharing, local)
            peers = sum(1 for other in sharing if other == local)
            os.environ["PDMPFLUX_HOST_THREADS"] = str(max(1, min(16, len(local) // max(peers, 1))))
            # (one rank alone needs >= 12 threads to beat the plain copy of the V rows -- the library's own default -- but
            # with several ranks the host's ingest bandwidth saturates (plain copy: 75 ms per step at 1 rank, 97 at 2,
            # 211 at 4), so moving half the bytes wins even with few threads per rank)
            os.environ["PDMPFLUX_VBITS"] = "1"
            dist.barrier()
        res = e2e(p, sampler, name, nch, n_ev, world, dev)
        if world > 1:
            worst = torch.tensor([res["ms_per_step"]], dtype=f64, device=dev)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            res["ms_per_step"] = float(worst)
            res["value"] = world * nch * n_ev / (float(worst) * 1e-3)
            res["h2d_bytes_per_step"] *= world; res["d2h_bytes_per_step"] *= world; res["history_bytes_per_step"] *= world
            res["timing"] += "; all %d ranks concurrently, max over ranks" % world
        line["e2e"] = res
    if rank == 0 and world == 1 and not args.no_extra:
        line["extra_workloads"] = extra_workloads(p, name, peak)
    if world > 1:
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, ALL_CPUS)  # the CPU baseline uses every host core
        line["cpu_baseline"] = cpu_baseline(name)[0]
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extra_workloads(p, main, peak):
    """Short device-resident runs (kernel time by CUDA events, best of 3 after a warm-up launch) of the other
    BASELINE.json configurations and of the headline config at 65536 chains; same byte accounting."""
    import torch
    out = {}
    todo = [(n, DEFAULT_CHAINS[n], DEFAULT_EVENTS[n]) for n in ("c1", "c2", "c3", "c4", "c5f", "c5b") if n != main]
    todo.append((main, 65536, 100))
    dev = torch.device("cuda")
    f64 = torch.float64
    for name, nch, n_ev in todo:
        cfgd = CONFIGS[name]; d = cfgd["d"]
        s = make_sampler(p, name)
        x0 = torch.full((nch, d), cfgd["x0"], dtype=f64, device=dev)
        v0 = torch.ones((nch, d), dtype=f64, device=dev) / (d ** 0.5 if cfgd["unit_v"] else 1.0)
        ch = p.DeviceChains(s, x0, v0, seed=2024)
        bufs = dict(X=torch.empty((nch, n_ev, d), dtype=f64, device=dev), V=torch.empty((nch, n_ev, d), dtype=f64, device=dev),
                    t=torch.empty((nch, n_ev), dtype=f64, device=dev), horizon=torch.empty((nch, n_ev), dtype=f64, device=dev),
                    ar=torch.empty((nch, n_ev), dtype=f64, device=dev),
                    error_value_ar=torch.empty((nch, n_ev, 5), dtype=f64, device=dev),
                    errored_bound=torch.empty((nch, n_ev), dtype=torch.int32, device=dev),
                    rejected=torch.empty((nch, n_ev), dtype=torch.int32, device=dev),
                    hitting_horizon=torch.empty((nch, n_ev), dtype=torch.int32, device=dev))
        view = p.device_history_view(n_ev, **bufs)
        st = torch.cuda.current_stream().cuda_stream
        ch.advance(n_ev, view, 0, st)
        best = float("inf")
        for _ in range(3):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); ch.advance(n_ev, view, 0, st); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        _, _, cnt = ch.status(); ch.close()
        gbs = nch * n_ev * bytes_per_event(d) / best / 1e9
        out[f"{name}@{nch}"] = {"workload": cfgd["desc"], "chains": nch, "events_per_chain": n_ev,
                                "events_per_s": nch * n_ev / best, "hbm_gbs": gbs, "hbm_frac": gbs / peak,
                                "bytes_per_event": bytes_per_event(d), "ms": best * 1e3}
        if name == "c4":
            # governing roofline is FP64 (north_star: DMMA): 2nd(2+2G) flop per bound build (z, w and the d x 2G product),
            # 2nd(2+1) per rate evaluation; counts come from the kernel's own counters (4 launches of n_ev events)
            n_rows, G = 100000, 10
            builds, rates = cnt[:, 0].sum() / 4.0, cnt[:, 1].sum() / 4.0
            flop = 2.0 * n_rows * d * ((2 + 2 * G) * builds + 3 * rates)
            out[f"{name}@{nch}"].update({"fp64_tflops": flop / best / 1e12, "fp64_peak_tflops": FP64_PEAK_TFLOPS,
                                         "fp64_frac": flop / best / 1e12 / FP64_PEAK_TFLOPS,
                                         "fp64_peak_source": "scratch/dmma_peak.cu DMMA microbenchmark on this pool's B200 (MEASURED_PEAKS.json has no FP64 entry)"})
        del bufs, view, ch
        torch.cuda.empty_cache()
    return out


def e2e(p, sampler, name, nch, n_ev, world, dev):
    """Same metric through the public C-ABI call with HOST buffers (pinned): every step copies the initial states
    host->device and the full history device->host inside the timed region.  Returns this rank's shard; the caller
    runs it on all ranks at once and takes the slowest."""
    import ctypes as C
    import numpy as np
    from pdmpflux_b200 import _lib as L
    lib = p.lib()
    cfgd = CONFIGS[name]
    d = cfgd["d"]
    n_sk = n_ev + 1

    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        L.check(lib.pdmpflux_host_alloc(C.byref(ptr), n))
        buf = (C.c_char * n).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), ptr

    arrs, ptrs = {}, []
    for f, shape, dt in (("X", (nch, n_sk, d), np.float64), ("V", (nch, n_sk, d), np.float64), ("t", (nch, n_sk), np.float64),
                         ("horizon", (nch, n_sk), np.float64), ("ar", (nch, n_sk), np.float64),
                         ("error_value_ar", (nch, n_sk, 5), np.float64), ("errored_bound", (nch, n_sk), np.int32),
                         ("rejected", (nch, n_sk), np.int32), ("hitting_horizon", (nch, n_sk), np.int32),
                         ("x0", (nch, d), np.float64), ("v0", (nch, d), np.float64)):
        arrs[f], ptr = pinned(shape, dt)
        ptrs.append(ptr)
    arrs["x0"][:] = cfgd["x0"]
    arrs["v0"][:] = 1.0 / (d ** 0.5 if cfgd["unit_v"] else 1.0)
    status = np.zeros(nch, dtype=np.int32)
    view = L.History(*(arrs[f].ctypes.data for f in ("X", "V", "t", "horizon", "ar", "error_value_ar", "errored_bound",
                                                     "rejected", "hitting_horizon")),
                     status.ctypes.data, None, None, n_sk, 0)

    def call(seed):
        L.check(lib.pdmpflux_sample_skeleton(sampler._handle, nch, n_sk, arrs["x0"].ctypes.data, arrs["v0"].ctypes.data,
                                             C.c_uint64(seed), 0, None, C.byref(view), None), status)

    import torch
    for w in range(4):  # warm-up calls: slab allocation, lazy module loading, PCIe / copy-engine ramp
        call(100 + w)
    torch.cuda.synchronize()
    # Each call is timed on its own and the median is reported: on these shared hosts the PCIe / host-memory leg is
    # occasionally 10-20x slower for a single call (other tenants), which a mean over 3 calls would inherit.
    reps, times = 7, []
    for i in range(reps):
        t0 = time.perf_counter()
        call(2 + i)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    times.sort()
    dt = times[len(times) // 2]
    ok = bool(np.isfinite(arrs["t"][:, -1]).all())
    for ptr in ptrs:
        lib.pdmpflux_host_free(ptr)
    h2d, d2h = C.c_int64(0), C.c_int64(0)
    lib.pdmpflux_last_transfer_bytes(C.byref(h2d), C.byref(d2h))   # counted by the library from the copies it issued
    return {"value": nch * n_ev / dt, "unit": "events/s", "h2d_bytes_per_step": int(h2d.value),
            "d2h_bytes_per_step": int(d2h.value), "history_bytes_per_step": nch * n_sk * bytes_per_event(d),
            "ms_per_step": dt * 1e3, "finite": ok,
            "timing": "median of %d individually timed calls (min %.1f ms, max %.1f ms)" % (reps, times[0] * 1e3, times[-1] * 1e3),
            "call": "pdmpflux_sample_skeleton (host buffers, pinned)"}


if __name__ == "__main__":
    main()

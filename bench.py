#!/usr/bin/env python3
"""bench.py -- batched bi-alignment throughput (cell-state updates per second) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (Cython)

One step = one pass of the hot path (forward fill + traceback) over one batch of synthetic pairs.
Workload: BASELINE.json config 3 ("100k protein pairs, length 200-500, random H/E/C structure,
max_shift 2, sharded over 8 GPUs") split evenly: 12 500 pairs per GPU and step, so N = 8 is exactly
config 3 and smaller N are its weak-scaling slices.  Pairs are independent: rank r aligns its LPT share,
no data-path collective.  `value` is measured with inputs resident in HBM; `e2e` goes through the public
host-buffer call (H2D of the sequence table and pair list, fill, traceback, D2H of scores and traces).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS_PER_GPU = 12500
MAX_SHIFT = 2
WORKLOAD = "cfg3: synthetic protein pairs len U{200..500}, H/E/C run structure, BLOSUM62, max_shift 2, score+traceback"


def cell_states_of(off, pa, pb, s):
    from bialign_b200.batch import cell_states

    lens = np.diff(off)
    return cell_states(lens[pa], lens[pb], s)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons of one GPU while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (oracle/_ref, Cython) when it was built, else the C oracle port
# ------------------------------------------------------------------------------------------------
def _ref_worker(job):
    seqA, seqB, strA, strB, params, want_trace = job
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import contextlib
    import io

    import bialignment  # the reference

    with contextlib.redirect_stdout(io.StringIO()):
        b = bialignment.BiAligner(seqA, seqB, strA, strB, nameA="A", nameB="B", **params)
        sc = int(b.optimize())
        if want_trace:
            b.traceback()
    return sc


def _port_worker(job):
    seqA, seqB, strA, strB, params, want_trace = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle

    return oracle.run(seqA, seqB, strA, strB, params, mode="codes")["score"]


def reference_available():
    d = os.path.join(ROOT, "oracle", "_ref")
    return os.path.isdir(d) and any(f.startswith("bialignment.") and f.endswith(".so") for f in os.listdir(d))


def cpu_sample(res, cls, off, pa, pb, params, trunc, npairs):
    """Bounded sample of the workload for the CPU arms: the first `npairs` pairs, both molecules cut to
    their first `trunc` residues (the reference runs ~7e4 cell-states/s/core; a full-length pair of
    this workload would take minutes per core)."""
    from bialign_b200 import workloads
    from bialign_b200.batch import cell_states

    jobs, cs = [], 0
    for p in range(npairs):
        a, sa = workloads.decode_protein(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[p]))
        if trunc:
            a, sa, b, sb = a[:trunc], sa[:trunc], b[:trunc], sb[:trunc]
        jobs.append((a, b, sa, sb, params, True))
        cs += int(cell_states(len(a), len(b), params["max_shift"]))
    return jobs, cs


class CpuArm:
    """A pool of host processes running the reference (or the C port) on a list of jobs.  The pool is created and its
    workers warmed (module import, one tiny alignment each) once, outside every timed window."""

    def __init__(self, cores, kind):
        import multiprocessing as mp

        self.kind, self.cores = kind, cores
        self.worker = _ref_worker if kind == "reference" else _port_worker
        self.pool = mp.get_context("fork").Pool(cores)
        tiny = ("ACDEF", "ACDF", "HHHEE", "HHEE", dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800,
                                                         gap_opening_cost=-150, gap_cost=-50, shift_cost=-150, max_shift=1), True)
        self.pool.map(self.worker, [tiny] * (2 * cores), chunksize=1)

    def run(self, jobs):
        t0 = time.perf_counter()
        scores = self.pool.map(self.worker, jobs, chunksize=1)
        return time.perf_counter() - t0, scores

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(res, cls, off, pa, pb, params):
    cores = os.cpu_count() or 1
    kind = "reference" if reference_available() else "port"
    if kind == "reference":
        trunc, npairs = 48, cores  # ~3.9e5 cell-states per pair -> ~6 s per core
    else:
        trunc, npairs = 0, cores * 2
    jobs, cs = cpu_sample(res, cls, off, pa, pb, params, trunc, npairs)
    arm = CpuArm(cores, kind)
    dt, _ = arm.run(jobs)
    arm.close()
    return {"value": cs / dt / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
            "sample": f"first {npairs} pairs of the workload" + (f", molecules cut to {trunc} residues" if trunc else "") +
                      f" ({cs} cell-states, {dt:.1f} s wall, warm pool of {cores} processes)"}


# ------------------------------------------------------------------------------------------------
def verify_shard(res, cls, off, pa, pb, params, scores, cols, toff, complete, nrescore=32):
    """Correctness of what the timed loops produced, checked with the CPU oracle (test infrastructure, used here only as
    the checker): every sampled trace is complete, ends at (n, m, n, m) and re-scores to the reported score; the
    smallest pair of the shard is recomputed from scratch (score and trace)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from bialign_b200 import workloads
    from bialign_b200.batch import trace_hex

    n = len(pa)
    ok, checked = True, 0
    for q in np.unique(np.linspace(0, n - 1, num=min(nrescore, n)).astype(np.int64)):
        a, sa = workloads.decode_protein(res, cls, off, int(pa[q]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[q]))
        v, end = oracle.eval_trace(a, b, sa, sb, params, trace_hex(cols, toff, q))
        ok &= bool(complete[q]) and v == int(scores[q]) and end == [len(a), len(b)] * 2
        checked += 1
    lens = np.diff(off)
    q = int(np.argmin((lens[pa] + 1) * (lens[pb] + 1)))
    a, sa = workloads.decode_protein(res, cls, off, int(pa[q]))
    b, sb = workloads.decode_protein(res, cls, off, int(pb[q]))
    r = oracle.run(a, b, sa, sb, params, mode="codes")
    ok &= r["score"] == int(scores[q]) and r["trace"] == trace_hex(cols, toff, q)
    return {"ok": bool(ok), "traces_rescored": checked, "pairs_recomputed_by_oracle": 1,
            "how": "oracle.eval_trace on sampled traces (score, end cell, completeness) + oracle.run on the smallest pair"}


def time_shape(al, res, cls, off, pa, pb, want_trace, reps, local_rank, barrier):
    """Device-resident timing of one workload shape on this rank: (best stats dict, wall ms of that run, clocks)."""
    eng = al.engine
    al.configure()
    eng.load_sequences(res, cls, off)
    eng.load_pairs(pa, pb)
    eng.run(want_trace=want_trace)  # warm-up (allocations, first-use)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    best, wall = None, 0.0
    t_begin, done = time.perf_counter(), 0
    while done < reps or time.perf_counter() - t_begin < 0.7:  # short shapes repeat until nvidia-smi has sampled the clocks
        t0 = time.perf_counter()
        eng.run(want_trace=want_trace)
        w = 1e3 * (time.perf_counter() - t0)
        st = eng.stats()
        done += 1
        if best is None or st["total_ms"] < best["total_ms"]:
            best, wall = st, w
    barrier()
    return best, wall, sampler.stop()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shapes", action="store_true", help="skip the per-shape block (cfg1, cfg2, cfg4, cfg5, non-affine) and the strong-scaling run")
    ap.add_argument("--warps", type=int, default=0, help="warps per CTA of the systolic kernel (0 = library default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, 1)
    if world != n_gpus:
        raise SystemExit(f"bench.py: --gpus {n_gpus} but WORLD_SIZE is {world}: launch N > 1 through torch.distributed.run "
                         "with one rank per GPU (the line's n_gpus, the workload size and the sharding must agree)")

    from bialign_b200 import workloads

    params = dict(workloads.PROTEIN_PARAMS, max_shift=MAX_SHIFT)
    total_pairs = args.pairs_per_gpu * n_gpus
    res, cls, off, pa, pb = workloads.protein_pairs(total_pairs, seed=3)
    config = {"workload": WORKLOAD, "pairs_per_step": total_pairs, "pairs_per_gpu": args.pairs_per_gpu,
              "sharding": f"LPT over {n_gpus} rank(s), no collective on the data path; results gathered on rank 0",
              "l2": "traceback-code stream per step (>= tens of GB) far exceeds the 126 MB L2; no explicit flush needed"}

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        kind = "reference" if reference_available() else "port"
        trunc, npairs = (40, cores) if kind == "reference" else (0, cores)
        jobs, cs = cpu_sample(res, cls, off, pa, pb, params, trunc, npairs)
        arm = CpuArm(cores, kind)  # processes forked and warmed before anything is timed
        times = []
        for it in range(args.warmup + args.steps):
            dt, _ = arm.run(jobs)
            if it >= args.warmup:
                times.append(dt)
        arm.close()
        ms = 1e3 * float(np.mean(times))
        val = cs / (ms * 1e-3) / 1e9
        sample = (f"each step = first {npairs} pairs of the workload" +
                  (f", molecules cut to {trunc} residues" if trunc else "") + f" ({cs} cell-states) on {cores} host cores")
        # the CPU arm times a bounded sample of the workload, not the workload: say so in its own config
        config = dict(config, reference_sample={"pairs": npairs, "truncated_to": trunc or None, "cell_states": cs,
                                                "comparable": "per-cell-state rate only (GCUPS); the GPU arm runs full-length pairs"})
        print(json.dumps({"impl": "reference", "metric": "batched bialign GCUPS (cell-states/s)", "value": val,
                          "unit": "GCUPS", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ---------------------------------------------------------------- this engine
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bialign_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    from bialign_b200 import _capi
    from bialign_b200.batch import BatchAligner, compact_shard, gather_scores, gather_traces, lpt_shards, pair_cost

    os.environ["BIALIGN_DEVICE"] = str(local_rank)
    dev = torch.device("cuda", local_rank)
    al = BatchAligner(device=local_rank, **params)
    if args.warps:
        al.set_option("warps_per_cta", args.warps)
    lens = np.diff(off)
    mine = lpt_shards(pair_cost(lens[pa], lens[pb], MAX_SHIFT), world)[rank]
    total_cs = int(cell_states_of(off, pa, pb, MAX_SHIFT).sum())
    # this rank's share of the sequence table and pair list (what it uploads every end-to-end step)
    res, cls, off, my_pa, my_pb = compact_shard(res, cls, off, pa[mine], pb[mine])
    my_cs = int(cell_states_of(off, my_pa, my_pb, MAX_SHIFT).sum())

    eng = al.engine
    al.configure()
    # --- device-resident leg: inputs already in HBM when the timed region starts
    eng.load_sequences(res, cls, off)
    eng.load_pairs(my_pa, my_pb)
    for _ in range(args.warmup):
        eng.run(want_trace=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev_ms, fill_ms, tb_ms, launches = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run(want_trace=True)
        st = eng.stats()
        ev_ms += st["total_ms"]
        fill_ms += st["fill_ms"]
        tb_ms += st["traceback_ms"]
        launches += st["kernel_launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    st = eng.stats()

    # --- end-to-end leg: host buffers in; scores + traces out and, with several ranks, gathered on rank 0 -- every step
    h2d = res.nbytes + cls.nbytes + off.nbytes + my_pa.nbytes + my_pb.nbytes
    d2h = 0
    scores = np.empty(len(my_pa), dtype=np.int64)

    def e2e_step():
        eng.align_batch(res, cls, off, my_pa, my_pb, want_trace=True, scores_out=scores)
        cols, toff, complete = eng.fetch_traces()
        nbytes = scores.nbytes + cols.nbytes + toff.nbytes + complete.nbytes  # what actually came back from the device
        if world > 1:  # the one cross-rank step of the path: a result gather (NCCL), no data-path collective
            full = gather_scores(mine, scores, total_pairs, device=dev)
            gcols, goff, gcomp = gather_traces(mine, cols, toff, complete, total_pairs, device=dev, dst=0)
            if rank == 0:
                nbytes += full.nbytes + gcols.nbytes
        return cols, toff, complete, nbytes

    e2e_step()  # warm (first-use allocations)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        cols, toff, complete, d2h = e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t1)
    verified = verify_shard(res, cls, off, my_pa, my_pb, params, scores, cols, toff, complete)
    verified["ok"] = bool(allsum(0.0 if verified["ok"] else 1.0) == 0.0)

    ev_ms_max, wall_ms_max, e2e_ms_max = allmax(ev_ms), allmax(wall_ms), allmax(e2e_ms)
    launches_all, h2d_all, d2h_all = allsum(launches), allsum(h2d), allsum(d2h)

    # --- roofline denominators: integer ALU, measured on this box
    add_rate, sms = _capi.microbench_int(local_rank, 0)
    fused_rate, _ = _capi.microbench_int(local_rank, 2)
    mnmx_rate, _ = _capi.microbench_int(local_rank, 1)
    int_peak = max(add_rate, mnmx_rate, 2.0 * fused_rate)  # algorithmic int ops/s (a fused add+max retires two)

    # --- every other named shape of BASELINE.json (device-resident, CUDA-event times, clocks sampled per shape)
    shapes, strong = None, None
    if not args.no_shapes:
        shapes = run_shapes(world, rank, local_rank, barrier, allmax, allsum, int_peak)
        # strong scaling: the whole of config 3 (100 000 pairs) at every N, one warm-up and one timed step
        sres, scls, soff, spa, spb = workloads.protein_pairs(PAIRS_PER_GPU * 8, seed=3)
        slens = np.diff(soff)
        smine = lpt_shards(pair_cost(slens[spa], slens[spb], MAX_SHIFT), world)[rank]
        s_cs = int(cell_states_of(soff, spa, spb, MAX_SHIFT).sum())
        sres, scls, soff, spa, spb = compact_shard(sres, scls, soff, spa[smine], spb[smine])
        best, wall, sclk = time_shape(al, sres, scls, soff, spa, spb, True, 1, local_rank, barrier)
        swall = allmax(wall)
        strong = {"workload": "all of cfg3: 100000 pairs at every N", "pairs": PAIRS_PER_GPU * 8, "n_gpus": n_gpus,
                  "ms_per_step": swall, "value": s_cs / swall / 1e6, "unit": "GCUPS", "steps": 1, "warmup": 1,
                  "sm_mhz": sclk.get("sm_mhz")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = wall_ms_max / args.steps  # barrier-to-barrier wall clock, max over ranks
    value = total_cs / (ms_per_step * 1e-3) / 1e9
    e2e_value = total_cs / (e2e_ms_max / args.steps * 1e-3) / 1e9

    fill_s = fill_ms / args.steps * 1e-3
    achieved = 30.0 * my_cs / fill_s  # 15 add + 15 max per cell-state
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "fill_traffic.json")) as fh:
            traffic = json.load(fh).get("dram_bytes_per_cell_state")
    except OSError:
        pass
    roofline = {"bound": "int_alu", "kernel": "fill_systolic_kernel<2,true>", "achieved": achieved / 1e12,
                "peak": int_peak / 1e12, "unit": "Tiop/s", "frac": achieved / int_peak,
                "peak_source": "measured here by ba_microbench_int (IADD %.2f, LOP3+VIMNMX %.2f, VIADDMNMX %.2f T thread-instr/s; MEASURED_PEAKS.json has no integer entry)"
                               % (add_rate / 1e12, mnmx_rate / 1e12, fused_rate / 1e12),
                "algorithmic_ops_per_cell_state": 30, "fill_ms_per_step": fill_ms / args.steps,
                "traceback_ms_per_step": tb_ms / args.steps, "kernel_share_of_step": fill_ms / max(ev_ms, 1e-9),
                "traffic": (traffic * my_cs if traffic else None),
                "hbm": {"achieved": st["code_bytes"] / fill_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": st["code_bytes"] / fill_s / 1e9 / hbm_peak,
                        "what": "traceback-code stores (a 6-byte slot per lane and iteration: 128 + 64 contiguous bytes per warp) during the fill",
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}}
    out = {"metric": "batched bialign GCUPS (cell-states/s)", "value": value, "unit": "GCUPS", "n_gpus": n_gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
           "device_ms_per_step": ev_ms_max / args.steps, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                   "ms_per_step": e2e_ms_max / args.steps,
                   "includes": "H2D of the sequence table and pair list, fill, traceback, D2H of scores and traces" +
                               (", NCCL gather of scores and traces on rank 0" if world > 1 else "")},
           "gpu_launches": int(launches_all), "kernel_kind": st["kernel_kind"], "waves_per_step": st["waves"],
           "verified": verified, "roofline": roofline}
    if shapes is not None:
        out["shapes"] = shapes
        out["strong_scaling"] = strong
    if n_gpus == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(res, cls, off, my_pa, my_pb, params)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_shapes(world, rank, local_rank, barrier, allmax, allsum, int_peak):
    """The other named shapes of BASELINE.json, each timed like the headline (device-resident inputs, CUDA events,
    clocks sampled while it runs).  Batches (cfg1, cfg4, non-affine) are weak-scaled: a fixed number of pairs per GPU,
    sharded like the headline.  Single long pairs (cfg2, cfg5) do not shard ("replicas only"): rank 0 runs
    them on its GPU and the other ranks wait.  Returns a list of dicts (identical on every rank)."""
    from bialign_b200 import workloads
    from bialign_b200.batch import BatchAligner, cell_states

    prot = workloads.PROTEIN_PARAMS
    out = []

    def record(name, al, res, cls, off, pa, pb, want_trace, reps, ops, sharded, note=None):
        affine = al.params["gap_opening_cost"] != 0
        active = sharded or rank == 0
        if active:
            best, wall, clk = time_shape(al, res, cls, off, pa, pb, want_trace, reps, local_rank, barrier if sharded else (lambda: None))
        else:
            best, wall, clk = {"cell_states": 0, "fill_ms": 0.0, "traceback_ms": 0.0, "total_ms": 0.0, "kernel_kind": 0,
                               "warps_per_cta": 0, "waves": 0, "code_bytes": 0}, 0.0, {}
        if not sharded:
            barrier()
        cs = allsum(float(best["cell_states"]))
        wall_max, fill_max, dev_max = allmax(wall), allmax(best["fill_ms"]), allmax(best["total_ms"])
        sm = allmax(float(clk.get("sm_mhz") or 0.0))
        reasons = sorted(set(clk.get("reasons", []))) if rank == 0 else []
        kk, wp = int(allmax(float(best["kernel_kind"]))), int(allmax(float(best["warps_per_cta"])))
        out.append({"shape": name, "pairs": int(allsum(float(len(pa) if active else 0))), "n_gpus": world if sharded else 1,
                    "want_trace": bool(want_trace), "model": "affine" if affine else "non-affine",
                    "cell_states": int(cs), "gcups": cs / wall_max / 1e6, "ms": wall_max, "device_ms": dev_max,
                    "fill_ms": fill_max, "gcups_fill": cs / fill_max / 1e6,
                    "frac": ops * cs / (fill_max * 1e-3) / int_peak / (world if sharded else 1),
                    "algorithmic_ops_per_cell_state": ops, "kernel_kind": kk, "warps_per_cta": wp,
                    "sm_mhz": sm, "clock_reasons": reasons, "scaling": "weak" if sharded else "replicas only (one pair, one GPU)",
                    **({"note": note} if note else {})})

    # cfg1: the README toy pair (42 aa, max_shift 1), 100 000 copies per GPU, score + traceback
    seqs = ["RAKLPLKEKKLTATANYHPGIRYIMTGYSAKYIYSSTYARFR", "KAKLPLKEKKLTRTANYHPGIRYIMTGYSAKRIYSSTYAYFR"]
    structs = ["CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEECCC", "HHHHHHHHHHHHCCCCCCTCEEEEEEECCCCCEEEEEEEECC"]
    al = BatchAligner(device=local_rank, max_shift=1, **prot)
    res, cls, off = al.encode(seqs, structs)
    record("cfg1: README protein toy pair x 100000 per GPU, max_shift 1, score+traceback", al, res, cls, off,
           np.zeros(100000, np.int32), np.ones(100000, np.int32), True, 2, 30, True)
    # cfg2: DNAPolymerase1 E. coli vs Xanthomonas (928 x 933), max_shift 1, full traceback
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "dnapol1.json")))
    al = BatchAligner(device=local_rank, **g["params"])
    res, cls, off = al.encode([g["seqA"], g["seqB"]], [g["strA"], g["strB"]])
    record("cfg2: DNAPolymerase1 928 x 933, max_shift 1, score+traceback", al, res, cls, off,
           np.array([0], np.int32), np.array([1], np.int32), True, 3, 30, False)
    # cfg4: RNA pairs of length 120 with supplied structures, max_shift 2, score only: 125 000 per GPU (1M on 8)
    al = BatchAligner(device=local_rank, max_shift=2, **workloads.RNA_PARAMS)
    res, cls, off, pa, pb = workloads.rna_pairs(125000, seed=4 + 100 * rank)
    record("cfg4: 125000 RNA pairs len 120 per GPU, max_shift 2, score only (16-bit pair mode)", al, res, cls, off, pa, pb,
           False, 2, 30, True)
    # non-affine model (the CLI default, gap_opening_cost 0) on the cfg3 slice: 12 500 pairs per GPU, score + traceback
    # (3000 pairs, the shape of earlier rounds, leave 1776 resident one-warp CTAs 1.7 pairs each: a third of that run is tail)
    al = BatchAligner(device=local_rank, max_shift=2, **dict(prot, gap_opening_cost=0, gap_cost=-200, shift_cost=-250))
    res, cls, off, pa, pb = workloads.protein_pairs(12500, seed=3 + 100 * rank)
    record("non-affine model: 12500 protein pairs 200-500 per GPU, max_shift 2, score+traceback (cells/s)", al, res, cls, off,
           pa, pb, True, 2, 26, True)
    # long pairs in a batch (about 90 fit one traceback-memory wave): a gang of CTAs per pair, several pairs per launch
    al = BatchAligner(device=local_rank, max_shift=2, **prot)
    res, cls, off, pa, pb = workloads.protein_pairs(400, lo=1900, hi=2100, seed=6 + 100 * rank)
    record("long pairs in a batch: 400 protein pairs 1900-2100 per GPU, max_shift 2, score+traceback (long-pair gangs)", al, res, cls,
           off, pa, pb, True, 2, 30, True)
    # scoring without a common divisor on the same pairs: value << tie bits does not fit 32 bits -> rebased two-launch run
    al = BatchAligner(device=local_rank, max_shift=2, **dict(prot, structure_weight=333, gap_opening_cost=-157, gap_cost=-49, shift_cost=-151))
    record("wide score range (structure_weight 333, costs -157/-49/-151: gcd 1): the same 400 pairs, score+traceback (rebased run)", al,
           res, cls, off, pa, pb, True, 2, 30, True,
           note="two launches (exact score-only + trace relative to row maxima); GCUPS counts the cell-states once")
    # cfg5: one 8192 x 8192 protein pair, max_shift 3, multi-CTA fill with traceback codes in HBM
    al = BatchAligner(device=local_rank, max_shift=3, **prot)
    res, cls, off, pa, pb = workloads.protein_pairs(1, lo=8192, hi=8192, seed=5)
    record("cfg5: one 8192 x 8192 protein pair, max_shift 3, score+traceback", al, res, cls, off, pa, pb, True, 2, 30, False)
    return out


if __name__ == "__main__":
    main()

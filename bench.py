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

    jobs, cs = [], 0
    from bialign_b200.batch import cell_states

    for p in range(npairs):
        a, sa = workloads.decode_protein(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[p]))
        if trunc:
            a, sa, b, sb = a[:trunc], sa[:trunc], b[:trunc], sb[:trunc]
        jobs.append((a, b, sa, sb, params, True))
        cs += int(cell_states(len(a), len(b), params["max_shift"]))
    return jobs, cs


def run_cpu_arm(jobs, cores, kind):
    import multiprocessing as mp

    worker = _ref_worker if kind == "reference" else _port_worker
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        scores = pool.map(worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return dt, scores


def cpu_baseline(res, cls, off, pa, pb, params, budget_s=20.0):
    cores = os.cpu_count() or 1
    kind = "reference" if reference_available() else "port"
    if kind == "reference":
        trunc, npairs = 48, cores  # ~3.9e5 cell-states per pair -> ~6 s per core
    else:
        trunc, npairs = 0, cores * 2
    jobs, cs = cpu_sample(res, cls, off, pa, pb, params, trunc, npairs)
    dt, _ = run_cpu_arm(jobs, cores, kind)
    return {"value": cs / dt / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
            "sample": f"first {npairs} pairs of the workload" + (f", molecules cut to {trunc} residues" if trunc else "") +
                      f" ({cs} cell-states, {dt:.1f} s wall, multiprocessing pool of {cores})"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--warps", type=int, default=0, help="warps per CTA of the systolic kernel (0 = library default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, 1)

    from bialign_b200 import workloads

    params = dict(workloads.PROTEIN_PARAMS, max_shift=MAX_SHIFT)
    total_pairs = args.pairs_per_gpu * n_gpus
    res, cls, off, pa, pb = workloads.protein_pairs(total_pairs, seed=3)
    config = {"workload": WORKLOAD, "pairs_per_step": total_pairs, "pairs_per_gpu": args.pairs_per_gpu,
              "sharding": f"LPT over {n_gpus} rank(s), no collective on the data path",
              "l2": "traceback-code stream per step (>= tens of GB) far exceeds the 126 MB L2; no explicit flush needed"}

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        kind = "reference" if reference_available() else "port"
        trunc, npairs = (40, cores) if kind == "reference" else (0, cores)
        jobs, cs = cpu_sample(res, cls, off, pa, pb, params, trunc, npairs)
        times = []
        for it in range(args.warmup + args.steps):
            dt, _ = run_cpu_arm(jobs, cores, kind)
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
        val = cs / (ms * 1e-3) / 1e9
        sample = (f"each step = first {npairs} pairs of the workload" +
                  (f", molecules cut to {trunc} residues" if trunc else "") + f" ({cs} cell-states) on {cores} host cores")
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

    from bialign_b200 import _capi
    from bialign_b200.batch import BatchAligner, compact_shard, lpt_shards, pair_cost

    os.environ["BIALIGN_DEVICE"] = str(local_rank)
    al = BatchAligner(device=local_rank, **params)
    if args.warps:
        al.set_option("warps_per_cta", args.warps)
    lens = np.diff(off)
    mine = lpt_shards(pair_cost(lens[pa], lens[pb], MAX_SHIFT), world)[rank]
    total_cs = int(cell_states_of(off, pa, pb, MAX_SHIFT).sum())
    # this rank's share of the sequence table and pair list (what it uploads every end-to-end step)
    res, cls, off, my_pa, my_pb = compact_shard(res, cls, off, pa[mine], pb[mine])
    lens = np.diff(off)
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

    # --- end-to-end leg: host buffers in, scores + traces out, every step
    h2d = res.nbytes + cls.nbytes + off.nbytes + my_pa.nbytes + my_pb.nbytes
    d2h = 0
    scores = np.empty(len(my_pa), dtype=np.int64)
    eng.align_batch(res, cls, off, my_pa, my_pb, want_trace=True, scores_out=scores)  # warm (first-use allocations)
    eng.fetch_traces()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        eng.align_batch(res, cls, off, my_pa, my_pb, want_trace=True, scores_out=scores)
        cols, toff, complete = eng.fetch_traces()
        d2h = scores.nbytes + int(np.sum(2 * (lens[my_pa] + lens[my_pb]) + 2)) + 5 * len(my_pa)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t1)

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

    ev_ms_max, wall_ms_max, e2e_ms_max = allmax(ev_ms), allmax(wall_ms), allmax(e2e_ms)
    launches_all, h2d_all, d2h_all = allsum(launches), allsum(h2d), allsum(d2h)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = wall_ms_max / args.steps  # barrier-to-barrier wall clock, max over ranks
    value = total_cs / (ms_per_step * 1e-3) / 1e9
    e2e_value = total_cs / (e2e_ms_max / args.steps * 1e-3) / 1e9

    # --- roofline of the dominant kernel (the fill): integer ALU, measured on this box
    add_rate, sms = _capi.microbench_int(local_rank, 0)
    fused_rate, _ = _capi.microbench_int(local_rank, 2)
    mnmx_rate, _ = _capi.microbench_int(local_rank, 1)
    int_peak = max(add_rate, mnmx_rate, 2.0 * fused_rate)  # algorithmic int ops/s (a fused add+max retires two)
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
                        "what": "traceback-code stores (8 B per band cell) during the fill",
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}}
    out = {"metric": "batched bialign GCUPS (cell-states/s)", "value": value, "unit": "GCUPS", "n_gpus": n_gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
           "device_ms_per_step": ev_ms_max / args.steps, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                   "ms_per_step": e2e_ms_max / args.steps},
           "gpu_launches": int(launches_all), "kernel_kind": st["kernel_kind"], "waves_per_step": st["waves"],
           "roofline": roofline}
    if n_gpus == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(res, cls, off, my_pa, my_pb, params)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the hot path: Ising couplings built/s (extraction) and spin flips/s (SA
sweeps) on B200, as a fraction of the measured HBM roofline, next to the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], the shape the metric is quoted on; it fits one GPU):
heisenberg_kagome_36-shaped U(1) basis (36 spins, 72 bonds), 10^7 sampled states PER GPU
(weak scaling: the global sorted basis has N x 10^7 states and every rank builds the CSR
rows of its contiguous row block against the full basis), synthetic log-normal amplitudes,
cluster-closed sampled subset (about a tenth of all candidates are hits).  A step is one
pass: [N>1: exchange X1 of basis words + amplitudes over NVLink peer memory] -> index
(first-position table + Bloom filter) -> single-pass extraction kernel (bit-plane applicability,
filter pre-sieve, exact search of the survivors, decoupled look-back, CSR written in place).
At N > 1 the steps are pipelined two deep by default (--pipeline 2): they are independent
extractions (one per cluster in the reference's experiment), so the copy engines gather the basis
of step k+1 while step k is indexed and extracted; --pipeline 1 runs exchange and extraction
strictly one after the other with X1 fused into the index kernel (asp_gather_index).
The annealing stage is timed separately on the same extracted model and reported under the
"anneal" key.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SYSTEM = "heisenberg_kagome_36"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")
CSRC = os.path.join(ROOT, "annealing-sign-problem_b200", "csrc")


def source_hash(files):
    """sha256 over the named kernel sources: a traffic figure measured with ncu belongs to one version of a kernel."""
    import hashlib

    h = hashlib.sha256()
    for name in files:
        h.update(open(os.path.join(CSRC, name), "rb").read())
    return h.hexdigest()


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind `key`, from profiles/traffic.json
    (written by tools/traffic_from_ncu.py out of an `ncu --set full` capture of this very workload).  The entry names the
    sources it was measured on; if they have changed since, the figure is stale and None is reported instead."""
    try:
        entry = json.load(open(TRAFFIC_FILE))[key]
    except Exception:  # noqa: BLE001
        return None, "no ncu capture recorded for %s in profiles/traffic.json" % key
    if entry.get("source_sha256") != source_hash(entry["sources"]):
        return None, "profiles/traffic.json entry for %s was measured on an older version of %s" % (key, ", ".join(entry["sources"]))
    return int(entry["dram_bytes"]), entry.get("capture", "")


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--states", type=int, default=10_000_000, help="sampled basis states per GPU")
    p.add_argument("--replicas", type=int, default=64, help="annealing replicas per GPU")
    p.add_argument("--sweeps", type=int, default=16, help="annealing sweeps per step")
    p.add_argument("--cpu-sample", type=int, default=200_000, help="states of the bounded CPU-baseline sample")
    p.add_argument("--exchange", default="peer", choices=["peer", "peer-sm", "peer-ce", "nccl"],
                   help="X1 at N > 1: 'peer' = row blocks in NVLink peer memory, pulled by the copy engines and indexed block by "
                        "block behind them; 'peer-sm' = one SM kernel pulls + indexes; 'nccl' = two all-gathers + index pass (the baseline)")
    p.add_argument("--pipeline", type=int, default=2, choices=[1, 2],
                   help="N > 1 with a peer exchange: 2 (default) = steps are independent extractions (one per cluster in the "
                        "reference's experiment): the copy engines gather the basis of step k+1 (asp_gather_blocks, no SM) while "
                        "step k is indexed and extracted; 1 = strictly serial, X1 fused with the index build in one kernel "
                        "(asp_gather_index)")
    p.add_argument("--pipeline-slots", type=int, default=2, choices=[2, 3],
                   help="private copies of the basis in the pipelined exchange: the exchange runs up to slots - 1 steps ahead")
    p.add_argument("--copy-streams", type=int, default=2, help="copy engines the pipelined exchange pulls with (blocks in flight at a time)")
    p.add_argument("--skip-anneal", action="store_true")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--skip-strong", action="store_true", help="N > 1: skip the strong-scaling measurement (10^7 states in total)")
    p.add_argument("--skip-configs", action="store_true", help="skip the other BASELINE.json configurations (the `configs` object)")
    p.add_argument("--skip-python", action="store_true", help="skip the end-to-end measurement through common.make_ising_model")
    return p.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every 5 ms DURING the timed region."""

    def __init__(self, gpu_uuid, gpu_index):
        self.uuid, self.index = gpu_uuid, gpu_index
        self.samples, self.reason_bits = [], 0
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.error = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode() if isinstance(self.uuid, str) else self.uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as exc:  # pragma: no cover
            self.error = repr(exc)

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception as exc:  # pragma: no cover
                self.error = repr(exc)
                return
            time.sleep(0.001)

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        names = []
        if self.thread is not None:
            nv = self.nv
            table = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4)]
            for name, attr, default in table:
                if self.reason_bits & int(getattr(nv, attr, default)):
                    names.append(name)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "samples": len(self.samples), "reasons": names}
        if self.error:
            out["error"] = self.error
        return out


def u1_operator(asp):
    """kagome_36 bond list on the U(1)-only basis (the symmetrised variant is
    integer-ALU-bound and measured by the tests, not the headline)."""
    cfg = asp.ls.load_config(asp.ls.system_path(SYSTEM))
    cfg["basis"]["symmetries"] = []
    cfg["basis"]["spin_inversion"] = None
    basis = asp.ls.SpinBasis.load_from_yaml(cfg["basis"])
    return asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], basis), cfg


# ------------------------------------------------------------------------------------------
# CPU legs (rank 0): the reference's own C (oracle/_ref) on a bounded sample
# ------------------------------------------------------------------------------------------
def cpu_sample_inputs(n_sample):
    """Inputs of the CPU legs, made WITHOUT the product: the oracle's numpy operator and its numpy twin of the
    synthetic generator (oracle/synthetic_np.py) -- same shape of workload (kagome_36 bond list on the U(1) basis,
    cluster-closed subset, ~12 % of the candidates are hits), smaller sample."""
    from oracle import synthetic_np
    from oracle.operator_np import OperatorNP, load_config, system_path

    cfg = load_config(system_path(SYSTEM))
    cfg["basis"]["symmetries"] = []
    cfg["basis"]["spin_inversion"] = None
    op_np = OperatorNP.from_config(cfg)
    spins = synthetic_np.cluster_closed_states(op_np, n_sample, 1234)
    psi = synthetic_np.synthetic_amplitudes(spins.shape[0], 1234)
    other_spins, other_coeffs, other_counts = op_np.apply_u64(spins)
    idx = np.clip(np.searchsorted(spins, other_spins), 0, spins.shape[0] - 1)
    other_psi = np.where(spins[idx] == other_spins, psi[idx], 0.0)
    return spins, psi, other_spins, other_coeffs, other_counts, other_psi


def cpu_extract_once(capi, inputs, impl):
    spins, psi, other_spins, other_coeffs, other_counts, other_psi = inputs
    from oracle.capi import pad512

    s512, o512 = pad512(spins), pad512(other_spins)  # the reference ABI takes 512-bit keys
    counts = np.ones(spins.shape[0], dtype=np.int64)
    t0 = time.perf_counter()
    rows, cols, vals, field = capi.build_matrix(s512, counts, psi, o512, other_coeffs, other_counts, other_psi, impl=impl)
    dt = time.perf_counter() - t0
    return rows.shape[0], dt


def cpu_anneal_once(capi, inputs, sweeps):
    """Oracle SA (our restatement -- the reference annealer is third-party Haskell, absent)."""
    from oracle import live_path

    spins, psi, other_spins, other_coeffs, other_counts, other_psi = inputs
    counts = np.ones(spins.shape[0], dtype=np.int64)
    rows, cols, vals, _ = capi.build_matrix(spins, counts, psi, other_spins, other_coeffs, other_counts, other_psi, impl="port64")
    n = spins.shape[0]
    indptr, indices, data = capi.canonical_csr(n, rows, cols, vals)
    cores = os.cpu_count() or 1
    reps = max(cores, 8)
    betas = live_path.default_betas(indptr, indices, data, None, sweeps)
    t0 = time.perf_counter()
    capi.anneal(indptr, indices, data, None, reps, betas, seed=1, threads=cores)
    dt = time.perf_counter() - t0
    return reps * sweeps * n / dt, cores, reps


def cpu_greedy_once(capi, inputs):
    """Oracle greedy solver (our restatement of the rules the reference preserves at common.py:298-438) on the sample model."""
    spins, psi, other_spins, other_coeffs, other_counts, other_psi = inputs
    counts = np.ones(spins.shape[0], dtype=np.int64)
    rows, cols, vals, _ = capi.build_matrix(spins, counts, psi, other_spins, other_coeffs, other_counts, other_psi, impl="port64")
    n = spins.shape[0]
    indptr, indices, data = capi.canonical_csr(n, rows, cols, vals)
    t0 = time.perf_counter()
    capi.greedy(indptr, indices, data, None)
    return n, time.perf_counter() - t0


def cpu_live_path_once(inputs):
    """The reference's LIVE extraction path (common.py:131-208: batched_apply -> searchsorted -> couplings -> scipy
    0.5 (M + M^T) -> COO) in the oracle's numpy/scipy restatement, neighbour generation included -- the second CPU
    figure BASELINE.md section 3 plans.  The reference's own file needs numba and /root/reference, neither of which is
    on the GPU box, so this is labelled "port"."""
    from oracle import live_path
    from oracle.operator_np import OperatorNP, load_config, system_path

    cfg = load_config(system_path(SYSTEM))
    cfg["basis"]["symmetries"] = []
    cfg["basis"]["spin_inversion"] = None
    op_np = OperatorNP.from_config(cfg)
    spins, psi = inputs[0], inputs[1]
    t0 = time.perf_counter()
    with np.errstate(divide="ignore"):
        model = live_path.make_ising_model(spins, op_np, log_psi=np.log(psi.astype(np.complex128)))
    dt = time.perf_counter() - t0
    return int(model.exchange.nnz), dt


def cpu_extract_threaded(capi, inputs, impl, threads):
    """The reference's C (re-entrant, no globals: cbits/build_matrix.c:22-53) on `threads` host threads: every thread gets
    a contiguous block of rows -- its own candidate slice and counts that are zero outside the block -- and its own
    output arrays, exactly what a caller with a thread pool around the reference's function would do (ctypes releases
    the GIL for the call).  -> (couplings, seconds of the slowest-to-finish pool)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle.capi import pad512

    spins, psi, other_spins, other_coeffs, other_counts, other_psi = inputs
    n = spins.shape[0]
    s512 = pad512(spins)
    counts = np.ones(n, dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(other_counts)])
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    jobs = []
    for t in range(threads):
        lo, hi = int(bounds[t]), int(bounds[t + 1])
        block_counts = np.zeros(n, dtype=np.int64)
        block_counts[lo:hi] = other_counts[lo:hi]
        sl = slice(int(offsets[lo]), int(offsets[hi]))
        jobs.append((pad512(other_spins[sl]), np.ascontiguousarray(other_coeffs[sl]), block_counts, np.ascontiguousarray(other_psi[sl])))

    def work(job):
        o512, coeffs, block_counts, o_psi = job
        return capi.build_matrix(s512, counts, psi, o512, coeffs, block_counts, o_psi, impl=impl)[0].shape[0]

    with ThreadPoolExecutor(max_workers=threads) as pool:
        t0 = time.perf_counter()
        nnz = sum(pool.map(work, jobs))
        dt = time.perf_counter() - t0
    return nnz, dt


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (cbits/build_matrix.c compiled where it
    lies, oracle/_ref) on a bounded sample of the same workload, on all host cores (one row block per thread)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import capi  # nothing of the product is imported on this arm

    capi.build()
    impl = "ref" if capi.have_ref() else "port"
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    inputs = cpu_sample_inputs(args.cpu_sample)
    nnz, times = 0, []
    for step in range(args.warmup + args.steps):
        nnz, dt = cpu_extract_threaded(capi, inputs, impl, cores)
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = nnz * args.steps / total
    nnz_1, dt_1 = cpu_extract_once(capi, inputs, impl)
    sample = ("%d-state kagome_36-shaped cluster-closed subset, %d candidates; the reference's C function (serial, re-entrant) called on %d "
              "threads, one contiguous row block each; time of the calls only (neighbour lists precomputed: the reference takes them from "
              "lattice_symmetries, absent here)" % (inputs[0].shape[0], inputs[2].shape[0], cores))
    line = {
        "impl": "reference", "metric": "ising_couplings_built_per_sec", "value": value, "unit": "couplings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "heisenberg_kagome_36-shaped U(1) basis, Ising extraction (bounded CPU sample)", "states": int(inputs[0].shape[0]),
                   "candidates": int(inputs[2].shape[0]), "couplings": int(nnz)},
        "cpu_baseline": {"value": value, "unit": "couplings/s", "cores": cores, "kind": "reference" if impl == "ref" else "port", "sample": sample,
                         "one_thread": {"value": nnz_1 / dt_1, "unit": "couplings/s"}},
        "e2e": {"value": value, "unit": "couplings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "candidates_per_sec": inputs[2].shape[0] * args.steps / total,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    import annealing_sign_problem_b200 as asp
    from annealing_sign_problem_b200 import common, synthetic
    from annealing_sign_problem_b200 import distributed as D
    from annealing_sign_problem_b200._lib import ffi, lib

    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    peak, peak_src = hbm_peak()
    op, cfg = u1_operator(asp)
    lib().asp_set_copy_streams(args.copy_streams)

    def measure_extraction(states_per_rank, with_e2e):
        """One full measurement of the extraction step on a basis of `states_per_rank` sampled states per rank:
        workload, exchange set-up, warm-up, the timed steps, the roofline figures and (optionally) the end-to-end leg."""
        # ---- workload: every rank samples its own cluster-closed subset; X1 + global sort --------
        mine = synthetic.cluster_closed_states(op, states_per_rank, 1000 + rank, dev)
        if world > 1:
            import torch.distributed as dist

            every = torch.empty(world * mine.shape[0], dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(every, mine)
            spins = synthetic._sorted_unique_unsigned(every)
            del every
        else:
            spins = mine
        del mine
        n_total = int(spins.shape[0])
        psi = synthetic.synthetic_amplitudes(n_total, 77, device=dev)
        row_begin, num_rows = D.block(n_total, rank, world)
        my_spins = spins[row_begin:row_begin + num_rows].clone()
        my_psi = psi[row_begin:row_begin + num_rows].clone()
        need = int(lib().asp_extract_csr_workspace_bytes(op.handle, n_total, num_rows))
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        # caller-sized outputs (the reference's C contract): the first pass finds the coupling count,
        # every later pass allocates that much (+1/16 slack) and makes ONE kernel launch
        first = common.extract_csr_device(op, spins, psi, row_begin, num_rows, workspace=workspace)
        nnz_known = int(first[1].numel())
        capacity = nnz_known + nnz_known // 16
        del first

        bounds = [D.block(n_total, r, world)[0] for r in range(world)] + [n_total]
        peer = None
        exchange_used = "none" if world == 1 else args.exchange
        if world > 1 and args.exchange != "nccl":
            try:
                peer = D.PeerBasis(max(bounds[r + 1] - bounds[r] for r in range(world)), dev,
                                   mode={"peer": "tma", "peer-sm": "sm", "peer-ce": "ce"}[args.exchange])
            except D.PeerMemoryUnavailable as exc:  # raised on every rank together: all ranks take the NCCL exchange
                if rank == 0:
                    print("bench: peer memory unavailable (%s); X1 falls back to NCCL all-gather" % exc, file=sys.stderr)
                exchange_used = "nccl (peer memory unavailable)"
        if peer is not None:
            peer.spins[:num_rows] = my_spins
            peer.psi[:num_rows] = my_psi

        def one_pass(timers=None):
            ex = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if (timers is not None and world > 1) else None
            indexed = False
            if peer is not None:  # X1 fused with the index build, over NVLink peer memory
                peer.begin_epoch()  # every peer has finished reading my previous block ...
                peer.publish()      # ... which a real caller would now have rewritten in place
                if ex:
                    ex[0].record()
                full_spins, full_psi = peer.gather_index(op, bounds, num_rows, workspace)
                if ex:
                    ex[1].record()
                    exchange.append(ex)
                peer.release()
                indexed = True
            elif world > 1:  # X1 through NCCL: two all-gathers, then asp_extract_csr indexes the result
                if ex:
                    ex[0].record()
                full_spins = D.all_gather_blocks(my_spins, n_total)
                full_psi = D.all_gather_blocks(my_psi, n_total)
                if ex:
                    ex[1].record()
                    exchange.append(ex)
            else:
                full_spins, full_psi = spins, psi
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timers is not None else None
            indptr = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
            indices = torch.empty(capacity, dtype=torch.int32, device=dev)
            data = torch.empty(capacity, dtype=torch.float64, device=dev)
            if ev:
                ev[0].record()
            # timed passes do not read the count back (h_nnz = NULL: no host round trip; the count is indptr[-1],
            # checked after the timed region) -- the capacity is known from the sizing pass
            nnz = ffi.new("uint64_t *") if timers is None else ffi.NULL
            extract = lib().asp_extract_csr_indexed if indexed else lib().asp_extract_csr
            common.check(extract(op.handle, n_total, common.ptr(full_spins, "uint64_t *"), common.ptr(full_psi, "double *"),
                                 row_begin, num_rows, common.ptr(workspace, "void *"), workspace.numel(), capacity,
                                 common.ptr(indptr, "int64_t *"), common.ptr(indices, "int32_t *"),
                                 common.ptr(data, "double *"), nnz, common.stream()))
            if ev:
                ev[1].record()
                timers.append(ev)
            if timers is not None:
                return indptr, indices, data
            m = int(nnz[0])
            return indptr, indices[:m], data[:m]

        pipelined = peer is not None and args.pipeline == 2
        if pipelined:
            s_exchange, s_compute = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def run_pipelined(count, timers=None):
            """Software pipeline over `count` independent steps (in the reference's experiment: one extraction per cluster).
            Exchange stream: [begin_epoch, publish, asp_gather_blocks, release] of the steps ahead -- the copy engines pull
            the row blocks over NVLink, no SM involved -- while the compute stream runs asp_extract_csr (index + extraction)
            of step k.  `depth` private copies of the basis; a copy is gathered into again only after the extraction that
            read it, so the exchange runs up to depth - 1 steps ahead of the extraction."""
            depth = args.pipeline_slots
            gathered = [torch.cuda.Event() for _ in range(depth)]
            extracted = [None] * depth
            fulls = [None] * depth
            here = torch.cuda.current_stream()
            s_exchange.wait_stream(here)
            s_compute.wait_stream(here)

            def exchange_step(k):
                with torch.cuda.stream(s_exchange):
                    if extracted[k % depth] is not None:
                        s_exchange.wait_event(extracted[k % depth])
                    ex = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timers is not None else None
                    peer.begin_epoch()
                    peer.publish()
                    if ex:
                        ex[0].record()
                    fulls[k % depth] = peer.gather_blocks(bounds, slot=k % depth)
                    if ex:
                        ex[1].record()
                        exchange.append(ex)
                    peer.release()
                    gathered[k % depth].record()

            out = None
            for k in range(min(depth - 1, count)):
                exchange_step(k)
            for k in range(count):
                if k + depth - 1 < count:
                    exchange_step(k + depth - 1)  # queued first: its flag kernels must not wait behind the extraction
                with torch.cuda.stream(s_compute):
                    s_compute.wait_event(gathered[k % depth])
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timers is not None else None
                    indptr = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
                    indices = torch.empty(capacity, dtype=torch.int32, device=dev)
                    data = torch.empty(capacity, dtype=torch.float64, device=dev)
                    if ev:
                        ev[0].record()
                    full_spins, full_psi = fulls[k % depth]
                    common.check(lib().asp_extract_csr(op.handle, n_total, common.ptr(full_spins, "uint64_t *"), common.ptr(full_psi, "double *"),
                                                       row_begin, num_rows, common.ptr(workspace, "void *"), workspace.numel(), capacity,
                                                       common.ptr(indptr, "int64_t *"), common.ptr(indices, "int32_t *"),
                                                       common.ptr(data, "double *"), ffi.NULL, common.stream()))
                    if ev:
                        ev[1].record()
                        timers.append(ev)
                    extracted[k % depth] = torch.cuda.Event()
                    extracted[k % depth].record()
                    out = (indptr, indices, data)
            here.wait_stream(s_compute)
            here.wait_stream(s_exchange)
            return out

        for _ in range(max(1, args.warmup)):  # at least one untimed pass: it also yields the coupling count
            out = one_pass()
        nnz_mine = int(out[1].numel())
        x1_parity = None
        if world > 1:
            # the rank's CSR rows behind the exchange, against the same rows built WITHOUT it (every rank still holds the
            # full basis the workload was made from): row starts, columns and values bit for bit, on every rank
            import torch.distributed as dist

            unsharded = common.extract_csr_device(op, spins, psi, row_begin, num_rows)
            same = all(torch.equal(a, b) for a, b in zip(out, unsharded))
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            x1_parity = bool(int(flag[0]))
            del unsharded
        del out
        if pipelined:
            out = run_pipelined(max(2, args.warmup))
            torch.cuda.synchronize()
            assert int(out[0][-1]) == nnz_mine
            del out
        uuid = getattr(torch.cuda.get_device_properties(dev), "uuid", None)
        sampler = ClockSampler("GPU-" + str(uuid) if uuid else "", local)
        if rank == 0:
            sampler.start()
        launches0 = int(lib().asp_kernel_launch_count())
        timers, kernel_only, exchange = [], [], []
        lib().asp_debug_time_extract_kernel(1)  # CUDA events around extract_csr_kernel alone, on its own stream
        D.barrier()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        if pipelined:
            out = run_pipelined(args.steps, timers)
        else:
            for _ in range(args.steps):
                out = one_pass(timers)
        end.record()
        D.barrier()
        torch.cuda.synchronize()
        total_ms = D.max_over_ranks(start.elapsed_time(end), dev)
        launches = int(lib().asp_kernel_launch_count()) - launches0
        clocks = sampler.stop() if rank == 0 else None
        lib().asp_debug_time_extract_kernel(0)
        call_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in timers]))  # memset + index kernel + extraction kernel
        kernel_only = [float(lib().asp_debug_extract_kernel_ms(k)) for k in range(min(args.steps, 64))]  # read AFTER the timed region
        kernel_ms = float(np.mean(kernel_only))                                 # extract_csr_kernel alone
        nnz_total = D.sum_over_ranks(float(nnz_mine), dev)
        candidates_mine = None
        value = nnz_total * args.steps / (total_ms * 1e-3)
        algo_bytes = 24.0 * num_rows + 20.0 * nnz_mine  # per row 24 B, per coupling 20 B
        traffic, traffic_note = (measured_traffic("extract_csr_kernel") if (states_per_rank == 10_000_000 and world == 1)
                                 else (None, "ncu captures are taken on one GPU at 10^7 states"))
        roofline = {
            "bound": "hbm", "kernel": "extract_csr_kernel", "achieved": algo_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peak,
            "unit": "GB/s", "frac": algo_bytes / (kernel_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms, "call_ms": call_ms,
            "exchange_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in exchange])) if exchange else 0.0,
            "frac_whole_call": algo_bytes / (call_ms * 1e-3) / 1e9 / peak,
            "note": "kernel_ms: CUDA events around extract_csr_kernel alone on its launching stream (asp_debug_time_extract_kernel; read after "
                    "the timed region); call_ms: events around the whole asp_extract_csr call (memset + index_block_kernel + extract_csr_kernel; "
                    "timed passes do not read the count back, it is checked afterwards); exchange_ms (N > 1; includes waiting for the slowest rank): "
                    "--exchange peer = asp_gather_index, ONE kernel that pulls every row block over NVLink peer memory (cp.async.bulk) AND builds "
                    "the index behind the transfer (so call_ms has no index pass); --exchange nccl = the two NCCL all-gathers; algorithmic bytes = "
                    "24 B/row + 20 B/coupling (SURVEY.md 8d); the kernel is bound by instruction issue, not by HBM (DESIGN.md 4.1, profiles/)",
        }
        if world > 1:  # every rank's own clocks (a step is in effect a barrier: the slowest rank sets the pace)
            import torch.distributed as dist

            mine_stats = {"kernel_ms": round(kernel_ms, 4), "call_ms": round(call_ms, 4), "exchange_ms": round(roofline["exchange_ms"], 4),
                          "step_ms": round(start.elapsed_time(end) / args.steps, 4)}
            every_stats = [None] * world
            dist.all_gather_object(every_stats, mine_stats)
            roofline["per_rank"] = {k: [st[k] for st in every_stats] for k in mine_stats}
        indptr, indices, data = out
        assert int(indptr[-1]) == nnz_mine, "timed pass produced a different coupling count"
        indices, data = indices[:nnz_mine], data[:nnz_mine]

        # ---- end to end through the C ABI with HOST (pinned) buffers ----------------------------
        e2e = None
        if with_e2e and not args.skip_e2e:
            h_spins = spins.cpu().pin_memory()
            h_psi = psi.cpu().pin_memory()
            h_indptr = torch.empty(num_rows + 1, dtype=torch.int32).pin_memory()  # scipy's index type below 2^31 couplings
            h_indices = torch.empty(capacity, dtype=torch.int32).pin_memory()
            h_data = torch.empty(capacity, dtype=torch.float64).pin_memory()

            if peer is not None:
                # sharded end to end: every rank uploads ONLY its own row block (pinned host -> its peer buffer), the
                # full basis arrives over NVLink (asp_gather_index), the CSR rows go back to pinned host memory
                h_my_spins = my_spins.cpu().pin_memory()
                h_my_psi = my_psi.cpu().pin_memory()

            def host_pass_sharded():
                peer.begin_epoch()
                peer.spins[:num_rows].copy_(h_my_spins, non_blocking=True)
                peer.psi[:num_rows].copy_(h_my_psi, non_blocking=True)
                peer.publish()
                full_spins, full_psi = peer.gather_index(op, bounds, num_rows, workspace)
                peer.release()
                assert common.extract_indexed_to_host(op, full_spins, full_psi, row_begin, num_rows, workspace, h_indptr, h_indices, h_data) == nnz_mine

            def host_pass_full():
                nnz = ffi.new("uint64_t *")
                common.check(lib().asp_extract_host_i32(op.handle, n_total, ffi.cast("uint64_t *", h_spins.data_ptr()),
                                                        ffi.cast("double *", h_psi.data_ptr()), row_begin, num_rows, capacity,
                                                        ffi.cast("int32_t *", h_indptr.data_ptr()), ffi.cast("int32_t *", h_indices.data_ptr()),
                                                        ffi.cast("double *", h_data.data_ptr()), nnz))
                assert int(nnz[0]) == nnz_mine

            host_pass = host_pass_sharded if peer is not None else host_pass_full
            e2e_steps = max(2, min(args.steps, 5))
            host_pass()
            D.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                host_pass()
            torch.cuda.synchronize()
            e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
            assert int(h_indptr[-1]) == nnz_mine
            one_call_ms = 1e3 * e2e_s / e2e_steps
            in_flight = 1
            if peer is None:
                # the same calls, two in flight (asp_extract_host_i32_submit / _join): the upload and extraction of call
                # k+1 ride under the download of call k (the host link is full duplex); independent extractions, as in the
                # reference's experiment (one per cluster).  Each job writes its own set of pinned output buffers.
                outs = [(h_indptr, h_indices, h_data),
                        (torch.empty_like(h_indptr).pin_memory(), torch.empty_like(h_indices).pin_memory(), torch.empty_like(h_data).pin_memory())]

                def submit(k):
                    job = ffi.new("asp_host_job **")
                    o_indptr, o_indices, o_data = outs[k % 2]
                    common.check(lib().asp_extract_host_i32_submit(
                        op.handle, n_total, ffi.cast("uint64_t *", h_spins.data_ptr()), ffi.cast("double *", h_psi.data_ptr()), row_begin,
                        num_rows, capacity, ffi.cast("int32_t *", o_indptr.data_ptr()), ffi.cast("int32_t *", o_indices.data_ptr()),
                        ffi.cast("double *", o_data.data_ptr()), job))
                    return job[0]

                def join(job):
                    nnz = ffi.new("uint64_t *")
                    common.check(lib().asp_extract_host_join(job, nnz))
                    assert int(nnz[0]) == nnz_mine

                def two_in_flight(count):
                    pending = submit(0)
                    for k in range(1, count):
                        following = submit(k)
                        join(pending)
                        pending = following
                    join(pending)

                two_in_flight(2)
                D.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                two_in_flight(e2e_steps)
                e2e_s2 = D.max_over_ranks(time.perf_counter() - t0, dev)
                assert int(outs[0][0][-1]) == nnz_mine and int(outs[1][0][-1]) == nnz_mine
                assert torch.equal(outs[0][1][:nnz_mine], outs[1][1][:nnz_mine]) and torch.equal(outs[0][2][:nnz_mine], outs[1][2][:nnz_mine])
                if e2e_s2 < e2e_s:
                    e2e_s, in_flight = e2e_s2, 2
                del outs
            e2e = {"value": nnz_total * e2e_steps / e2e_s, "unit": "couplings/s",
                   "h2d_bytes_per_step": int((num_rows if peer is not None else n_total) * 16),
                   "d2h_bytes_per_step": int((num_rows + 1) * 4 + nnz_mine * 12),
                   "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "calls_in_flight": in_flight, "ms_per_step_one_call_at_a_time": one_call_ms,
                   "api": ("per rank: own row block pinned host -> peer buffer, asp_gather_index over NVLink, asp_extract_indexed_to_host_i32 "
                           "(row chunks copied back while the next chunk is extracted); bytes are per rank" if peer is not None else
                           "asp_extract_host_i32 (include/asp_b200.h; int32 row starts and columns, f64 values: scipy's CSR types), pinned host "
                           "buffers, row chunks copied back while the next chunk is extracted; two calls in flight (asp_extract_host_i32_submit / _join) when that is faster" + ("; every rank uploads the full basis; bytes are per rank" if world > 1 else ""))}
            del h_spins, h_psi, h_indptr, h_indices, h_data

        if peer is not None:
            peer.close()
        return {"value": value, "total_ms": total_ms, "launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e,
                "nnz_total": nnz_total, "n_total": n_total, "num_rows": num_rows, "need": need, "exchange_used": exchange_used,
                "pipelined": pipelined, "csr": (indptr, indices, data), "x1_parity": x1_parity, "kernel_ms": kernel_ms, "call_ms": call_ms}

    m = measure_extraction(args.states, True)
    value, total_ms, launches, clocks, roofline, e2e = m["value"], m["total_ms"], m["launches"], m["clocks"], m["roofline"], m["e2e"]
    nnz_total, n_total, num_rows, need, exchange_used, pipelined = (m[k] for k in ("nnz_total", "n_total", "num_rows", "need", "exchange_used", "pipelined"))
    indptr, indices, data = m["csr"]
    strong = None
    if world > 1 and not args.skip_strong:
        # BASELINE.json configs[3] as written: 10^7 sampled states IN TOTAL, row blocks over the N ranks (strong scaling)
        ms = measure_extraction(max(args.states // world, 1), False)
        strong = {"value": ms["value"], "unit": "couplings/s", "ms_per_step": ms["total_ms"] / args.steps, "states_total": ms["n_total"],
                  "rows_per_gpu": ms["num_rows"], "couplings_total": int(ms["nnz_total"]), "kernel_ms": ms["kernel_ms"], "call_ms": ms["call_ms"],
                  "exchange_ms": ms["roofline"]["exchange_ms"], "x1_parity": ms["x1_parity"],
                  "note": "same step as the headline (exchange + index + extraction) on a basis of %d states in total" % ms["n_total"]}
        del ms

    # ---- annealing stage on the extracted model (replicas shard over ranks) ------------------
    def measure_anneal(ham, replicas, betas, steps, warmup, traffic_key=None, e0=None):
        """`steps` annealing runs of `replicas` replicas per rank (disjoint random streams per rank, X2 at the end of each
        run); -> the spin-flip rate with both rooflines: the bytes the row-sharing kernel must move, and SURVEY.md 8d's
        formula (one CSR stream per replica per sweep)."""
        n_model = ham.size
        t0 = time.perf_counter()
        plan = asp.sa.AnnealPlan(ham)
        torch.cuda.synchronize()
        plan_ms = 1e3 * (time.perf_counter() - t0)
        escale = asp.sa.energy_scale(ham)
        sweeps = int(betas.shape[0])
        state = {}

        def anneal_pass(seed):
            bits, energies = plan.anneal_device(replicas, betas, seed, escale=escale, replica_offset=rank * ((replicas + 31) // 32 * 32))
            state["energies"] = energies
            best = int(torch.argmin(energies))
            return D.reduce_best(float(energies[best]), bits[best])  # X2

        for w in range(warmup):
            anneal_pass(w)
        launches_a0 = int(lib().asp_kernel_launch_count())
        D.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for k in range(steps):
            e_best, _, _ = anneal_pass(100 + k)
        a1.record()
        D.barrier()
        torch.cuda.synchronize()
        a_ms = D.max_over_ranks(a0.elapsed_time(a1), dev)
        flips = float(replicas) * world * sweeps * n_model * steps
        groups = (replicas + 31) // 32
        # what the kernel has to stream per sweep and group of 32 replicas (its rows are shared by the 32 replicas of a
        # warp): the relabelled CSR (12 B per coupling, the diagonal is not stored), per 4-position task one first-entry
        # pointer and one row-boundary word (8 + 16 B = 6 B per spin), the 32-replica spin words read and written back
        # (8 B per spin), the fields (8 B per spin) only when there are any; neighbour words count as cache hits
        per_spin = 14.0 + (8.0 if np.any(ham.field) else 0.0)
        required = groups * sweeps * (12.0 * plan.nnz + per_spin * plan.n_padded) * steps
        survey = replicas * sweeps * (12.0 * plan.nnz + 16.0 * n_model + 8.0) * steps  # one stream per replica
        traffic, traffic_note = measured_traffic(traffic_key) if (traffic_key and world == 1) else (None, "no ncu capture for this configuration")
        out = {
            "metric": "spin_flips_per_sec", "value": flips / (a_ms * 1e-3), "unit": "proposals/s", "ms_per_step": a_ms / steps,
            "config": {"spins": n_model, "couplings": int(plan.nnz), "replicas_per_gpu": replicas, "sweeps_per_step": sweeps,
                       "colour_classes": plan.num_classes, "plan_ms": plan_ms, "best_energy": e_best},
            "roofline": {"bound": "hbm", "kernel": "sa_sweep_kernel", "achieved": required / (a_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": required / (a_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_note,
                         "bytes_per_launch": required / steps,
                         "survey_formula": {"achieved": survey / (a_ms * 1e-3) / 1e9, "frac": survey / (a_ms * 1e-3) / 1e9 / peak,
                                            "note": "SURVEY.md 8d prices one CSR stream per REPLICA per sweep; the kernel reads a row once "
                                                    "per 32 replicas, so this figure can exceed 1 and is not a roofline fraction"},
                         "note": "bytes = groups of 32 replicas x sweeps x (12 B per stored coupling + %d B per spin): the stream the "
                                 "row-sharing kernel cannot avoid; the kernel is bound by instruction issue (DESIGN.md 4.3)" % per_spin},
            "gpu_launches": int(lib().asp_kernel_launch_count()) - launches_a0,
        }
        if e0 is not None:
            es = state["energies"]
            out["config"]["exact_ground_state_energy"] = e0
            out["config"]["replicas_at_E0"] = float((((es - e0) / e0).abs() <= 1e-12).double().mean())
        return out

    def strided_betas(ham, sweeps):
        """`sweeps` inverse temperatures spread over the WHOLE default ladder (every 8th rung of an 8x longer anneal): hot,
        critical and frozen sweeps in the proportions of a full run."""
        if sweeps <= 1:
            return asp.sa.default_betas(ham, 1)
        return np.ascontiguousarray(asp.sa.default_betas(ham, sweeps * 8)[3::8])

    anneal = None
    if not args.skip_anneal:
        if world > 1:
            # every rank anneals the full single-GPU-sized model of rank 0's shape: rebuild locally
            a_spins = synthetic.cluster_closed_states(op, args.states, 1000, dev)
            a_psi = synthetic.synthetic_amplitudes(a_spins.shape[0], 77, device=dev)
            del indptr, indices, data
            indptr, indices, data = common.extract_csr_device(op, a_spins, a_psi)
        n_model = int(indptr.shape[0] - 1)

        class _Shape:  # Hamiltonian wants a .shape; skip the host scipy copy at this size
            shape = (n_model, n_model)

        ham = asp.sa.Hamiltonian(_Shape(), np.zeros(n_model), _device_csr=(indptr, indices, data, None))
        anneal = measure_anneal(ham, args.replicas, strided_betas(ham, args.sweeps), args.steps, args.warmup,
                                traffic_key="sa_sweep_kernel" if (args.states == 10_000_000 and args.replicas == 64 and args.sweeps == 16) else None)
        anneal["config"]["schedule"] = "%d sweeps: every 8th rung of the %d-sweep default ladder (hot to frozen)" % (args.sweeps, 8 * args.sweeps)
        if world == 1:
            # what the reference's 32- and 36-spin runs call instead of the annealer (common.py:249-250): the greedy solver
            # on the same model (asp_greedy_solve: plan + maximum spanning forest + descent; SURVEY 8f N1, parity unpinned)
            t0 = time.perf_counter()
            g_plan = asp.sa.AnnealPlan(ham)
            torch.cuda.synchronize()
            g_plan_ms = 1e3 * (time.perf_counter() - t0)
            g_plan.greedy_device()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                g_bits, g_energy, g_rounds, g_sweeps = g_plan.greedy_device()
            g1.record()
            torch.cuda.synchronize()
            g_ms = g0.elapsed_time(g1) / 3
            anneal["greedy"] = {"metric": "greedy_spins_per_sec", "value": n_model / (g_ms * 1e-3), "unit": "spins/s", "ms_per_solve": g_ms,
                                "plan_ms": g_plan_ms, "merge_rounds": g_rounds, "descent_sweeps": g_sweeps, "energy": float(g_energy),
                                "energy_of_the_anneal": anneal["config"]["best_energy"],
                                "api": "asp_greedy_solve on the plan of the same model (configuration and energy on the device)"}
            del g_plan, g_bits
        del ham
    del indptr, indices, data

    # ---- the other BASELINE.json configurations ------------------------------------------------
    def exact_ground_state(operator):
        """Lowest eigenpair of the operator in its own (possibly symmetrised) basis: matrix elements from the device
        batched_apply, Lanczos (scipy, a library call outside the path) on the host."""
        import scipy.sparse
        import scipy.sparse.linalg

        basis = operator.basis
        d_states = basis.states_device()
        n = int(d_states.shape[0])
        other, coeffs, counts = operator.batched_apply_device(d_states)
        cols = basis.batched_index_device(other)
        rows = torch.repeat_interleave(torch.arange(n, device=dev), counts)
        h = scipy.sparse.coo_matrix((coeffs.cpu().numpy(), (rows.cpu().numpy(), cols.cpu().numpy())), shape=(n, n)).tocsr()
        w, v = scipy.sparse.linalg.eigsh(h, k=2, which="SA", tol=1e-13, v0=np.random.default_rng(0).standard_normal(n))
        k = int(np.argmin(w))
        return float(w[k]), np.ascontiguousarray(v[:, k])

    def full_basis_config(system, replicas, sweeps, steps):
        """Exact ground state -> make_ising_model (the reference's seam) -> replica annealing with the default ladder."""
        operator = asp.load_hamiltonian(asp.ls.system_path(system))
        e0, psi0 = exact_ground_state(operator)
        states = operator.basis.states
        with np.errstate(divide="ignore"):
            log_psi = np.log(psi0.astype(np.complex128))
            asp.make_ising_model(states, operator, log_psi=log_psi)  # warm-up (allocations, first launches)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            model = asp.make_ising_model(states, operator, log_psi=log_psi)
            torch.cuda.synchronize()
            make_ms = 1e3 * (time.perf_counter() - t0)
        ham = model.ising_hamiltonian
        a = measure_anneal(ham, replicas, asp.sa.default_betas(ham, sweeps), steps, 1, e0=e0)
        return {"workload": "%s, full %sbasis, exact ground state (Lanczos), make_ising_model + %d replicas x %d sweeps" % (
                    system, "symmetrised " if operator.basis.is_symmetrised else "", replicas, sweeps),
                "states": int(model.size), "couplings": int(ham.exchange.nnz), "make_ising_model_ms": make_ms,
                "couplings_per_sec_python_seam": ham.exchange.nnz / (make_ms * 1e-3),
                "energy_of_exact_signs_minus_E0": float(ham.energy(model.initial_signs) - e0), "anneal": a}

    def sampled_config(system, states, replicas, sweeps, steps, symmetrised=False):
        """Cluster-closed sampled subset -> device extraction (kernel time from CUDA events) -> replica annealing."""
        if symmetrised:
            operator = asp.load_hamiltonian(asp.ls.system_path(system))
            c_spins = synthetic.representative_cluster_states(operator, states, 5, dev)
        else:
            cfg_c = asp.ls.load_config(asp.ls.system_path(system))
            cfg_c["basis"]["symmetries"], cfg_c["basis"]["spin_inversion"] = [], None
            operator = asp.ls.Operator.load_from_yaml(cfg_c["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg_c["basis"]))
            c_spins = synthetic.cluster_closed_states(operator, states, 5, dev)
        n = int(c_spins.shape[0])
        c_psi = synthetic.synthetic_amplitudes(n, 5, device=dev)
        common.extract_csr_device(operator, c_spins, c_psi)
        lib().asp_debug_time_extract_kernel(1)
        calls, kernels = [], []
        for _ in range(steps):
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record()
            c_indptr, c_indices, c_data = common.extract_csr_device(operator, c_spins, c_psi)
            e1_.record()
            torch.cuda.synchronize()
            calls.append(e0_.elapsed_time(e1_))
            kernels.append(float(lib().asp_debug_last_extract_kernel_ms()))
        lib().asp_debug_time_extract_kernel(0)
        nnz = int(c_indices.numel())
        call_ms = float(np.median(calls))
        out = {"workload": "%s%s, %d sampled states (cluster-closed), extraction + %d replicas per GPU x %d sweeps" % (
                   system, " symmetrised (orbit representatives)" if symmetrised else "-shaped U(1) basis", n, replicas, sweeps),
               "states": n, "couplings": nnz,
               "extract": {"value": nnz / (call_ms * 1e-3), "unit": "couplings/s", "call_ms": call_ms,
                           "note": "whole extract_csr_device call (allocation of the outputs included)"}}
        algo = 24.0 * n + 20.0 * nnz
        if not symmetrised:
            kernel_ms = float(np.median(kernels))
            out["extract"]["kernel_ms"] = kernel_ms
            out["extract"]["roofline"] = {"bound": "hbm", "kernel": "extract_csr_kernel", "achieved": algo / (kernel_ms * 1e-3) / 1e9, "peak": peak,
                                          "unit": "GB/s", "frac": algo / (kernel_ms * 1e-3) / 1e9 / peak, "traffic": None}
        else:
            out["extract"]["roofline"] = {"bound": "hbm", "kernel": "apply_kernel + legacy_build_kernel + canonicalize_kernel",
                                          "achieved": algo / (call_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": algo / (call_ms * 1e-3) / 1e9 / peak, "traffic": None,
                                          "note": "integer-ALU bound: every candidate walks its symmetry orbit (DESIGN.md 4.2)"}
        if replicas:
            class _S:
                shape = (n, n)

            c_ham = asp.sa.Hamiltonian(_S(), np.zeros(n), _device_csr=(c_indptr, c_indices, c_data, None))
            out["anneal"] = measure_anneal(c_ham, replicas, strided_betas(c_ham, sweeps), steps, 1)
        return out

    configs = None
    if not args.skip_configs:
        configs = {}
        if world == 1:
            configs["cfg1_j1j2_square_4x4"] = full_basis_config("j1j2_square_4x4", 64, 5120, 2)
            configs["cfg2_heisenberg_kagome_18"] = full_basis_config("heisenberg_kagome_18", 1024, 1600, 2)
            configs["cfg3_sk_32_1"] = sampled_config("sk_32_1", 1_000_000, 4096, 4, 2)
            configs["cfg4_heisenberg_kagome_36_symmetrised"] = sampled_config("heisenberg_kagome_36", 1_000_000, 0, 0, 2, symmetrised=True)
        # configs[4]: replicas over the GPUs (64 per GPU, disjoint random streams, X2 at the end of every run)
        configs["cfg5_heisenberg_pyrochlore_2x2x2"] = sampled_config("heisenberg_pyrochlore_2x2x2", 10_000_000, 64, 16, 2)
        torch.cuda.empty_cache()

    # ---- end to end through the PYTHON seam (common.make_ising_model, the call a user of the reference makes) ----
    e2e_python = None
    if rank == 0 and world == 1 and not args.skip_python:
        e2e_python = {}
        for n_py in (1_000_000, args.states):
            p_spins = synthetic.cluster_closed_states(op, n_py, 1000, dev)
            h_spins_py = p_spins.cpu().numpy().view(np.uint64)
            h_log_psi = np.log(synthetic.synthetic_amplitudes(int(p_spins.shape[0]), 77).numpy().astype(np.complex128))
            del p_spins
            asp.make_ising_model(h_spins_py, op, log_psi=h_log_psi)
            t0 = time.perf_counter()
            model = asp.make_ising_model(h_spins_py, op, log_psi=h_log_psi)
            dt = time.perf_counter() - t0
            e2e_python["%d_states" % n_py] = {"value": model.ising_hamiltonian.exchange.nnz / dt, "unit": "couplings/s", "ms": 1e3 * dt,
                                              "couplings": int(model.ising_hamiltonian.exchange.nnz),
                                              "api": "common.make_ising_model(spins, operator, log_psi) with numpy inputs -> IsingModel with a scipy COO "
                                                     "exchange matrix (sort + unique, extraction, 0.5 (M + M^T), COO rows on the device; results in host memory)"}
            del model
        torch.cuda.empty_cache()

    # ---- CPU baseline on rank 0 (N = 1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        from oracle import capi

        capi.build()
        impl = "ref" if capi.have_ref() else "port"
        inputs = cpu_sample_inputs(args.cpu_sample)
        host_cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
        nnz_1, dt_1 = cpu_extract_once(capi, inputs, impl)
        nnz_s, dt = cpu_extract_threaded(capi, inputs, impl, host_cores)
        flips_s, cores, reps = cpu_anneal_once(capi, inputs, 4)
        cpu = {"value": nnz_s / dt, "unit": "couplings/s", "cores": host_cores, "kind": "reference" if impl == "ref" else "port",
               "sample": "%d-state subset of the same shape, %d candidates; cbits/build_matrix.c (serial, re-entrant C, 512-bit keys) called on "
                         "%d threads, one contiguous row block each; calls only, neighbour lists precomputed" % (
                             inputs[0].shape[0], inputs[2].shape[0], host_cores),
               "one_thread": {"value": nnz_1 / dt_1, "unit": "couplings/s"},
               "candidates_per_sec": inputs[2].shape[0] / dt,
               "anneal": {"value": flips_s, "unit": "proposals/s", "cores": cores, "kind": "port",
                          "sample": "oracle/anneal_port.c, %d replicas x 4 sweeps on the %d-spin sample model" % (reps, inputs[0].shape[0])}}
        try:  # the greedy solver's CPU restatement on the same sample model
            g_n, g_dt = cpu_greedy_once(capi, inputs)
            cpu["greedy"] = {"value": g_n / g_dt, "unit": "spins/s", "cores": 1, "kind": "port",
                             "sample": "oracle/greedy_port.c (Kruskal with a signed union-find + descent) on the %d-spin sample model" % g_n}
        except Exception as exc:  # noqa: BLE001
            cpu["greedy"] = {"unavailable": repr(exc)[:200]}
        try:  # an extra figure: it must never cost the headline
            nnz_l, dt_l = cpu_live_path_once(inputs)
            cpu["live_path"] = {"value": nnz_l / dt_l, "unit": "couplings/s", "cores": 1, "kind": "port",
                                "sample": "oracle/live_path.py (numpy/scipy restatement of common.py:131-208) on the same %d states, "
                                          "neighbour generation included" % inputs[0].shape[0]}
        except Exception as exc:  # noqa: BLE001
            cpu["live_path"] = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "ising_couplings_built_per_sec", "value": value, "unit": "couplings/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "heisenberg_kagome_36-shaped U(1) basis (36 spins, 72 bonds), %d sampled states per GPU, "
                                   "cluster-closed subset, Ising extraction to CSR" % args.states,
                       "states_total": n_total, "rows_per_gpu": num_rows, "couplings_total": int(nnz_total),
                       "candidates_per_row": 37.0, "parallelism": "row blocks x%d, basis exchanged every step (%s)%s" % (
                           world, exchange_used, "; steps pipelined two deep: the copy engines gather the basis of step k+1 over NVLink while "
                           "step k is indexed and extracted (independent extractions, two private copies)" if pipelined else ""),
                       "l2": "inputs (%.0f MB) larger than L2" % ((n_total * 16 + need) / 1e6)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "anneal": anneal,
            "configs": configs, "e2e_python": e2e_python, "x1_parity": m["x1_parity"], "scaling_strong": strong,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

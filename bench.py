#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 alignment engine (contract: see the task statement).

Workload (BASELINE.json configs[1], "C2"): GASAL2 local + semi-global alignment WITH start/end
positions on synthetic 150 bp x 600 bp read/window pairs, 1,000,000 pairs PER GPU (weak scaling).
One "step" = one pass of the hot path over that batch in BOTH modes (local WITH_START, then semi-global
HEAD=TAIL=TARGET WITH_START).  Metric: GCUPS = sum(query_len * target_len) of the forward DP cells
(both modes) / time / 1e9 (same definition as the suite's own sw-benchmark).

  value : device-timed (CUDA events on the storages' streams), inputs already resident in HBM; the
          timed region is pack -> DP (forward + start pass) per mode.
  e2e   : same metric through the reference-facing C ABI (gasal_aln_async / gasal_is_aln_async_done)
          with HOST buffers: pinned H2D of the sequences + lens/offsets and D2H of the results inside the
          timed region, chunked over 2 storages per mode so copies overlap kernels (as test_prog.cpp does).
  --impl reference : the reference's own kernels (oracle/_ref, built from /root/reference sources for
          the host) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

QLEN, TLEN = 150, 600
SCORING = (1, 4, 6, 1)


def make_workload(n_pairs: int, seed: int):
    from genomics_gpu_b200 import synth
    chunks = []
    step = 100_000
    for i, lo in enumerate(range(0, n_pairs, step)):
        chunks.append(synth.config2(min(step, n_pairs - lo), seed=seed * 1000 + i))
    if len(chunks) == 1:
        return chunks[0]
    q8, t8 = (QLEN + 7) // 8 * 8, (TLEN + 7) // 8 * 8
    n = sum(c.n for c in chunks)
    return synth.PairBatch(np.concatenate([c.q_bytes for c in chunks]), np.concatenate([c.t_bytes for c in chunks]),
                           (np.arange(n, dtype=np.uint64) * q8).astype(np.uint32), (np.arange(n, dtype=np.uint64) * t8).astype(np.uint32),
                           np.full(n, QLEN, dtype=np.uint32), np.full(n, TLEN, dtype=np.uint32))


WORKLOAD = "C2: GASAL2 local + semi-global, WITH_START, 150x600 bp synthetic pairs"


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled
    in-process every 5 ms (the timed region of the default run is well under a second, too short for `nvidia-smi -lms` to
    start up on an 8-GPU box); nvidia-smi is the fallback when the NVML binding is missing."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []          # (sm_mhz, power_w, reasons_bitmask)
        self.sm_max = None
        self.nvml = None
        self.stop_flag = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = nvml_handle_for_cuda_index(pynvml, gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    pw = None
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    try:
                        rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    except Exception:
                        rs = 0
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml is not None:
            self.stop_flag.set(); self.thread.join(timeout=1)
            # NVML reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal_slowdown, 0x40 hw_thermal_slowdown
            bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            sm = [x[0] for x in self.samples]
            loaded_idx = [i for i, v in enumerate(sm) if v > 0.5 * max(sm)] if sm else []
            reasons = sorted({nm for i in loaded_idx for nm, b in bits.items() if self.samples[i][2] & b})
            power = [x[1] for x in self.samples if x[1] is not None]
            return {"sm_mhz": statistics.median([sm[i] for i in loaded_idx]) if loaded_idx else None, "sm_max_mhz": self.sm_max,
                    "power_w_max": max(power) if power else None, "samples": len(sm), "source": "nvml, 5 ms poll", "reasons": reasons}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # samples under load only (the box idles at ~120 MHz between launches)
        loaded = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "source": "nvidia-smi -lms 50", "reasons": sorted(reasons)}


def nvml_handle_for_cuda_index(pynvml, cuda_index: int):
    """NVML handle of the CUDA device `cuda_index` of this process (PCI bus id match, so CUDA_VISIBLE_DEVICES is honoured)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(cuda_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(cuda_index), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(cuda_index).pci_device_id
        return pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
    except Exception:
        return pynvml.nvmlDeviceGetHandleByIndex(cuda_index)


def bind_to_gpu_numa_node(cuda_index: int):
    """Run this rank on the CPUs NVML reports as local to its GPU, so the pinned staging pages of the e2e arm are allocated
    (first touch) on the GPU's NUMA node.  Returns the previous affinity set (restored before the CPU baseline), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = nvml_handle_for_cuda_index(pynvml, cuda_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        prev = os.sched_getaffinity(0)
        want = cpus & prev
        if want and want != prev:
            os.sched_setaffinity(0, want)
            return prev
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------------
def dpx_peak_ops_per_s():
    """R = measured lane-ops/s of the DPX/ALU integer pipe of one B200 (profiles/dpx_microbench_r01.json)."""
    p = os.path.join(ROOT, "profiles", "dpx_microbench_r01.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["R_alu_lane_ops_per_s"]), "measured (profiles/dpx_microbench_r01.json)"
    except Exception:
        return 148 * 64 * 1.965e9, "a-priori 148 SM x 64 lanes/clk x 1.965 GHz (microbenchmark file missing)"


def ncu_traffic(kernel_key: str, pairs_per_launch: int):
    """DRAM bytes per launch of the kernel from the committed `ncu --set full` capture (profiles/ncu_summary_r02.json: bytes per pair of
    one 65536-pair launch), scaled to this run's pairs per launch"""
    p = os.path.join(ROOT, "profiles", "ncu_summary_r02.json")
    try:
        with open(p) as f:
            return json.load(f)[kernel_key]["dram_bytes_per_pair"] * pairs_per_launch
    except Exception:
        return None


def cpu_baseline_run(batch, n_sample: int, threads: int):
    """The reference's own kernels compiled for the host (kind 'reference'), else our C port (kind 'port')."""
    from oracle import pyoracle as po
    sub = batch.slice(0, min(n_sample, batch.n))
    orc = po.Oracle()
    pq, pt = orc.pack(sub.q_bytes), orc.pack(sub.t_bytes)
    if po.RefHost.available(1024):
        eng, kind = po.RefHost(1024), "reference"
    else:
        eng, kind = orc, "port"
    t0 = time.perf_counter()
    for algo in (po.LOCAL, po.SEMI_GLOBAL):
        kw = dict(algo=algo, start=po.WITH_START, scoring=SCORING, n_threads=threads)
        if kind == "port":
            kw["max_query_len"] = 1024
        eng.align(pq, pt, sub.q_lens, sub.t_lens, sub.q_offs, sub.t_offs, **kw)
    dt = time.perf_counter() - t0
    return 2 * sub.cells / dt / 1e9, kind, sub.n, dt


def ref_gpu_baseline(batch, n_sample: int):
    """The reference's own CUDA library recompiled for sm_100 (oracle/_ref/ref_gpu_driver_mq*), same GPU, same workload
    sample, through its own API with its example's 2-stream pipeline; batch size = half the sample (its GPU-filling
    setting) -- a reported baseline (north_star), not the judged reference arm."""
    from oracle import pyoracle as po
    if not (po.RefGpu.available(160) and po.RefGpu.available(608)):
        return {"unavailable": "oracle/_ref/ref_gpu_driver_mq160/608 not built"}
    sub = batch.slice(0, min(n_sample, batch.n))
    out = {"pairs": sub.n, "batch_size": sub.n // 2, "streams": 2}
    t_ev = t_wall = 0.0
    try:
        for name, algo, mq in (("local", po.LOCAL, 160), ("semi_global", po.SEMI_GLOBAL, 608)):
            r = po.RefGpu(mq).run(sub, algo=algo, start=po.WITH_START, scoring=SCORING, batch_size=sub.n // 2, reps=3)["_timing"]
            out[name] = {"max_query_len": mq, "event_ms": r["event_ms_sum"], "wall_ms": r["wall_ms"], "gcups_events": r["gcups_events"], "gcups_wall": r["gcups_wall"]}
            t_ev += r["event_ms_sum"]; t_wall += r["wall_ms"]
    except Exception as e:  # the baseline must never break the bench line
        return {"unavailable": f"reference GPU driver failed: {e}"[:200]}
    out["value"] = 2 * sub.cells / (t_ev * 1e-3) / 1e9
    out["value_wall"] = 2 * sub.cells / (t_wall * 1e-3) / 1e9
    out["unit"] = "GCUPS"
    out["timing"] = "value: CUDA events around each gasal_aln_async batch (H2D + kernels + D2H), summed; value_wall: host wall clock incl. its per-sequence host fill"
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.ref_pairs
    batch = make_workload(n_sample, seed=2)
    for _ in range(args.warmup):
        cpu_baseline_run(batch, max(n_sample // 10, 1000), threads)
    t_total, cells = 0.0, 0
    kind = "port"
    for _ in range(args.steps):
        g, kind, n, dt = cpu_baseline_run(batch, n_sample, threads)
        t_total += dt; cells += 2 * batch.slice(0, n).cells
    value = cells / t_total / 1e9
    line = {"metric": "GCUPS", "impl": "reference", "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "modes_per_step": ["local WITH_START", "semi_global TARGET/TARGET WITH_START"], "scoring": list(SCORING),
                       "pairs_per_step": n_sample, "note": "bounded sample of the same workload per step (the CPU needs minutes for the full 1 M pairs)"},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": threads, "kind": kind,
                             "sample": f"{n_sample} pairs of the C2 workload per step, both modes, reference kernels compiled for the host (OpenMP)"},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "pairs_per_s": 2 * n_sample * args.steps / t_total}
    print(json.dumps(line))


def nvb_extra(n: int):
    """The suite's second alignment benchmark shape (NvB/sw-benchmark/sw-benchmark.cu:560-640: Gotoh 2/-1/-2/-1 global / semi-global /
    local + edit-distance semi-global, scores only) through include/nvbio_b200.h; device-timed, best of 3."""
    from genomics_gpu_b200 import nvb, synth
    pats, poffs, texts, toffs = synth.nvb_jobs(n, 150, 150, 600, 600, seed=31, sub=0.02, indel=0.0)
    out = {}
    with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
        cells = st.cells()
        for name, al in (("gotoh_global", nvb.aligner(nvb.GOTOH, nvb.GLOBAL, 2, -1, -2, -1)), ("gotoh_semi_global", nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1)),
                         ("gotoh_local", nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1)), ("ed_semi_global", nvb.aligner(nvb.EDIT_DISTANCE, nvb.SEMI_GLOBAL))):
            best = 1e30
            for _ in range(3):
                st.enact(al); best = min(best, st.last_ms())
            out[name] = {"kernel": st.kernel(), "dp_ms": best, "gcups_dp": cells / best / 1e6}
        # BatchedAlignmentTraceback (batched.h:436-451): full flow matrix + walk, Gotoh local / semi-global
        for name, al in (("traceback_gotoh_local", nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1)), ("traceback_gotoh_semi_global", nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1))):
            best = 1e30
            for _ in range(2):
                st.traceback(al); best = min(best, st.last_ms())
            out[name] = {"kernel": st.kernel(), "ms": best, "gcups": cells / best / 1e6, "alignments_per_s": n / best * 1e3,
                         "flow_bytes_per_launch_set": 0.5 * cells}
    # BatchedBandedAlignmentScore<31> / BatchedBandedAlignmentTraceback<31> (batched.h:337-357, :463-478): 150 bp reads in 150 + 31 + 3 bp windows
    pats, poffs, texts, toffs = synth.nvb_banded_jobs(n, 150, 150, 31, seed=32, sub=0.02, indel=0.002, slack=3)
    with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
        band_cells = float(np.diff(poffs).astype(np.int64).sum()) * 31
        al = nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1)
        best = 1e30
        for _ in range(3):
            st.enact_banded(al, 31); best = min(best, st.last_ms())
        out["banded31_gotoh_semi_global"] = {"kernel": st.kernel(), "dp_ms": best, "band_gcups": band_cells / best / 1e6, "jobs_per_s": n / best * 1e3}
        best = 1e30
        for _ in range(3):
            st.traceback(al, 31); best = min(best, st.last_ms())
        out["banded31_traceback_gotoh_semi_global"] = {"kernel": st.kernel(), "ms": best, "band_gcups": band_cells / best / 1e6, "alignments_per_s": n / best * 1e3}
    return out


# ---------------------------------------------------------------------------------------------------------
def run_extras(local_rank: int, n_fwd: int):
    """Device-timed GCUPS of the other BASELINE.json configurations (parity-test cases, reported for context only; bounded
    samples): forward-only kernels on the C2 shape, C3 (global 16k x 16k, wavefront kernel), C4 (banded k=64 + CIGAR),
    C5 (mixed 100 bp - 10 kb, local WITH_START)."""
    from genomics_gpu_b200 import capi, synth
    from genomics_gpu_b200.aligner import Aligner
    out = {}

    def timed(batch, reps=3, **kw):
        mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
        with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, scoring=SCORING, device=local_rank, **kw) as al:
            best = None
            for _ in range(reps):
                al.align(batch)
                ms = al.last_ms(0)
                best = ms if best is None or ms[2] < best[2] else best
            return {"kernel": al.kernel_name(0), "pairs": batch.n, "dp_ms": best[2], "gcups_dp": batch.cells / (best[2] * 1e-3) / 1e9}

    try:
        b2 = synth.config2(n_fwd, seed=2)
        out["C2_forward_only"] = {"local": timed(b2, algo=capi.LOCAL), "semi_global": timed(b2, algo=capi.SEMI_GLOBAL), "global": timed(b2, algo=capi.GLOBAL)}
        # the other semiglobal_skipping_head / _tail variants of the reference API (semiglobal_kernel_template.h:39-388): all on the s16x2 lanes
        out["C2_semi_global_head_tail_variants_with_start"] = {
            f"head_{hn}_tail_{tn}": timed(b2, algo=capi.SEMI_GLOBAL, head=h, tail=t, start=capi.WITH_START)
            for (hn, h), (tn, t) in ((("none", capi.NONE), ("target", capi.TARGET)), (("query", capi.QUERY), ("target", capi.TARGET)),
                                     (("target", capi.TARGET), ("query", capi.QUERY)), (("both", capi.BOTH), ("both", capi.BOTH)))}
        # KSW seed extension (algo = KSW, host_seed_scores; ksw_kernel_template.h:46-199): one thread per pair, literal semantics
        try:
            bk = synth.make_extension_pairs(n_fwd, 150, 200)          # the read is a mutated copy of the window's first 150 bases: the extension runs the whole read
            with Aligner(max_query_len=157, max_target_len=207, max_n_alns=bk.n, n_streams=1, scoring=SCORING, device=local_rank, algo=capi.KSW) as al:
                seeds = np.full(bk.n, 20, dtype=np.uint32)
                best = None
                for _ in range(3):
                    rk = al.align(bk, seeds=seeds)
                    ms = al.last_ms(0)
                    best = ms if best is None or ms[2] < best[2] else best
                out["KSW_extension_150x200_seed20"] = {"kernel": al.kernel_name(0), "pairs": bk.n, "dp_ms": best[2], "pairs_per_s": bk.n / (best[2] * 1e-3),
                                                       "gcups_dp_full_rectangle": bk.cells / (best[2] * 1e-3) / 1e9, "mean_query_end": float(rk["query_batch_end"].mean()),
                                                       "note": "gcups over the full 150 x 200 rectangle; ksw_extend itself narrows the band row by row"}
        except Exception as exc:  # noqa: BLE001
            out["KSW_extension_150x200_seed20"] = {"error": repr(exc)}
        out["C3_global_16k_x_16k"] = {"n128": timed(synth.config3(128, 16384, seed=3), algo=capi.GLOBAL), "n32": timed(synth.config3(32, 16384, seed=3), algo=capi.GLOBAL),
                                      "n16": timed(synth.config3(16, 16384, seed=3), algo=capi.GLOBAL)}
        b4 = synth.config4(n_fwd, seed=4)
        out["C4_banded_k64_250bp"] = {"score_end": timed(b4, algo=capi.BANDED, k_band=64), "with_cigar": timed(b4, algo=capi.BANDED, k_band=64, start=capi.WITH_TB)}
        out["C5_mixed_100bp_10kb_local_with_start"] = timed(synth.config5(4000, seed=5), algo=capi.LOCAL, start=capi.WITH_START)
        # long reads in flanked windows, semi-global (read global, window free at both ends): the s16x2 wavefront lanes (k_wave16_semi)
        lr = synth.long_read_windows(600, 2000, 9000, 1500, seed=5)
        out["long_reads_2kb_9kb_semi_global"] = {"score_end": timed(lr, algo=capi.SEMI_GLOBAL), "with_start": timed(lr, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)}
        out["C1_10k_pairs_150x300_local"] = config1_extra(local_rank)
        # reads with uncalled bases: pairs with an N leave the main s16x2 launch for its N-aware second pass (one more LOP3 per cell pair)
        out["C2_local_with_start_vs_fraction_of_pairs_with_N"] = {
            f"{int(100 * f)}%": timed(synth.with_n_bases(b2, f, seed=17) if f else b2, algo=capi.LOCAL, start=capi.WITH_START) for f in (0.0, 0.01, 0.05, 0.20)}
        out["real_reads_reference_fixture_x66"] = real_reads_extra(timed)
        out["C4_local_250bp_with_cigar_s16x2"] = timed(b4, algo=capi.LOCAL, start=capi.WITH_TB)
        out["NW_linear_full_matrix_N32000"] = nw_extra(32000)
        out["nvbio_style_batched_scoring_262144x150x600"] = nvb_extra(n_fwd)
        out["note"] = "gcups_dp counts q_len x t_len forward cells per pair (banded and start passes add time, not cells); device-timed DP stage, best of 3"
    except Exception as e:  # context only: never break the bench line
        out["error"] = str(e)[:200]
    return out


def config1_extra(local_rank: int):
    """BASELINE config 1: 10,000 pairs 150 x 300, local, score + end: the reference's kernels on the host cores (the CPU path the
    config names) beside the engine on the same pairs, results compared field by field."""
    from genomics_gpu_b200 import capi, synth
    from genomics_gpu_b200.aligner import Aligner
    from oracle import pyoracle as po
    b = synth.config1(10_000, seed=1)
    orc = po.Oracle()
    pq, pt = orc.pack(b.q_bytes), orc.pack(b.t_bytes)
    eng, kind = (po.RefHost(1024), "reference kernels compiled for the host") if po.RefHost.available(1024) else (orc, "port (oracle/gotoh_compat.c)")
    threads = len(os.sched_getaffinity(0)) or 1
    kw = dict(algo=po.LOCAL, scoring=SCORING, n_threads=threads)
    if not po.RefHost.available(1024):
        kw["max_query_len"] = 1024
    t0 = time.perf_counter(); want = eng.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, **kw); cpu_s = time.perf_counter() - t0
    with Aligner(algo=capi.LOCAL, max_query_len=157, max_target_len=307, max_n_alns=b.n, n_streams=1, scoring=SCORING, device=local_rank) as al:
        best = None
        for _ in range(3):
            got = al.align(b); ms = al.last_ms(0)
            best = ms if best is None or ms[0] < best[0] else best
        kname = al.kernel_name(0)
    same = all(np.array_equal(got[k], want[k]) for k in ("aln_score", "query_batch_end", "target_batch_end"))
    return {"pairs": b.n, "cells": b.cells, "cpu": {"kind": kind, "cores": threads, "seconds": cpu_s, "gcups": b.cells / cpu_s / 1e9},
            "gpu": {"kernel": kname, "pipeline_ms": best[0], "dp_ms": best[2], "gcups_pipeline": b.cells / best[0] / 1e6, "gcups_dp": b.cells / best[2] / 1e6},
            "results_identical": bool(same)}


def real_reads_extra(timed):
    """The reference's own 20,000 real read / window pairs (GASAL2/test_prog/*.fasta.gz; 4,000 of them are committed as
    tests/golden/real4000_inputs.npz), replicated to 264,000 pairs in their original (unsorted, ragged: 150 bp vs 152-277 bp) order."""
    from genomics_gpu_b200 import capi
    from genomics_gpu_b200.synth import PairBatch
    g = np.load(os.path.join(ROOT, "tests", "golden", "real4000_inputs.npz"))
    def ascii_of(words):                  # 4-bit words -> one ASCII byte per base with the same low nibble (first base in bits 31:28)
        sh = np.array([28, 24, 20, 16, 12, 8, 4, 0], dtype=np.uint32)
        return (((words[:, None] >> sh[None, :]) & 15) | 0x40).astype(np.uint8).reshape(-1)
    q1, t1 = ascii_of(g["packed_q"]), ascii_of(g["packed_t"])
    reps = 66
    qb, tb = np.tile(q1, reps), np.tile(t1, reps)
    qo = (g["q_offs"].astype(np.uint64)[None, :] + (np.arange(reps, dtype=np.uint64) * np.uint64(q1.size))[:, None]).reshape(-1).astype(np.uint32)
    to = (g["t_offs"].astype(np.uint64)[None, :] + (np.arange(reps, dtype=np.uint64) * np.uint64(t1.size))[:, None]).reshape(-1).astype(np.uint32)
    b = PairBatch(qb, tb, qo, to, np.tile(g["q_lens"], reps), np.tile(g["t_lens"], reps))
    has_n = float(np.mean([(b"N" in b.query(i)) or (b"N" in b.target(i)) for i in range(0, 4000)]))
    return {"pairs_with_N_fraction": has_n, "local": timed(b, algo=capi.LOCAL), "local_with_start": timed(b, algo=capi.LOCAL, start=capi.WITH_START),
            "semi_global_with_start": timed(b, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)}


def nw_extra(n: int):
    """The suite's NW benchmark shape (Non-CDP/NW/main.cu:67, N = 32000): full int32 matrix, device-timed, against the HBM write
    roofline (4 B per cell); with the reference's own nw_gpu3 on the same GPU when oracle/_ref/nw_ref_gpu travelled here."""
    import re
    import tempfile
    from genomics_gpu_b200 import nw as nwmod, synth
    ref, qry = synth.nw_benchmark_pair(n, seed=7)
    best = None
    for it in range(4):
        _, ms, _ = nwmod.nw_matrix(ref, qry, want_matrix=False)
        if it and (best is None or ms[1] < best):
            best = ms[1]
    nwmod.release()
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f)["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    cells = float(n) * n
    out = {"N": n, "kernel_ms": best, "gcups": cells / best / 1e6, "write_gbs": 4 * cells / best / 1e6, "frac_of_hbm_peak": 4 * cells / best / 1e6 / hbm,
           "bound": "wavefront critical path (2N/4 dependent tile steps), not HBM: see DESIGN.md K6"}
    exe = os.path.join(ROOT, "oracle", "_ref", "nw_ref_gpu")
    if os.path.exists(exe):
        try:
            with tempfile.TemporaryDirectory() as td:      # the reference appends runtimes_*.txt to its cwd
                p = subprocess.run([exe, "-N", str(n), "-3"], cwd=td, capture_output=True, text=True, timeout=300)
            txt = re.sub(r"\x1b\[[0-9;]*m", "", p.stdout)
            m = re.search(r"GPU kernel time \(version 3\): ([0-9.]+) ms", txt)
            c = re.search(r"CPU time: ([0-9.]+) ms", txt)
            out["reference_nw_gpu3_ms"] = float(m.group(1)) if m else None
            out["reference_nw_cpu_ms_1_core"] = float(c.group(1)) if c else None
            out["reference_verified_itself"] = "Verification succeeded" in txt
        except Exception as e:
            out["reference_error"] = str(e)[:100]
    return out


# ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from genomics_gpu_b200 import capi
    from genomics_gpu_b200.aligner import Aligner, pack_host
    from genomics_gpu_b200.synth import PairBatch

    lib = capi.load()
    lib.gasal_set_device(local_rank, 0)
    if os.environ.get("GASALX_HOST_WC") == "1":          # experiment switch: write-combined pinned sequence pages
        lib.gasalx_set_host_write_combined(1)
    prev_affinity = bind_to_gpu_numa_node(local_rank) if args.numa_bind else None
    n = args.pairs
    batch = make_workload(n, seed=2 + rank)
    cells_mode = batch.cells
    q_bytes, t_bytes = int(batch.q_bytes.size), int(batch.t_bytes.size)
    modes = [("local", capi.LOCAL), ("semi_global", capi.SEMI_GLOBAL)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only arm: one storage per mode holding the whole batch, inputs resident in HBM as 4-bit words ------------
    res_al = {}
    for name, algo in modes:
        al = Aligner(algo=algo, start=capi.WITH_START, scoring=SCORING, max_query_len=QLEN + 7, max_target_len=TLEN + 7, max_n_alns=n,
                     n_streams=1, device=local_rank)
        al.fill(0, batch)
        al.launch(0, q_bytes, t_bytes, n)      # uploads the inputs once; afterwards they are resident in HBM
        al.wait(0)
        res_al[name] = al

    def resident_step():
        out = {}
        for name, _ in modes:
            al = res_al[name]
            s = al.vec.a[0]
            s.current_n_alns = n
            lib.gasalx_aln_device_resident(C.byref(s), q_bytes, t_bytes, n, C.byref(al.params))
            al.wait(0)
            out[name] = al.last_ms(0)
        return out

    for _ in range(args.warmup):
        resident_step()
    barrier()
    per_mode = {name: [] for name, _ in modes}
    for _ in range(args.steps):
        r = resident_step()
        for k, v in r.items():
            per_mode[k].append(v)
    barrier()
    kernel_ms = sum(v[0] for vs in per_mode.values() for v in vs)            # CUDA-event time of the DP kernels (no copies)
    kernel_names = {name: res_al[name].kernel_name(0) for name, _ in modes}
    for al in res_al.values():
        al.close()

    # ---- pack arm: the on-GPU 4-bit pack kernel (reference layout: ASCII pages, gasal_align.cu:190-201) on the same sequences ----
    lib.gasalx_set_host_packing(0)
    pal = Aligner(algo=capi.LOCAL, start=capi.WITH_START, scoring=SCORING, max_query_len=QLEN + 7, max_target_len=TLEN + 7, max_n_alns=n, n_streams=1, device=local_rank)
    pal.fill(0, batch); pal.launch(0, q_bytes, t_bytes, n); pal.wait(0)
    pack_times = []
    for it in range(args.warmup + min(args.steps, 10)):
        s = pal.vec.a[0]; s.current_n_alns = n
        lib.gasalx_aln_device_resident(C.byref(s), q_bytes, t_bytes, n, C.byref(pal.params)); pal.wait(0)
        if it >= args.warmup:
            pack_times.append(pal.last_ms(0)[1])
    pal.close()
    lib.gasalx_set_host_packing(1)

    # ---- pipeline arm (the HEADLINE): host buffers -> results through the reference-facing C ABI ------------------------------
    # The batch is split into chunks, one storage (pinned page + stream) per chunk and mode.  The pinned pages are filled ONCE with
    # gasal_host_batch_add (which packs to 4 bits while it copies: the default host-packing mode); every step re-submits them with
    # gasal_aln_async, i.e. the timed region holds the pinned H2D of all sequences + lens/offsets, the kernels and the D2H of
    # the results, pipelined over the storages' streams as a GASAL2 caller would.  Two clocks on the SAME region:
    #   value : SURVEY 8(d) primary -- CUDA events, first H2D enqueue .. last D2H complete of the step (gasalx_span_ms), summed;
    #   e2e   : host wall clock bracketed by barrier + cuda synchronize, including the host-side read of every result array.
    n_chunks = args.e2e_chunks
    chunk = (n + n_chunks - 1) // n_chunks
    subs = [batch.slice(lo, min(lo + chunk, n)) for lo in range(0, n, chunk)]

    def pipeline_arm(kind: str, steps: int, warm: int):
        """kind: 'default' (host-packed pages), 'ascii' (reference layout, k_pack on the GPU), 'ispacked' (caller-packed words)"""
        packed = kind == "ispacked"
        lib.gasalx_set_host_packing(0 if kind == "ascii" else 1)
        staged = [PairBatch(pack_host(sb.q_bytes).view(np.uint8), pack_host(sb.t_bytes).view(np.uint8), sb.q_offs, sb.t_offs, sb.q_lens, sb.t_lens)
                  for sb in subs] if packed else subs
        if any(int(sb.q_bytes.size) % 8 or int(sb.t_bytes.size) % 8 for sb in staged):     # gasal_aln_async wants multiples of 8 bytes
            lib.gasalx_set_host_packing(1)
            return None
        als = {}
        for name, algo in modes:
            al = Aligner(algo=algo, start=capi.WITH_START, scoring=SCORING, max_query_len=QLEN + 7, max_target_len=TLEN + 7, max_n_alns=chunk,
                         n_streams=len(subs), device=local_rank, is_packed=packed)
            for si, sb in enumerate(staged):
                al.fill(si, sb)
            als[name] = al
        lib.gasalx_set_host_packing(1)
        csum = np.zeros(1, dtype=np.int64)
        group = (C.POINTER(capi.GpuStorage) * (2 * len(staged)))()
        k = 0
        for name, _ in modes:
            for si in range(len(staged)):
                group[k] = C.pointer(als[name].vec.a[si]); k += 1
        span = (C.c_float * 1)()
        spans = []

        def step():
            # all storages of both modes are put in flight before the first wait, as a caller with 2 x 16 storages would: the
            # H2D of later chunks (and of the second mode) overlaps the kernels of earlier ones
            for name, _ in modes:
                al = als[name]
                for si, sb in enumerate(staged):
                    s = al.vec.a[si]
                    lib.gasalx_host_batch_rearm(C.byref(s), int(sb.q_bytes.size), int(sb.t_bytes.size), sb.n)
                    al.launch(si, int(sb.q_bytes.size), int(sb.t_bytes.size), sb.n)
            for name, _ in modes:
                al = als[name]
                for si, sb in enumerate(staged):
                    al.wait(si)
                    r = al.results(si, sb.n)          # host_res arrays (D2H already done on the stream)
                    csum[0] += int(r["aln_score"].sum())
            if lib.gasalx_span_ms(group, len(group), span) == 0:
                spans.append(float(span[0]))

        for _ in range(warm):
            step()
        spans.clear(); csum[0] = 0
        launches0 = sum(al.launches(si) for al in als.values() for si in range(len(staged)))
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        launches = sum(al.launches(si) for al in als.values() for si in range(len(staged))) - launches0
        host_packed = bool(lib.gasalx_storage_host_packed(C.byref(als["local"].vec.a[0])))
        for al in als.values():
            al.close()
        seq_bytes = sum(int(sb.q_bytes.size) + int(sb.t_bytes.size) for sb in staged)
        if host_packed:
            seq_bytes //= 2
        return {"wall_s": wall, "span_ms": sum(spans), "n_spans": len(spans), "checksum": int(csum[0]), "h2d_seq_bytes": 2 * seq_bytes, "launches": launches, "clocks": clocks}

    main_arm = pipeline_arm("default", args.steps, max(3, args.warmup))
    clocks = main_arm["clocks"]
    side_steps = max(3, args.steps // 2)
    ascii_arm = pipeline_arm("ascii", side_steps, 2) if args.e2e_ascii else None
    ispacked_arm = pipeline_arm("ispacked", side_steps, 2) if args.e2e_packed else None

    # ---- reductions over ranks: time = max, work = sum -------------------------------------------------------
    t = torch.tensor([main_arm["span_ms"], 1e3 * main_arm["wall_s"], kernel_ms,
                      1e3 * ascii_arm["wall_s"] if ascii_arm else 0.0, 1e3 * ispacked_arm["wall_s"] if ispacked_arm else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    span_ms_max, wall_ms_max, kernel_ms_max, ascii_ms_max, ispacked_ms_max = [float(x) for x in t.tolist()]
    total_cells_per_step = 2 * cells_mode * world            # weak scaling: every rank has its own n pairs
    value = total_cells_per_step * args.steps / (span_ms_max * 1e-3) / 1e9
    e2e_value = total_cells_per_step * args.steps / (wall_ms_max * 1e-3) / 1e9
    kernel_only = total_cells_per_step * args.steps / (kernel_ms_max * 1e-3) / 1e9

    sharded = run_sharded(args, rank, world, local_rank) if args.sharded else None

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant DP kernel (by time in the step) ---------------------------------------------
    R, R_src = dpx_peak_ops_per_s()
    dp_ms = {name: statistics.mean(v[2] for v in per_mode[name]) for name, _ in modes}
    top = max(dp_ms, key=dp_ms.get)
    ops_per_cell = 3.0 if kernel_names[top] == "s16x2" else 6.0
    peak_gcups = R / ops_per_cell / 1e9
    achieved = cells_mode / (dp_ms[top] * 1e-3) / 1e9
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0
    pack_bytes = 1.5 * (q_bytes + t_bytes)
    pack_gbs = pack_bytes / (statistics.mean(pack_times) * 1e-3) / 1e9

    if prev_affinity:
        os.sched_setaffinity(0, prev_affinity)      # the CPU baseline uses every host core
    cpu_threads = len(os.sched_getaffinity(0)) or 1
    cpu_val, cpu_kind, cpu_n, cpu_dt = cpu_baseline_run(batch, args.cpu_pairs, cpu_threads)
    refgpu = ref_gpu_baseline(batch, args.refgpu_pairs) if args.refgpu_pairs > 0 else {"unavailable": "disabled"}
    extras = run_extras(local_rank, 262_144) if args.extras else {"skipped": True}
    if sharded is not None:
        extras["sharded"] = sharded

    def side(arm, ms_max, note):
        if not arm:
            return None
        return {"value": total_cells_per_step * side_steps / (ms_max * 1e-3) / 1e9, "unit": "GCUPS", "steps": side_steps, "h2d_bytes_per_step": arm["h2d_seq_bytes"] + 2 * 16 * n,
                "score_checksum_per_step_matches": bool(arm["checksum"] * args.steps == main_arm["checksum"] * side_steps), "note": note}

    line = {
        "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": span_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "s16x2/int32 (DPX integer)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": n,
                   "modes_per_step": ["local WITH_START", "semi_global TARGET/TARGET WITH_START"], "scoring": list(SCORING),
                   "l2_policy": "inputs larger than L2 (376 MB of packed sequences per mode and step)",
                   "kernels": kernel_names, "storages_per_mode": len(subs),
                   "timing": "value: CUDA events, first H2D enqueue .. last D2H complete of each step over all storages (SURVEY 8d primary), "
                             "summed over the steps, max over ranks; kernel_only: DP kernels alone on HBM-resident inputs"},
        "pairs_per_s": 2 * n * world * args.steps / (span_ms_max * 1e-3),
        "kernel_only": {"value": kernel_only, "unit": "GCUPS", "ms_per_step": kernel_ms_max / args.steps,
                        "note": "DP kernels (forward + start pass) on inputs resident in HBM, no copies: explains `value`, is not the headline"},
        "per_mode": {name: {"dp_ms": dp_ms[name], "gcups_dp_only": cells_mode / (dp_ms[name] * 1e-3) / 1e9} for name, _ in modes},
        "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": main_arm["h2d_seq_bytes"] + 2 * 16 * n, "d2h_bytes_per_step": 2 * 20 * n,
                "wall_ms_per_step": wall_ms_max / args.steps, "chunks_per_mode": len(subs), "score_checksum": main_arm["checksum"],
                "timing": "host wall clock bracketed by barrier + cuda synchronize around the same steps as `value`; every result array is read on the host",
                "host_pages": "4-bit words packed by gasal_host_batch_add while it copies the caller's ASCII (default host packing); pages filled once, re-submitted every step",
                "pairs_per_s": 2 * n * world * args.steps / (wall_ms_max * 1e-3),
                "ascii_pages": side(ascii_arm, ascii_ms_max, "reference layout: 1 B/base ASCII pages over PCIe, 4-bit pack kernel on the GPU (gasalx_set_host_packing(0))"),
                "ispacked_api": side(ispacked_arm, ispacked_ms_max, "the reference API's isPacked input: the CALLER hands over 4-bit words (src/ctors.cpp:64-67)")},
        "gpu_launches": int(main_arm["launches"]),
        "clocks": clocks,
        "roofline": {"bound": "dpx-integer-pipe", "kernel": f"{top} DP ({kernel_names[top]})", "achieved": achieved, "peak": peak_gcups, "unit": "GCUPS",
                     "frac": achieved / peak_gcups, "traffic": ncu_traffic(top, n // max(1, len(subs))), "algorithmic_bytes": 396 * (n // max(1, len(subs))),
                     "peak_def": f"R / {ops_per_cell:g} ops per cell, R = {R:.4g} lane-ops/s, {R_src}",
                     "note": "compute-bound integer DP (SURVEY 8d): compulsory HBM traffic is < 0.01 B/cell; the HBM-bound stage is pack, below. "
                             "achieved counts FORWARD cells only while the launch also runs the start pass; "
                             "other_configs.C2_forward_only has the same kernels without it"},
        "roofline_pack": {"bound": "hbm", "achieved": pack_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": pack_gbs / hbm_peak,
                          "traffic": ncu_traffic("pack", n // max(1, len(subs))), "bytes_def": "1.5 B per base (1 read + 0.5 written)",
                          "note": "k_pack is on the path only for ASCII pages (host packing off); timed here on the same sequences"},
        "cpu_baseline": {"value": cpu_val, "unit": "GCUPS", "cores": cpu_threads, "kind": cpu_kind,
                         "sample": f"first {cpu_n} pairs of the same workload, both modes, {cpu_dt:.1f} s"},
        "ref_gpu_baseline": refgpu,
        "other_configs": extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_sharded(args, rank, world, local_rank):
    """STRONG-scaling job through the sharder (north_star: "length-balanced sharding across 8 B200", BASELINE config 5): 200,000
    mixed-length pairs (100 bp - 10 kb, local WITH_START).  Rank 0 plans the length-balanced assignment (gasalx_plan_batches: length
    binning + LPT on DP cells), every rank aligns its shard through the library batcher (gasalx_align_all on its own GPU), and the
    results are gathered on rank 0 in the caller's order -- no data-path collective.  Every rank builds the same seeded input (8,000
    distinct synthetic pairs x 25) instead of shipping 1 GB of sequence through the rendezvous.  Also C3 at 16 pairs per GPU."""
    import torch
    import torch.distributed as dist
    from genomics_gpu_b200 import capi, synth
    from genomics_gpu_b200.aligner import align_all
    from genomics_gpu_b200.sharding import shard_indices
    from genomics_gpu_b200.synth import PairBatch
    out = {}
    distinct, reps = args.sharded_distinct, args.sharded_reps
    base = synth.config5(distinct, seed=5)
    qb, tb = np.tile(base.q_bytes, reps), np.tile(base.t_bytes, reps)
    qo = (base.q_offs.astype(np.uint64)[None, :] + (np.arange(reps, dtype=np.uint64) * np.uint64(base.q_bytes.size))[:, None]).reshape(-1)
    to = (base.t_offs.astype(np.uint64)[None, :] + (np.arange(reps, dtype=np.uint64) * np.uint64(base.t_bytes.size))[:, None]).reshape(-1)
    ql, tl = np.tile(base.q_lens, reps), np.tile(base.t_lens, reps)
    n = int(ql.size)
    cells = (ql.astype(np.int64) * tl.astype(np.int64))
    plan = [None]
    if rank == 0:
        plan[0] = shard_indices(ql, tl, world, max_batch_pairs=2048)
    if world > 1:
        dist.broadcast_object_list(plan, src=0)
    idx = plan[0][rank]
    mine = PairBatch(qb, tb, qo[idx], to[idx], ql[idx], tl[idx])
    keys = ("aln_score", "query_batch_end", "target_batch_end", "query_batch_start", "target_batch_start")

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    best = None
    for it in range(3):                              # one warm-up (storages, pinned pages), two timed passes: best counts
        sync()
        t0 = time.perf_counter()
        res = align_all(mine, algo=capi.LOCAL, start=capi.WITH_START, scoring=SCORING, n_gpus=1, gpu_ids=[local_rank], streams_per_gpu=4) if mine.n else \
            {k: np.zeros(0, dtype=np.int32) for k in keys}
        payload = np.stack([res[k] for k in keys]) if mine.n else np.zeros((5, 0), dtype=np.int32)
        gathered = [None] * world
        if world > 1:
            dist.gather_object((idx, payload), gathered if rank == 0 else None, dst=0)       # the host gather: 20 bytes per pair
        else:
            gathered[0] = (idx, payload)
        full = None
        if rank == 0:
            full = np.zeros((5, n), dtype=np.int32)
            for i2, p2 in gathered:
                full[:, i2] = p2
        sync()
        dt = time.perf_counter() - t0
        if it and (best is None or dt < best[0]):
            best = (dt, full)
    tt = torch.tensor([best[0]], dtype=torch.float64, device="cuda")
    per_rank_cells = torch.zeros(world, dtype=torch.float64, device="cuda"); per_rank_cells[rank] = float(cells[idx].sum())
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(per_rank_cells, op=dist.ReduceOp.SUM)
    lib = capi.load(); lib.gasalx_batcher_release()
    prc = per_rank_cells.tolist()
    if rank == 0:
        out["C5_sharded_strong_scaling"] = {
            "pairs": n, "distinct_pairs": distinct, "cells": int(cells.sum()), "seconds": float(tt[0]), "gcups": float(cells.sum()) / float(tt[0]) / 1e9,
            "scaling": "strong (total work fixed as N grows)", "per_rank_cells": [int(x) for x in prc],
            "imbalance_max_over_mean": max(prc) / (sum(prc) / world), "score_checksum": int(best[1][0].astype(np.int64).sum()),
            "coord_checksum": int(best[1][1:].astype(np.int64).sum()),
            "timing": "wall clock, barrier to barrier: plan already known; per-rank gasalx_align_all (length-binned batches, pinned streaming, results in shard order) + host gather on rank 0",
            "note": "score_checksum / coord_checksum are the same at every N when the shards reproduce the single-GPU results"}
    # C3: 16 pairs of 16k x 16k per GPU (the per-GPU share of BASELINE config 3 on 8 GPUs), weak
    from genomics_gpu_b200.aligner import Aligner
    b3 = synth.config3(16, 16384, seed=3 + rank)
    with Aligner(algo=capi.GLOBAL, max_query_len=16384 + 7, max_target_len=16384 + 7, max_n_alns=16, n_streams=1, scoring=SCORING, device=local_rank) as al:
        best3 = None
        for it in range(3):
            sync()
            al.align(b3)
            ms = al.last_ms(0)
            best3 = ms[2] if best3 is None or ms[2] < best3 else best3
        kname = al.kernel_name(0)
    t3 = torch.tensor([best3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    if rank == 0:
        out["C3_16_pairs_per_gpu"] = {"kernel": kname, "pairs_per_gpu": 16, "dp_ms_max_over_ranks": float(t3[0]), "gcups": world * b3.cells / (float(t3[0]) * 1e-3) / 1e9, "scaling": "weak"}
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--numa-bind", type=int, default=1, help="bind each rank to the CPUs local to its GPU (NVML affinity)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per mode")
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--e2e-packed", type=int, default=1, help="also time the pipeline arm with caller-packed (isPacked) host pages")
    ap.add_argument("--e2e-ascii", type=int, default=1, help="also time the pipeline arm with ASCII host pages (host packing off, pack kernel on the GPU)")
    ap.add_argument("--sharded", type=int, default=1, help="also run the strong-scaling LPT-sharded mixed-length job (C5, 200k pairs) over the ranks and C3 at 16 pairs per GPU")
    ap.add_argument("--sharded-distinct", type=int, default=8000, help="distinct synthetic pairs of the sharded job")
    ap.add_argument("--sharded-reps", type=int, default=25, help="the sharded job replicates its distinct pairs this many times (8000 x 25 = 200,000 pairs)")
    ap.add_argument("--cpu-pairs", type=int, default=600_000, help="bounded sample for the cpu_baseline leg (~15 s on 16 cores)")
    ap.add_argument("--ref-pairs", type=int, default=200_000, help="pairs per step of the --impl reference arm (~5 s per step on 16 cores)")
    ap.add_argument("--extras", type=int, default=1, help="also time the other BASELINE.json configurations on rank 0 (context only)")
    ap.add_argument("--refgpu-pairs", type=int, default=262_144, help="sample for the reference-CUDA-build baseline on the same GPU (0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- query bases/s (PML + chain statistics) of the B200 path, next to the reference's CPU pml_query.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c1|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (backward move-structure traversal, PML + chain id per base) over one batch
of synthetic reads.  Workload at every N: BASELINE.json configs[1] per GPU -- a synthetic 32-haplotype x 10 Mbp
pangenome (+ reverse complements, 0.1 % SNP/indel divergence), marks = tunnels, sub-sample 10, and 10 M x 150 bp
reads with 1 % substitutions -- index replicated per GPU, each rank traversing its own 10 M reads (weak scaling,
no data-path collective).  Prints ONE JSON line (rank 0).

Keys beyond the driver contract:
  value     whole-job bases/s with reads + table resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       same metric through the C-ABI call colbwt_query with HOST (pinned) buffers in and DENSE arrays out: read packing
            (host or device), H2D, traversal, and the results' way back (plain copies, the compact form expanded by the host
            threads, or both alternating: the library measures and says which in `packing` / `transport`) all inside the timed
            region; byte counts from the library (colbwt_index_last_bytes).  Under torchrun also `alone_value` (rank 0 running
            the same call while the others idle) and `efficiency_vs_alone`
  e2e_compact  the same reads through colbwt_query_compact (match bit per base + non-zero chain ids: what a consumer that
            needs no dense arrays takes), expanded once on the host after the timed region for the parity check
  roofline  the traversal kernel against the measured HBM stream peak (MEASURED_PEAKS.json), algorithmic bytes =
            35.25 B/base (32 B gathered sector + 0.25 B read in + 3 B written); `gather` adds the
            measured random-32-B-sector rate over a buffer of the index's size and the fraction of it achieved
  cpu_baseline  the reference's own col_pml::query_pml (oracle/_ref, compiled from the reference sources) -- or the C
            port when that is absent -- on the host cores, on a bounded sample of the same reads
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_BASE = 35.25   # SURVEY.md section 8d / DESIGN.md (32 B gathered + 0.25 B in + 3 B out; kept fixed even when PML is written as u8)

WORKLOADS = {
    # name: (haplotypes, genome bp, snp, indel, reads, read_len, sub, ins, del, len_sigma, tree)
    "c1": dict(H=4, G=1_000_000, snp=1e-3, indel=0.0, reads=100_000, read_len=150, sub=0.01, ins=0.0, dele=0.0, len_sigma=0.0, tree=False),
    "c2": dict(H=32, G=10_000_000, snp=9e-4, indel=1e-4, reads=10_000_000, read_len=150, sub=0.01, ins=0.0, dele=0.0, len_sigma=0.0, tree=False),
    "c2small": dict(H=32, G=1_000_000, snp=9e-4, indel=1e-4, reads=2_000_000, read_len=150, sub=0.01, ins=0.0, dele=0.0, len_sigma=0.0, tree=False),
    "c3small": dict(H=64, G=5_000_000, snp=1e-3, indel=1e-4, reads=400_000, read_len=10_000, sub=0.02, ins=0.015, dele=0.015, len_sigma=0.5, tree=True),
}
# configs[2] as near its stated size as the tooling's GPU suffix sorter allows (ranks are packed in 31 bits and the working set must fit 180 GB
# of HBM): 64 haplotypes x 12 Mbp (+ reverse complements: n = 1.54e9); 125 k reads of ~10 kbp per GPU = the stated 1 M reads on 8 GPUs
WORKLOADS["c3"] = dict(WORKLOADS["c3small"], G=12_000_000, reads=125_000)
WORKLOADS["c3big"] = dict(WORKLOADS["c3small"], G=16_000_000, reads=125_000)   # n = 2.05e9: just under the sorter's 2^31 limit
for _s in (1, 100):   # configs[3]: sub-sample sweep on the 32-haplotype index (tunnel marking; `all` mode is covered by golden fixtures)
    WORKLOADS[f"c4_s{_s}"] = dict(WORKLOADS["c2"], split_rate=_s)
for _s in (1, 10, 100):   # non-tunnel marking: marks come from the product's own GPU col_split (-m all), table from from_primaries
    WORKLOADS[f"c4_all_s{_s}"] = dict(WORKLOADS["c2"], split_rate=_s, mode="all")
WORKLOADS["c5small"] = dict(synthetic_rows=250_000_000, mean_len=16.0, reads=10_000_000, read_len=150, sub=0.01)
# configs[4]: 1e9 rows (16 GB packed table per replica), thresholds fitted to a real index (synthdata.pipeline.synth_move_table
# snap=...), 100 M mixed reads over 8 GPUs = per GPU 12.4 M x 150 bp at 1 % + 125 k x 10 kbp at 5 % substitutions (3.1 Gbases)
WORKLOADS["c5"] = dict(synthetic_rows=1_000_000_000, mean_len=16.0, snap=0.854, reads=12_400_000, read_len=150, sub=0.01,
                       long_reads=125_000, long_len=10_000, long_sub=0.05)
WORKLOADS["c5mid"] = dict(WORKLOADS["c5"], synthetic_rows=250_000_000, reads=6_200_000, long_reads=62_500)
WORKLOADS["tiny"] = dict(H=4, G=50_000, snp=1e-3, indel=0.0, reads=5_000, read_len=150, sub=0.01, ins=0.0, dele=0.0, len_sigma=0.0, tree=False)
WORKLOAD_TEXT = {
    "tiny": "4-haplotype x 50 kbp toy (CPU self-test of bench.py only)",
    "c4_s1": "configs[3]: the configs[1] pangenome marked with tunnels -s 1 (densest chain ids), 10M x 150 bp reads",
    "c4_s100": "configs[3]: the configs[1] pangenome marked with tunnels -s 100 (sparsest chain ids), 10M x 150 bp reads",
    "c4_all_s1": "configs[3]: the configs[1] pangenome marked with `-m all -s 1` by colbwt_col_split, 10M x 150 bp reads",
    "c4_all_s10": "configs[3]: the configs[1] pangenome marked with `-m all -s 10` by colbwt_col_split, 10M x 150 bp reads",
    "c4_all_s100": "configs[3]: the configs[1] pangenome marked with `-m all -s 100` by colbwt_col_split, 10M x 150 bp reads",
    "c5": "configs[4]: 1e9 directly synthesised move rows (16 GB packed table per replica, thresholds fitted to a real index), per GPU 12.4M x 150 bp (1% subst.) + 125k x 10 kbp (5% subst.) LF-walk reads = 100M mixed reads over 8 GPUs",
    "c5mid": "configs[4] at a quarter of its size: 2.5e8 fitted synthetic rows, per GPU 6.2M x 150 bp + 62.5k x 10 kbp LF-walk reads",
    "c5small": "configs[4] scaled to 2.5e8 directly synthesised move rows (4 GB packed table, DRAM-resident), 10M x 150 bp LF-walk reads, 1% substitutions",
    "c1": "configs[0]: 4-haplotype x 1 Mbp pangenome (+revcomp), tunnels -s 10, 100k x 150 bp reads",
    "c2": "configs[1]: 32-haplotype x 10 Mbp pangenome (+revcomp, 0.1% SNP/indel divergence), tunnels -s 10, 10M x 150 bp reads, 1% substitutions",
    "c2small": "configs[1] scaled down 10x in genome length and 5x in reads (smoke runs only)",
    "c3": "configs[2]: 64-haplotype x 12 Mbp tree-structured pangenome (+revcomp, n = 1.54e9: the largest the tooling's GPU suffix sorter fits in 180 GB; stated 50 Mbp), per GPU 125k x ~10 kbp (log-normal) nanopore-like reads at 5% error = the stated 1M reads on 8 GPUs",
    "c3big": "configs[2]: 64-haplotype x 16 Mbp tree-structured pangenome (+revcomp, n = 2.05e9), per GPU 125k x ~10 kbp nanopore-like reads at 5% error",
    "c3small": "configs[2] scaled down: 64-haplotype x 5 Mbp tree-structured pangenome, 400k x 10 kbp (log-normal lengths) nanopore-like reads at 5% error",
}


# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed regions run."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.active = threading.Event()
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].isdigit() else device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------------------
def cache_dir():
    d = os.environ.get("COLBWT_BENCH_CACHE", "/tmp/colbwt_bench_cache")
    os.makedirs(d, exist_ok=True)
    return d


def build_workload(name: str, device: str, verbose: bool):
    """Generates (or loads from the per-box cache) the index file and the text the reads are drawn from."""
    from synthdata import pangenome as P, pipeline as PL
    w = WORKLOADS[name]
    stem = os.path.join(cache_dir(), name)
    meta_path = stem + ".meta.json"
    if "synthetic_rows" in w:
        return build_synthetic_table(name, device, verbose)
    if not os.path.exists(meta_path):
        t0 = time.time()
        haps = P.make_haplotypes(w["G"], w["H"], snp=w["snp"], indel=w["indel"], seed=1, tree=w["tree"])
        idx = PL.build_index(haps, with_revcomp=True, split_rate=w.get("split_rate", 10), min_mum=20, device=device, verbose=verbose)
        if w.get("mode") == "all":
            # the tooling only marks tunnels; `-m all` comes from the product's GPU build chain on the same primaries
            import col_bwt_b200 as cb
            PL.write_reference_inputs(stem + ".fa", idx)
            cb.col_split(stem + ".fa", "all", w.get("split_rate", 10))
            t = cb.ColPml.from_primaries(stem + ".fa")
            t.save(stem + ".col_pml")
            idx["columns"] = {"n": t.n, "ch": np.zeros(t.r, np.uint8), "bwt_r": t.bwt_r, "col_id": np.zeros(int(t.stats.marked_rows), np.uint8) + 1}
            t.close()
        else:
            PL.write_col_pml(stem + ".col_pml", idx["columns"])
        np.save(stem + ".text.npy", idx["text"])
        np.save(stem + ".seq_starts.npy", idx["seq_starts"])
        cols = idx["columns"]
        meta = {"n": int(cols["n"]), "r": int(cols["ch"].size), "bwt_r": int(cols["bwt_r"]), "mums": int(idx["mum_len"].size),
                "marked_rows": int((cols["col_id"] > 0).sum()), "build_s": time.time() - t0}
        with open(meta_path + ".tmp", "w") as f:
            json.dump(meta, f)
        os.replace(meta_path + ".tmp", meta_path)
    meta = json.load(open(meta_path))
    return stem + ".col_pml", np.load(stem + ".text.npy", mmap_mode="r"), np.load(stem + ".seq_starts.npy"), meta


def synth_workload(name: str, rank: int, device: str, want_file: bool, verbose: bool):
    """configs[4]-style workload, generated where it is used: EVERY rank synthesises the same move table (same seed) in its
    own HBM and walks its own reads on it (seed by rank), so that nothing is generated by one rank while the others wait and
    no 18 GB file has to be read back by each of them; the rows are handed to the library in device memory
    (colbwt_index_from_rows with a device pointer).  Only the rank that runs the CPU checker also writes the table as a
    `.col_pml` file (the reference's loader reads files).  Returns (rows tensor, n, seqs, off, path or None)."""
    import torch
    from synthdata import pipeline as PL
    w = WORKLOADS[name]
    t0 = time.time()
    rows, n, cols = PL.synth_move_table(w["synthetic_rows"], mean_len=w["mean_len"], device=device, snap=w.get("snap", 0.0))
    seqs, off = PL.walk_reads(cols, w["reads"], w["read_len"], sub=w["sub"], seed=6 + rank)
    if w.get("long_reads"):
        s2, o2 = PL.walk_reads(cols, w["long_reads"], w["long_len"], sub=w["long_sub"], seed=1006 + rank, chunk_reads=1 << 18)
        seqs = np.concatenate([seqs, s2])
        off = np.concatenate([off, o2[1:] + off[-1]]).astype(np.uint64)
        del s2, o2
    del cols
    torch.cuda.empty_cache()
    path = None
    if want_file:
        path = os.path.join(cache_dir(), name + ".col_pml")
        with open(path, "wb") as f:
            f.write(np.array([rows.shape[0], n, rows.shape[0], rows.shape[0]], dtype="<u8").tobytes())
            step = 1 << 24
            for a in range(0, rows.shape[0], step):
                f.write(rows[a:a + step].cpu().numpy().tobytes())
    if verbose:
        print(f"[bench] rank {rank}: synthetic table {name}: {rows.shape[0]} rows, n={n}, {off.size - 1} reads, {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
    return rows, n, seqs, off, path


def build_synthetic_table(name: str, device: str, verbose: bool):
    """configs[4]-style workload: a directly synthesised move table (no text).  The table file and one set of LF-walk
    reads per rank seed are cached; `text` is returned as None and `seq_starts` carries the cache stem instead."""
    import torch
    from synthdata import pipeline as PL
    w = WORKLOADS[name]
    stem = os.path.join(cache_dir(), name)
    meta_path = stem + ".meta.json"
    if not os.path.exists(meta_path):
        t0 = time.time()
        rows, n, cols = PL.synth_move_table(w["synthetic_rows"], mean_len=w["mean_len"], device=device)
        with open(stem + ".col_pml", "wb") as f:
            f.write(np.array([rows.shape[0], n, rows.shape[0], rows.shape[0]], dtype="<u8").tobytes())
            step = 1 << 24
            for a in range(0, rows.shape[0], step):
                f.write(rows[a:a + step].cpu().numpy().tobytes())
        del rows
        torch.cuda.empty_cache()
        for seed_rank in range(int(os.environ.get("WORLD_SIZE", "1"))):
            seqs, off = PL.walk_reads(cols, w["reads"], w["read_len"], sub=w["sub"], seed=6 + seed_rank)
            np.save(stem + f".reads{seed_rank}.npy", seqs)
        meta = {"n": int(n), "r": int(w["synthetic_rows"]), "bwt_r": int(w["synthetic_rows"]), "mums": 0, "marked_rows": -1,
                "build_s": time.time() - t0}
        del cols
        torch.cuda.empty_cache()
        with open(meta_path + ".tmp", "w") as f:
            json.dump(meta, f)
        os.replace(meta_path + ".tmp", meta_path)
        if verbose:
            print(f"[bench] synthetic table {name}: {meta}", file=sys.stderr, flush=True)
    return stem + ".col_pml", None, stem, json.load(open(meta_path))


def make_reads(name: str, text, seq_starts, rank: int, n_reads: int | None, device: str):
    from synthdata import pangenome as P
    w = WORKLOADS[name]
    n = n_reads or w["reads"]
    if text is None:   # synthetic table: cached LF-walk reads (seq_starts is the cache stem)
        path = f"{seq_starts}.reads{rank}.npy"
        if not os.path.exists(path):
            path = f"{seq_starts}.reads0.npy"
        seqs = np.load(path)[: n * w["read_len"]]
        return np.ascontiguousarray(seqs), np.arange(n + 1, dtype=np.uint64) * np.uint64(w["read_len"])
    return P.sample_reads_device(np.asarray(text), seq_starts, n, w["read_len"], sub=w["sub"], ins=w["ins"], dele=w["dele"],
                                 len_sigma=w["len_sigma"], seed=2 + rank, device=device)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(workload: str):
    """dram bytes per launch of the traversal kernel from the committed ncu capture of this workload, if any:
    {"bytes": dram__bytes_read.sum + dram__bytes_write.sum of one launch, "from": which report}.  Not measured in this run
    (hardware counters need ncu, and a number taken under ncu is never a bench value): the line says where it comes from."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        v = json.load(open(p)).get(workload)
        if isinstance(v, dict):
            return v
        if v:
            return {"bytes": v, "from": "profiles/traffic.json (ncu --set full capture of this kernel on this workload, another box of the pool)"}
    return None


def cpu_reference(path: str):
    """The CPU checker/baseline: the reference's own code if oracle/_ref travelled here, else the C port."""
    import oracle
    if oracle.have_ref():
        return oracle.Reference(path), "reference"
    return oracle.Oracle(path), "port"


def time_cpu(ref, kind, seqs, off, threads, target_s):
    """Reference on a bounded sample: calibrate on 2k reads, then size the sample for ~target_s seconds."""
    n_reads = off.size - 1
    k0 = min(n_reads, 2000)

    def run(k):
        o = off[: k + 1]
        s = seqs[: int(o[-1])]
        t0 = time.perf_counter()
        if kind == "reference":
            ref.query_batch(s, o, threads=threads, want_output=False)
        else:
            ref.query_batch(s, o, want_output=False)
        return int(o[-1]), time.perf_counter() - t0
    b0, t0 = run(k0)
    rate = b0 / max(t0, 1e-6)
    k = int(min(n_reads, max(k0, target_s * rate / max(1, b0 / k0))))
    return k, run


# ------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(obj) -> None:
    """The one JSON line goes to the process's original stdout; everything else any library prints to fd 1 (NCCL's
    version banner, for one) has been redirected to stderr by main()."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def pml_width_for(cb, max_len: int) -> int:
    width = cb.PML_U8 if max_len < 256 else (cb.PML_U16 if max_len < 65536 else cb.PML_U32)
    return max(width, int(os.environ.get("COLBWT_BENCH_PML_WIDTH", "0")))   # development: measure a wider PML type


def pml_invariants(pml_d, off, n_inv):
    """Size-independent invariants on a large slice of the measured output: PML[j] is 0 or PML[j+1]+1
    inside a read, and never exceeds the bases left of the read."""
    sl = pml_d[: int(off[n_inv])].astype(np.int64)
    nxt = np.empty_like(sl)
    nxt[:-1] = sl[1:]
    lens = np.diff(off[: n_inv + 1]).astype(np.int64)
    ends = off[1: n_inv + 1].astype(np.int64)[lens > 0] - 1
    nxt[ends] = 0
    left = np.repeat(off[1: n_inv + 1].astype(np.int64), lens) - np.arange(sl.size)
    return bool((((sl == 0) | (sl == nxt + 1)) & (sl <= left)).all())


def cli_baselines(path: str, seqs, off, cpu_seconds: float):
    """the reference's own pml_query executable, end to end (index load, FASTA parse, traversal,
    text output), as shipped (MULTI_THREAD on: two std::thread per mismatching base, common.hpp:50 / col_bwt.hpp:537-546)
    and with that one macro off -- both compiled from the reference sources into oracle/_ref.  Bounded samples."""
    import subprocess
    import tempfile
    import oracle
    out = {}
    prefix = path[: -len(".col_pml")] if path.endswith(".col_pml") else path
    for key, exe, budget_bases in (("nomt_e2e", "pml_query_nomt", 30_000_000), ("as_shipped", "pml_query", 225_000)):
        binp = oracle.ref_bin(exe)
        if not os.path.exists(binp):
            continue
        k = int(max(1, min(off.size - 1, np.searchsorted(off, np.uint64(budget_bases), side="right") - 1)))
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "sample.fa")
            with open(fa, "wb") as f:
                for i in range(k):
                    f.write(b">r%d\n" % i)
                    f.write(seqs[int(off[i]): int(off[i + 1])].tobytes())
                    f.write(b"\n")
            # index load alone (a query file with one empty read), subtracted to leave parse + traversal + text output
            empty = os.path.join(td, "empty.fa")
            open(empty, "wb").write(b">e\n\n")
            t0 = time.perf_counter()
            subprocess.run([binp, prefix, "-p", empty], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
            t_load = time.perf_counter() - t0
            t0 = time.perf_counter()
            r = subprocess.run([binp, prefix, "-p", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=max(60.0, 20 * cpu_seconds))
            t_all = time.perf_counter() - t0
        bases = int(off[k] - off[0])
        out[key] = {"value": bases / max(1e-9, t_all - t_load), "unit": "bases/s", "cores": 1 if key == "nomt_e2e" else "1 + 2 transient threads per mismatch",
                    "sample": f"first {k} reads ({bases} bases): oracle/_ref/{exe} PREFIX -p sample.fa, {t_all:.2f} s wall minus {t_load:.2f} s index load; rc {r.returncode}",
                    "with_index_load": bases / t_all}
    return out


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("COLBWT_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=None, help="reads per GPU (default: the workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--check-reads", type=int, default=200_000, help="reads compared with the CPU checker after the timed runs (0 = every read)")
    ap.add_argument("--inproc", type=int, default=0, help="ONE process driving this many GPUs through colbwt_index_load(..., n_devices) (no torchrun)")
    ap.add_argument("--sweep-out", default=None, help="under torchrun: also measure with only the first 1, 2, 4 ... ranks active and append those JSON lines to this file")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    have_gpu = torch.cuda.is_available()

    if a.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(a, have_gpu)

    if not have_gpu:
        emit({"error": "no CUDA device: the B200 path has no CPU fallback"})
        return 1
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import col_bwt_b200 as cb
    dev = f"cuda:{local}"
    n_inproc = max(1, a.inproc)
    if n_inproc > 1 and world > 1:
        emit({"error": "--inproc is a single-process mode; do not combine it with torchrun"})
        return 1
    synthetic = "synthetic_rows" in WORKLOADS[a.workload] and "snap" in WORKLOADS[a.workload]
    if synthetic:
        if n_inproc > 1:
            emit({"error": "--inproc is not wired for the directly synthesised workloads"})
            return 1
        rows_dev, n_syn, seqs, off, path = synth_workload(a.workload, rank, dev, rank == 0, a.verbose)
        if a.reads:
            seqs, off = seqs[: int(off[a.reads])], off[: a.reads + 1]
    else:
        # rank 0 generates into the box-local cache, the others load it
        if rank == 0:
            path, text, seq_starts, meta = build_workload(a.workload, dev, a.verbose)
        barrier()
        if rank != 0:
            path, text, seq_starts, meta = build_workload(a.workload, dev, a.verbose)
    if synthetic:
        pass
    elif n_inproc > 1:   # one process, N replicas: N ranks' worth of reads in one call
        parts = [make_reads(a.workload, text, seq_starts, r, a.reads, dev) for r in range(n_inproc)]
        seqs = np.concatenate([p[0] for p in parts])
        off = np.concatenate([[0]] + [p[1][1:].astype(np.uint64) + np.uint64(sum(int(q[1][-1]) for q in parts[:i])) for i, p in enumerate(parts)]).astype(np.uint64)
        del parts
    else:
        seqs, off = make_reads(a.workload, text, seq_starts, rank, a.reads, dev)
    torch.cuda.empty_cache()
    n_reads, n_bases = off.size - 1, int(off[-1])
    max_len = int(np.diff(off).max())
    width = pml_width_for(cb, max_len)
    pml_dtype = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]

    if synthetic:
        tbl = cb.ColPml.from_device_rows(rows_dev.data_ptr(), int(rows_dev.shape[0]), int(rows_dev.shape[0]), int(n_syn), devices=[local])
        del rows_dev
        torch.cuda.empty_cache()
    else:
        # COLBWT_BENCH_INPROC_SAME_GPU=1: all replicas on GPU 0 (exercises the in-process multi-replica path on a 1-GPU box)
        same = os.environ.get("COLBWT_BENCH_INPROC_SAME_GPU") == "1"
        tbl = cb.ColPml.load(path, devices=[local] if n_inproc == 1 else ([0] * n_inproc if same else list(range(n_inproc))))
    st = tbl.stats
    sampler = ClockSampler(local)
    sampler.start()

    # ---- kernel-only: reads + table resident in HBM ------------------------------------------------------------
    if n_inproc == 1:
        batches = [tbl.batch(seqs, off, width)]
    else:   # one batch per replica (its share of the reads), traversed concurrently from one thread per device
        cuts = [int(n_reads * d / n_inproc) for d in range(n_inproc + 1)]
        batches = [tbl.batch(seqs[int(off[cuts[d]]): int(off[cuts[d + 1]])], off[cuts[d]: cuts[d + 1] + 1] - off[cuts[d]], width, device_slot=d) for d in range(n_inproc)]

    def run_batches(iters):
        if len(batches) == 1:
            return batches[0].run(iters)
        res = [0.0] * len(batches)

        def one(i):
            res[i] = batches[i].run(iters)
        th = [threading.Thread(target=one, args=(i,)) for i in range(len(batches))]
        [t.start() for t in th]
        [t.join() for t in th]
        return max(res)

    for _ in range(a.warmup):
        run_batches(1)
    barrier()
    sampler.active.set()
    t0 = time.perf_counter()
    ms_per_step = run_batches(a.steps)          # CUDA events around K back-to-back traversals on the launching stream
    barrier()
    wall_kernel = time.perf_counter() - t0
    sampler.active.clear()
    ms_per_step = max_over_ranks(ms_per_step)
    value = world * n_bases / (ms_per_step * 1e-3)

    # ---- parity check against the oracle (not timed; checker only) -------------------------------------------------
    if len(batches) == 1:
        pml_d, cid_d = batches[0].download()
    else:
        dl = [b.download() for b in batches]
        pml_d, cid_d = np.concatenate([x[0] for x in dl]), np.concatenate([x[1] for x in dl])
        del dl
    k = n_reads if a.check_reads <= 0 else min(n_reads, a.check_reads)
    if (n_inproc > 1 or synthetic) and k < n_reads:   # a sample that touches every replica's share / both read classes: a stride over the reads
        sel = np.unique(np.linspace(0, n_reads - 1, k).astype(np.int64))
    else:
        sel = np.arange(k)
    # the CPU checker reads the table from a file; for the directly synthesised workloads only rank 0 has written one (18 GB),
    # so only rank 0 checks there -- every rank traverses the same table with the same code
    have_checker = path is not None
    ref, kind = cpu_reference(path) if have_checker else (None, "none")
    lens_sel = (off[sel + 1] - off[sel]).astype(np.int64)
    pos_sel = np.repeat(off[sel].astype(np.int64), lens_sel) + (np.arange(int(lens_sel.sum())) - np.repeat(np.cumsum(lens_sel) - lens_sel, lens_sel))
    s_seqs = seqs[pos_sel]
    s_off = np.concatenate([[0], np.cumsum(lens_sel)]).astype(np.uint64)
    if have_checker:
        want = (ref.query_batch(s_seqs, s_off, threads=max(1, (os.cpu_count() or 1) // world)) if kind == "reference" else ref.query_batch(s_seqs, s_off))
    else:
        want = (pml_d[pos_sel].astype(np.uint32), cid_d[pos_sel].copy())   # no checker on this rank: the streamed call is compared with the device-resident pass
    parity = bool(np.array_equal(pml_d[pos_sel].astype(np.uint32), want[0]) and np.array_equal(cid_d[pos_sel], want[1]))
    n_inv = min(n_reads, 500_000)
    properties = pml_invariants(pml_d, off, n_inv)
    mismatch_frac = float((pml_d[: min(n_bases, 50_000_000)] == 0).mean())
    cid_frac = float((cid_d[: min(n_bases, 50_000_000)] > 0).mean())
    del pml_d, cid_d

    # ---- end to end through the C-ABI with host buffers ----------------------------------------------------------------
    h_seqs = cb.PinnedArray(n_bases, np.uint8)
    h_seqs.array[:] = seqs
    h_pml = cb.PinnedArray(n_bases, pml_dtype)
    h_cid = cb.PinnedArray(n_bases, np.uint8)

    def e2e_dense():
        tbl.query(h_seqs.array, off, width, out=(h_pml.array, h_cid.array))

    def timed(fn, active_ranks):
        """K calls of fn on the first `active_ranks` ranks (the others wait at the barriers); seconds per call, max over ranks."""
        barrier()
        t0 = time.perf_counter()
        if rank < active_ranks:
            for _ in range(a.steps):
                fn()
        dt = (time.perf_counter() - t0) / a.steps if rank < active_ranks else 0.0
        barrier()
        return max_over_ranks(dt)

    # warm-up: the staging allocation, then one large call per mode (host / device packing x dense / compact transport):
    # the library measures each once and keeps the fastest (query.cu, tasks.h: choose_mode)
    for _ in range(max(10, a.warmup)):   # first call allocates the staging; eight (packing x transport) modes are each tried once
        e2e_dense()
    sampler.active.set()
    e2e_s = timed(e2e_dense, world)
    sampler.active.clear()
    h2d, d2h = tbl.last_bytes               # what the library asked the copy engines to move in the last call
    e2e_value = world * n_bases / e2e_s
    e2e_parity = bool(np.array_equal(h_pml.array[pos_sel].astype(np.uint32), want[0]) and np.array_equal(h_cid.array[pos_sel], want[1]))
    device_pack = tbl.last_packing == "device"
    transport = tbl.last_transport
    # same call with fewer ranks active: what one rank gets with the host to itself, and the curve in between
    e2e_curve, kernel_curve = {}, {}
    if world > 1:
        ks = sorted({1} | ({kk for kk in (2, 4) if kk < world} if a.sweep_out else set()))
        for kk in ks:
            e2e_curve[kk] = kk * n_bases / timed(e2e_dense, kk)
            if a.sweep_out:   # the device-resident traversal with only the first kk ranks active (ranks share nothing there)
                barrier()
                ms_k = max_over_ranks(run_batches(a.steps) if rank < kk else 0.0)
                kernel_curve[kk] = kk * n_bases / (ms_k * 1e-3)

    # ---- the compact result (one match bit per base + sparse chain ids): the same call for consumers that need no dense arrays
    c_cap = int(cb._L.colbwt_compact_bound(off.ctypes.data, n_reads))
    c_cap = min(c_cap, max(1 << 20, int(0.45 * n_bases) + (64 << 20)))   # typical results are ~0.3 B/base; the bound is 1.27
    h_comp = cb.PinnedArray(c_cap, np.uint8)
    comp = None
    try:
        used = 0

        def e2e_compact():
            nonlocal used
            used = tbl.query_compact(h_seqs.array, off, out=h_comp.array).size
        for _ in range(4):
            e2e_compact()
        c_s = timed(e2e_compact, world)
        c_pack = tbl.last_packing
        c_h2d, c_d2h = tbl.last_bytes
        cp, cc = cb.compact_expand(h_comp.array[:used], off, width)
        c_parity = bool(np.array_equal(cp[pos_sel].astype(np.uint32), want[0]) and np.array_equal(cc[pos_sel], want[1]))
        del cp, cc
        comp = {"value": world * n_bases / c_s, "unit": "bases/s", "s_per_step": c_s, "d2h_bytes_per_step": int(c_d2h), "result_bytes": int(used),
                "h2d_bytes_per_step": int(c_h2d),
                "packing": c_pack, "parity_after_host_expand": c_parity,
                "api": "colbwt_query_compact (pinned host buffers): match bit per base + non-zero chain ids; colbwt_compact_expand rebuilds the dense arrays"}
        if world > 1:
            c_curve = {kk: kk * n_bases / timed(e2e_compact, kk) for kk in sorted(e2e_curve)}
            comp["alone_value"] = c_curve[1]
            comp["efficiency_vs_alone"] = round(comp["value"] / (world * comp["alone_value"]), 4)
            comp["active_ranks_curve"] = {str(kk): v for kk, v in c_curve.items()}
    except cb.ColBwtError as e:
        comp = {"error": str(e)}
    sampler.stop_flag.set()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline ----------------------------------------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    per_gpu_bases_s = n_bases / n_inproc / (ms_per_step * 1e-3)
    achieved = per_gpu_bases_s * ALGO_BYTES_PER_BASE / 1e9
    table_bytes = int(st.r) * 16
    s_rand = cb.gather_bench(max(table_bytes, 1 << 20), 1 << 28, False, local)
    s_dep = cb.gather_bench(max(table_bytes, 1 << 20), 1 << 26, True, local)
    traffic = profiled_traffic(a.workload)
    roofline = {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic["from"] if traffic else None,
                "peak_source": peak_src, "kernel": f"k_traverse<packed,u{8 * width}>",
                "algorithmic_bytes_per_base": ALGO_BYTES_PER_BASE,
                "algorithmic_bytes_per_base_as_written": 32 + 0.25 + width + 1,
                "frac_as_written": round(per_gpu_bases_s * (32 + 0.25 + width + 1) / 1e9 / peak, 4),
                "gather": {"table_bytes": table_bytes, "random_sector_rate_per_s": s_rand, "dependent_sector_rate_per_s": s_dep,
                           "frac_of_random_sector_rate": round(per_gpu_bases_s / s_rand, 4)}}
    if traffic:   # what the DRAM actually moved (ncu capture of this workload) at this run's launch time
        dram_gbs = traffic["bytes"] / (ms_per_step * 1e-3) / 1e9
        roofline["dram"] = {"gbs": round(dram_gbs, 1), "frac_of_peak": round(dram_gbs / peak, 4),
                            "note": "every L2 miss of a 16-byte row gather fills a 128-byte line (DESIGN.md section 4)"}

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------------------------
    cpu = None
    if world == 1 and n_inproc == 1 and a.cpu_seconds > 0:
        cores = os.cpu_count() or 1
        threads = cores if kind == "reference" else 1
        kk, run = time_cpu(ref, kind, seqs, off, threads, a.cpu_seconds)
        b, t = run(kk)
        cpu = {"value": b / t, "unit": "bases/s", "cores": threads, "kind": kind,
               "sample": f"first {kk} reads ({b} bases) of the same batch, {t:.2f} s; MULTI_THREAD off (output-identical, SURVEY.md 6)"}
        if threads > 1:   # the same harness on one core, a ~2 s sample
            k1, run1 = time_cpu(ref, kind, seqs, off, 1, min(2.0, a.cpu_seconds))
            b1, t1 = run1(k1)
            cpu["one_core"] = {"value": b1 / t1, "unit": "bases/s", "sample": f"first {k1} reads ({b1} bases), {t1:.2f} s"}
        if kind == "reference":
            try:
                cpu.update(cli_baselines(path, seqs, off, a.cpu_seconds))
            except Exception as e:   # the executables are optional extras of the baseline, never of the GPU numbers
                cpu["cli_error"] = repr(e)

    e2e = {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "s_per_step": e2e_s,
           "bytes_are": "per GPU, counted by the library from the copies it enqueues (colbwt_index_last_bytes)",
           "pml_bytes": width, "packing": "device" if device_pack else "host", "transport": transport,
           "api": "colbwt_query (host pinned buffers in, dense PML + chain-id arrays out; read packing inside the timed region)"}
    if world > 1:
        e2e["alone_value"] = e2e_curve[1]
        e2e["efficiency_vs_alone"] = round(e2e_value / (world * e2e_curve[1]), 4)
        e2e["active_ranks_curve"] = {str(kk): v for kk, v in sorted(e2e_curve.items())}
        e2e["note"] = ("alone_value = rank 0 running the same call while the other ranks idle (same per-rank host threads); "
                       "efficiency_vs_alone = value / (n_gpus x alone_value): what sharing the host costs")
    out = {
        "metric": "query bases/sec (PML + chain stats)", "value": value, "unit": "bases/s", "n_gpus": world * n_inproc, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/u16/u8 integer", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[a.workload], "name": a.workload, "reads_per_gpu": n_reads // n_inproc, "bases_per_gpu": n_bases // n_inproc,
                   "index": {"n": int(st.n), "rows": int(st.r), "bwt_runs": int(st.bwt_r), "marked_rows": int(st.marked_rows),
                             "exact_search_rows": int(st.slow_rows), "hbm_bytes": int(st.device_bytes)},
                   "parallelism": (f"index replicated x{world}, reads sharded, no collective" if n_inproc == 1 else
                                   f"ONE process, index replicated on {n_inproc} GPUs (colbwt_index_load n_devices={n_inproc}), one feeder thread per GPU, no collective"),
                   "host": {"cpus": os.cpu_count(), "packing_threads_per_process": max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))},
                   "l2": "no flush needed: table + per-step outputs exceed the 126 MB L2" if table_bytes + n_bases * (width + 1) > (200 << 20) else "working set fits L2 (small workload)",
                   "mismatch_step_frac": round(mismatch_frac, 4), "cid_nonzero_frac": round(cid_frac, 4)},
        "e2e": e2e, "e2e_compact": comp,
        "gpu_launches": sum(b.launches for b in batches) * a.steps,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary(),
        "parity_vs_oracle": {"kernel": parity, "e2e": e2e_parity, "reads_checked": int(sel.size), "checker": kind,
                             "pml_invariants_on_first_reads": {"reads": n_inv, "ok": properties}},
        "wall_s_kernel_region": wall_kernel,
    }
    if a.sweep_out and world > 1:
        # the same job with only the first k ranks active (the others idle at the barriers): one line per k, same keys as the
        # full line where they apply; per-rank host threads stay those of the N-rank launch (stated in config.host)
        with open(a.sweep_out, "a") as f:
            for kk, v in sorted(e2e_curve.items()):
                line = {"metric": out["metric"], "value": kernel_curve.get(kk), "unit": "bases/s", "n_gpus": kk, "launched_ranks": world, "steps": a.steps,
                        "scaling": "weak", "config": out["config"], "e2e": {"value": v, "unit": "bases/s"},
                        "e2e_compact": {"value": comp["active_ranks_curve"][str(kk)], "unit": "bases/s"} if comp and "active_ranks_curve" in comp else None,
                        "parity_vs_oracle": out["parity_vs_oracle"]}
                f.write(json.dumps(line) + "\n")
            f.write(json.dumps(out) + "\n")
    emit(out)
    if world > 1:
        dist.destroy_process_group()
    ok = parity and e2e_parity and properties and (comp is None or comp.get("parity_after_host_expand", True))
    return 0 if ok else 2


def run_reference(a, have_gpu):
    """--impl reference: the reference's own CPU implementation on this box's host cores, same workload/metric."""
    dev = "cuda:0" if have_gpu else "cpu"
    cap = min(a.reads or WORKLOADS[a.workload]["reads"], 400_000 if WORKLOADS[a.workload]["read_len"] < 1000 else 6000)
    if "snap" in WORKLOADS[a.workload]:   # directly synthesised table: generated on the GPU, written out for the reference's loader
        import torch
        rows_dev, _n, seqs, off, path = synth_workload(a.workload, 0, dev, True, a.verbose)
        del rows_dev
        torch.cuda.empty_cache()
        seqs, off = seqs[: int(off[cap])], off[: cap + 1]
    else:
        path, text, seq_starts, meta = build_workload(a.workload, dev, a.verbose)
        # a bounded sample of the same reads (rank 0's batch); enough for the calibration to size the steps
        seqs, off = make_reads(a.workload, text, seq_starts, 0, cap, dev)
    ref, kind = cpu_reference(path)
    cores = os.cpu_count() or 1
    threads = cores if kind == "reference" else 1
    kk, run = time_cpu(ref, kind, seqs, off, threads, max(2.0, a.cpu_seconds))
    for _ in range(a.warmup):
        run(max(1, kk // 4))
    times, bases = [], 0
    for _ in range(a.steps):
        b, t = run(kk)
        bases = b
        times.append(t)
    s_per_step = float(np.mean(times))
    v = bases / s_per_step
    out = {
        "impl": "reference", "metric": "query bases/sec (PML + chain stats)", "value": v, "unit": "bases/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 integer (CPU)", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[a.workload], "name": a.workload, "reads_per_step": kk, "bases_per_step": bases},
        "cpu_baseline": {"value": v, "unit": "bases/s", "cores": threads, "kind": kind,
                         "sample": f"{kk} reads ({bases} bases) of the workload per step; col_pml::query_pml on {threads} host threads sharing one table; MULTI_THREAD off (output-identical)"},
        "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""bench.py — throughput of the STAR alignment hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

A "step" = one pass of the whole hot path (prep -> MMP seed search -> windows -> stitch/extend -> select -> pack) over
one chunk of synthetic 2x100 bp read pairs.

  value  : read pairs / s with the chunk already resident in HBM (star_gpu_upload_chunk once, star_gpu_map_resident per step)
  e2e    : the same metric through the reference-facing C-ABI call star_gpu_map_chunk with PINNED HOST buffers
           (host->device copy of the sequences and device->host copy of all results inside the timed region)
  roofline: MMP seed-search kernel, algorithmic bytes (SURVEY.md §8d formula, counts from the instrumented ORACLE on a
           sample, cross-checked against the kernel's own counters) / CUDA-event duration of that kernel
  cpu_baseline: the UNMODIFIED reference (oracle/_ref/STAR) on all host cores on a bounded sample of the same workload

Workload: BASELINE.json configs[1] (GRCh38 + GENCODE) cannot be built here: no real genome exists in the container or on the
GPU box and there is no network.  The bench therefore uses the survey's self-contained tier: seeded synthetic
chr21-sized genome (46.7 Mb, 3 chromosomes, repeat families, 2000-gene annotation -> sjdb), index built at bench time by the
reference's own genomeGenerate.  `config.workload` names it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU; reads are sharded, weak scaling)
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF_STAR = os.path.join(ROOT, "oracle", "_ref", "STAR")
METRIC = "reads/sec (2x100 bp PE)"
UNIT = "read pairs/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def prepare_genome(workdir, preset):
    """genome.fa + annot.gtf + idx/ (reference genomeGenerate).  Returns (chrs, trs, idx_dir)."""
    import synth
    chrs = synth.make_genome(preset)
    trs = synth.make_annotation(chrs, preset)
    idx = os.path.join(workdir, "idx")
    if not os.path.exists(os.path.join(idx, "SAindex")):
        os.makedirs(idx, exist_ok=True)
        synth.write_fasta(chrs, os.path.join(workdir, "genome.fa"))
        synth.write_gtf(chrs, trs, os.path.join(workdir, "annot.gtf"))
        nb = {"tiny": 7, "small": 9, "chr21": 11}[preset]
        t0 = time.time()
        subprocess.check_call([REF_STAR, "--runMode", "genomeGenerate", "--genomeDir", "idx", "--genomeFastaFiles", "genome.fa",
                               "--sjdbGTFfile", "annot.gtf", "--sjdbOverhang", "99", "--genomeSAindexNbases", str(nb),
                               "--runThreadN", str(min(64, os.cpu_count() or 8)), "--outFileNamePrefix", "gen_"],
                              cwd=workdir, stdout=subprocess.DEVNULL)
        log("index built in %.1f s" % (time.time() - t0))
    return chrs, trs, idx


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def reference_run(workdir, idx, fq1, fq2, threads, tag, repeat=1):
    """Wall time of the unmodified reference on `repeat` concatenated copies of (fq1, fq2) and of an index-load-only run."""
    if repeat > 1:
        fq1 = ",".join([fq1] * repeat)
        fq2 = ",".join([fq2] * repeat)
    out = os.path.join(workdir, tag)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    t0 = time.time()
    subprocess.check_call([REF_STAR, "--genomeDir", idx, "--readFilesIn", fq1, fq2, "--outFileNamePrefix", out + "/", "--runThreadN", str(threads),
                           "--outSAMtype", "SAM"], stdout=subprocess.DEVNULL)
    dt = time.time() - t0
    # index-load-only run (wall clock minus a --readMapNumber 1 run)
    out0 = os.path.join(workdir, tag + "_load")
    shutil.rmtree(out0, ignore_errors=True)
    os.makedirs(out0)
    t0 = time.time()
    subprocess.check_call([REF_STAR, "--genomeDir", idx, "--readFilesIn", fq1, fq2, "--outFileNamePrefix", out0 + "/", "--runThreadN", str(threads),
                           "--readMapNumber", "1"], stdout=subprocess.DEVNULL)
    dt0 = time.time() - t0
    shutil.rmtree(out, ignore_errors=True)
    shutil.rmtree(out0, ignore_errors=True)
    return dt, dt0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("STAR_B200_BENCH_PAIRS", 1 << 20)), help="read pairs per GPU per step")
    ap.add_argument("--ref-pairs", type=int, default=int(os.environ.get("STAR_B200_BENCH_REF_PAIRS", 1_000_000)))
    ap.add_argument("--ref-repeat", type=int, default=int(os.environ.get("STAR_B200_BENCH_REF_REPEAT", 2)),
                    help="the reference arm maps this many concatenated copies of the sample (measured on the 128-core box: 124-129 k pairs/s for 2 M, 4 M and 16 M pairs alike - the reference is bound by its serial FASTQ chunker, so the bounded sample is representative)")
    ap.add_argument("--preset", default=os.environ.get("STAR_B200_BENCH_PRESET", "chr21"))
    ap.add_argument("--mm", type=float, default=0.005)
    ap.add_argument("--workdir", default=os.environ.get("STAR_B200_BENCH_DIR", "/tmp/star_b200_bench"))
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import synth
    workdir = os.path.join(a.workdir, a.preset)
    os.makedirs(workdir, exist_ok=True)
    workload = "synthetic %s-sized genome (46.7 Mb, 3 chr, sjdbOverhang 99, SAindexNbases 11), 2x100 bp PE, %.1f%% subst" % (a.preset, a.mm * 100)
    host_cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        chrs, trs, idx = prepare_genome(workdir, a.preset)
        m1, m2 = synth.make_reads(chrs, trs, a.ref_pairs, read_len=100, mm=a.mm, seed=4242)
        fq1, fq2 = os.path.join(workdir, "ref_1.fq"), os.path.join(workdir, "ref_2.fq")
        synth.write_fastq(m1, fq1)
        synth.write_fastq(m2, fq2)
        times = []
        rep = max(1, a.ref_repeat)
        for s in range(a.warmup + a.steps):
            dt, dt0 = reference_run(workdir, idx, fq1, fq2, host_cores, "refrun", repeat=rep)
            if s >= a.warmup:
                times.append(max(1e-3, dt - dt0))
            log("reference step %d: %.2f s total, %.2f s load-only" % (s, dt, dt0))
        t = float(np.mean(times))
        v = a.ref_pairs * rep / t
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic",
                "config": {"workload": workload, "pairs_per_step": a.ref_pairs * rep, "threads": host_cores,
                           "timing": "wall clock of the full STAR run minus a --readMapNumber 1 (index load) run"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": host_cores, "kind": "reference",
                                 "sample": "%d x %d pairs of the same workload, oracle/_ref/STAR --runThreadN %d" % (rep, a.ref_pairs, host_cores)},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if rank == 0:
        chrs, trs, idx = prepare_genome(workdir, a.preset)
    if world > 1:
        dist.barrier()
    if rank != 0:
        chrs, trs, idx = prepare_genome(workdir, a.preset)

    import star_b200 as sb
    lib = sb.load_library()
    index = sb.Index(lib, idx)
    n = a.pairs
    m1, m2 = synth.make_reads(chrs, trs, n, read_len=100, mm=a.mm, seed=1000 + rank)
    seq, off, _, nm = sb.pack_reads([m1, m2])
    eng = sb.Engine(lib, index, max_reads=n, device=local_rank)
    n_out = eng.n_out
    # pinned host buffers for the e2e leg
    pin_seq = torch.empty(seq.nbytes, dtype=torch.uint8, pin_memory=True)
    pin_seq.numpy()[:] = seq
    pin_off = torch.empty(off.nbytes, dtype=torch.uint8, pin_memory=True)
    pin_off.numpy().view(np.uint64)[:] = off
    pin_res = torch.empty(n * sb.capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    pin_al = torch.empty(n * 2 * sb.capi.ALIGN_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)   # 2 alignments / read of head-room
    res_np = pin_res.numpy().view(sb.capi.RESULT_DTYPE)
    al_np = pin_al.numpy().view(sb.capi.ALIGN_DTYPE)
    ab = sb.capi.AlignBatch()
    ab.reads = res_np.ctypes.data
    ab.aligns = al_np.ctypes.data
    ab.alignsCapacity = al_np.shape[0]
    seq_p = pin_seq.numpy()
    off_p = pin_off.numpy().view(np.uint64)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: chunk resident in HBM
    eng.upload(seq_p, off_p, n, nm)
    for _ in range(a.warmup):
        eng.map_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.star_gpu_launch_count()
    sync_all()
    t0 = time.perf_counter()
    dev_ms, seed_ms, stitch_ms = 0.0, 0.0, 0.0
    last = None
    for _ in range(a.steps):
        st = eng.map_resident()
        dev_ms += st.ms_total
        seed_ms += st.ms_seed
        stitch_ms += st.ms_stitch
        last = st
    sync_all()
    wall = time.perf_counter() - t0
    launches = lib.star_gpu_launch_count() - launches0
    # ---- e2e: host buffers through star_gpu_map_chunk
    for _ in range(max(1, a.warmup // 2)):
        eng.map_chunk(seq_p, off_p, n, nm, out=(res_np, al_np, ab))
    sync_all()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(a.steps):
        _, al_out, st2 = eng.map_chunk(seq_p, off_p, n, nm, out=(res_np, al_np, ab))
        h2d, d2h = st2.h2d_bytes, st2.d2h_bytes
    sync_all()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.finish()

    # max over ranks (device timing) and the single collective of the path: the Log.final.out counters
    t_dev = dev_ms / 1e3
    t_val = max(t_dev, 0.0)
    if world > 1:
        tt = torch.tensor([t_val, wall, wall_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_val, wall, wall_e2e = [float(x) for x in tt.tolist()]
        counters = torch.tensor([n * a.steps, int((res_np["unmapType"] < 0).sum())], dtype=torch.int64, device="cuda")
        dist.all_reduce(counters)
    total_pairs = n * world * a.steps
    value = total_pairs / t_val
    e2e_value = total_pairs / wall_e2e

    if rank == 0:
        # ---- roofline of the MMP seed-search kernel
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(pk):
            peaks = json.load(open(pk))
            peak_gbs, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        gstrand = int(index.view.contents.GstrandBit)
        w_sai, w_sa = (gstrand + 3) / 8.0, (gstrand + 1) / 8.0
        # counts from the instrumented ORACLE on a sample (tests assert the kernel's own counters are identical)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_capi as oc
        ns = min(n, 20000)
        oe = oc.OracleEngine(oc.load_oracle(), index)
        t0 = time.perf_counter()
        _, _, st_o = oe.map_chunk(seq[: int(off[ns * 2])].copy(), off[: ns * 2 + 1].copy(), ns, nm)
        t_oracle = time.perf_counter() - t0
        oe.close()
        b_pair_oracle = (st_o.mmp_sai_words * w_sai + st_o.mmp_compare_calls * w_sa + st_o.mmp_bases_examined * 1.0) / ns
        b_pair_kernel = (last.mmp_sai_words * w_sai + last.mmp_compare_calls * w_sa + last.mmp_bases_examined * 1.0) / n
        seed_s = seed_ms / a.steps / 1e3
        achieved = b_pair_oracle * n / seed_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "seed_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if int(tj.get("pairs_per_launch", -1)) == n:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        roofline = {"bound": "hbm", "kernel": "seed_search_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_pair": b_pair_oracle,
                    "algorithmic_bytes_per_pair_kernel_counters": b_pair_kernel, "kernel_ms": seed_s * 1e3,
                    "stitch_kernel_ms": stitch_ms / a.steps}
        # ---- cpu baseline: the unmodified reference on all host cores, bounded sample
        cpu = None
        if world > 1:
            cpu = None   # the CPU baseline is timed at N=1 only (rank 0)
        elif os.path.exists(REF_STAR):
            rp = min(a.ref_pairs, n)
            rep = max(1, a.ref_repeat)
            fq1, fq2 = os.path.join(workdir, "cpu_1.fq"), os.path.join(workdir, "cpu_2.fq")
            synth.write_fastq(m1[:rp], fq1)
            synth.write_fastq(m2[:rp], fq2)
            dt, dt0 = reference_run(workdir, idx, fq1, fq2, host_cores, "cpubase", repeat=rep)
            cpu = {"value": rp * rep / max(1e-3, dt - dt0), "unit": UNIT, "cores": host_cores, "kind": "reference",
                   "sample": "%d x the first %d pairs of the step's chunk, oracle/_ref/STAR --runThreadN %d, wall %.2f s minus %.2f s index load" % (rep, rp, host_cores, dt, dt0),
                   "oracle_port_1thread_pairs_per_s": ns / t_oracle}
        else:
            cpu = {"value": ns / t_oracle, "unit": UNIT, "cores": 1, "kind": "port", "sample": "%d pairs, oracle/star_oracle.cpp, 1 thread" % ns}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": t_val / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64",
                "data": "synthetic",
                "config": {"workload": workload, "pairs_per_gpu_per_step": n, "read_definition": "one 2x100 pair = one read (STAR 'Number of input reads')",
                           "l2_policy": "inputs larger than L2 (SA 392 MB + reads %d MB per step, L2 126 MB)" % (seq.nbytes >> 20),
                           "parallelism": "reads sharded across %d GPU(s), index replicated, one NCCL allreduce of the counters" % world,
                           "mapped_fraction": float((res_np["unmapType"] < 0).mean()), "overflow_tier_reads_per_step": int(last.slow_path_reads),
                           "flat_path_reads_per_step": int(last.heavy_reads), "flat_path_ms": float(last.ms_heavy),
                           "wall_s_value_leg": wall},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_e2e / a.steps * 1e3},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    eng.close()
    index.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

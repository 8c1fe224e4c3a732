"""N>1 path on CPU: world_size 2 over gloo.  Each rank maps a contiguous slice of the reads (with the oracle-driven test CLI,
there is no GPU here), the counters are allreduced, rank 0 merges; the merged outputs must equal the single-process reference
goldens byte for byte (SAM order = input order, global junction collapse + filters, summed Log.final.out counters)."""
import os
import subprocess
import sys

import conftest as cf
import oracle_capi as oc

ROOT = cf.ROOT


def test_two_rank_sharded_run_equals_reference(oracle, lib, golden, tmp_path):
    out = str(tmp_path) + "/"
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29611",
           "-m", "star_b200.dist", "--cli", oc.ORACLE_CLI, "--",
           "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq"),
           "--outFileNamePrefix", out, "--runThreadN", "2", "--quantMode", "GeneCounts", "--outReadsUnmapped", "Fastx"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    whole = str(tmp_path) + "/whole_"   # gene counts and unmapped reads of the merged run = those of a single-process run
    subprocess.check_call([oc.ORACLE_CLI] + cmd[cmd.index("--genomeDir"):cmd.index("--outFileNamePrefix")] + ["--outFileNamePrefix", whole, "--quantMode", "GeneCounts",
                          "--outReadsUnmapped", "Fastx"], stdout=subprocess.DEVNULL)
    for f in ("ReadsPerGene.out.tab", "Unmapped.out.mate1", "Unmapped.out.mate2"):
        assert open(out + f, "rb").read() == open(whole + f, "rb").read(), f
    ref = os.path.join(golden, "ref_std")
    assert cf.sam_body(out + "Aligned.out.sam") == cf.sam_body(os.path.join(ref, "Aligned.out.sam"))
    assert open(out + "SJ.out.tab", "rb").read() == open(os.path.join(ref, "SJ.out.tab"), "rb").read()
    assert cf.log_counters(out + "Log.final.out") == cf.log_counters(os.path.join(ref, "Log.final.out"))
    # both shards really mapped a slice
    n0 = len(cf.sam_body(out + "shard0.Aligned.out.sam"))
    n1 = len(cf.sam_body(out + "shard1.Aligned.out.sam"))
    assert n0 > 0 and n1 > 0 and n0 + n1 == len(cf.sam_body(out + "Aligned.out.sam"))


def test_world_size_one_is_the_plain_run(oracle, lib, golden, tmp_path):
    """`torchrun --nproc-per-node 1 -m star_b200.dist` (the N=1 point of a scaling run): no shard files, no merge — the final files are
    the single process' own (found by the 2-GPU measurement of round 2: the N=1 launch looked for shard0.shard.bin)."""
    out = str(tmp_path) + "/"
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1", "--master-port", "29617",
           "-m", "star_b200.dist", "--cli", oc.ORACLE_CLI, "--",
           "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq"),
           "--outFileNamePrefix", out, "--runThreadN", "2"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = os.path.join(golden, "ref_std")
    assert cf.sam_body(out + "Aligned.out.sam") == cf.sam_body(os.path.join(ref, "Aligned.out.sam"))
    assert open(out + "SJ.out.tab", "rb").read() == open(os.path.join(ref, "SJ.out.tab"), "rb").read()
    assert cf.log_counters(out + "Log.final.out") == cf.log_counters(os.path.join(ref, "Log.final.out"))
    assert os.path.exists(out + "dist_timing.json")


def test_shard_arguments():
    from star_b200 import dist
    a = dist.shard_args(["--genomeDir", "g", "--outFileNamePrefix", "o/x_"], 3, 8, device=3)
    assert a[a.index("--outFileNamePrefix") + 1] == "o/x_shard3."
    assert a[a.index("--gpuShardIndex") + 1] == "3" and a[a.index("--gpuShardCount") + 1] == "8" and a[a.index("--gpuDevice") + 1] == "3"


def test_two_rank_twopass_equals_reference(oracle, lib, golden, twopass_golden, tmp_path):
    """--twopassMode Basic over 2 ranks: the 1st-pass junction records of both shards are all-gathered (gloo), every rank inserts the same
    global list into its replica of the index and maps its slice again; merged outputs = the single-process reference run."""
    out = str(tmp_path) + "/"
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29613",
           "-m", "star_b200.dist", "--cli", oc.ORACLE_CLI, "--",
           "--genomeDir", os.path.join(twopass_golden, "idx0"), "--readFilesIn", os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq"),
           "--outFileNamePrefix", out, "--runThreadN", "2", "--twopassMode", "Basic", "--sjdbInsertSave", "All"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = os.path.join(twopass_golden, "A_novel")
    assert cf.sam_body(out + "Aligned.out.sam") == cf.sam_body(os.path.join(ref, "Aligned.out.sam"))
    assert open(out + "SJ.out.tab", "rb").read() == open(os.path.join(ref, "SJ.out.tab"), "rb").read()
    assert cf.log_counters(out + "Log.final.out") == cf.log_counters(os.path.join(ref, "Log.final.out"))
    assert open(out + "_STARpass1/SJ.out.tab", "rb").read() == open(os.path.join(ref, "_STARpass1/SJ.out.tab"), "rb").read()
    assert cf.log_counters(out + "_STARpass1/Log.final.out") == cf.log_counters(os.path.join(ref, "_STARpass1/Log.final.out"))
    import hashlib
    for shard in ("shard0.", "shard1."):   # both replicas of the index were rebuilt to the reference's bytes
        for line in open(os.path.join(ref, "_STARgenome/sha256.txt")):
            name, digest = line.split()
            assert hashlib.sha256(open(out + shard + "_STARgenome/" + name, "rb").read()).hexdigest() == digest, (shard, name)


def test_two_rank_bysjout_twopass_equals_reference(oracle, lib, golden, twopass_golden, tmp_path):
    """--outFilterType BySJout with --twopassMode Basic over 2 ranks: two exchanges (1st-pass junctions, then the junction records of
    all reads between the BySJout stages, both all-gathered over gloo); merged outputs = the single-process reference run
    (held reads of every shard after the 1st-stage records of all shards)."""
    out = str(tmp_path) + "/"
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29615",
           "-m", "star_b200.dist", "--cli", oc.ORACLE_CLI, "--",
           "--genomeDir", os.path.join(twopass_golden, "idx0"), "--readFilesIn", os.path.join(golden, "hard_1.fq"), os.path.join(golden, "hard_2.fq"),
           "--outFileNamePrefix", out, "--runThreadN", "2", "--outFilterType", "BySJout", "--twopassMode", "Basic",
           "--outSAMattributes", "NH", "HI", "AS", "nM", "XS", "--sjdbInsertSave", "All"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = os.path.join(twopass_golden, "S4_bysjout_twopass")
    assert cf.sam_body(out + "Aligned.out.sam") == cf.sam_body(os.path.join(ref, "Aligned.out.sam"))
    assert open(out + "SJ.out.tab", "rb").read() == open(os.path.join(ref, "SJ.out.tab"), "rb").read()
    assert cf.log_counters(out + "Log.final.out") == cf.log_counters(os.path.join(ref, "Log.final.out"))
    assert len(cf.sam_body(out + "shard0.Aligned.out.stage2.sam")) + len(cf.sam_body(out + "shard1.Aligned.out.stage2.sam")) > 0

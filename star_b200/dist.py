"""Multi-GPU driver: one process per GPU (torchrun), reads sharded by contiguous slices, index replicated per GPU.

There is no collective on the data path.  Collectives, all after mapping:
  * allreduce(sum) of the 24 Log.final.out counters (NCCL on GPUs; gloo in the CPU test-suite),
  * a barrier; the junction records and SAM shards are files on the node's filesystem and are merged by rank 0 through the
    C-ABI helper star_host_merge_shards (global collapse + the neighbour-distance filter need the complete sorted list).

  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m star_b200.dist -- --genomeDir idx --readFilesIn r_1.fq r_2.fq --outFileNamePrefix out/
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

N_COUNTERS = 24


def _prefix(argv):
    for i, a in enumerate(argv):
        if a == "--outFileNamePrefix" and i + 1 < len(argv):
            return argv[i + 1]
    return "./"


def shard_args(argv, rank, world, device=None):
    """Command line of shard `rank`: same arguments, shard slice, shard output prefix."""
    pre = _prefix(argv)
    out = [a for a in argv]
    if "--outFileNamePrefix" in out:
        i = out.index("--outFileNamePrefix")
        out[i + 1] = pre + "shard%d." % rank
    else:
        out += ["--outFileNamePrefix", pre + "shard%d." % rank]
    out += ["--gpuShardIndex", str(rank), "--gpuShardCount", str(world)]
    if device is not None:
        out += ["--gpuDevice", str(device)]
    return out


def read_shard_counters(prefix, rank):
    with open(prefix + "shard%d.shard.bin" % rank, "rb") as f:
        return np.frombuffer(f.read(8 * N_COUNTERS), dtype=np.uint64).copy()


def run_sharded(argv, cli=None, backend=None):
    """Runs under torchrun (RANK/WORLD_SIZE/LOCAL_RANK set).  cli: None = the CUDA engine in-process (star_cli_main);
    or the path of an executable with the same command line (the test-suite passes the oracle-driven CLI)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = cli is None
    if backend is None:
        backend = "nccl" if use_cuda else "gloo"
    if use_cuda:
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    import star_b200 as sb
    lib = sb.load_library()
    lib.star_host_merge_shards.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_void_p]
    prog = "STAR"
    sargv = shard_args(argv, rank, world, device=local_rank if use_cuda else None)
    os.makedirs(os.path.dirname(_prefix(argv)) or ".", exist_ok=True)
    if cli is None:
        arr = (C.c_char_p * (len(sargv) + 1))(*([prog.encode()] + [a.encode() for a in sargv]))
        rc = lib.star_cli_main(len(sargv) + 1, arr)
    else:
        rc = subprocess.call([cli] + sargv, stdout=subprocess.DEVNULL)
    ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int64, device="cuda" if use_cuda else "cpu")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        dist.destroy_process_group()
        return rc or 1
    # the one collective of the path: the Log.final.out counters
    cnt = read_shard_counters(_prefix(argv), rank).astype(np.int64)
    t = torch.from_numpy(cnt)
    if use_cuda:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.barrier()
    rc = 0
    if rank == 0:
        total = t.cpu().numpy().astype(np.uint64)
        margv = [prog] + list(argv)
        arr = (C.c_char_p * len(margv))(*[a.encode() for a in margv])
        rc = lib.star_host_merge_shards(len(margv), arr, world, total.ctypes.data)
    dist.barrier()
    dist.destroy_process_group()
    return rc


def main():
    args = sys.argv[1:]
    cli = None
    if args and args[0] == "--cli":
        cli = args[1]
        args = args[2:]
    if args and args[0] == "--":
        args = args[1:]
    sys.exit(run_sharded(args, cli=cli))


if __name__ == "__main__":
    main()

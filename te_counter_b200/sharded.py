"""
One BAM counted by all GPUs of a node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        -m te_counter_b200.sharded --glb genes_tes.glb --bam in.bam --mode se -o out.tsv
    ... --mode sc --whitelist 3M-february-2018.txt --maxcells 10000 [--strand] -o out.tsv

Every rank builds the same `measureTE` the reference's bin/te_count builds (bin/te_count:77-120) and calls the same
method; inside, the ranks split the file by byte ranges (shard.py).  Rank 0 writes the outputs, which are byte for byte
those of a one-GPU run.  With fewer GPUs than ranks the ranks share GPUs and talk over gloo (tests).
"""
import argparse
import logging
import os
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m te_counter_b200.sharded")
    ap.add_argument("--glb", required=True)
    ap.add_argument("--bam", required=True)
    ap.add_argument("--mode", choices=["se", "pe", "sc"], default="se")
    ap.add_argument("-q", "--qual", type=int, default=20)
    ap.add_argument("--whitelist")
    ap.add_argument("--maxcells", type=int, default=10000)
    ap.add_argument("--strand", action="store_true")
    ap.add_argument("-o", "--out", required=True)
    args = ap.parse_args(argv)
    import torch
    import torch.distributed as dist
    from . import measureTE
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpu = torch.cuda.device_count()
    device = local % max(1, n_gpu)
    torch.cuda.set_device(device)
    if world > 1:
        if n_gpu >= world:
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        else:
            dist.init_process_group("gloo")
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING, format="%(levelname)-8s: %(message)s")
    log = logging.getLogger("te_count")
    mte = measureTE(sys.path[0], args.qual, device=device)
    mte.bind_genome(args.glb)
    if args.mode == "sc":
        res = mte.sc_parse_bamse(args.bam, UMIS=True, whitelistfilename=args.whitelist, strand=args.strand, log=log,
                                 label=os.path.basename(args.out), maxcells=args.maxcells)
        if rank == 0:
            mte.sc_save_result(res, args.out, maxcells=args.maxcells, log=log)
    else:
        mte.load_genome()
        res = (mte.parse_bampe if args.mode == "pe" else mte.parse_bamse)(args.bam, log=log)
        if rank == 0:
            mte.save_result_bulk(res, args.out, log=log)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

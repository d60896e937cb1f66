#!/usr/bin/env python3
"""Launcher with the reference's command line (run.py:12-69 of dom0015/surrDAMH):

    python3 run.py <configuration> <N> [csv]

<configuration> is a JSON file or a bare name resolved as examples/<name>.json; <N> is the reference's
number of samplers -- here the number of lockstep chains per GPU unless the JSON carries the additive
key "no_chains".  Where the reference builds an `mpirun` line of N sampler processes + solver pool +
collector, this runs one process per visible GPU (re-executing itself under torch.distributed.run
when there is more than one) with all chains of a GPU in lockstep.  `csv` writes the reference's
saved_samples/ CSV layout (small chain counts only).
"""
import json
import os
import subprocess
import sys

# the asynchronous collector adds ~8 CUDA streams; with the default 8 hardware launch queues two of them share a queue and serialise
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    if len(sys.argv) < 3:
        print(__doc__)
        return 2
    conf_name, N = sys.argv[1], int(sys.argv[2])
    want_csv = "csv" in sys.argv[3:]
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu > 1 and "RANK" not in os.environ and "--single" not in sys.argv:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ngpu), "--master-addr",
               "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from surrdamh_b200.configuration import Configuration
    from surrdamh_b200.sampler import LockstepSampler
    from surrdamh_b200.trace_csv import CsvTraceWriter
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C = Configuration(N, conf_name)
    no_chains = C.no_chains if "no_chains" in C.conf else N
    sink = CsvTraceWriter(C.saved_samples_name) if want_csv else None
    with torch.cuda.stream(torch.cuda.Stream()):
        smp = LockstepSampler(C, no_chains=no_chains, rank=rank, world=world, device=torch.device("cuda", local), trace_sink=sink)
        for r in smp.run():
            if rank == 0:
                s = r.summary()
                print("--- SAMPLER alg%03d%s --- chains/GPU %d x %d GPUs, steps %d --- acc/rej/prerej: %d %d %d --- %.3f s, %d refits"
                      % (s["stage"], s["type"], s["chains"], world, s["steps"], s["accepted"], s["rejected"], s["prerejected"],
                         s["seconds"], s["refits"]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

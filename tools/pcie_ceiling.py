"""Plain-copy ceiling of the box: every rank moves pinned host memory to its GPU and back at the same time, no kernel.

    python tools/pcie_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/pcie_ceiling.py [--mb 1024] [--reps 10] [--bind]

Prints one JSON line (rank 0): per-rank and aggregate GB/s for H2D alone, D2H alone and both directions at once, max-over-ranks
timed with CUDA events between barriers.  The e2e leg of bench.py moves 28 B in and 41 B out per env-step, so
    ceiling env-steps/s = min(h2d_both / 28, d2h_both / 41)
is what ml4ca_env_step_host can reach on this box at this rank count (bench.py measures the same thing in-line as
e2e.copy_ceiling with the bench's own buffer sizes)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--bind", action="store_true", help="pin the rank to the CPUs NVML reports as local to its GPU")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    note = "not bound"
    if args.bind:                       # (bench.py cannot be imported for this: it re-points file descriptor 1)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & set(os.sched_getaffinity(0))
            if cpus:
                os.sched_setaffinity(0, cpus)
                note = "bound to %d CPUs local to GPU %d" % (len(cpus), local)
        except Exception as e:  # noqa: BLE001
            note = "not bound (%r)" % (e,)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(do_in, do_out):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_in.wait_event(e0), s_out.wait_event(e0)
        for _ in range(args.reps):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        a, b = torch.cuda.Event(), torch.cuda.Event()
        a.record(s_in), b.record(s_out)
        torch.cuda.current_stream().wait_event(a), torch.cuda.current_stream().wait_event(b)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return n * args.reps / float(t.item()) / 1e9      # GB/s per rank and direction

    timed(True, True)
    h2d, d2h, both = timed(True, False), timed(False, True), timed(True, True)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mb_per_copy": args.mb, "cpu_binding": note,
                          "h2d_alone_GBs_per_rank": h2d, "d2h_alone_GBs_per_rank": d2h, "each_direction_when_both_GBs_per_rank": both,
                          "aggregate_both_directions_GBs": 2 * both * world,
                          "ceiling_env_steps_per_s": world * both * 1e9 / 41.0,
                          "ceiling_note": "41 B per env-step device-to-host is the larger of the two streams (28 B the other way)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Bare pinned-host -> device copy ceiling of the box, one process per GPU (same launch as bench.py):
  python tools/h2d_microbench.py                      # one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_microbench.py
Every rank copies `--mb` megabytes from pinned memory `--iters` times (device events, max over ranks) in three shapes: one
contiguous copy, the strided copy ll_set_scans_* issues (B rows of the scan length out of a [B][stride] buffer), and 64
separate row copies (what round 1 issued).  Rank 0 prints one JSON line; bench.py's e2e H2D bytes per second can be read
against `aggregate_GBps`."""
import argparse, json, os
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=92)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rows, stride = 64, 131072 * 16
valid = args.mb * 1000 * 1000 // rows // 16 * 16          # bytes per row actually copied
host = torch.empty((rows, stride), dtype=torch.uint8).pin_memory()
devb = torch.empty((rows, stride), dtype=torch.uint8, device=dev)
flat_h = torch.empty(rows * valid, dtype=torch.uint8).pin_memory()
flat_d = torch.empty(rows * valid, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return rows * valid * args.iters / (float(ms[0]) * 1e-3) / 1e9


def contiguous():
    with torch.cuda.stream(stream):
        flat_d.copy_(flat_h, non_blocking=True)


from cuda import cudart


def strided():   # what ll_set_scans_* issues: one cudaMemcpy2DAsync, rows of `valid` bytes out of a [rows][stride] buffer
    err, = cudart.cudaMemcpy2DAsync(devb.data_ptr(), stride, host.data_ptr(), stride, valid, rows,
                                    cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, stream.cuda_stream)
    assert int(err) == 0, err


def per_row():
    with torch.cuda.stream(stream):
        for r in range(rows):
            devb[r, :valid].copy_(host[r, :valid], non_blocking=True)


res = {k: timed(f) for k, f in (("contiguous", contiguous), ("strided_2d", strided), ("per_row_64", per_row))}
if rank == 0:
    print(json.dumps({"n_gpus": world, "mb_per_copy": rows * valid / 1e6, "per_rank_GBps": res,
                      "aggregate_GBps": {k: v * world for k, v in res.items()},
                      "note": "pinned host -> device, device-event timed, max over ranks; aggregate = per-rank rate of the slowest rank x ranks"}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()

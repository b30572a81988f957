"""aggregate host<->device copy rate of the box with every GPU copying at once (run under torchrun, one rank per GPU):
pinned 1 GiB buffers, H2D + D2H concurrently on two streams, first with the process where the launcher put it, then
again after binding the rank to the CPUs NVML names as local to its GPU (fresh pinned buffers: first touch on that node).
Explains the e2e line of bench.py at N > 1: that number is bounded by what this prints."""
import os, time, json
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30


def measure(tag):
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(2)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both(reps):
        for _ in range(reps):
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    both(2); torch.cuda.synchronize()
    res = {}
    for mode in ("alone", "all"):
        # "alone": ranks take turns; "all": every rank at once
        rates = []
        for turn in range(world if mode == "alone" else 1):
            if world > 1: dist.barrier()
            torch.cuda.synchronize()
            if mode == "all" or turn == rank:
                t0 = time.perf_counter(); both(8); torch.cuda.synchronize(); dt = time.perf_counter() - t0
                rates.append(8 * n / dt / 1e9)
            if world > 1: dist.barrier()
        res[mode] = rates[0]
    t = torch.tensor([res["alone"], res["all"]], device="cuda", dtype=torch.float64)
    if world > 1:
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
    else:
        g = [t]
    if rank == 0:
        alone = [round(float(x[0]), 1) for x in g]; allr = [round(float(x[1]), 1) for x in g]
        print(json.dumps({"placement": tag, "n_gpus": world, "GBps_per_direction_alone": alone, "GBps_per_direction_all_at_once": allr,
                          "aggregate_all_at_once": round(sum(allr), 1)}), flush=True)


measure("as launched: cpus " + str(len(os.sched_getaffinity(0))))
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = {64 * i + b for i, wv in enumerate(words) for b in range(64) if (wv >> b) & 1}
    cpus &= os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
        measure("bound to the GPU's NVML cpu set: cpus " + str(len(cpus)))
    elif rank == 0:
        print(json.dumps({"placement": "nvml returned an empty cpu set"}))
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print(json.dumps({"placement": "nvml affinity unavailable: " + repr(e)}))
if world > 1:
    dist.destroy_process_group()

"""host<->device copy rates of this box (pinned memory): alone and both directions at once."""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
print(f"H2D alone {n/t(h2d)/1e9:.1f} GB/s   D2H alone {n/t(d2h)/1e9:.1f} GB/s   both at once {n/t(both)/1e9:.1f} GB/s per direction")
def chunks(k):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_a[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_b[i*c:(i+1)*c], non_blocking=True)
for k in (8, 32, 128):
    print(f"both, {k} chunks of {n//k>>20} MiB: {n/t(lambda: chunks(k))/1e9:.1f} GB/s per direction")

# the copy pattern of oip_pan_pipeline_host on C2: 16 blocks, 3 x 33.5 MB up, 99 MB down per block, 3 slots, events
import numpy as np
blk_in, blk_out, nblk = 2048 * 8192 * 2, 2048 * 24176 * 2, 16
hin = [torch.empty(blk_in * nblk, dtype=torch.uint8).pin_memory() for _ in range(3)]
hout = torch.empty(blk_out * nblk, dtype=torch.uint8).pin_memory()
din = [[torch.empty(blk_in, dtype=torch.uint8, device="cuda") for _ in range(3)] for _ in range(3)]
dout = [torch.empty(blk_out, dtype=torch.uint8, device="cuda") for _ in range(3)]
sc = torch.cuda.Stream()
def pattern(with_kernel):
    evs = []
    for b in range(nblk):
        s = b % 3
        with torch.cuda.stream(s1):
            for i in range(3): din[s][i].copy_(hin[i][b * blk_in:(b + 1) * blk_in], non_blocking=True)
            e_in = torch.cuda.Event(); e_in.record(s1)
        with torch.cuda.stream(sc):
            sc.wait_event(e_in)
            if with_kernel: dout[s][:blk_in].copy_(din[s][0], non_blocking=True)
            e_c = torch.cuda.Event(); e_c.record(sc)
        with torch.cuda.stream(s2):
            s2.wait_event(e_c)
            hout[b * blk_out:(b + 1) * blk_out].copy_(dout[s], non_blocking=True)
for wk in (False, True):
    dt = t(lambda: pattern(wk), reps=3)
    print(f"pipeline pattern (kernel={wk}): {dt*1e3:.1f} ms  ({blk_in*3*nblk/dt/1e9:.1f} GB/s up, {blk_out*nblk/dt/1e9:.1f} GB/s down)")

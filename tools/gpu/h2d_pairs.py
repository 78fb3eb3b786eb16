"""Which GPUs share a host link? Pinned-host -> device copy bandwidth of every GPU alone, of every PAIR concurrently and of
all GPUs together (one process, one copy stream per device, 256 MB per copy). A pair that shares a PCIe switch uplink /
root port gets about half of the solo rate each.   python tools/gpu/h2d_pairs.py"""
import itertools, subprocess, torch

n = torch.cuda.device_count()
MB = 256
host = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]

def rate(ids, reps=4):
    ev = {}
    for i in ids:
        with torch.cuda.device(i), torch.cuda.stream(streams[i]):
            dev[i].copy_(host[i], non_blocking=True)
    for i in ids:
        torch.cuda.synchronize(i)
    for i in ids:
        with torch.cuda.device(i), torch.cuda.stream(streams[i]):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(streams[i])
            for _ in range(reps):
                dev[i].copy_(host[i], non_blocking=True)
            b.record(streams[i])
            ev[i] = (a, b)
    out = {}
    for i in ids:
        torch.cuda.synchronize(i)
        out[i] = reps * MB * 1.048576 / ev[i][0].elapsed_time(ev[i][1])   # GB/s
    return out

solo = {i: rate([i])[i] for i in range(n)}
print("solo GB/s:", {i: round(v, 1) for i, v in solo.items()})
print("pairs (GB/s each; * = both below 75 % of solo):")
for i, j in itertools.combinations(range(n), 2):
    r = rate([i, j])
    flag = "*" if r[i] < 0.75 * solo[i] and r[j] < 0.75 * solo[j] else " "
    print(f"  {flag} ({i},{j}) {r[i]:5.1f} {r[j]:5.1f}")
r = rate(list(range(n)))
print("all together:", {i: round(v, 1) for i, v in r.items()}, "sum", round(sum(r.values()), 1))
try:
    print(subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current", "--format=csv"],
                         capture_output=True, text=True).stdout)
    print(subprocess.run("lspci -tv 2>/dev/null | grep -i -B3 nvidia | head -60", shell=True, capture_output=True, text=True).stdout)
    print(subprocess.run("lscpu | grep -i 'numa\\|socket\\|model name'", shell=True, capture_output=True, text=True).stdout)
except Exception as e:
    print(e)

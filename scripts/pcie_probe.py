"""PCIe probe: H2D alone, D2H alone, both at once (pinned memory, 156 MB each: one step's worth of the host entry point)."""
# usage: python scripts/pcie_probe.py [device] [bind]   (bind: allocate the pinned buffers NUMA-local, sagnn_b200.hostmem)
import contextlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sagnn_b200 import hostmem
devi = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 0
torch.cuda.set_device(devi)
info = {}
n = 155_556_864
with (hostmem.near_gpu(devi, info) if "bind" in sys.argv else contextlib.nullcontext()):
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.zero_(); h_out.zero_()
print("device %d, pinned pages bound near the GPU: %s" % (devi, info or "no"))
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for name, a, b in (() if "concurrent" in sys.argv else (("H2D alone", 1, 0), ("D2H alone", 0, 1), ("both at once", 1, 1))):
    run(a, b, 2); t = run(a, b)
    print("%-13s %.2f ms  %.1f GB/s per direction" % (name, t * 1e3, n / t / 1e9))
c = n // 6
def chunked(reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for k in range(6):
            with torch.cuda.stream(s1): d_in[k*c:(k+1)*c].copy_(h_in[k*c:(k+1)*c], non_blocking=True)
            with torch.cuda.stream(s2): h_out[k*c:(k+1)*c].copy_(d_out[k*c:(k+1)*c], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
if "concurrent" not in sys.argv: chunked(2); t = chunked(); print("%-13s %.2f ms  %.1f GB/s per direction" % ("both, 6 chunks", t * 1e3, n / t / 1e9))

# ---- concurrent mode: python scripts/pcie_probe.py concurrent N  -> the same duplex copy on N GPUs at once
# (children start at a common wall-clock time); shows whether the GPUs' host links are independent
if "concurrent" in sys.argv:
    import subprocess
    if "child" in sys.argv:
        t_start = float(sys.argv[sys.argv.index("child") + 1])
        run(1, 1, 2)
        while time.time() < t_start:
            pass
        t = run(1, 1, 40)
        print("CHILD %d %.3f" % (devi, n / t / 1e9))
    else:
        N = int(sys.argv[sys.argv.index("concurrent") + 1])
        t0 = time.time() + 25.0
        ps = [subprocess.Popen([sys.executable, __file__, str(i), "concurrent", "child", repr(t0)], stdout=subprocess.PIPE, text=True)
              for i in range(N)]
        rates = []
        for p in ps:
            out = p.communicate()[0]
            rates += [float(l.split()[2]) for l in out.splitlines() if l.startswith("CHILD")]
        print("%d GPUs copying both ways at once: per GPU %s GB/s per direction, sum %.1f GB/s per direction"
              % (N, " ".join("%.1f" % r for r in rates), sum(rates)))

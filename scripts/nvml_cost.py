import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml as n
from cmh_b200 import engine
from cmh_b200.index import HammingIndex
n.nvmlInit(); h = n.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, 100_000_000, 64, dev); q = engine.synth_codes(4001, 0, 8192, 64, dev)
idx = HammingIndex(db, 0, nd_total=db.n)
for _ in range(2): idx.search_packed(q, 1000)
torch.cuda.synchronize()
def step_ms(poll=None, period=0.05):
    import threading
    stop = threading.Event()
    def loop():
        while not stop.is_set():
            poll(); stop.wait(period)
    th = None
    if poll: th = threading.Thread(target=loop, daemon=True); th.start()
    time.sleep(0.2); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): idx.search_packed(q, 1000)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5 * 1e3
    stop.set()
    if th: th.join()
    return round(dt, 2)
fns = {"none": None,
       "clock_sm": lambda: n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM),
       "power": lambda: n.nvmlDeviceGetPowerUsage(h),
       "reasons": lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h),
       "all3": lambda: (n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), n.nvmlDeviceGetPowerUsage(h), n.nvmlDeviceGetCurrentClocksEventReasons(h))}
for k, f in fns.items():
    if f:
        t0 = time.perf_counter(); [f() for _ in range(20)]; per = (time.perf_counter() - t0) / 20 * 1e3
    else: per = 0
    print(k, "call ms:", round(per, 3), "step ms @50ms:", step_ms(f, 0.05), "@200ms:", step_ms(f, 0.2))

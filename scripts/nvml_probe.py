import time, pynvml, torch
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(8192, 8192, device="cuda")
def t(f, n=30):
    ts=[]
    for _ in range(n):
        a=time.perf_counter(); f(); ts.append((time.perf_counter()-a)*1e3)
    return "%.3f / %.3f ms (median/max)" % (sorted(ts)[len(ts)//2], max(ts))
for _ in range(3): (x@x)
print("clock idle   ", t(lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
print("reasons idle ", t(lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
for _ in range(200): y = x@x
print("clock busy   ", t(lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
for _ in range(200): y = x@x
print("reasons busy ", t(lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
torch.cuda.synchronize()

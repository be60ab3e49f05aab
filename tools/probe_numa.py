import pynvml, os
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("gpus", n, "cpus allowed", len(os.sched_getaffinity(0)), "of", os.cpu_count())
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    try:
        aff = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        print(i, "cpu affinity words", [hex(a) for a in aff])
    except Exception as e:
        print(i, "affinity err", e)
    try:
        print(i, "numa", pynvml.nvmlDeviceGetNumaNodeId(h))
    except Exception as e:
        print(i, "numa err", type(e).__name__, e)

import time, os
t=time.perf_counter(); s=0
for i in range(5_000_000): s+=i*i
print("host loop 5M iters: %.3f s, cpus %d, load %s" % (time.perf_counter()-t, os.cpu_count(), os.getloadavg()))

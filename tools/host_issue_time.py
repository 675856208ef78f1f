"""Host time to ISSUE one rebuild step (no sync inside) against its device time: is the step launch-bound?
   python tools/host_issue_time.py [workload]"""
import os, sys, time
sys.path.insert(0, '.')
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else 'baby'
dev = torch.device('cuda:0')
job = bench.RebuildJob(name, 'bf16', None, dev, 0, 1, 0, 'weak')
for _ in range(5):
    job.step_device()
torch.cuda.synchronize()
for streams in ("2", "1"):
    os.environ["DIFFMM_STREAMS"] = streams
    for _ in range(3):
        job.step_device()
    torch.cuda.synchronize()
    host, devt = [], []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0.record()
        job.step_device()
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3); devt.append(e0.elapsed_time(e1))
    host.sort(); devt.sort()
    print(f"{name} streams={streams}: host issue {host[len(host)//2]:.3f} ms/step, device {devt[len(devt)//2]:.3f} ms/step (start-to-end, empty queue at start)")

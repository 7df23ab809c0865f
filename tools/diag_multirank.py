"""Per-phase device time of the distributed step (diagnostic; run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from unigasfoam_b200 import cases
from unigasfoam_b200.cloud import UniGasCloud
from unigasfoam_b200.exchange import SlotExchanger

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = cases.couette(rank=rank, n_ranks=world)
cl = case.make_cloud(UniGasCloud, device=local)
ex = SlotExchanger(cl, case.mesh, rank, world, slot_capacity=16384, cuda=True)
st = torch.cuda.ExternalStream(cl.stream())
names = ["move", "pack", "sendrecv", "unpack+move", "allreduce+item", "round2", "finish"]
acc = dict.fromkeys(names, 0.0); host = dict.fromkeys(names, 0.0)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(st); return e
import ctypes as C
keep = []
for step in range(25):
    marks = [ev()]; t = [time.perf_counter()]
    cl.move(); ex.begin_step(); marks.append(ev()); t.append(time.perf_counter())
    with torch.cuda.stream(st):
        cl.migratePackSlots(ex._sp, ex.cap); marks.append(ev()); t.append(time.perf_counter())
        ops = [dist.P2POp(dist.isend, ex.send[k], b) for k, b in ex.sends] + [dist.P2POp(dist.irecv, ex.recv[k], b) for k, b in ex.recvs]
        for r in dist.batch_isend_irecv(ops): r.wait()
        marks.append(ev()); t.append(time.perf_counter())
        cl.migrateUnpackSlots(ex._rp, ex.cap); cl.moveReceived(); marks.append(ev()); t.append(time.perf_counter())
        if os.environ.get("DIAG_SYNC", "1") == "1":
            tot = ex._inflight.clone(); dist.all_reduce(tot); n = int(tot.item())
        marks.append(ev()); t.append(time.perf_counter())
        if os.environ.get("DIAG_ROUNDS", "1") == "2":
            ex._round()
        marks.append(ev()); t.append(time.perf_counter())
    cl.finishStep(); marks.append(ev()); t.append(time.perf_counter())
    if os.environ.get("DIAG_SYNC", "1") == "1" or step == 24:
        st.synchronize()
    keep.append((marks, t))
if True:
  for step, (marks, t) in enumerate(keep):
    if step >= 5:
        for i, k in enumerate(names):
            acc[k] += marks[i].elapsed_time(marks[i + 1]); host[k] += (t[i + 1] - t[i]) * 1e3
if rank == 0:
    print("phase            device_ms  host_ms  (mean over 20 steps)")
    for k in names: print(f"{k:16s} {acc[k]/20:8.3f} {host[k]/20:8.3f}")
    print("sum device", sum(acc.values()) / 20, "wall per step", keep[5][0][0].elapsed_time(keep[24][0][-1]) / 20)
dist.destroy_process_group()

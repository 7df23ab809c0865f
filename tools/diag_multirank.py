"""Per-phase device time of the distributed step with the NVLink peer-memory transfer (diagnostic; run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from unigasfoam_b200 import cases
from unigasfoam_b200.cloud import UniGasCloud
from unigasfoam_b200.exchange import PeerExchanger

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
meta = dist.new_group(backend="gloo")
case = cases.couette(rank=rank, n_ranks=world)
cl = case.make_cloud(UniGasCloud, device=local, parcelCapacity=int(1.5 * case.n_parcels) + 4096)
ex = PeerExchanger(cl, case.mesh, rank, world, slot_capacity=6000, meta_group=meta)
st = torch.cuda.ExternalStream(cl.stream())
names = ["move", "pack1", "wait+unpack1", "move_received1", "pack2", "wait+unpack2", "move_received2", "finish"]
acc = dict.fromkeys(names, 0.0)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(st); return e
keep = []
NS = 30
for step in range(NS):
    marks = [ev()]
    cl.move(); ex.begin_step(); marks.append(ev())
    for _ in range(2):
        buf = ex.round_no % 2; epoch = ex.round_no + 1; ex.round_no += 1
        slots = [d[0] + buf * d[1] for d in ex.dst]; flags = [d[2] + buf * d[3] for d in ex.dst]
        cl.migratePackPeer(slots, flags, ex.cap, epoch); marks.append(ev())
        cl.migrateUnpackPeer(ex.base + buf * ex.nproc * ex.slot_bytes, ex.base + 2 * ex.nproc * ex.slot_bytes + buf * ex.nproc * 8, ex.cap, epoch); marks.append(ev())
        cl.moveReceived(); marks.append(ev())
    cl.finishStep(); marks.append(ev())
    keep.append(marks)
st.synchronize()
for marks in keep[10:]:
    for i, k in enumerate(names):
        acc[k] += marks[i].elapsed_time(marks[i + 1])
n = NS - 10
out = [None] * world
dist.all_gather_object(out, {k: round(v / n, 4) for k, v in acc.items()}, group=meta)
if rank == 0:
    for r, o in enumerate(out):
        print("rank", r, o, "sum", round(sum(o.values()), 4))
    print("wall per step", keep[10][0].elapsed_time(keep[-1][-1]) / n)
dist.barrier(group=meta)
dist.destroy_process_group()

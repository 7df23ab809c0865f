"""Multi-rank parcel migration: the transfer loop of OpenFOAM's Cloud::move.

One process per GPU / subdomain.  After `cloud.move()` every rank packs the parcels that stopped on its
processor patches (device-side, in index order), the packed records travel with ONE variable-split
all_to_all per round over the data group (NCCL over NVLink on GPUs, gloo in CPU tests), are appended on
the receiving side and resume tracking from their stepFraction; rounds repeat until no rank has anything
in flight - the same termination rule as the reference's transfer loop.

Only plumbing lives here (torch.distributed); packing, unpacking and tracking are libugf kernels.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._capi import UGF_MIGRATE_STRIDE as STRIDE


class _DevArray:
    """Raw device pointer exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def _sender_tag(tag):
    return tag


def _receiver_view(tag):
    """The tag the *sender* used for the patch that matches a local patch with `tag`."""
    return ("i",) if tag[0] == "i" else ("c", tag[2], tag[1])


class Exchanger:
    def __init__(self, cloud, mesh, rank, world, data_group=None, meta_group=None, cuda=False):
        self.cloud, self.mesh, self.rank, self.world = cloud, mesh, rank, world
        self.data_group, self.meta_group, self.cuda = data_group, meta_group, cuda
        self.proc = [(i, p.partner, tuple(p.tag)) for i, p in enumerate(mesh.patches) if p.kind == "processor"]
        # message order between a pair of ranks: by the sender's patch tag
        self.send_order = {b: sorted([q for q in self.proc if q[1] == b], key=lambda q: _sender_tag(q[2])) for b in range(world)}
        self.recv_order = {b: sorted([q for q in self.proc if q[1] == b], key=lambda q: _receiver_view(q[2])) for b in range(world)}
        k = torch.tensor([max(len(v) for v in self.send_order.values()) if self.send_order else 0], dtype=torch.int64)
        dist.all_reduce(k, op=dist.ReduceOp.MAX, group=meta_group)
        self.K = max(int(k.item()), 1)
        self.rounds = 0
        self.sent = 0
        self._stream = None
        if cuda:
            self._stream = torch.cuda.ExternalStream(cloud.stream())

    def _wrap(self, ptr, n):
        """float64 tensor over n doubles at a raw pointer (host for the oracle, device for libugf)."""
        if n == 0:
            return torch.empty(0, dtype=torch.float64, device="cuda" if self.cuda else "cpu")
        addr = C.cast(ptr, C.c_void_p).value
        if self.cuda:
            return torch.as_tensor(_DevArray(addr, n), device="cuda")
        return torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,)))

    def exchange(self):
        """One transfer round.  Returns the number of parcels in flight over all ranks (0 = done)."""
        counts = self.cloud.migrateCounts()
        # counts[b*K + slot]: what this rank sends to rank b through its slot-th patch facing b
        vec = torch.zeros(self.world * self.K, dtype=torch.int64)
        for b in range(self.world):
            for slot, (patch, _, _) in enumerate(self.send_order[b]):
                vec[b * self.K + slot] = int(counts[patch])
        allv = [torch.zeros_like(vec) for _ in range(self.world)]
        dist.all_gather(allv, vec, group=self.meta_group)
        allv = torch.stack(allv).view(self.world, self.world, self.K)  # [sender, receiver, slot]
        total = int(allv.sum().item())
        if total == 0:
            return 0
        self.rounds += 1
        ctx = torch.cuda.stream(self._stream) if self.cuda else _Null()
        with ctx:
            chunks, in_splits = [], []
            for b in range(self.world):
                nb = 0
                for patch, _, _ in self.send_order[b]:
                    buf, n = self.cloud.migratePack(patch)
                    if n:
                        chunks.append(self._wrap(buf, n * STRIDE))
                    nb += n
                in_splits.append(nb * STRIDE)
            self.sent += sum(in_splits) // STRIDE
            dev = "cuda" if self.cuda else "cpu"
            send = torch.cat(chunks) if chunks else torch.empty(0, dtype=torch.float64, device=dev)
            out_splits = [int(allv[b, self.rank].sum().item()) * STRIDE for b in range(self.world)]
            recv = torch.empty(sum(out_splits), dtype=torch.float64, device=dev)
            dist.all_to_all_single(recv, send, out_splits, in_splits, group=self.data_group)
            off = 0
            for b in range(self.world):
                for slot, (patch, _, _) in enumerate(self.recv_order[b]):
                    n = int(allv[b, self.rank, slot].item())
                    if n:
                        ptr = C.cast(recv.data_ptr() + off * 8, C.POINTER(C.c_double))
                        self.cloud.migrateUnpack(patch, ptr, n)
                        off += n * STRIDE
            self.cloud.moveReceived()
            if self.cuda:
                self._stream.synchronize()  # recv must outlive the kernels reading it
        return total


class SlotExchanger:
    """Fixed-slot transfer rounds without per-round host bookkeeping.

    Every processor patch owns one slot (header + slot_capacity records) of a persistent send and a persistent
    receive buffer.  A round is: device-side pack of all patches (two passes over the cell ids) -> one grouped
    NCCL send/recv with the neighbours only -> device-side unpack (counts come from the slot headers) -> resume
    tracking -> all-reduce of the device-resident in-flight counter, the only value the host reads back.
    """

    def __init__(self, cloud, mesh, rank, world, slot_capacity, group=None, cuda=False):
        self.cloud, self.mesh, self.rank, self.world, self.group, self.cuda = cloud, mesh, rank, world, group, cuda
        self.cap = int(slot_capacity)
        proc = [(i, p.partner, tuple(p.tag)) for i, p in enumerate(mesh.patches) if p.kind == "processor"]
        self.slot_of = {patch: k for k, (patch, _, _) in enumerate(proc)}
        stride = (self.cap + 1) * STRIDE
        dev = "cuda" if cuda else "cpu"
        self.send = torch.zeros(max(len(proc), 1), stride, dtype=torch.float64, device=dev)
        self.recv = torch.zeros(max(len(proc), 1), stride, dtype=torch.float64, device=dev)
        # second and later rounds of a step only carry parcels that were received in this step: small slots
        self.cap2 = max(256, self.cap // 16)
        self.send2 = torch.zeros(max(len(proc), 1), (self.cap2 + 1) * STRIDE, dtype=torch.float64, device=dev)
        self.recv2 = torch.zeros(max(len(proc), 1), (self.cap2 + 1) * STRIDE, dtype=torch.float64, device=dev)
        peers = sorted({b for _, b, _ in proc})
        # order of messages between a pair of ranks: by the sender's patch tag (see Exchanger)
        self.sends = [(self.slot_of[q[0]], b) for b in peers for q in sorted([q for q in proc if q[1] == b], key=lambda q: _sender_tag(q[2]))]
        self.recvs = [(self.slot_of[q[0]], b) for b in peers for q in sorted([q for q in proc if q[1] == b], key=lambda q: _receiver_view(q[2]))]
        self.rounds = 0
        self._stream = torch.cuda.ExternalStream(cloud.stream()) if cuda else None
        ptr = cloud.migrateInflightPtr()
        if cuda:
            self._inflight = torch.as_tensor(_DevI64(ptr), device="cuda")
        else:
            self._inflight = torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(1,)))
        self._sp = C.cast(self.send.data_ptr(), C.POINTER(C.c_double))
        self._rp = C.cast(self.recv.data_ptr(), C.POINTER(C.c_double))
        self._sp2 = C.cast(self.send2.data_ptr(), C.POINTER(C.c_double))
        self._rp2 = C.cast(self.recv2.data_ptr(), C.POINTER(C.c_double))
        self._round_in_step = 0

    def begin_step(self):
        """Call after cloud.move(): the next round is the first of the step (full-size slots)."""
        self._round_in_step = 0

    def _round(self):
        self.rounds += 1
        first = self._round_in_step == 0
        self._round_in_step += 1
        send, recv, sp, rp, cap = (self.send, self.recv, self._sp, self._rp, self.cap) if first else \
                                  (self.send2, self.recv2, self._sp2, self._rp2, self.cap2)
        self.cloud.migratePackSlots(sp, cap)
        ops = [dist.P2POp(dist.isend, send[k], b, self.group) for k, b in self.sends]
        ops += [dist.P2POp(dist.irecv, recv[k], b, self.group) for k, b in self.recvs]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.cloud.migrateUnpackSlots(rp, cap)
        self.cloud.moveReceived()

    def exchange(self):
        """One transfer round; returns the number of parcels still in flight over all ranks afterwards."""
        ctx = torch.cuda.stream(self._stream) if self.cuda else _Null()
        with ctx:
            self._round()
            total = self._inflight.clone()
            dist.all_reduce(total, group=self.group)
            return int(total.item())

    # -- fixed-round mode: no host synchronisation inside the step ---------------------------------------------
    def exchange_fixed(self, n_rounds, exact_tail=False, max_rounds=512):
        """n_rounds transfer rounds back to back with no host synchronisation and no global collective; whatever
        is still in flight afterwards is added to a device-resident tally that check_settled() reduces over all
        ranks.  A parcel left in flight (n_rounds too small for this decomposition) is therefore reported loudly
        at the next check, never dropped silently.

        exact_tail: after the n_rounds unsynchronised rounds the in-flight count is reduced once and further rounds follow
        until it is zero - the reference's termination rule with one collective per step instead of one per round, for
        decompositions whose round count has a tail (block corners, an axis between symmetry planes)."""
        ctx = torch.cuda.stream(self._stream) if self.cuda else _Null()
        with ctx:
            for _ in range(n_rounds):
                self._round()
            if exact_tail:
                total = self._inflight.clone()
                dist.all_reduce(total, group=self.group)
                left = int(total.item())
                extra = 0
                while left > 0:
                    extra += 1
                    if extra > max_rounds:
                        raise RuntimeError("parcel migration did not settle")
                    self._round()
                    total = self._inflight.clone()
                    dist.all_reduce(total, group=self.group)
                    left = int(total.item())
                self.tail_rounds = getattr(self, "tail_rounds", 0) + extra
                return
            if not hasattr(self, "_left"):
                self._left = torch.zeros_like(self._inflight)
            self._left += self._inflight

    def check_settled(self):
        """Collective: raises if any rank had parcels waiting on a processor patch after a fixed-round step."""
        if not hasattr(self, "_left"):
            return
        ctx = torch.cuda.stream(self._stream) if self.cuda else _Null()
        with ctx:
            total = self._left.clone()
            dist.all_reduce(total, group=self.group)
            n = int(total.item())
            self._left.zero_()
        if n != 0:
            raise RuntimeError(f"{n} parcels were still waiting on processor patches after the fixed number of transfer "
                               "rounds: raise the round count for this decomposition")


class PeerExchanger(SlotExchanger):
    """Fixed-slot transfer rounds over NVLink peer memory: the pack kernel writes each processor patch's records
    straight into the matching receive slot in the neighbour's HBM and raises a flag there; the neighbour's
    unpack waits for the flag on the device.  No send buffer, no NCCL call and no host synchronisation on the data
    path (torch.distributed is only used once, to pass the cudaIpc handles and the slot maps around).

    Receive region per rank: 2 buffers x nProc slots x (cap + 1) records, then 2 x nProc flag words.  Round R (counted
    over the whole run) uses buffer R % 2 and epoch R + 1: a slot is rewritten two rounds later, after the round in
    between has been acknowledged by both sides (ugf.h)."""

    def __init__(self, cloud, mesh, rank, world, slot_capacity, group=None, meta_group=None):
        super().__init__(cloud, mesh, rank, world, slot_capacity, group=group, cuda=True)
        self.send = self.recv = self.send2 = self.recv2 = None  # the NCCL path's staging buffers are not needed
        nproc = len(self.slot_of)
        self.nproc = nproc
        self.slot_bytes = (self.cap + 1) * STRIDE * 8
        self.region_bytes = 2 * nproc * self.slot_bytes + 2 * nproc * 8
        self.base, handle = cloud.peerAlloc(max(self.region_bytes, 64))
        # what every rank publishes: its handle, its slot count and, per sending peer, the local slots in the order
        # the peer's messages arrive (same pairing rule as the NCCL path: i-th send to b <-> i-th receive from a)
        recv_from = {}
        for k, b in self.recvs:
            recv_from.setdefault(b, []).append(k)
        pub = [None] * world
        dist.all_gather_object(pub, dict(handle=handle, nproc=nproc, slot_bytes=self.slot_bytes, recv_from=recv_from), group=meta_group)
        self.peer_base = {}
        for b in sorted({b for _, b in self.sends}):
            if pub[b]["slot_bytes"] != self.slot_bytes:
                raise RuntimeError("neighbouring ranks disagree on the migration slot capacity")
            self.peer_base[b] = self.base if b == rank else cloud.peerOpen(pub[b]["handle"])
        # destination (slot address for buffer 0, flag address for buffer 0, strides to buffer 1) of my slot k
        self.dst = [None] * nproc
        ordinal = {}
        for k, b in self.sends:
            i = ordinal.get(b, 0)
            ordinal[b] = i + 1
            kb = pub[b]["recv_from"][rank][i]
            nb = pub[b]["nproc"]
            base = self.peer_base[b]
            self.dst[k] = (base + kb * self.slot_bytes, nb * self.slot_bytes, base + 2 * nb * self.slot_bytes + kb * 8, nb * 8)
        self.round_no = 0
        dist.barrier(group=meta_group)  # every region is mapped before the first round writes into it

    def _round(self):
        self.rounds += 1
        self._round_in_step += 1
        buf = self.round_no % 2
        epoch = self.round_no + 1
        self.round_no += 1
        slots = [d[0] + buf * d[1] for d in self.dst]
        flags = [d[2] + buf * d[3] for d in self.dst]
        self.cloud.migratePackPeer(slots, flags, self.cap, epoch)
        self.cloud.migrateUnpackPeer(self.base + buf * self.nproc * self.slot_bytes,
                                     self.base + 2 * self.nproc * self.slot_bytes + buf * self.nproc * 8, self.cap, epoch)
        self.cloud.moveReceived()


class ProcessorHalo:
    """Cell values across processor faces: `halo(a)` takes this rank's values on its processor faces ([nProcFaces, k] in
    patch order - the owner-cell values for fvc::interpolate / FaceCellWave) and returns what the neighbour ranks hold on
    the same faces.  One message per neighbour rank and call (the patches facing that rank concatenated in tag order, as
    the migration does), every 10-100 steps (hybrid-mask and adaptation smoothing), so this is plumbing, not a
    kernel: host arrays over the process group (staged through the GPU when the group is NCCL).  Every rank of the group
    must make the same sequence of calls."""

    def __init__(self, mesh, rank, world, group=None, cuda=False):
        self.rank, self.group, self.cuda = rank, group, cuda
        off, self.proc = 0, []
        for p in mesh.patches:  # the order FaceOperators lists processor faces in (empty patches are skipped there too)
            if p.kind == "processor" and p.size:
                self.proc.append((off, p.size, p.partner, tuple(p.tag)))
                off += p.size
        self.n = off
        peers = sorted({q[2] for q in self.proc})
        self.send_order = {b: sorted([q for q in self.proc if q[2] == b], key=lambda q: _sender_tag(q[3])) for b in peers}
        self.recv_order = {b: sorted([q for q in self.proc if q[2] == b], key=lambda q: _receiver_view(q[3])) for b in peers}
        self.calls = 0

    def __call__(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.ndim != 2 or a.shape[0] != self.n:
            raise ValueError(f"halo: expected [{self.n}, k] values on the processor faces, got {a.shape}")
        self.calls += 1
        out = np.empty_like(a)
        dev = "cuda" if self.cuda else "cpu"
        ops, recv = [], {}
        for b in self.send_order:
            snd = torch.from_numpy(np.concatenate([a[o:o + n] for o, n, _, _ in self.send_order[b]])).to(dev)
            recv[b] = torch.empty_like(snd)
            ops.append(dist.P2POp(dist.isend, snd, b, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv[b], b, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for b, buf in recv.items():
            buf, at = buf.cpu().numpy(), 0
            for o, n, _, _ in self.recv_order[b]:
                out[o:o + n] = buf[at:at + n]
                at += n
        return out


def group_reducers(group=None, cuda=False):
    """(reduce_min, reduce_max) over the ranks of a process group for UniGasDynamicAdapter / FaceOperators."""
    def make(op):
        def red(v):
            t = torch.tensor([float(v)], dtype=torch.float64, device="cuda" if cuda else "cpu")
            dist.all_reduce(t, op=op, group=group)
            return float(t.item())
        return red
    return make(dist.ReduceOp.MIN), make(dist.ReduceOp.MAX)


def max_imbalance(n_local, group=None, cuda=False):
    """uniGasDynamicLoadBalancing::calculate (U/dynamicLoadBalancing/uniGasDynamicLoadBalancing.C:48-68) over the ranks of a
    decomposed run: max |N_rank - N / nRanks| / (N / nRanks) in per cent - the `Maximum imbalance` line of the reference's
    log.  n_local = cloud.size() of this rank."""
    dev = "cuda" if cuda else "cpu"
    tot = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    ideal = float(tot.item()) / dist.get_world_size(group)
    worst = torch.tensor([abs(float(n_local) - ideal)], dtype=torch.float64, device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    return 100.0 * float(worst.item()) / ideal if ideal > 0 else 0.0


class _DevI64:
    def __init__(self, ptr):
        self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def evolve_distributed(cloud, exchanger, n_steps=1, inflow=False, max_rounds=64, fixed_rounds=None, exact_tail=False):
    """uniGasCloud::evolve across ranks: the phases of ugf_step with the transfer loop after the move.  Both
    exchangers return 0 from exchange() once no rank has parcels in flight (the reference's termination rule).
    fixed_rounds (SlotExchanger / PeerExchanger): run exactly that many rounds without host synchronisation and verify
    quiescence one step later; with exact_tail the count is checked once per step after them and further rounds follow
    until nothing is in flight (exact like the reference, one collective per step)."""
    for _ in range(n_steps):
        if inflow:
            cloud.controlBeforeMove()
        cloud.move()
        if hasattr(exchanger, "begin_step"):
            exchanger.begin_step()
        if fixed_rounds:
            if exact_tail:
                exchanger.exchange_fixed(fixed_rounds, exact_tail=True, max_rounds=max_rounds)
            else:
                exchanger.exchange_fixed(fixed_rounds)
            cloud.finishStep()
            continue
        for _r in range(max_rounds):
            if exchanger.exchange() == 0:
                break
        else:
            raise RuntimeError("parcel migration did not settle")
        cloud.finishStep()
    if fixed_rounds and not exact_tail:
        exchanger.check_settled()


class LocalSubdomains:
    """All subdomains of a decomposed case inside ONE process: the transfer loop of Cloud::move between clouds that live in
    the same address space (host pointers), with the exact termination rule.  Used by bench.py's reference arm - the CPU
    restatement of an N-rank run on the host's cores, where a process per rank would only split the same cores - and by CPU
    tests of the per-rank block builders; with libugf clouds: all subdomains on ONE GPU (one handle each), which runs the
    processor-patch path of the move, the pack / unpack kernels and the resumed tracks without a second device.
    clouds[r] / meshes[r] = rank r."""

    def __init__(self, clouds, meshes):
        self.clouds, self.meshes = clouds, meshes
        self.links = []  # (rank, patch, peer rank, peer patch)
        for r, m in enumerate(meshes):
            for pi, p in enumerate(m.patches):
                if p.kind != "processor":
                    continue
                peer = p.partner
                want = _receiver_view(tuple(p.tag)) if p.tag else ("i",)
                qi = p.peer_patch
                if qi < 0:
                    cands = [k for k, q in enumerate(meshes[peer].patches)
                             if q.kind == "processor" and q.partner == r and (tuple(q.tag) if q.tag else ("i",)) == want]
                    if len(cands) != 1:
                        raise RuntimeError(f"rank {r} patch {p.name}: no unique matching processor patch on rank {peer}")
                    qi = cands[0]
                if meshes[peer].patches[qi].size != p.size:
                    raise RuntimeError(f"processor patches {p.name} / {meshes[peer].patches[qi].name} differ in size")
                self.links.append((r, pi, peer, qi))
        self.rounds = 0

    def exchange(self):
        """One transfer round over all subdomains; returns the number of parcels that were in flight."""
        total = 0
        for cl in self.clouds:
            total += int(cl.migrateCounts().sum())
        if total == 0:
            return 0
        self.rounds += 1
        packed = []
        for r, pi, peer, qi in self.links:  # pack everything first: a parcel received in this round waits for the next one
            buf, n = self.clouds[r].migratePack(pi)
            if n and getattr(self.clouds[r], "migrate_buffers", "host") == "device":
                # libugf clouds on one GPU: the packed records stay in the sender's device buffer (valid until its next pack on
                # that patch, i.e. the next round) and the receiver reads them from there
                packed.append((peer, qi, buf, n))
            elif n:
                rec = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(n * STRIDE,)).copy()
                packed.append((peer, qi, rec, n))
        for peer, qi, rec, n in packed:
            ptr = rec if not isinstance(rec, np.ndarray) else rec.ctypes.data_as(C.POINTER(C.c_double))
            self.clouds[peer].migrateUnpack(qi, ptr, n)
        for cl in self.clouds:
            cl.moveReceived()
        return total

    def evolve(self, n_steps=1, inflow=False, max_rounds=64):
        for _ in range(n_steps):
            for cl in self.clouds:
                if inflow:
                    cl.controlBeforeMove()
                cl.move()
            for _r in range(max_rounds):
                if self.exchange() == 0:
                    break
            else:
                raise RuntimeError("parcel migration did not settle")
            for cl in self.clouds:
                cl.finishStep()

    def size(self):
        return sum(cl.size() for cl in self.clouds)

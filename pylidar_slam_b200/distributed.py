"""Multi-GPU plumbing: one process per GPU; every rank holds the whole local
map (kd) or all model maps (projective) and reduces a shard of the correspondences; the 30
normal-equation accumulators are all-reduced once per ICP iteration inside the library (NCCL on the
context's stream).  torch.distributed only carries the rendezvous (the 128-byte ncclUniqueId)."""
import ctypes as C
import glob
import os

import numpy as np


def query_shard(num_queries: int, rank: int, world: int) -> np.ndarray:
    """kd map: rank r reduces queries r, r + world, ... (the kd kernels' q_begin/q_stride)."""
    return np.arange(rank, num_queries, world)


def pixel_shard(num_pixels: int, rank: int, world: int):
    """projective map: rank r reduces pixels [hw*r/world, hw*(r+1)/world) (proj_icp_iter_kernel)."""
    return num_pixels * rank // world, num_pixels * (rank + 1) // world


def find_nccl() -> str:
    import torch
    cands = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib",
                                          "libnccl.so*")))
    return cands[0] if cands else "libnccl.so.2"


def broadcast_unique_id(make_id, dist, rank: int, device=None) -> bytes:
    """Rank 0 creates the 128-byte id with `make_id()`; every rank returns the same bytes."""
    import torch
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.tensor(list(make_id()), dtype=torch.uint8)
    if device is not None:
        uid = uid.to(device)
    dist.broadcast(uid, 0)
    return bytes(uid.cpu().tolist())


def _all_agree(dist, ok: bool, device) -> bool:
    """True iff `ok` on EVERY rank (one tiny MIN all-reduce): ranks must leave a failed attempt together."""
    import torch
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(int(flag.cpu()[0]) == 1)


def init_comm(ctx, dist, rank: int, world: int, device, mode: str = "p2p") -> str:
    """Connects `ctx` to its peer ranks (collective over all ranks) and returns the mode in use.

    mode "p2p" : one-shot all-reduce over NVLink peer memory fused into the solve kernel (CUDA IPC handles
                 all-gathered with torch.distributed); mode "nccl": ncclAllReduce on the library's stream.
    If any rank cannot export or map the IPC handles (IPC disabled in a container, no peer access between two
    GPUs), ALL ranks agree on it and fall back to the NCCL mode together instead of hanging or crashing."""
    if mode == "p2p":
        import logging

        import torch
        buf = (C.c_ubyte * 64)()
        err = None
        try:
            ctx.call("pls_comm_p2p_handle", world, buf)
        except Exception as e:  # noqa: BLE001  (the failure is reported below, by every rank)
            err = e
        if _all_agree(dist, err is None, device):
            mine = torch.tensor(list(buf), dtype=torch.uint8, device=device)
            gathered = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(world)]
            dist.all_gather(gathered, mine)
            raw = b"".join(bytes(t.cpu().tolist()) for t in gathered)
            try:
                ctx.call("pls_comm_p2p_init", world, rank, raw)
            except Exception as e:  # noqa: BLE001
                err = e
            if _all_agree(dist, err is None, device):
                dist.barrier()
                return "p2p"
            if err is None:
                ctx.call("pls_comm_destroy")  # this rank mapped its peers, another one could not
        if rank == 0 or err is not None:
            logging.warning("plslam_b200: peer-to-peer exchange unavailable (%s); all ranks use the NCCL all-reduce",
                            err if err is not None else "a peer rank failed")
    nccl_path = find_nccl().encode()

    def make_id():
        buf = (C.c_ubyte * 128)()
        st = ctx.lib.pls_comm_unique_id(nccl_path, buf)
        assert st == 0, "pls_comm_unique_id failed"
        return bytes(buf)

    raw = broadcast_unique_id(make_id, dist, rank, device)
    ctx.call("pls_comm_init", world, rank, raw, nccl_path)
    return "nccl"


def parse_cpulist(text: str) -> set:
    """'0-31,64-95' -> {0, ..., 31, 64, ..., 95} (the sysfs / cpuset list format)."""
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def format_cpulist(cpus) -> str:
    out, run = [], []
    for c in sorted(cpus) + [None]:
        if run and (c is None or c != run[-1] + 1):
            out.append(str(run[0]) if len(run) == 1 else f"{run[0]}-{run[-1]}")
            run = []
        if c is not None:
            run.append(c)
    return ",".join(out)


def pin_to_gpu_numa(device_index: int, sysfs: str = "/sys", min_cpus: int = 8):
    """One process per GPU on a two-socket box: keeps every thread of this process (the Python thread, the CUDA driver's
    helpers, later pools) on the CPUs of the socket the GPU hangs off (`local_cpulist` of its PCI device), so that launches,
    pinned host buffers (first touch) and the GPU's writes into mapped host memory do not cross the inter-socket link --
    what `numactl --cpunodebind` does from outside.  Call it right after torch.cuda.set_device and BEFORE any pinned
    allocation.  Returns the CPU list it pinned to, or None when there is nothing to do or anything is unavailable
    (no sysfs entry, a cpuset that already is narrower, fewer than `min_cpus` CPUs left): never raises."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, "bus", "pci", "devices", bdf, "local_cpulist")) as fh:
            local = parse_cpulist(fh.read())
        current = os.sched_getaffinity(0)
        want = local & current
        if len(want) < min_cpus or want == current:
            return None
        for tid in os.listdir("/proc/self/task"):
            try:
                os.sched_setaffinity(int(tid), want)
            except OSError:
                pass
        return format_cpulist(want)
    except Exception:  # noqa: BLE001  (an optimisation, never a reason to fail)
        return None


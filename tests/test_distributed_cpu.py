"""World-size-2 gloo tests (CPU) of the N>1 host logic: shard partitions cover every correspondence
exactly once, the all-reduced shard accumulators equal the unsharded normal equations (same Gauss-Newton
step), and the unique-id rendezvous delivers identical bytes to every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pylidar_slam_b200.distributed import broadcast_unique_id, pixel_shard, query_shard


def test_shards_partition_exactly():
    for n in (0, 1, 7, 32514, 262144):
        for world in (1, 2, 4, 8):
            q = np.concatenate([query_shard(n, r, world) for r in range(world)])
            assert sorted(q.tolist()) == list(range(n))
            edges = [pixel_shard(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def _accumulators(ref, tgt, nrm, scheme, sigma):
    """The 30 accumulators (21 upper JtWJ, 6 JtWr, sum (w r)^2, sum r^2, count) in float64."""
    from oracle import icp_oracle as orc
    x = torch.zeros(1, 6, dtype=ref.dtype)
    J = orc.p2plane_jacobian(x, tgt, nrm)[0]
    r = orc.p2plane_residual(x, tgt, ref, nrm)[0]
    w = orc.ls_weights(scheme, sigma, r.unsqueeze(0), tgt, ref)[0]
    if w.dim() == 0 or w.numel() == 1:
        w = torch.ones_like(r) * w.reshape(-1)[0]
    wj = (J * w.unsqueeze(-1)).double()
    wr = (r * w).double()
    H = wj.t() @ wj
    acc = [H[a, b] for a in range(6) for b in range(a, 6)] + list(wj.t() @ wr) + [(wr * wr).sum(), (r.double() ** 2).sum(),
                                                                                 torch.tensor(float(r.numel()), dtype=torch.float64)]
    return torch.stack([torch.as_tensor(v, dtype=torch.float64) for v in acc])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    N = 5001
    tgt = torch.randn(1, N, 3) * 10
    nrm = torch.randn(1, N, 3)
    nrm /= nrm.norm(dim=-1, keepdim=True)
    ref = tgt + 0.01 * torch.randn(1, N, 3) + torch.tensor([0.05, -0.02, 0.01])
    full = _accumulators(ref, tgt, nrm, "geman_mcclure", 0.3)
    idx = torch.from_numpy(query_shard(N, rank, world))
    part = _accumulators(ref[:, idx], tgt[:, idx], nrm[:, idx], "geman_mcclure", 0.3)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)               # what ncclAllReduce does on the device
    lo, hi = pixel_shard(N, rank, world)
    part2 = _accumulators(ref[:, lo:hi], tgt[:, lo:hi], nrm[:, lo:hi], "geman_mcclure", 0.3)
    dist.all_reduce(part2, op=dist.ReduceOp.SUM)
    uid = broadcast_unique_id(lambda: bytes(range(128)), dist, rank)
    ok = bool(torch.allclose(part, full, rtol=1e-10, atol=1e-9)) and bool(torch.allclose(part2, full, rtol=1e-10, atol=1e-9))
    ok = ok and uid == bytes(range(128)) and int(part[29]) == N
    out[rank] = ok
    dist.destroy_process_group()


def test_sharded_reduction_equals_unsharded_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


class _FakeCommCtx:
    """Stands in for a pls context in the rendezvous tests: records the calls, fails where told to."""

    def __init__(self, fail_at=None):
        self.fail_at, self.calls = fail_at, []

        class _Lib:
            @staticmethod
            def pls_comm_unique_id(path, buf):
                for i in range(128):
                    buf[i] = i
                return 0
        self.lib = _Lib()

    def call(self, name, *args):
        self.calls.append(name)
        if name == self.fail_at:
            raise RuntimeError(f"{name}: simulated failure")
        if name == "pls_comm_p2p_handle":
            for i in range(64):
                args[1][i] = (i + 1) % 256
        return 0


def _comm_worker(rank, world, port, scenario, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pylidar_slam_b200.distributed import init_comm
    fail = {"ok": None, "handle_fails_on_1": "pls_comm_p2p_handle" if rank == 1 else None,
            "map_fails_on_0": "pls_comm_p2p_init" if rank == 0 else None}[scenario]
    ctx = _FakeCommCtx(fail)
    mode = init_comm(ctx, dist, rank, world, torch.device("cpu"), mode="p2p")
    out[rank] = (mode, list(ctx.calls))
    dist.destroy_process_group()


def test_p2p_rendezvous_falls_back_to_nccl_on_every_rank_together():
    """init_comm: if ANY rank cannot export or map the CUDA-IPC handles, all ranks agree (MIN all-reduce) and take the
    NCCL mode together; a rank that had already mapped its peers releases them first."""
    for scenario, want in (("ok", "p2p"), ("handle_fails_on_1", "nccl"), ("map_fails_on_0", "nccl")):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        out = mp.Manager().dict()
        mp.spawn(_comm_worker, args=(2, port, scenario, out), nprocs=2, join=True)
        assert out[0][0] == out[1][0] == want, (scenario, dict(out))
        if want == "nccl":
            assert out[0][1][-1] == out[1][1][-1] == "pls_comm_init"
        if scenario == "map_fails_on_0":
            assert "pls_comm_destroy" in out[1][1] and "pls_comm_destroy" not in out[0][1]
        if scenario == "handle_fails_on_1":
            assert "pls_comm_p2p_init" not in out[0][1] and "pls_comm_p2p_init" not in out[1][1]


def test_numa_pinning_helper(tmp_path, monkeypatch):
    """pin_to_gpu_numa: the GPU's PCI `local_cpulist` intersected with the current affinity, applied to every thread of
    the process; a no-op (None) when the entry is missing, when nothing would change or when too few CPUs would be left."""
    import types
    import pylidar_slam_b200.distributed as D
    assert D.parse_cpulist("0-31,64-95\n") == set(range(32)) | set(range(64, 96))
    assert D.parse_cpulist("3") == {3} and D.parse_cpulist("") == set()
    assert D.format_cpulist(set(range(32)) | set(range(64, 96)) | {100}) == "0-31,64-95,100"
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1B, pci_device_id=0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "local_cpulist").write_text("0-31,64-95\n")
    calls = []
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(128)))
    monkeypatch.setattr(os, "sched_setaffinity", lambda tid, cpus: calls.append((tid, set(cpus))))
    got = D.pin_to_gpu_numa(0, sysfs=str(tmp_path))
    assert got == "0-31,64-95"
    tids = {int(t) for t in os.listdir("/proc/self/task")}
    assert {t for t, _ in calls} == tids and all(c == set(range(32)) | set(range(64, 96)) for _, c in calls)
    calls.clear()
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(32)) | set(range(64, 96)))   # already there
    assert D.pin_to_gpu_numa(0, sysfs=str(tmp_path)) is None and not calls
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(28, 36)))                    # 4 CPUs would be left
    assert D.pin_to_gpu_numa(0, sysfs=str(tmp_path)) is None and not calls
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(128)))
    assert D.pin_to_gpu_numa(0, sysfs=str(tmp_path / "nowhere")) is None and not calls               # no sysfs entry
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: (_ for _ in ()).throw(RuntimeError("no device")))
    assert D.pin_to_gpu_numa(0, sysfs=str(tmp_path)) is None and not calls

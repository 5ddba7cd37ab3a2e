"""CPU coverage of the N > 1 path (world_size 2, gloo): the shard plan, the torch.distributed binding of p2g_allgather_fn
(host buffers, in-place and out-of-place), and the property the sharding rests on -- the leaves a rank owns are whole LDE
cosets whose Merkle subtrees end exactly at its cap entries, so all-gathering per-rank subtree caps gives the oracle's cap."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan(p2g):
    plan = p2g.sharding.shard_plan(20, 3, 4, 8)
    assert [p["leaves"] for p in plan][:2] == [(0, 1 << 20), (1 << 20, 2 << 20)]
    assert plan[3]["cosets"] == (3, 4) and plan[3]["cap_entries"] == (6, 8)
    assert p2g.sharding.shard_plan(22, 3, 4, 2)[1]["cosets"] == (4, 8)
    assert p2g.sharding.query_owner((5 << 20) + 17, 20, 3, 8) == 5
    for bad in (3, 16, 0):
        with pytest.raises(ValueError):
            p2g.sharding.shard_plan(20, 3, 4, bad)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        from conftest import load_product
        from oracle import corc
        p2g = load_product()
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        grp = p2g.sharding.TorchDistGroup(device=None)
        cb = grp.callback()
        # 1. the callback as the library calls it: out-of-place, then in-place (send = recv + rank * bytes)
        n = 4096 + 24
        mine = np.full(n, rank + 1, dtype=np.uint8)
        out = np.zeros(n * world, dtype=np.uint8)
        assert cb(None, mine.ctypes.data, out.ctypes.data, n, 0) == 0
        assert all((out[r * n:(r + 1) * n] == r + 1).all() for r in range(world))
        buf = np.zeros(n * world, dtype=np.uint8)
        buf[rank * n:(rank + 1) * n] = 7 * (rank + 1)
        assert cb(None, buf.ctypes.data + rank * n, buf.ctypes.data, n, 0) == 0
        assert all((buf[r * n:(r + 1) * n] == 7 * (r + 1)).all() for r in range(world))
        # 2. coset decomposition: rank r hashes only its leaves, subtree caps are all-gathered -> the oracle's cap
        log_n, rate_bits, cap_height, ncols = 6, 3, 4, 11
        rng = np.random.default_rng(99)
        coeffs = rng.integers(0, 0xFFFFFFFF00000001, size=(ncols, 1 << log_n), dtype=np.uint64)
        lde = corc.lde(coeffs, rate_bits)                          # [ncols][8N], leaf order
        for hasher, hs in (("keccak25", 25), ("poseidon", 32)):
            full = corc.merkle_cap(lde, cap_height, hasher)
            plan = p2g.sharding.shard_plan(log_n, rate_bits, cap_height, world)[rank]
            lo, hi = plan["leaves"]
            logw = world.bit_length() - 1
            part = corc.merkle_cap(np.ascontiguousarray(lde[:, lo:hi]), cap_height - logw, hasher)
            part = np.ascontiguousarray(part, dtype=np.uint8)
            got = np.zeros(part.size * world, dtype=np.uint8)
            assert cb(None, part.ctypes.data, got.ctypes.data, part.size, 0) == 0
            assert bytes(got) == bytes(np.asarray(full, dtype=np.uint8)), hasher
            c0, c1 = plan["cap_entries"]
            assert bytes(part) == bytes(np.asarray(full, dtype=np.uint8)[c0 * hs:c1 * hs])
        # 2b. the id of the library-owned NCCL communicator travels from rank 0 over the process group (no GPU needed for the id itself)
        ng = p2g.sharding.NcclGroup.from_torch_dist(device=None)
        ids = [None] * world
        dist.all_gather_object(ids, ng.unique_id)
        assert (ng.rank, ng.world) == (rank, world) and len(ng.unique_id) == 128 and ids[0] == ids[1] and any(ids[0])
        # 3. the leaves of a rank are whole cosets: coset z of the LDE = size-N transform with shift g * w^bitrev(z)
        assert (hi - lo) % (1 << log_n) == 0 and lo // (1 << log_n) == plan["cosets"][0]
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))


def test_allgather_binding_and_cap_decomposition_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_thread_group_allgather_host_buffers(p2g):
    """The in-process rank group used by the single-GPU sharded tests: host-buffer exchange, out-of-place and in-place."""
    world, n = 4, 1000
    group = p2g.sharding.ThreadGroup(world)

    def rank_main(rank, member):
        cb = member.callback()
        mine = np.full(n, rank + 1, dtype=np.uint8)
        out = np.zeros(n * world, dtype=np.uint8)
        assert cb(None, mine.ctypes.data, out.ctypes.data, n, 0) == 0
        buf = np.zeros(n * world, dtype=np.uint8)
        buf[rank * n:(rank + 1) * n] = 10 + rank
        assert cb(None, buf.ctypes.data + rank * n, buf.ctypes.data, n, 0) == 0
        return out, buf
    for out, buf in group.run(rank_main):
        assert all((out[r * n:(r + 1) * n] == r + 1).all() and (buf[r * n:(r + 1) * n] == 10 + r).all() for r in range(world))


def test_thread_group_propagates_rank_errors(p2g):
    group = p2g.sharding.ThreadGroup(2)

    def rank_main(rank, member):
        if rank == 1:
            raise RuntimeError("rank 1 failed")
        member.allgather(0, 0, 0, False)      # would wait for rank 1 forever without the barrier abort
    with pytest.raises(Exception):
        group.run(rank_main)

"""Host bindings of the library's exchange callback (`p2g_allgather_fn`, include/p2g.h) for coset-sharded proofs.

One process per GPU: `TorchDistGroup` binds the callback to `torch.distributed` (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  `ThreadGroup` runs the ranks as threads of one process (several handles on one or more devices) and is
used to exercise the sharded path on a single GPU.
"""
import ctypes as C
import threading

from . import lib as _lib


def shard_plan(degree_bits, rate_bits, cap_height, world):
    """Which leaves / cosets / cap entries each rank owns (mirrors p2g_circuit_create_sharded)."""
    logw = world.bit_length() - 1
    if world < 1 or (1 << logw) != world or logw > min(rate_bits, cap_height):
        raise ValueError("world must be a power of two <= 2^min(rate_bits, cap_height)")
    lde = 1 << (degree_bits + rate_bits)
    per, nz, ncap = lde // world, (1 << rate_bits) // world, (1 << min(cap_height, degree_bits + rate_bits)) // world
    return [{"rank": r, "leaves": (r * per, (r + 1) * per), "cosets": (r * nz, (r + 1) * nz),
             "cap_entries": (r * ncap, (r + 1) * ncap)} for r in range(world)]


def query_owner(index, degree_bits, rate_bits, world):
    return index // ((1 << (degree_bits + rate_bits)) // world)


class _DevBuf:
    """A raw device pointer as a __cuda_array_interface__ object (zero-copy view for torch)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _tensor(ptr, nbytes, device):
    import torch
    if device is None:
        return torch.frombuffer((C.c_ubyte * nbytes).from_address(ptr), dtype=torch.uint8)
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=f"cuda:{device}")


class _Group:
    rank = 0
    world = 1

    def callback(self):
        """The C function pointer handed to p2g_circuit_create_sharded (kept alive by the group)."""
        if getattr(self, "_cb", None) is None:
            def cb(user, send, recv, nbytes, is_device):
                try:
                    return self.allgather(send, recv, nbytes, bool(is_device))
                except Exception:  # never unwind into C
                    import traceback
                    traceback.print_exc()
                    return -1
            self._cb = _lib.ALLGATHER_FN(cb)
        return self._cb


class TorchDistGroup(_Group):
    """All-gather over an initialised torch.distributed process group (NCCL or gloo)."""

    def __init__(self, device=None, group=None):
        import torch.distributed as dist
        self.dist, self.group, self.device = dist, group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nccl = dist.get_backend(group) == "nccl"
        self.calls, self.bytes = 0, 0

    def allgather(self, send, recv, nbytes, is_device):
        import torch
        self.calls += 1
        self.bytes += nbytes * self.world
        if is_device:
            s, r = _tensor(send, nbytes, self.device), _tensor(recv, nbytes * self.world, self.device)
            self.dist.all_gather_into_tensor(r, s, group=self.group)
            torch.cuda.current_stream(self.device).synchronize()
            return 0
        s, r = _tensor(send, nbytes, None), _tensor(recv, nbytes * self.world, None)
        if self.nccl:   # host buffers (caps, opened rows: a few KB) are staged through the device
            sd = s.to(f"cuda:{self.device}")
            rd = torch.empty(nbytes * self.world, dtype=torch.uint8, device=sd.device)
            self.dist.all_gather_into_tensor(rd, sd, group=self.group)
            r.copy_(rd.cpu())
        else:
            out = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(self.world)]
            self.dist.all_gather(out, s.clone(), group=self.group)
            r.copy_(torch.cat(out))
        return 0


NCCL_UNIQUE_ID_BYTES = 128


def preload_nccl():
    """Inside a Python process PyTorch brings its own libnccl.so.2 (nvidia-nccl wheel), newer than the system one, and a process can
    hold only one library of that SONAME: load PyTorch's copy first so that libp2g (which takes whatever is already loaded) and a
    later `import torch` agree.  A host without PyTorch (the Rust CLI) simply gets the system libnccl."""
    import importlib.util
    import os
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        path = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(path):
            C.CDLL(path, mode=C.RTLD_GLOBAL)
            return path
    return None


def nccl_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it and hands the bytes to the other ranks)."""
    preload_nccl()
    buf = C.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
    _lib.check(_lib.lib().p2g_nccl_unique_id(buf))
    return buf.raw


class NcclGroup(_Group):
    """The communicator lives inside libp2g (p2g_circuit_create_sharded_nccl): no callback, every exchange is an ncclAllGather on
    the handle's stream.  One instance per circuit handle (a communicator is created per handle); `unique_id` = the 128 bytes from
    nccl_unique_id() on rank 0."""
    in_library = True

    def __init__(self, rank, world, device, unique_id):
        if len(unique_id) != NCCL_UNIQUE_ID_BYTES:
            raise ValueError("unique_id must be the 128 bytes of an ncclUniqueId")
        preload_nccl()
        self.rank, self.world, self.device, self.unique_id = rank, world, device, bytes(unique_id)

    @classmethod
    def from_torch_dist(cls, device, group=None):
        """Rank 0 draws the id, torch.distributed (any backend) carries it to the other ranks: plumbing only."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        cuda = dist.get_backend(group) == "nccl"
        t = torch.zeros(NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device=f"cuda:{device}" if cuda else "cpu")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0, group=group)
        return cls(rank, world, device, bytes(t.cpu().numpy().tobytes()))


class ThreadGroup:
    """`world` ranks as threads of this process.  group.member(r) is rank r's binding; all ranks must make the same calls."""

    def __init__(self, world, devices=None):
        self.world = world
        self.devices = devices if devices is not None else [0] * world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def member(self, rank):
        g = self

        class Member(_Group):
            pass
        m = Member()
        m.rank, m.world, m.device = rank, self.world, self.devices[rank]

        def allgather(send, recv, nbytes, is_device):
            g.slots[rank] = (send, nbytes, is_device)
            g.barrier.wait()
            for p in range(g.world):
                ps, pn, pd = g.slots[p]
                assert pn == nbytes and pd == is_device, "ranks disagree on the exchange"
                if ps == recv + p * nbytes and p == rank:
                    continue   # in place
                if is_device:
                    import torch
                    dst = _tensor(recv + p * nbytes, nbytes, m.device)
                    dst.copy_(_tensor(ps, nbytes, g.devices[p]))
                else:
                    C.memmove(recv + p * nbytes, ps, nbytes)
            if is_device:
                import torch
                torch.cuda.synchronize(m.device)
            g.barrier.wait()
            return 0
        m.allgather = allgather
        return m

    def run(self, fn):
        """Runs fn(rank, member) on one thread per rank; returns the results in rank order, re-raising the first error."""
        out, err = [None] * self.world, []

        def work(r):
            try:
                out[r] = fn(r, self.member(r))
            except BaseException as e:   # noqa: BLE001
                err.append(e)
                self.barrier.abort()
        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        return out

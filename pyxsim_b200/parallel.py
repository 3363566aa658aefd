"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for the only collectives
the path has -- the scalar count reductions of pyxsim/photon_list.py:471-472,826.

Cells/particles shard over ranks with no data-path exchange (the reference deals yt chunks
round-robin over MPI ranks, photon_list.py:401, and writes one file per rank, :214-216);
each rank keeps its own photon/event shard ``{prefix}.{rank:04d}``.
"""
import os

import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def rank():
    return dist.get_rank() if is_initialized() else 0


def size():
    return dist.get_world_size() if is_initialized() else 1


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op when
    WORLD_SIZE is unset or 1).  Binds this process to ``cuda:LOCAL_RANK`` when CUDA exists."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.is_available():
        # more ranks than GPUs (tests on a one-GPU box): ranks share devices round-robin
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
    if ws <= 1 or is_initialized():
        return
    if backend is None:
        # (PXB_DIST_BACKEND=gloo lets several ranks share one GPU in tests; NCCL wants one each)
        backend = os.environ.get("PXB_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend)


def bind_host_to_gpu(device_index=None):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity), so that the pinned
    host buffers it allocates afterwards -- first touch -- and the threads that fill them sit on
    the GPU's NUMA node.  Best effort: returns what it found and did, never raises.  (On a
    virtualised host that reports one NUMA node for every GPU this changes nothing.)"""
    info = {"bound": False}
    try:
        import pynvml

        if not torch.cuda.is_available():
            return info
        if device_index is None:
            device_index = torch.cuda.current_device()
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{props.uuid}".encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        info["cpus_before"] = len(os.sched_getaffinity(0))
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()[-12:]
            try:
                info["numa_node"] = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
            except Exception:
                info["numa_node"] = None
        pynvml.nvmlDeviceSetCpuAffinity(h)
        info["cpus_after"] = len(os.sched_getaffinity(0))
        info["bound"] = True
    except Exception as e:  # no NVML, no permission, ...: run unbound
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info


def _device():
    if is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_device():
    """Device on which tensors handed to the collectives must live."""
    return _device()


def allreduce_sum(value):
    """Sum of an integer over ranks (comm.mpi_allreduce in the reference)."""
    if size() == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def allgather_counts(values):
    """All-gather a short list of int64 counters: returns a [size, len(values)] tensor on the
    host.  Used to derive global offsets (exclusive prefix over ranks) for concatenated output."""
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=_device())
    if size() == 1:
        return t.cpu().reshape(1, -1)
    out = [torch.empty_like(t) for _ in range(size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu()


def barrier():
    if size() > 1:
        dist.barrier()


def shard_bounds(n, block=256, r=None, s=None):
    """Block-cyclic ownership of ``n`` cells: rank r owns blocks r, r+s, r+2s ...  Returns the
    list of (start, stop) ranges; photon-rich core cells are spread over all ranks.  Small
    blocks on purpose: a block that spans many rows of a uniform grid turns into thick slabs, and
    ``s`` slabs per period can put a cluster core on two ranks (4096-cell blocks of a 256^3 grid
    on 4 GPUs: 37.7 ms per step instead of 26.4)."""
    r = rank() if r is None else r
    s = size() if s is None else s
    nblocks = (n + block - 1) // block
    return [(b * block, min((b + 1) * block, n)) for b in range(r, nblocks, s)]


def gather_arrays(arrays, dst=0):
    """Variable-length gather of same-length 1-D tensors from every rank to rank ``dst`` in rank
    order (the optional "final event-list concatenation"; by default lists stay sharded per rank
    like the reference's per-rank files).  Counts travel with one all-gather, payloads with
    point-to-point send/recv (NCCL over NVLink on GPUs, gloo in the CPU test).  Returns the
    concatenated dict on ``dst`` and None elsewhere."""
    keys = sorted(arrays)
    n_local = int(arrays[keys[0]].shape[0]) if keys else 0
    if size() == 1:
        return dict(arrays)
    counts = allgather_counts([n_local])[:, 0].tolist()
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + int(c))
    me = rank()
    if me == dst:
        out = {}
        for k in keys:
            t = arrays[k]
            buf = torch.empty(offs[-1], dtype=t.dtype, device=t.device)
            buf[offs[me]:offs[me + 1]] = t
            for r in range(size()):
                if r != dst and counts[r] > 0:
                    dist.recv(buf[offs[r]:offs[r + 1]], src=r)
            out[k] = buf
        return out
    for k in keys:
        if n_local > 0:
            dist.send(arrays[k].contiguous(), dst=dst)
    return None

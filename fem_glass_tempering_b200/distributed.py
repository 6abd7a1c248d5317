"""Element partition of the structured plates over the GPUs of one node.

One process per GPU (torch.distributed for rendezvous only).  The plate is cut into contiguous x-slabs of
hexahedron/quad columns; with the lattice numbering of mesh.py/fe.py (x slowest) every owned or ghost set is a
contiguous index range, so the ghost-dof forward scatter (TVP:351) is a handful of contiguous NCCL send/recv
pairs with the two slab neighbours — no pack/unpack kernels, no index lists.

  DG: one ghost column of cells on each side (the SIP facet terms need the neighbour cell's dofs).
  CG: one ghost column on the LEFT only: a rank owns the node planes [c0, c1) of its columns, and integrates
      every cell touching an owned node (its own columns plus the column left of plane c0), so the operator
      result on owned nodes is complete without a reverse scatter.  Plane c1 is a ghost owned by the right
      neighbour.
The viscoelastic update runs on owned + ghost nodes redundantly and needs no communication.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from . import mesh as _mesh


def env_ranks():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str | None = None):
    """torch.distributed rendezvous from RANK/WORLD_SIZE/MASTER_* (torchrun); no-op for one process."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_ranks()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa_node(device: int) -> dict:
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node and the
    node's cpulist), BEFORE any pinned host buffer is allocated: cudaHostAlloc'ed pages are first touched by this process,
    so they land in memory local to the GPU's root complex.  With all ranks on node 0 the per-step device->host output
    copies of 8 GPUs (16 GB per step) share one memory controller and one inter-socket link.  Returns what was done
    (bench.py prints it); silently does nothing where sysfs has no answer."""
    info = {"device": device, "numa_node": None, "cpus": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        info["pci"] = bdf
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(numa_node=node, cpus=len(cpus))
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)[:120]
    return info


def make_context(rank: int, world: int, device: int) -> _lib.Context:
    """Library context with its own NCCL communicator; the unique id travels through torch.distributed."""
    if world == 1:
        return _lib.Context(device)
    import torch
    import torch.distributed as dist
    uid = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    return _lib.Context(device, rank, world, uid[0])


def column_range(n_cols: int, rank: int, world: int):
    base, rem = divmod(n_cols, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def slab_partition(dim: int, n, lengths, family: str, degree: int, rank: int, world: int):
    """Local mesh (with ghost columns) + the partition dict ThermalOperator/ThermoViscoProblem take.

    n = global cell counts per axis, lengths = global plate size.  Returns (mesh, partition, info)."""
    assert dim in (2, 3), "slab partition is defined for the 2-D and 3-D plates"
    nx = n[0]
    assert nx >= world, "fewer cell columns than ranks"
    c0, c1 = column_range(nx, rank, world)
    gl = 1 if rank > 0 else 0
    gr = 1 if (rank < world - 1 and family == "DG") else 0
    xr = (c0 - gl, c1 + gr)
    if dim == 2:
        m = _mesh.rectangle_mesh(n[0], n[1], lengths[0], lengths[1], x_range=xr)
        col_cells = 2 * n[1]
        plane = (n[1] + 1) if degree == 1 else (2 * n[1] + 1)
    else:
        m = _mesh.box_mesh(n[0], n[1], n[2], lengths[0], lengths[1], lengths[2], x_range=xr)
        col_cells = 6 * n[1] * n[2]
        plane = (n[1] + 1) * (n[2] + 1) if degree == 1 else (2 * n[1] + 1) * (2 * n[2] + 1)
    m.comm = _mesh._Comm(rank, world)
    n_ld = {1: dim + 1, 2: (dim + 1) * (dim + 2) // 2}[degree]
    L = np.asarray(lengths, dtype=np.float64)

    def exterior_mask(mid):
        tol = 1e-9 * L.max()
        on = np.zeros(mid.shape[0], dtype=bool)
        for a in range(dim):
            on |= (np.abs(mid[:, a]) < tol) | (np.abs(mid[:, a] - L[a]) < tol)
        return on

    halo = []
    if family == "DG":
        cell_lo, cell_hi = gl * col_cells, (gl + c1 - c0) * col_cells
        own_lo, own_hi = cell_lo * n_ld, cell_hi * n_ld
        blk = col_cells * n_ld
        if rank > 0:
            halo.append((rank - 1, own_lo, blk, 0, blk))
        if rank < world - 1:
            halo.append((rank + 1, own_hi - blk, blk, own_hi, blk))
        owned_points = (c1 - c0) * col_cells * n_ld
        own_cells = (cell_lo, cell_hi)
        global_offset, global_n = c0 * blk, nx * blk
    else:
        gp = degree                                        # ghost planes on the left: 1 (P1) or 2 (P2)
        planes_owned = (c1 - c0) * degree + (1 if rank == world - 1 else 0)
        own_lo = gl * gp * plane
        own_hi = own_lo + planes_owned * plane
        cell_lo, cell_hi = 0, m.n_cells
        own_cells = (gl * col_cells, (gl + c1 - c0) * col_cells)
        if rank > 0:
            halo.append((rank - 1, own_lo, plane, 0, gp * plane))
        if rank < world - 1:
            halo.append((rank + 1, own_hi - gp * plane, gp * plane, own_hi, plane))
        owned_points = (c1 - c0) * col_cells * n_ld
        global_offset, global_n = c0 * degree * plane, (nx * degree + 1) * plane
    part = dict(cell_lo=cell_lo, cell_hi=cell_hi, own_lo=own_lo, own_hi=own_hi, exterior_mask=exterior_mask,
                halo=halo, own_cell_lo=own_cells[0], own_cell_hi=own_cells[1],
                global_own_offset=global_offset, global_n_dofs=global_n)    # where the owned range sits in the unpartitioned numbering
    info = dict(columns=(c0, c1), ghost_left=gl, ghost_right=gr, owned_cell_points=owned_points,
                owned_nodes=own_hi - own_lo)
    return m, part, info

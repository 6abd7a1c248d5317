"""Per-step field output of the time loop — what the reference does with four VTXWriters (T, phi, Tf, xi) and one
XDMFFile (sigma) on every step (ThermoViscoProblem.py:246-276, 357-364, 614-620).

ADIOS2/HDF5 are not available here, so the container format differs (one .npy per field and step + index.json +
the dof coordinates; in a partitioned run one such set PER RANK holding the rank's owned nodes, see RankFiles), but the schedule is the reference's: the five fields are captured after the viscoelastic
update and BEFORE T_prev <- T_cur (TVP:374-379).  A device-side snapshot decouples the compute stream
from PCIe: the device->host copies run on a side stream into pinned double buffers while the next step computes,
and the files are written by a background thread.  Enable with  problem.output_dir = "output"  before
problem.setup(); `problem.host_mirror = HostMirror(problem)` keeps the host copies without writing files.
"""
from __future__ import annotations

import json
import os
import queue
import threading

import numpy as np


class HostMirror:
    """Pinned host copies of the five output fields, refreshed once per step without stalling the compute stream.

    capture(): device->device snapshot into staging buffers on the compute stream (0.7 ms for 2 GB), then the
    device->host copies run on a side stream while the next step computes; T is copied first so that a host-side
    consumer of the temperature does not wait for the 9x larger stress.  Two host slots alternate."""
    FIELDS = ("T", "phi", "Tf", "xi", "sigma")            # TVP:357-361

    def __init__(self, problem, slots: int = 2):
        import torch
        self._torch = torch
        p = problem
        self.functions = {"T": p.functions_current["T"], "phi": p.functions["phi"], "Tf": p.functions_current["Tf"],
                          "xi": p.functions["xi"], "sigma": p.functions_next["sigma"]}
        self.FIELDS = HostMirror.FIELDS
        if getattr(p, "mechanics", None) is not None:       # extension: the accumulated displacement travels with the stress
            self.functions["displacement"] = p.functions["displacement"]
            self.FIELDS = HostMirror.FIELDS + ("displacement",)
        self._device = p._device
        self._stream = torch.cuda.Stream(device=self._device)
        self._staging = {k: torch.empty_like(f.x.array) for k, f in self.functions.items()}
        self._staged = torch.cuda.Event()
        self._drained = None                                   # last D2H that read the staging buffers
        self._slots = []
        for _ in range(max(2, slots)):
            bufs = {k: torch.empty(f.x.array.shape, dtype=torch.float64, pin_memory=True) for k, f in self.functions.items()}
            self._slots.append({"bufs": bufs, "events": {k: torch.cuda.Event() for k in bufs}, "free": threading.Event()})
            self._slots[-1]["free"].set()
        self._n = 0
        self.bytes_per_capture = sum(f.x.array.numel() * 8 for f in self.functions.values())

    def capture(self) -> int:
        torch = self._torch
        k = self._n % len(self._slots)
        slot = self._slots[k]
        slot["free"].wait()                                    # a consumer (writer thread) is done with this slot
        cur = torch.cuda.current_stream(self._device)
        if self._drained is not None:
            cur.wait_event(self._drained)                      # the previous snapshot has left the staging buffers
        for name, f in self.functions.items():
            self._staging[name].copy_(f.x.array, non_blocking=True)
        self._staged.record(cur)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(self._staged)
            for name in self.FIELDS:                           # T first
                slot["bufs"][name].copy_(self._staging[name], non_blocking=True)
                slot["events"][name].record(self._stream)
        self._drained = slot["events"][self.FIELDS[-1]]
        self._n += 1
        return k

    def field(self, slot: int, name: str):
        """Pinned host tensor of one field of a captured slot (blocks until its copy has landed)."""
        self._slots[slot]["events"][name].synchronize()
        return self._slots[slot]["bufs"][name]

    def wait(self, slot: int) -> dict:
        s = self._slots[slot]
        for ev in s["events"].values():
            ev.synchronize()
        return s["bufs"]

    def hold(self, slot: int) -> None:
        self._slots[slot]["free"].clear()

    def release(self, slot: int) -> None:
        self._slots[slot]["free"].set()


class RankFiles:
    """File layout of one rank's share of the output (no CUDA involved, so the CPU tests can exercise it).

    One process: `<field>_<step>.npy` + `index.json`, whole arrays.  Partitioned run (one process per GPU): every rank
    writes ONLY the nodes it owns — `<field>_<step>_r<rank>.npy`, `index_r<rank>.json`, `dof_coordinates_<field>_r<rank>.npy`
    — and records where that range sits in the unpartitioned numbering, so read_series() reassembles the global field.
    The reference writes one collective file set through mesh.comm (TVP:246-276); ranks never share a file here."""

    def __init__(self, directory: str, rank: int, size: int, fields: dict, owned: tuple | None = None, global_offset: int = 0,
                 global_n: int | None = None):
        """fields: key -> dict(block_size, name, n_nodes); owned = (lo, hi) local node range (None: everything)."""
        self.dir, self.rank, self.size = directory, int(rank), int(size)
        os.makedirs(directory, exist_ok=True)
        self.suffix = f"_r{self.rank}" if self.size > 1 else ""
        self.owned = owned
        self.index = {"rank": self.rank, "size": self.size, "fields": fields, "steps": [],
                      "owned_nodes": list(owned) if owned else None, "global_node_offset": int(global_offset),
                      "global_n_nodes": int(global_n) if global_n is not None else None}

    def rows(self, array: np.ndarray, block_size: int) -> np.ndarray:
        if self.owned is None:
            return array
        return array[self.owned[0] * block_size: self.owned[1] * block_size]

    def save_coordinates(self, key: str, xc: np.ndarray) -> None:
        np.save(os.path.join(self.dir, f"dof_coordinates_{key}{self.suffix}.npy"), xc if self.owned is None else xc[self.owned[0]:self.owned[1]])

    def save_step(self, n: int, t: float, arrays: dict) -> None:
        files = {}
        for k, a in arrays.items():
            name = f"{k}_{n:06d}{self.suffix}.npy"
            np.save(os.path.join(self.dir, name), self.rows(a, self.index["fields"][k]["block_size"]))
            files[k] = name
        self.index["steps"].append({"step": n, "t": t, "files": files})

    def close(self) -> None:
        with open(os.path.join(self.dir, f"index{self.suffix}.json"), "w") as fh:
            json.dump(self.index, fh, indent=1)


class FieldWriter:
    def __init__(self, directory: str, problem, slots: int = 2):
        self.dir = directory
        self._mirror = HostMirror(problem, slots)
        fn = self._mirror.functions
        self._n = 0
        comm, part = problem.mesh.comm, getattr(problem, "_partition", None)
        owned, offset, total = None, 0, None
        if part is not None and comm.size > 1:
            if not getattr(problem, "_same_space", True):
                raise NotImplementedError("per-rank output of a partitioned run needs fe_config['T'] == fe_config['sigma']")
            owned = (int(part["own_lo"]), int(part["own_hi"]))
            offset, total = int(part.get("global_own_offset", 0)), part.get("global_n_dofs")
        fields = {k: {"block_size": f.function_space.block_size, "name": f.name, "n_nodes": f.function_space.n_nodes}
                  for k, f in fn.items()}
        self._files = RankFiles(directory, comm.rank, comm.size, fields, owned, offset, total)
        for key in ("T", "sigma") + (("displacement",) if "displacement" in fn else ()):
            self._files.save_coordinates(key, fn[key].function_space.tabulate_dof_coordinates())
        self._q: queue.Queue = queue.Queue()
        self._err = None
        self._thread = threading.Thread(target=self._drain, daemon=True)
        self._thread.start()

    def write(self, t: float) -> None:
        if self._err is not None:
            raise self._err
        slot = self._mirror.capture()
        self._mirror.hold(slot)                              # until the files are on disk
        self._q.put((self._n, float(t), slot))
        self._n += 1

    def _drain(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            n, t, slot = item
            try:
                bufs = self._mirror.wait(slot)
                self._files.save_step(n, t, {k: buf.numpy() for k, buf in bufs.items()})
            except Exception as e:  # noqa: BLE001
                self._err = e
            finally:
                self._mirror.release(slot)

    def close(self) -> None:
        self._q.put(None)
        self._thread.join()
        self._files.close()
        if self._err is not None:
            raise self._err


def read_series(directory: str, field: str):
    """(times [n], values [n, n_nodes * block_size]) of one field written by FieldWriter; the per-rank files of a partitioned
    run are put back together in the unpartitioned node numbering."""
    single = os.path.join(directory, "index.json")
    if os.path.exists(single):
        paths = [single]
    else:
        import glob
        paths = sorted(glob.glob(os.path.join(directory, "index_r*.json")))
        if not paths:
            raise FileNotFoundError(f"no index.json / index_r*.json in {directory}")
    parts = []
    for path in paths:
        with open(path) as fh:
            index = json.load(fh)
        steps = sorted(index["steps"], key=lambda s: s["step"])
        parts.append((index.get("global_node_offset", 0), np.array([s["t"] for s in steps]),
                      np.stack([np.load(os.path.join(directory, s["files"][field])) for s in steps]), index))
    if len(parts) == 1:
        return parts[0][1], parts[0][2]
    parts.sort(key=lambda p: p[0])
    size = parts[0][3]["size"]
    assert len(parts) == size, f"{len(parts)} rank files found, the run had {size} ranks"
    t = parts[0][1]
    assert all(np.array_equal(t, p[1]) for p in parts), "ranks wrote different time stamps"
    bs = parts[0][3]["fields"][field]["block_size"]
    pos = 0
    for off, _, v, _ in parts:
        assert off * bs == pos, "owned ranges of the ranks do not tile the global numbering"
        pos += v.shape[1]
    total = parts[0][3].get("global_n_nodes")
    assert total is None or pos == total * bs
    return t, np.concatenate([p[2] for p in parts], axis=1)

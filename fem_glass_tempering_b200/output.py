"""Per-step field output of the time loop — what the reference does with four VTXWriters (T, phi, Tf, xi) and one
XDMFFile (sigma) on every step (ThermoViscoProblem.py:246-276, 357-364, 614-620).

ADIOS2/HDF5 are not available here, so the container format differs (one .npy per field and step + index.json +
the dof coordinates), but the schedule is the reference's: the five fields are captured after the viscoelastic
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


class FieldWriter:
    def __init__(self, directory: str, problem, slots: int = 2):
        self.dir = directory
        os.makedirs(directory, exist_ok=True)
        self._mirror = HostMirror(problem, slots)
        fn = self._mirror.functions
        self._n = 0
        self._index = {"fields": {k: {"block_size": f.function_space.block_size, "name": f.name} for k, f in fn.items()},
                       "steps": []}
        for key in ("T", "sigma"):
            np.save(os.path.join(directory, f"dof_coordinates_{key}.npy"), fn[key].function_space.tabulate_dof_coordinates())
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
                files = {}
                for k, buf in bufs.items():
                    name = f"{k}_{n:06d}.npy"
                    np.save(os.path.join(self.dir, name), buf.numpy())
                    files[k] = name
                self._index["steps"].append({"step": n, "t": t, "files": files})
            except Exception as e:  # noqa: BLE001
                self._err = e
            finally:
                self._mirror.release(slot)

    def close(self) -> None:
        self._q.put(None)
        self._thread.join()
        with open(os.path.join(self.dir, "index.json"), "w") as fh:
            json.dump(self._index, fh, indent=1)
        if self._err is not None:
            raise self._err


def read_series(directory: str, field: str):
    """(times [n], values [n, n_nodes * block_size]) of one field written by FieldWriter."""
    with open(os.path.join(directory, "index.json")) as fh:
        index = json.load(fh)
    steps = sorted(index["steps"], key=lambda s: s["step"])
    return (np.array([s["t"] for s in steps]),
            np.stack([np.load(os.path.join(directory, s["files"][field])) for s in steps]))

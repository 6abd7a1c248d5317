"""Per-step field output of the time loop — what the reference does with four VTXWriters (T, phi, Tf, xi) and one
XDMFFile (sigma) on every step (ThermoViscoProblem.py:246-276, 357-364, 614-620).

ADIOS2/HDF5 are not available here, so the container format differs (one .npy per field and step + index.json +
the dof coordinates), but the schedule is the reference's: the five fields are captured after the viscoelastic
update and BEFORE T_prev <- T_cur (TVP:374-379).  The device->host copies run on a side stream into
pinned double buffers and the files are written by a background thread, so the compute stream never waits for
the disk.  Enable with  problem.output_dir = "output"  before  problem.setup().
"""
from __future__ import annotations

import json
import os
import queue
import threading

import numpy as np


class FieldWriter:
    FIELDS = ("T", "phi", "Tf", "xi", "sigma")            # TVP:357-361

    def __init__(self, directory: str, problem, slots: int = 2):
        import torch
        self._torch = torch
        self.dir = directory
        os.makedirs(directory, exist_ok=True)
        p = problem
        self._fn = {"T": p.functions_current["T"], "phi": p.functions["phi"], "Tf": p.functions_current["Tf"],
                    "xi": p.functions["xi"], "sigma": p.functions_next["sigma"]}
        self._device = p._device
        self._stream = torch.cuda.Stream(device=self._device)
        self._slots = []
        for _ in range(max(2, slots)):
            bufs = {k: torch.empty(f.x.array.shape, dtype=torch.float64, pin_memory=True) for k, f in self._fn.items()}
            self._slots.append({"bufs": bufs, "copied": torch.cuda.Event(), "free": threading.Event()})
            self._slots[-1]["free"].set()
        self._n = 0
        self._index = {"fields": {k: {"block_size": f.function_space.block_size, "name": f.name}
                                  for k, f in self._fn.items()}, "steps": []}
        for key in ("T", "sigma"):
            np.save(os.path.join(directory, f"dof_coordinates_{key}.npy"), self._fn[key].function_space.tabulate_dof_coordinates())
        self._q: queue.Queue = queue.Queue()
        self._err = None
        self._thread = threading.Thread(target=self._drain, daemon=True)
        self._thread.start()

    def write(self, t: float) -> None:
        torch = self._torch
        if self._err is not None:
            raise self._err
        slot = self._slots[self._n % len(self._slots)]
        slot["free"].wait()                                  # its previous contents are on disk
        slot["free"].clear()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self._device))
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)                   # the fields of this step are complete
            for k, f in self._fn.items():
                slot["bufs"][k].copy_(f.x.array, non_blocking=True)
            slot["copied"].record(self._stream)
        # the compute stream must not overwrite the fields before the copy has read them
        torch.cuda.current_stream(self._device).wait_event(slot["copied"])
        self._q.put((self._n, float(t), slot))
        self._n += 1

    def _drain(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            n, t, slot = item
            try:
                slot["copied"].synchronize()
                files = {}
                for k, buf in slot["bufs"].items():
                    name = f"{k}_{n:06d}.npy"
                    np.save(os.path.join(self.dir, name), buf.numpy())
                    files[k] = name
                self._index["steps"].append({"step": n, "t": t, "files": files})
            except Exception as e:  # noqa: BLE001
                self._err = e
            finally:
                slot["free"].set()

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

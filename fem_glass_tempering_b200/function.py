"""Minimal stand-ins for dolfinx.fem.{FunctionSpace, Function, Constant}: the reference-facing objects
that own the device arrays of the hot path (ThermoViscoProblem.py:61-173).

A Function's `.x.array` is a flat float64 torch CUDA tensor in dolfinx's blocked layout
`array[node*bs + comp]`.  PyTorch is buffer ownership only; the arithmetic happens in libsurroglas_b200.
"""
from __future__ import annotations

import numpy as np

from . import fe


class Constant:
    """dolfinx.fem.Constant: `.value` is a numpy array (scalar or 1-D)."""

    def __init__(self, mesh, value):
        self.mesh = mesh
        self.value = np.asarray(value, dtype=np.float64)

    def __float__(self):
        return float(self.value)

    def __len__(self):
        return len(self.value)

    def __iter__(self):
        return iter(self.value.tolist())

    def __getitem__(self, i):
        return float(self.value[i])


class FiniteElementInfo:
    """What the reference keeps in self.finiteElements[...] (family/degree/value shape)."""

    def __init__(self, family: str, cell: str, degree: int, shape: tuple):
        self._family, self.cell, self._degree, self._shape = family, cell, degree, tuple(shape)

    def family(self) -> str:
        return "Lagrange" if self._family == "CG" else "Discontinuous Lagrange"   # TVP:284,308

    def degree(self) -> int:
        return self._degree

    def value_shape(self) -> tuple:
        return self._shape


class _ElementView:
    def __init__(self, scalar: fe.ScalarSpace):
        self._el = scalar.element

    def interpolation_points(self) -> np.ndarray:
        return self._el.interpolation_points()


class FunctionSpace:
    """A scalar node set (fe.ScalarSpace) times a value shape."""

    def __init__(self, mesh, scalar: fe.ScalarSpace, shape: tuple = ()):
        self.mesh, self.scalar, self.shape = mesh, scalar, tuple(shape)
        self.block_size = int(np.prod(shape)) if shape else 1
        self.element = _ElementView(scalar)

    @property
    def n_nodes(self) -> int:
        return self.scalar.n_nodes

    def tabulate_dof_coordinates(self) -> np.ndarray:
        return self.scalar.tabulate_dof_coordinates()


class _Vector:
    def __init__(self, fn):
        self._fn = fn

    @property
    def array(self):
        if self._fn._array is None:
            raise RuntimeError(f"Function '{self._fn.name}' is not materialised (ThermoViscoProblem(materialize='minimal'))")
        return self._fn._array

    def scatter_forward(self) -> None:
        """Owner -> ghost update (TVP:351)."""
        if self._fn._scatter is not None and self._fn._array is not None:
            self._fn._scatter(self._fn._array, self._fn.function_space.block_size)


class Function:
    def __init__(self, V: FunctionSpace, name: str | None = None, *, device=None, allocate: bool = True, alias=None,
                 scatter=None):
        import torch
        self.function_space, self.name = V, name or "f"
        self._scatter = scatter
        if alias is not None:
            self._array = alias._array
        elif allocate:
            self._array = torch.zeros(V.n_nodes * V.block_size, dtype=torch.float64, device=device)
        else:
            self._array = None
        self.x = _Vector(self)

    @property
    def materialised(self) -> bool:
        return self._array is not None

    def interpolate(self, u) -> None:
        """interpolate(callable): u(x) with x of shape (3, n_nodes) like dolfinx; returns (n,) or (bs, n).
        interpolate(PointwiseExpression): evaluate one reference Expression (slow, reference-shaped path)."""
        import torch
        if hasattr(u, "evaluate_into"):
            u.evaluate_into(self)
            return
        V = self.function_space
        xc = V.tabulate_dof_coordinates()
        x3 = np.zeros((3, xc.shape[0]))
        x3[: xc.shape[1]] = xc.T
        vals = np.asarray(u(x3), dtype=np.float64)
        vals = vals.reshape(-1, xc.shape[0])                     # (bs, n)
        assert vals.shape[0] == V.block_size, "callable returned the wrong value size"
        self.x.array.copy_(torch.from_numpy(np.ascontiguousarray(vals.T).ravel()))

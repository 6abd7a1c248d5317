"""One small run of hot path (B) on cuda:0 for __graft_entry__.smoke(): two implicit-Euler steps of the 3-D DG1 plate
(class-table operator, Chebyshev-preconditioned CG, inexact Newton) and of a 2-D CG2 plate, checked against the
assembled CPU oracle (test infrastructure, imported here only as the checker)."""
import numpy as np


def run() -> None:
    import torch
    from . import _lib, fe
    from . import mesh as msh
    from .thermal_op import ThermalOperator
    from oracle import thermal_oracle as to
    from oracle.visco_oracle import MAIN_PARAMS

    ctx = _lib.Context(0)
    for dim, family, degree, cheb in ((3, "DG", 1, 3), (2, "CG", 2, 0)):
        m = msh.box_mesh(6, 6, 3, 6.0, 6.0, 3.0) if dim == 3 else msh.rectangle_mesh(9, 5, 9.0, 5.0)
        space = fe.ScalarSpace(m, family, degree)
        op = ThermalOperator(ctx, space, MAIN_PARAMS, 0.1, cheb_degree=cheb)
        orc = to.ThermalOracle(m.x, m.cells, space.dofmap, space.element.nodes, family, degree, MAIN_PARAMS, 0.1)
        T_o = np.full(space.n_nodes, 800.0)
        T_d = torch.from_numpy(T_o.copy()).to("cuda:0")
        Tp_d = T_d.clone()
        for _ in range(2):
            T_o, _, ok = orc.newton(T_o, T_o.copy())
            st = op.timestep(T_d, Tp_d)
            assert ok and st.converged == 1
            Tp_d.copy_(T_d)
        err = float(np.max(np.abs(T_d.cpu().numpy() - T_o)) / np.max(np.abs(T_o)))
        assert err <= 1e-10, (family, degree, err)
        info = op.class_info()
        print(f"smoke thermal: {family}{degree} d={dim} n_dofs={space.n_nodes} class tables={info['active']} "
              f"chebyshev={op.chebyshev_info()['degree']} newton={st.newton_its} pcg={st.lin_its} rel err T={err:.1e}")
        op.close()
    ctx.close()

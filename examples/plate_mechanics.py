"""A 3-D plate quench with the optional equilibrium step (model_params["mechanics"]) and the corrected
physics: the same API as main.py, plus the displacement and the equilibrated stress profile through the thickness.
Run from the repository root on a machine with a B200:  python examples/plate_mechanics.py [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_glass_tempering_b200 import ThermoViscoProblem          # noqa: E402
from fem_glass_tempering_b200 import mesh as msh                 # noqa: E402

n = (48, 48, 8)                                                   # cubes of 1 mm, six tetrahedra each
mesh = msh.plate_mesh(3, n, tuple(float(k) for k in n))
fe_config = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
model_params = {
    "f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "T_0": 800.0, "alpha": 1.0, "htc": 280.1,
    "rho": 2500.0, "cp": 1433.0, "k": 1.0, "H": 627.8e3, "Tb": 869.0e0, "Rg": 8.314,
    "alpha_solid": 9.10e-6, "alpha_liquid": 25.10e-6, "Tf_init": 873.0,
    "sip_penalty": 6.0,                                           # the reference's 5.0 is not coercive on tetrahedra
    "physics": "corrected",                                       # history = partial stress, structural strain, exact exponentials
    "mechanics": {"fixed": "symmetry", "rtol": 1e-8},             # held on the planes x = 0, y = 0, z = 0: an eighth model
}
model = ThermoViscoProblem(mesh_path="", config=fe_config, time=(0.0, 50.0), dt=0.1, model_parameters=model_params, mesh=mesh,
                           verbose=False, materialize="minimal")
model.n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
model.setup(dirichlet_bc=False)
model.solve()

u = model.functions["displacement"].x.array.view(-1, 3)
sig = model.functions_next["sigma"].x.array.view(-1, 3, 3)
xc = model.functionSpaces["sigma"].tabulate_dof_coordinates()
import numpy as np                                                # noqa: E402
mid = (np.abs(xc[:, 0] - 4.0) < 4.0) & (np.abs(xc[:, 1] - 4.0) < 4.0)      # a column near the symmetry axis
z, sxx = xc[mid, 2], sig[:, 0, 0].cpu().numpy()[mid]
print(f"{model.n_steps} steps, last equilibrium solve: {model.mechanics.last_iters} PCG iterations, "
      f"max |u| = {float(u.abs().max()):.4e} mm")
for zz in sorted(set(np.round(z, 6))):
    print(f"  z = {zz:4.1f} mm   mean sigma_xx = {sxx[np.round(z, 6) == zz].mean(): .4e}")

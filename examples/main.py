"""The reference's driver script (main.py) against the B200 package: same dict configuration, same three calls.
Run from the repository root on a machine with a B200:  python examples/main.py [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_glass_tempering_b200 import ThermoViscoProblem, create_mesh  # noqa: E402

jit_options = {"cffi_extra_compile_args": ["-O3", "-march=native"]}   # accepted for API parity, unused
time = (0.0, 50.0)
dt = 0.1
mesh_path = "mesh1d.msh"
if not os.path.exists(mesh_path):
    create_mesh(path=mesh_path)

fe_config = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "CG", "degree": 1}}
model_params = {
    "f": 0.0, "epsilon": 0.93, "sigma": 5.670e-8, "T_ambient": 600.0, "T_0": 800.0, "alpha": 1.0, "htc": 280.1,
    "rho": 2500.0, "cp": 1433.0, "k": 1.0, "H": 627.8e3, "Tb": 869.0e0, "Rg": 8.314,
    "alpha_solid": 9.10e-6, "alpha_liquid": 25.10e-6, "Tf_init": 873.0,
}

model = ThermoViscoProblem(mesh_path=mesh_path, config=fe_config, time=time, dt=dt,
                           model_parameters=model_params, jit_options=jit_options, verbose=False)
if len(sys.argv) > 1:
    model.n_steps = int(sys.argv[1])
model.setup(dirichlet_bc=False)
model.solve()
T = model.functions_current["T"].x.array
print(f"T surface = {float(T[0]):.6f} K, T centre = {float(T[T.numel() // 2]):.6f} K, "
      f"max |sigma| = {float(model.functions_next['sigma'].x.array.nan_to_num().abs().max()):.6e}")

"""Long-run stability check of the 3-D DG plate (48x48x8 mm, 1 mm cubes): python tools/small_diag.py <sip_penalty> [steps]
Prints Newton/PCG counts and the temperature range per step.  With the reference's penalty 5.0 the range explodes after
~10 steps (non-coercive SIP form on tetrahedra); with 6.0 it decays monotonically towards T_ambient."""
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from fem_glass_tempering_b200 import ThermoViscoProblem  # noqa: E402
from fem_glass_tempering_b200 import mesh as msh  # noqa: E402

pen = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
mesh = msh.plate_mesh(3, (48, 48, 8), (48.0, 48.0, 8.0))
cfg = {"T": {"element": "DG", "degree": 1}, "sigma": {"element": "DG", "degree": 1}}
prob = ThermoViscoProblem(mesh_path="", time=(0.0, 50.0), dt=0.1, config=cfg, model_parameters=dict(bench.MAIN_PARAMS, sip_penalty=pen),
                          mesh=mesh, materialize="minimal", verbose=False)
prob.setup(dirichlet_bc=False)
print("sip_penalty", pen, "chebyshev", prob._thermal_op.chebyshev_info())
for i in range(steps):
    try:
        prob.solve_timestep(t=0.0)
    except AssertionError as e:
        print("step", i, "FAILED", str(e)[:160])
        break
    st = prob.solver.last_stats
    T = prob.functions_current["T"].x.array
    if i % 5 == 4 or i < 3:
        print(f"step {i:3d}: newton {st.newton_its} pcg {st.lin_its:3d}  T in [{float(T.min()):.3f}, {float(T.max()):.3f}]")

/*
 * surroglas_b200.h — C ABI of the B200-native SurroGlas per-timestep hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces a call the reference
 * (pzimbrod/fem-glass-tempering) makes into dolfinx/PETSc; the reference site is
 * cited on each declaration as  TVP = ThermoViscoProblem.py, VM =
 * ViscoelasticModel.py, TM = ThermalModel.py.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (SG_E_*); the message is
 *     available from sg_last_error() (thread-local).  Nothing throws across the ABI.
 *   - every data pointer is a DEVICE pointer owned by the caller (PyTorch); the
 *     library never frees caller memory.  float64 everywhere, int32 indices.
 *   - arrays use the dolfinx blocked layout array[node*bs + comp] (TVP:82-101).
 *   - `stream` is a cudaStream_t passed as void*; all compute calls are asynchronous
 *     on it unless stated otherwise.
 *   - plans/operators are opaque handles created/destroyed by paired calls; the
 *     library keeps no global mutable state.
 */
#ifndef SURROGLAS_B200_H
#define SURROGLAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_MAX_TERMS 16

enum {
    SG_OK = 0,
    SG_E_INVALID = -1,   /* bad argument */
    SG_E_CUDA = -2,      /* CUDA runtime error */
    SG_E_NOCONV = -3,    /* solver did not converge (TVP:390 assert(converged)) */
    SG_E_NCCL = -4,
    SG_E_UNSUPPORTED = -5
};

int sg_version(void);
/* kernels launched by this library in this process so far (bench.py: gpu_launches) */
int64_t sg_launch_count(void);
const char *sg_last_error(void);

/* ------------------------------------------------------------------ context */
typedef struct sg_ctx sg_ctx;

/* One context per process/GPU (the reference is one MPI rank per process,
 * TVP:27-28 MPI.COMM_WORLD).  nccl_unique_id: 128-byte ncclUniqueId obtained from
 * sg_nccl_unique_id() on rank 0 and broadcast by the host (torch.distributed), or
 * NULL when nranks == 1. */
int sg_nccl_unique_id(void *out128);
int sg_ctx_create(int device, int rank, int nranks, const void *nccl_unique_id, sg_ctx **out);
int sg_ctx_destroy(sg_ctx *ctx);
int sg_ctx_sm_count(const sg_ctx *ctx);

/* ------------------------------------------------- (A) viscoelastic update */

/* VM:9-84: Prony tableaux and material constants; dt from VM:88. */
typedef struct {
    int32_t dim;      /* VM:17 */
    int32_t n_terms;  /* VM:16 tableau_size (<= SG_MAX_TERMS) */
    double H, Rg, Tb; /* VM:75-79 */
    double alpha_solid, alpha_liquid; /* VM:81-83 */
    double dt;
    double m[SG_MAX_TERMS], lambda_m[SG_MAX_TERMS];   /* VM:19-34 */
    double g[SG_MAX_TERMS], lambda_g[SG_MAX_TERMS];   /* VM:35-50 */
    double k[SG_MAX_TERMS], lambda_k[SG_MAX_TERMS];   /* VM:51-68 */
    int32_t mode;     /* SG_VISCO_REFERENCE (0): the reference's expressions exactly as executed, quirks included;
                         SG_VISCO_CORRECTED (1): the scheme its comments cite (csrc/visco.cu), fused call only */
    double chi;       /* VM:15; only the corrected scheme reads it */
} sg_visco_params;

enum { SG_VISCO_REFERENCE = 0, SG_VISCO_CORRECTED = 1 };

/* Phases = the four reference methods; a phase only controls which outputs are
 * WRITTEN, every input it needs is recomputed from T_cur/T_prev/Tf. */
enum {
    SG_PHASE_TF = 1,      /* TVP:393-407  _solve_Tf            */
    SG_PHASE_STRAIN = 2,  /* TVP:409-423  _solve_strains       */
    SG_PHASE_SHIFT = 4,   /* TVP:426-435  _solve_shifted_time  */
    SG_PHASE_STRESS = 8,  /* TVP:438-452  _solve_stress        */
    SG_PHASE_ALL = 15
};

/* The dolfinx Functions of TVP:106-173 the chain reads/writes.  Pointers in the
 * "optional" block may be NULL: the quantity is then kept in registers/shared
 * memory and never written to HBM. */
typedef struct {
    /* T space (bs 1 unless noted) */
    const double *T_cur;    /* functions_current["T"]  */
    const double *T_prev;   /* functions_previous["T"] */
    double *Tf_partial;     /* bs N; in/out: previous on entry, current on exit (TVP:466-470) */
    double *Tf;             /* functions_current["Tf"] == functions_previous["Tf"] (TVP:480-482) */
    double *phi;            /* functions["phi"] */
    double *xi;             /* functions["xi"]  */
    /* sigma space */
    double *s_tilde;        /* bs N*d*d; in/out (TVP:552,559) */
    double *sigma_tilde;    /* bs N*d*d; in/out (TVP:571,578) */
    double *sigma;          /* bs d*d;   functions_next["sigma"] (TVP:591) */
    /* optional, full-materialisation mode */
    double *T_next, *phi_next;                                   /* TVP:524,533 */
    double *thermal_strain, *total_strain, *deviatoric_strain;   /* bs d*d; TVP:492,504,516 */
    double *ds_partial, *dsigma_partial;                         /* bs N*d*d; TVP:549,568 */
    double *s_partial, *sigma_partial;                           /* bs N*d*d; TVP:555,574 */
} sg_visco_fields;

/* Cross-space evaluation (fe_config["T"] != fe_config["sigma"], e.g. main.py:24-27
 * DG1 -> CG1): value of a T-space function at sigma node s is
 *   sum_j weights[local_point[s]*n_ld + j] * array[dofs[s*n_ld + j]]
 * where (cell, local_point) is the LAST cell that dolfinx's interpolation visits
 * for that node (last writer wins), and zero weights are skipped like FFCx's
 * table compression does. */
typedef struct {
    int32_t n_ld;                /* dofs per cell of the T element */
    int32_t n_points;            /* interpolation points per cell of the sigma element */
    const int32_t *dofs;         /* [n_sigma_nodes * n_ld] T-space dof indices of the winner cell */
    const uint8_t *local_point;  /* [n_sigma_nodes] */
    const double *weights;       /* [n_points * n_ld] T basis at the sigma points */
} sg_visco_gather;

typedef struct sg_visco_plan sg_visco_plan;
int sg_visco_plan_create(sg_ctx *ctx, const sg_visco_params *params, sg_visco_plan **out);
int sg_visco_plan_destroy(sg_visco_plan *plan);

/* Replaces the 17 Function.interpolate(Expression) calls + 7 copies of
 * TVP:370-373 (definitions VM:111-228) when the T and sigma spaces share their
 * node set: one fused pass over n_nodes. */
int sg_visco_update(sg_visco_plan *plan, int64_t n_nodes, const sg_visco_fields *f,
                    uint32_t phases, void *stream);

/* Cross-space variants: T-space quantities (phi, Tf_partial, Tf, T_next,
 * phi_next, xi) over the T nodes ... */
int sg_visco_update_scalar(sg_visco_plan *plan, int64_t n_T_nodes, const sg_visco_fields *f,
                           uint32_t phases, void *stream);
/* ... and sigma-space quantities over the sigma nodes, reading T_cur, T_prev, Tf
 * and xi through `gather`. */
int sg_visco_update_tensor(sg_visco_plan *plan, int64_t n_sigma_nodes, const sg_visco_fields *f,
                           const sg_visco_gather *gather, uint32_t phases, void *stream);

/* Algorithmic HBM bytes per node of sg_visco_update for the given materialisation
 * (used by bench.py for the roofline). */
int64_t sg_visco_bytes_per_node(const sg_visco_params *params, const sg_visco_fields *f, uint32_t phases);


/* --------------------------------------------- (B) heat-equation operator + solve */

/* Mesh, element tables and physics of the thermal problem.  Replaces
 * NonlinearProblem(F, u) (TVP:331) built from the weak form TVP:293-325 with the
 * constants of ThermalModel (TM:18-27).  Index/geometry arrays are DEVICE pointers
 * that must outlive the operator; the reference tables are HOST pointers and are
 * copied.  Local cells [cell_lo, cell_hi) are the ones this rank integrates over;
 * dofs [own_lo, own_hi) are the ones it owns (reductions run over them). */
typedef struct {
    int32_t dim, degree;
    int32_t family;                 /* 0 = CG ('Lagrange'), 1 = DG ('Discontinuous Lagrange'), TVP:284,308 */
    int64_t n_cells, cell_lo, cell_hi;
    int64_t n_dofs, own_lo, own_hi;
    const int32_t *dofmap;          /* CG: [n_ld][n_cells] (SoA); DG: NULL (dof = cell*n_ld + i) */
    const double *geom;             /* [dim*dim + 2][n_cells]: Jinv[a][c] = d xi_a/d x_c, |detJ|, CellDiameter h */
    const int32_t *nbr;             /* DG: [dim+1][n_cells] neighbour across local facet f, -1 = none */
    const int32_t *nbinfo;          /* DG: [n_cells], 5 bits per facet: nb_facet | perm_id << 2 */
    int64_t n_bfacets;              /* exterior facets (ds) */
    const int32_t *bf_cell, *bf_facet;
    const double *bf_area;
    /* reference tables (host), see fem_glass_tempering_b200/fe.py: operator_tables() */
    int32_t n_ld, nqc, nqf, nqb, n_perm;
    const double *mass, *load, *cq_w, *cq_grad;
    const double *fq_w, *fq_val, *fq_grad;
    const int32_t *fq_perm;
    const double *bq_w, *bq_val;
    /* physics: main.py:29-55 / TM:18-27; penalty = 5.0 (TVP:313) */
    double dt, alpha, f, sigma, epsilon, htc, T_ambient, penalty;
    /* Cells whose element contribution x_K . A_K x_K enters the fused x.Ax reduction of the solver: every
     * global cell must be in this range on exactly one rank (CG slabs integrate a ghost column too).
     * own_cell_lo == own_cell_hi means [cell_lo, cell_hi). */
    int64_t own_cell_lo, own_cell_hi;
    int32_t flags;                  /* SG_THERMAL_* */
} sg_thermal_desc;

enum {
    SG_THERMAL_NO_CLASSES = 1,      /* never use the local-matrix class tables (general per-cell-geometry kernel only) */
    SG_THERMAL_GENERAL_RESIDUAL = 2,/* evaluate the residual with the per-cell-geometry kernel even when the tables exist */
    SG_THERMAL_NO_STENCIL = 8       /* CG spaces: never use the row-stencil form of the Jacobian apply (see
                                       sg_thermal_stencil_info); the cell-centric class kernel scatters instead */
};

typedef struct sg_thermal_op sg_thermal_op;
int sg_thermal_op_create(sg_ctx *ctx, const sg_thermal_desc *desc, sg_thermal_op **out);
int sg_thermal_op_destroy(sg_thermal_op *op);

/* F(T; v) of TVP:293-325 (what NonlinearProblem.F assembles). */
int sg_thermal_residual(sg_thermal_op *op, const double *T, const double *T_prev, double *F, void *stream);
/* y = J(T_lin) x, J = dF/dT (what NonlinearProblem.J assembles + PETSc MatMult). */
int sg_thermal_jac_apply(sg_thermal_op *op, const double *T_lin, const double *x, double *y, void *stream);
/* diag J(T_lin) (Jacobi preconditioner; the reference uses GAMG, TVP:344). */
int sg_thermal_jac_diag(sg_thermal_op *op, const double *T_lin, double *diag, void *stream);
/* Local-matrix classes found at creation (cells with equal shape and neighbourhood share one
 * precomputed element matrix; sg_thermal_jac_apply then reads no geometry).  Returns 1 when the class
 * tables are in use, 0 when the mesh has too many classes and the general kernel runs, <0 on error. */
int sg_thermal_class_info(const sg_thermal_op *op, int32_t *n_geometry, int32_t *n_self, int32_t *n_facet);
/* CG spaces on a mesh whose rows repeat (lattice-numbered plates): the Jacobian apply runs in gather form, one 16-bit
 * class id per row and the class's (column offset, coefficient) list in shared memory - what PETSc's assembled MatMult
 * does with a CSR matrix (TVP:340-346), without the matrix.  Returns 1 and the table sizes when that form is in use,
 * 0 when the cell-centric kernels run, <0 on error. */
int sg_thermal_stencil_info(const sg_thermal_op *op, int32_t *n_classes, int32_t *n_entries, int32_t *max_nnz);
/* Optional timing of the Jacobian-apply cell kernel with CUDA event pairs on its launch stream
 * (measurement harness only; at most `capacity` launches are recorded after each enable). */
int sg_thermal_profile(sg_thermal_op *op, int32_t enable, int32_t capacity);
int sg_thermal_profile_read(sg_thermal_op *op, int64_t *n_launches, double *ms_total);
/* kind 0 = Jacobian-apply cell kernel, 1 = fused Chebyshev step of the DG solver; launches that returned
 * immediately (solve already converged) are excluded. */
int sg_thermal_profile_read_kind(sg_thermal_op *op, int32_t kind, int64_t *n_launches, double *ms_total);
int64_t sg_thermal_cheb_step_bytes(const sg_thermal_op *op, int32_t first);
/* Algorithmic HBM bytes of one sg_thermal_jac_apply (roofline numerator). */
int64_t sg_thermal_apply_bytes(const sg_thermal_op *op);

/* Ghost-dof forward scatter (TVP:351 x.scatter_forward()).  The element partition is
 * by x-slabs with lattice numbering, so every exchange is a contiguous range. */
typedef struct {
    int32_t peer;
    int64_t send_offset, send_count;   /* in nodes; multiplied by block_size */
    int64_t recv_offset, recv_count;
} sg_halo_segment;
typedef struct sg_halo_plan sg_halo_plan;
int sg_halo_plan_create(sg_ctx *ctx, int32_t n_segments, const sg_halo_segment *segments, sg_halo_plan **out);
int sg_halo_plan_destroy(sg_halo_plan *plan);
int sg_halo_forward(sg_halo_plan *plan, double *vec, int32_t block_size, void *stream);
/* Optional NVLink peer-memory transport for block_size 1 exchanges and the solver's 1-2 double all-reduces (one
 * process per GPU, all on one NVSwitch domain; replaces PETSc's MPI halo + MPI_Allreduce inside KSP cg, TVP:339-346):
 * every rank allocates ONE IPC-exported block holding flags, all-reduce slots, mailboxes and - when workspace_doubles > 0
 * - the solver's vector workspace.  Vectors inside that workspace are exchanged by storing the boundary rows straight
 * into the neighbour's ghost rows (one kernel, no receive-side copy; the consuming operator kernel waits for the flag
 * before its boundary strips only), any other vector goes through the mailboxes; the all-reduces run inside the last
 * block of the reducing kernels.  No NCCL launch on the data path.  sg_halo_peer_alloc returns 1 and a 64-byte
 * cudaIpcMemHandle_t when the plan qualifies (0 otherwise, <0 on error); the host gathers the handles of ALL ranks
 * (rank order, 64 bytes each) and, per rank, three int64 {vector stride = n_dofs, recv_offset of the rows arriving from
 * the rank below, recv_offset of the rows arriving from above} and passes both tables to sg_halo_peer_open, or NULLs
 * if any rank returned 0.  sg_halo_peer_workspace returns the workspace to hand to sg_thermal_solver_create (NULL: none).
 * Waits are bounded: a neighbour that never arrives turns into SG_E_NCCL at the next host synchronisation, not a hang. */
int sg_halo_peer_alloc(sg_halo_plan *plan, void *handle64, int64_t workspace_doubles);
int sg_halo_peer_open(sg_halo_plan *plan, const void *handles, const int64_t *layout3);
double *sg_halo_peer_workspace(const sg_halo_plan *plan);
int sg_halo_uses_peer_memory(const sg_halo_plan *plan);

/* Newton + Jacobi-PCG time-step solver (NewtonSolver.solve, TVP:334-346,389). */
typedef struct {
    double newton_rtol;   /* 1e-12, TVP:336 */
    double newton_atol;   /* 1e-10, dolfinx default */
    int32_t newton_max_it;/* 50, dolfinx default */
    double lin_rtol;      /* final linear-residual target of the step, relative to |F(T_0)| */
    double lin_atol;      /* absolute floor of that target */
    int32_t lin_max_it;
    double forcing_eta;   /* > 0: inexact Newton, first forcing term eta_1 (Eisenstat-Walker choice 2 afterwards);
                             0: every PCG solve runs to the final target (PETSc-like fixed tolerance) */
} sg_newton_opts;

typedef struct {
    int32_t newton_its, lin_its, converged;
    double dx_norm_first, dx_norm_last, lin_rel_res_last;
} sg_newton_stats;

typedef struct sg_thermal_solver sg_thermal_solver;
int64_t sg_thermal_solver_workspace_doubles(const sg_thermal_op *op);
int sg_thermal_solver_create(sg_thermal_op *op, double *workspace, sg_halo_plan *halo, sg_thermal_solver **out);
int sg_thermal_solver_destroy(sg_thermal_solver *s);
/* DG spaces with the class tables in use: precondition CG with the degree-`degree` Chebyshev polynomial in
 * M^-1 J (M = element mass blocks) instead of M^-1 alone: one CG iteration then does degree + 1 operator
 * applications (fused with the polynomial recurrence, J z never goes to memory) but only one set of CG vector
 * updates and reductions.  [lo, hi] bounds the spectrum of M^-1 J; hi <= 0: estimated at first use from the Ritz
 * values of 24 Lanczos (CG) steps (hi = 1.05 theta_max).  Returns 1 when enabled, 0 when not available (CG spaces,
 * general kernel) or degree == 0.  If the polynomial turns out not to be positive definite the interval is widened
 * (x1.3, twice) and the solve retried; after that the solver falls back to M^-1 for good. */
int sg_thermal_solver_set_chebyshev(sg_thermal_solver *s, int32_t degree, double lo, double hi);
int sg_thermal_solver_get_chebyshev(const sg_thermal_solver *s, int32_t *degree, double *lo, double *hi);
/* Solve J(T_lin) x = b with Jacobi-PCG from x = 0; blocks until converged (KSP 'cg', TVP:343). */
int sg_pcg_solve(sg_thermal_solver *s, const double *T_lin, const double *b, double *x, double rtol, double atol,
                 int32_t max_it, int32_t *iters, double *rel_res, void *stream);
/* One implicit-Euler step: Newton on F(T) = 0 starting from T (in/out), TVP:384-391.  Returns
 * SG_E_NOCONV when the incremental criterion is not met (the reference asserts, TVP:390). */
int sg_thermal_timestep(sg_thermal_solver *s, double *T, const double *T_prev, const sg_newton_opts *opts,
                        sg_newton_stats *stats, void *stream);

/* ------------------------------------------ (C) mechanical equilibrium (extension)
 *
 * The reference has no equilibrium solve: VM:135-139 sets total_strain = -thermal_strain (a fully restrained body).
 * Behind model_parameters["mechanics"] the framework solves, once per time step, for the displacement increment du in
 * the vector-P1 space on the mesh vertices for which the stresses of the reference's own Prony chain (VM:176-228),
 * evaluated with total_strain = eps(du) - thermal_strain, are in weak equilibrium.  The chain is linear in the strain:
 *     sigma = sigma0 + 2 G_eff dev(eps(du)) + K_eff tr(eps(du)) I        at every sigma node,
 * sigma0 = what sg_visco_update wrote (the reference's stress), G_eff/K_eff = the chain's own per-node factors summed over
 * the Prony terms.  Matrix-free Jacobi-PCG on the GPU (csrc/mech.cu); eps(du) is constant per cell and a sigma node takes
 * the strain of winner_cell[node] (the last cell touching it: the rule dolfinx's interpolate applies to cell-wise
 * discontinuous expressions, TVP:456-591).  Single GPU. */
typedef struct {
    int32_t dim;
    int64_t n_vertices, n_cells;
    const double *coords;         /* device [n_vertices, dim]   mesh.geometry.x */
    const int32_t *cells;         /* device [n_cells, dim+1]    vertex ids */
    const uint8_t *fixed;         /* device [n_vertices*dim]    1 = this displacement component is held at zero */
    int32_t n_ld_sigma;           /* nodes per cell of the sigma space */
    int64_t n_sigma_nodes;
    const int32_t *sigma_dofmap;  /* device [n_cells, n_ld_sigma] */
    const double *sigma_weights;  /* HOST   [n_ld_sigma]: int phi_l / |K| of the sigma element */
    const int32_t *winner_cell;   /* device [n_sigma_nodes] */
} sg_mech_desc;

/* what sg_mech_correct adds eps(du) to; sigma and mech_strain are required, the rest may be NULL */
typedef struct {
    double *sigma;                /* bs d*d, in/out: sigma0 on entry */
    double *mech_strain;          /* bs d*d, out: eps(du) at the sigma nodes */
    double *total_strain, *deviatoric_strain;                   /* bs d*d, in/out (VM:135-146) */
    double *ds_partial, *dsigma_partial;                        /* bs N*d*d, in/out (VM:176-191) */
    double *s_partial, *sigma_partial;                          /* bs N*d*d, in/out (VM:212-221) */
    double *s_tilde, *sigma_tilde;  /* bs N*d*d, in/out; SG_VISCO_CORRECTED only (there the history IS the partial stress) */
} sg_mech_fields;

typedef struct sg_mech_op sg_mech_op;
int sg_mech_op_create(sg_ctx *ctx, const sg_mech_desc *desc, sg_mech_op **out);
int sg_mech_op_destroy(sg_mech_op *op);
/* G_eff, K_eff [n_nodes] from xi at the sigma nodes, with the factors of VM:176-191 (or of the corrected scheme) */
int sg_mech_coefficients(sg_visco_plan *plan, int64_t n_nodes, const double *xi_sigma, double *G_eff, double *K_eff,
                         void *stream);
/* cell means of the nodal moduli; must precede sg_mech_apply / sg_mech_solve whenever G_eff/K_eff changed */
int sg_mech_set_moduli(sg_mech_op *op, const double *G_eff, const double *K_eff, void *stream);
/* y = (P A P + I - P) x: the tangent stiffness with the held components replaced by the identity */
int sg_mech_apply(sg_mech_op *op, const double *x, double *y, void *stream);
/* b = -P B^T sigma0 (the out-of-balance force of the restrained stress state) */
int sg_mech_rhs(sg_mech_op *op, const double *sigma0, double *b, void *stream);
/* Jacobi-PCG for A du = b(sigma0) starting from the du passed in (e.g. the previous step's increment); blocks until
 * |r| <= max(rtol |b|, atol) or max_it (-> SG_E_NOCONV) */
int sg_mech_solve(sg_mech_op *op, const double *sigma0, double *du, double rtol, double atol, int32_t max_it,
                  int32_t *iters, double *rel_res, void *stream);
int sg_mech_correct(sg_mech_op *op, sg_visco_plan *plan, const double *du, const double *xi_sigma, const double *G_eff,
                    const double *K_eff, const sg_mech_fields *f, void *stream);
/* algorithmic bytes of one sg_mech_apply (cell ids, gradients, moduli, gathers and scatters counted once per cell) */
int64_t sg_mech_apply_bytes(const sg_mech_op *op);

#ifdef __cplusplus
}
#endif
#endif /* SURROGLAS_B200_H */

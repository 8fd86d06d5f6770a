/* cutfemx_b200.h -- C ABI of libcutfemx_b200: the sm_100a cut-cell hot path of CutFEMx.
 *
 * This is the drop-in boundary.  Every entry point replaces one
 * seam of the reference's host code; the reference file:line it stands in for is cited
 * on each declaration (paths relative to the CutFEMx source tree).  The reference's C++
 * host library and nanobind layer stay as they are and call these functions with the flat
 * arrays they already hold (DOLFINx host layouts); INTEGRATION.md shows the edits.
 *
 * Conventions
 *  - plain C types only; all arrays are caller-owned; `memspace` says where a pointer
 *    lives (CFX_HOST: the library copies to / from the GPU; CFX_DEVICE: borrowed as is).
 *  - every function returns 0 on success or a negative cfx_status; the message is
 *    available from cfx_last_error().  The host patch rethrows it as std::runtime_error,
 *    which is what the reference throws at the same places (e.g. cut.cpp:97-106).
 *  - one context per (process, GPU); calls on one context are not re-entrant (the
 *    reference is single-threaded per MPI rank, SURVEY.md section 8b "Threading").
 *  - there is NO CPU fallback: without a CUDA device cfx_ctx_create fails.
 */
#ifndef CUTFEMX_B200_H
#define CUTFEMX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfx_ctx cfx_ctx;
typedef struct cfx_list cfx_list;       /* device-resident int32 array (entity ids / facet rows) */
typedef struct cfx_rules cfx_rules;     /* device-resident cutcells::quadrature::QuadratureRules */
typedef struct cfx_pattern cfx_pattern; /* device-resident CSR pattern + values (la::MatrixCSR) */
typedef struct cfx_form cfx_form;       /* integral table of one Form (Form.h:46-89)            */
typedef int cfx_status;

enum { CFX_OK = 0, CFX_ERR_INVALID = -1, CFX_ERR_CUDA = -2, CFX_ERR_STATE = -3, CFX_ERR_RANGE = -4,
       CFX_ERR_UNSUPPORTED = -5 };
enum { CFX_HOST = 0, CFX_DEVICE = 1 };
/* cell types: value == number of vertices (mesh/convert.h:14-38 maps the DOLFINx enum) */
enum { CFX_TRIANGLE = 3, CFX_TETRAHEDRON = 4 };
/* cutcells::cell::domain codes written by cfx_update (cut.cpp:292-321) */
enum { CFX_DOMAIN_UNSET = 0, CFX_DOMAIN_INSIDE = 1, CFX_DOMAIN_INTERSECTED = 2, CFX_DOMAIN_OUTSIDE = 3 };
/* cutcells::Relation (cut.cpp:323-342) */
enum { CFX_REL_LT = 0, CFX_REL_LE = 1, CFX_REL_GT = 2, CFX_REL_GE = 3, CFX_REL_EQ = 4 };
enum { CFX_MAX_LEVEL_SETS = 4, CFX_MAX_SPACES = 4, CFX_MAX_CLAUSES = 32, CFX_MAX_CONSTANTS = 8 };

/* Hand-written element-kernel families (the role of the runintgen/FFCx generated
 * tabulate_tensor functions called at assemble_matrix_impl.h:142,:323,:534 and
 * assemble_vector_impl.h:103,:211,:333; forms: python/demo/demo_poisson.py:183-201). */
enum {
  CFX_K_LAPLACE = 1,        /* rank 2, cells:           c0 * inner(grad u, grad v)                                  */
  CFX_K_MASS = 2,           /* rank 2, cells:           c0 * u * v                                                  */
  CFX_K_NITSCHE = 3,        /* rank 2, interface rules: -(grad u.n) v - (grad v.n) u + c0/h u v   (needs normals)   */
  CFX_K_GHOST_GRAD_JUMP = 4,/* rank 2, interior facets: c0 * avg(h) * jump(grad u,n) * jump(grad v,n)               */
  CFX_K_SOURCE = 5,         /* rank 1, cells:           c0 * v                                                      */
  CFX_K_NITSCHE_RHS = 6,    /* rank 1, interface rules: -(grad v.n) c1 + c0/h c1 v                (needs normals)   */
  CFX_K_ONE = 7,            /* rank 0, cells / rules:   c0  (volume, area, perimeter)                               */
  /* blocked (vector) Lagrange spaces, block size == gdim (demo_elasticity.py:213-238): */
  CFX_K_ELASTICITY = 8,     /* rank 2, cells:           inner(sigma(u), eps(v)), sigma = 2 c0 eps + c1 tr(eps) I        */
  CFX_K_SOURCE_VEC = 9,     /* rank 1, cells:           inner((c0, c1, c2), v)                                         */
  CFX_K_NITSCHE_VEC = 11,   /* rank 2, interface rules on a blocked space: -(sigma(u) n).v - (sigma(v) n).u
                               + c2 (2 c0 + c1)/h u.v, sigma = 2 c0 eps + c1 tr(eps) I  (needs normals)             */
  CFX_K_SQUARE_FN = 10      /* rank 0, cells / rules:   c0 * w^2, w a Function of the form's (scalar) space -- the error
                               functional (uh - u_exact)**2 of demo_poisson.py:213 with w = uh - I(u_exact); the dof
                               values are restricted to each entity's cell like pack_coefficients does
                               (pack_form.h:68-158), set with cfx_form_set_coefficient                                 */
  /* CFX_K_GHOST_GRAD_JUMP on a blocked space acts on every component: c0 avg(h) inner(jump(grad u, n), jump(grad v, n)) */
};

/* ------------------------------------------------------------------ context */
/* `stream` is a cudaStream_t (NULL = legacy default stream). */
cfx_status cfx_ctx_create(int device, void* stream, cfx_ctx** out);
void cfx_ctx_destroy(cfx_ctx* ctx);
const char* cfx_last_error(const cfx_ctx* ctx); /* ctx may be NULL: last error of the calling thread */
cfx_status cfx_sync(cfx_ctx* ctx);
int cfx_version(void);
/* number of kernel launches issued through this context since creation (bench.py "gpu_launches"); a graph launch
 * counts the kernel nodes of the graph */
int64_t cfx_launch_count(const cfx_ctx* ctx);
/* bytes of device memory the context holds (mesh mirrors, static tables, result buffers, cached blocks) */
int64_t cfx_device_bytes(const cfx_ctx* ctx);
/* diagnostics of a bound space: [0] rows the fast gather paths could not handle, [1] static rows without a contribution
 * list (largest counts in the eager assemblies so far, -1 = none yet), [2], [3] learned capacities of the active-row
 * and band-row lists.  A deferred-size step takes two launch decisions from [0] and [1] and verifies them on the
 * device. */
cfx_status cfx_space_counters(const cfx_ctx* ctx, int space, int64_t out[4]);
/* forget them (after an untypical assembly such as the all-facets pattern behind a static exchange plan) */
cfx_status cfx_space_forget(cfx_ctx* ctx, int space);

/* ------------------------------------------------------------------ deferred sizes and CUDA graphs
 * The reference returns every intermediate result to the host (numpy arrays: locate_entities cut.cpp:877-924,
 * runtime_quadrature :1311-1335, ghost_penalty_facets cut.py:340-380), so a time step of demo_moving_poisson.py:69-107
 * is a chain of host round trips.  In DEFERRED-SIZE mode the calls of that chain -- cfx_update, cfx_locate_entities,
 * cfx_runtime_quadrature, cfx_evaluate_normals (out = NULL), cfx_ghost_penalty_facets, cfx_facet_integration_rows_list,
 * cfx_form_add_*_integral_list, cfx_create_sparsity, cfx_assemble_matrix / _vector / _system -- leave the sizes of
 * their results on the device when they are handed an object to REUSE (lists, rules, matrices from an earlier step,
 * whose buffers then act as capacities, `margin` = extra fraction allocated beyond what a step needed): nothing
 * synchronises, and the first query of a size (cfx_list_size, cfx_rules_sizes, cfx_pattern_sizes) or any fetch makes
 * the sizes known to the host again.  A result that does not fit its capacity raises an error at that point (or at
 * cfx_check) instead of being truncated silently; repeat the step with deferred mode off to grow the buffers.
 * Every other entry point may be called in either mode; it synchronises as before. */
cfx_status cfx_set_deferred(cfx_ctx* ctx, int on, double margin /* < 0: keep the current one (default 0.125) */);
/* read the device-side error flag (synchronises): capacity exceeded, invalid entity index, entry not in pattern */
cfx_status cfx_check(cfx_ctx* ctx);
/* Capture the calls between begin and end into a CUDA graph instead of executing them (deferred-size mode only, after
 * the same calls have run once on the same objects; no size query / fetch in between), then replay the whole step
 * with one launch per time step: the level-set values are re-read from where they were bound, every object the
 * captured calls produced is refreshed in place.  The objects must stay alive, and must not be passed to other
 * calls that would reallocate them, while the graph exists. */
typedef struct cfx_graph cfx_graph;
cfx_status cfx_graph_begin(cfx_ctx* ctx);
cfx_status cfx_graph_end(cfx_ctx* ctx, cfx_graph** out);
cfx_status cfx_graph_launch(cfx_ctx* ctx, cfx_graph* g);
int64_t cfx_graph_kernel_nodes(const cfx_graph* g);
void cfx_graph_free(cfx_ctx* ctx, cfx_graph* g);
/* Lanes: independent call sequences of one step -- in python/demo/demo_poisson.py:156-175 the volume rules, the
 * interface rules with their normals, the ghost-penalty facets and the cell lists all depend on cutfemx.update only
 * -- may be issued on streams of their own: the calls between cfx_lane_begin(ctx, k) (k = 1..3) and cfx_lane_end are
 * ordered after everything issued on the main stream so far and run concurrently with the other lanes and with
 * later main-stream work; cfx_lane_join orders the main stream after every lane (call it before anything consumes a
 * lane's results).  Inside a graph capture the lanes become parallel branches of the graph.  Each lane has its own
 * stream, free-block cache and scan scratch; lanes do not nest.  Results are bit-identical with or without lanes. */
cfx_status cfx_lane_begin(cfx_ctx* ctx, int lane);
cfx_status cfx_lane_end(cfx_ctx* ctx);
cfx_status cfx_lane_join(cfx_ctx* ctx);
/* on = 0: every lane is the main stream (the library's internal forks -- static | band | inactive rows of the pattern
 * and of the matrix gather -- included): one kernel at a time, for per-kernel timing.  Default on (off when the
 * environment has CFX_NO_LANES). */
cfx_status cfx_set_lanes(cfx_ctx* ctx, int on);

/* ------------------------------------------------------------------ mesh views
 * replaces build_mesh_view, cut.cpp:500-538: x is geometry().x() (stride 3),
 * x_dofmap the geometry dofmap (n_cells_total rows of nv int32).  Only the first
 * n_cells_owned cells are cut (cell index_map size_local(), cut.cpp:507,532). */
cfx_status cfx_mesh_bind(cfx_ctx* ctx, const double* x, int64_t n_nodes, const int32_t* x_dofmap,
                         int64_t n_cells_owned, int64_t n_cells_total, int cell_type, int gdim, int memspace);
/* topology()->connectivity(tdim,tdim-1) as (n_cells_total, tdim+1) int32 and, optionally,
 * connectivity(tdim-1,tdim) as an AdjacencyList (offsets n_facets+1, data); when f2c is NULL
 * it is derived on the device.  Used by cut.py:340-380, cut.cpp:926-994, wrappers/cut.cpp:54-115. */
cfx_status cfx_topology_bind(cfx_ctx* ctx, const int32_t* c2f, const int32_t* f2c_offsets, const int32_t* f2c,
                             int64_t n_facets, int64_t n_owned_facets, int memspace);

/* ------------------------------------------------------------------ level sets
 * replaces build_level_set_function, cut.cpp:593-636 (dofmap span + dof_values span).
 * dofmap == NULL means "the geometry dofmap" (P1 on vertex numbering). With memspace
 * CFX_HOST and pin_host != 0 the value array is page-locked for the life of the binding
 * so that cfx_update re-reads it at full PCIe rate (it is the array cut.cpp:854-855 re-binds). */
cfx_status cfx_levelset_bind(cfx_ctx* ctx, int ls, const int32_t* dofmap, int nd, int degree, const double* values,
                             int64_t n_dofs, int memspace, int pin_host);
/* Forget level set `ls`: the caller's value array (and its page-locking) is no longer referenced.  A CutData that
 * goes away calls this for its level sets -- the reference's CutData owns its level-set view (cut.cpp:742-786). */
cfx_status cfx_levelset_unbind(cfx_ctx* ctx, int ls);
/* cutfemx::update, cut.cpp:845-868 -> cutcells::cut: re-read every level set's values and
 * classify all owned cells (cut.cpp:292-321: all dofs < 0 inside, all > 0 outside, else intersected). */
/* cutfemx.cut(level_set, entities, entity_dim = tdim) (python/cutfemx/cut.py:186-249, cut.cpp:500-538): restrict
 * the hosts of the cut to the listed OWNED cells; every other cell is in no domain (no selector matches it, it
 * gets no rule).  cells == NULL restores "all cells".  Takes effect at the next cfx_update.  Facet hosts
 * (entity_dim < tdim, cut.cpp:540-591) are not on the accelerated path yet. */
cfx_status cfx_set_host_cells(cfx_ctx* ctx, const int32_t* cells, int64_t n, int memspace);
cfx_status cfx_update(cfx_ctx* ctx);
cfx_status cfx_counts(cfx_ctx* ctx, int ls, int64_t counts[3]); /* inside, intersected, outside */
cfx_status cfx_domain_fetch(cfx_ctx* ctx, int ls, int8_t* out, int memspace);

/* ------------------------------------------------------------------ selectors
 * cutfemx::locate_entities, cut.cpp:877-924.  The selector is the compiled DNF of
 * cutcells::SelectionExpr: term t is the AND of clauses term_offsets[t]..term_offsets[t+1],
 * terms are OR-ed; clause k is "level set clause_ls[k]  clause_rel[k]  0".
 * Result: ascending owned cell ids.  If *out is non-NULL the list object is reused. */
cfx_status cfx_locate_entities(cfx_ctx* ctx, int n_terms, const int32_t* term_offsets, const int32_t* clause_ls,
                               const int32_t* clause_rel, cfx_list** out);
int64_t cfx_list_size(const cfx_list* l);
const int32_t* cfx_list_device_ptr(const cfx_list* l);
cfx_status cfx_list_fetch(cfx_ctx* ctx, const cfx_list* l, int32_t* out, int memspace);
void cfx_list_free(cfx_ctx* ctx, cfx_list* l);

/* ------------------------------------------------------------------ run-time quadrature
 * cutfemx::runtime_quadrature, cut.cpp:1311-1335 -> cutcells::select_part +
 * cutcells::output::quadrature_rules, for a single-clause selector "ls rel 0".
 * Points are parent-cell reference coordinates, weights are physical.
 * If *inout is non-NULL its buffers are reused. */
cfx_status cfx_runtime_quadrature(cfx_ctx* ctx, int ls, int relation, int order, cfx_rules** inout);
cfx_status cfx_rules_sizes(const cfx_rules* r, int64_t* npts, int64_t* nrules, int* tdim);
/* export in the reference's container layout (runtime_quadrature.h:107-137,
 * wrappers/cut.cpp:185-226): points AoS (npts,tdim), weights, offsets (nrules+1), parent_map. */
cfx_status cfx_rules_fetch(cfx_ctx* ctx, const cfx_rules* r, double* points_aos, double* weights, int32_t* offsets,
                           int32_t* parent_map, int memspace);
/* RuntimeSurfaceProvenance (runtime_quadrature.h:30-43), filled by make_surface_provenance (cut.cpp:1273-1308) for
 * straight codimension-one rules of a single "ls = 0" selector: per rule the index of its cut entity among the cut
 * entities, the parent entity, the local id of the zero entity inside the cut entity and the entity's dimension.
 * *level_set_index = -1 (arrays untouched) for any other selector -- the reference's empty provenance.  The arrays hold
 * nrules entries each; pass NULL arrays to query the level-set index only. */
cfx_status cfx_rules_surface_provenance(cfx_ctx* ctx, const cfx_rules* r, int32_t* level_set_index,
                                        int32_t* cut_cell_ids, int32_t* parent_cell_ids,
                                        int32_t* local_zero_entity_ids, int32_t* dimensions, int memspace);
/* RuntimeQuadrature::physical_points, runtime_quadrature.h:102-221: SoA (gdim, npts). */
cfx_status cfx_rules_physical_points(cfx_ctx* ctx, const cfx_rules* r, double* out_soa, int memspace);
void cfx_rules_free(cfx_ctx* ctx, cfx_rules* r);
/* sub-simplex rule tables: query the built-in one / override it (e.g. with basix::make_quadrature) */
cfx_status cfx_simplex_rule(int dim, int order, int* npts, double* points, double* weights, int capacity);
cfx_status cfx_set_simplex_rule(cfx_ctx* ctx, int dim, int order, int npts, const double* points,
                                const double* weights);

/* cutfemx::level_set::evaluate_normals, level_set/normal.h:39-188 (floor 1e-14 at :176-177):
 * sign * grad(phi)/|grad(phi)| at the rules' points, AoS (npts, gdim) float64.  The result is
 * also kept with the rules for the CFX_K_NITSCHE* kernels. out may be NULL. */
cfx_status cfx_evaluate_normals(cfx_ctx* ctx, int ls, cfx_rules* r, double sign, double* out_aos, int memspace);
/* cutfemx::level_set::evaluate_values, level_set/value.h:34-119 */
cfx_status cfx_evaluate_values(cfx_ctx* ctx, int ls, const cfx_rules* r, double* out, int memspace);

/* ------------------------------------------------------------------ facets as hosts of the cut
 * cutfemx.cut(level_set, facets, entity_dim = tdim - 1): build_entity_mesh_view (cut.cpp:540-591) +
 * build_entity_level_sets (cut.cpp:1022-1063, fem/entity_dofmap.cpp:11-88).  The listed facets are classified by the
 * level-set values at their vertices (every bound P1 level set, current values); locate returns FACET ids in the
 * order of the list (host_parent_index, cut.cpp:344-359); the rules have tdim = mesh tdim - 1, points in the facet's
 * reference coordinates (vertices in ascending vertex number), physical weights and parent_map = facet ids
 * (test_cut_api.py:171-188, :349-367, :424-501).  cfx_rules_physical_points / _fetch / _sizes work on them; they
 * cannot be attached to cell integrals.  Needs cfx_topology_bind.  Relation "=" gives the interface inside each cut
 * facet: a segment with its physical length (triangle facets) or the cut point with weight 1 (segment facets). */
typedef struct cfx_ecut cfx_ecut;
cfx_status cfx_cut_facets(cfx_ctx* ctx, const int32_t* facets, int64_t n, int memspace, cfx_ecut** inout);
cfx_status cfx_ecut_locate(cfx_ctx* ctx, const cfx_ecut* e, int n_terms, const int32_t* term_offsets,
                           const int32_t* clause_ls, const int32_t* clause_rel, cfx_list** out);
cfx_status cfx_ecut_runtime_quadrature(cfx_ctx* ctx, const cfx_ecut* e, int ls, int relation, int order,
                                       cfx_rules** inout);
/* Exterior / boundary facet integrals with run-time rules (`... * ds(subdomain_data=rules)`): the reference maps the
 * facet rules to (cell, local facet) rows and to the parent cell's reference coordinates before the generated kernel
 * runs (_facet_payload_with_rows + facet_runtime_quadrature_payload, python/cutfemx/_runintgen_adapter.py:605-680).
 * This call returns, as ordinary cell-hosted rules, the rules of `facet_rules` whose facet is local facet
 * `local_facet` of its first cell: points in the cell's reference coordinates, physical weights, parent_map = cells
 * (each cell at most once per local facet index, so tdim + 1 calls cover every rule).  The result feeds
 * cfx_form_add_cell_integral with any point-evaluated kernel family (CFX_K_MASS, CFX_K_SOURCE, CFX_K_ONE, ...). */
cfx_status cfx_rules_facets_to_cells(cfx_ctx* ctx, const cfx_rules* facet_rules, int local_facet, cfx_rules** inout);
void cfx_ecut_free(cfx_ctx* ctx, cfx_ecut* e);

/* ------------------------------------------------------------------ ghost-penalty facets
 * cutfemx.ghost_penalty_facets, python/cutfemx/cut.py:340-380: owned interior facets of the
 * cells intersected by level set `cut_ls` whose two cells are both in (cut cells U selector). */
cfx_status cfx_ghost_penalty_facets(cfx_ctx* ctx, int cut_ls, int n_terms, const int32_t* term_offsets,
                                    const int32_t* clause_ls, const int32_t* clause_rel, int include_ghosts,
                                    cfx_list** out);
/* cutfemx::interior_facets_for_cells, cut.cpp:926-994 */
cfx_status cfx_interior_facets_for_cells(cfx_ctx* ctx, const int32_t* cells, int64_t n, int memspace,
                                         int include_ghosts, cfx_list** out);
/* facet_integration_rows("interior_facet"), wrappers/cut.cpp:54-115: (cell0, lf0, cell1, lf1) per facet */
cfx_status cfx_facet_integration_rows(cfx_ctx* ctx, const int32_t* facets, int64_t n, int memspace, cfx_list** out);
/* the same for a facet list that lives on the device (e.g. the result of cfx_ghost_penalty_facets), whose length
 * may be deferred */
cfx_status cfx_facet_integration_rows_list(cfx_ctx* ctx, const cfx_list* facets, cfx_list** out);

/* ------------------------------------------------------------------ function spaces
 * dolfinx::fem::DofMap of a Lagrange space of `degree` (1|2) on the bound mesh:
 * dofmap (n_cells_total, nd) int32 (scalar dofs), block size bs (1, or gdim for vector spaces: matrices
 * then hold row-major bs x bs blocks per pattern entry, vectors bs values per dof -- wrappers/fem.cpp:330-385,
 * assemble_vector_impl.h:109-120), index_map size_local / +num_ghosts. */
cfx_status cfx_space_bind(cfx_ctx* ctx, int space, const int32_t* dofmap, int nd, int bs, int degree,
                          int64_t n_dofs_owned, int64_t n_dofs_total, int memspace);

/* ------------------------------------------------------------------ forms (Form.h:46-89,119-677)
 * One integral = (kernel family, entity list, constants, run-time rules as custom_data), the
 * tuple python/cutfemx/fem.py:346-351 hands to create_form_*.  A cell integral over the mixed
 * measure subdomain_data=[cells, rules] (demo_poisson.py:165-167) passes both `cells`
 * (standard compile-time quadrature) and `rules` (loop index -> rule). */
cfx_status cfx_form_create(cfx_ctx* ctx, int space, int rank, cfx_form** out);
cfx_status cfx_form_add_cell_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* cells, int64_t n_cells,
                                      int memspace, cfx_rules* rules, const double* constants, int n_constants);
cfx_status cfx_form_add_interior_facet_integral(cfx_ctx* ctx, cfx_form* f, int kernel, const int32_t* rows4,
                                                int64_t n_facets, int memspace, const double* constants,
                                                int n_constants);
/* The same with the entity lists given as library lists (results of cfx_locate_entities / cfx_facet_integration_rows*),
 * borrowed until the form is freed; their lengths may be deferred. */
cfx_status cfx_form_add_cell_integral_list(cfx_ctx* ctx, cfx_form* f, int kernel, const cfx_list* cells, cfx_rules* rules,
                                           const double* constants, int n_constants);
cfx_status cfx_form_add_interior_facet_integral_list(cfx_ctx* ctx, cfx_form* f, int kernel, const cfx_list* rows4,
                                                     const double* constants, int n_constants);
/* Exterior-facet integral over facet-hosted run-time rules (ufl.Measure("ds", subdomain_data=rules) on rules from a
 * facet-hosted cut, test_cut_api.py:504-527).  Shipped family: CFX_K_ONE on a rank-0 form -- c0 times the measure of
 * the selected part of the host facets. */
cfx_status cfx_form_add_exterior_facet_integral(cfx_ctx* ctx, cfx_form* f, int kernel, cfx_rules* rules,
                                                const double* constants, int n_constants);
/* The ordinary Function coefficient of a form (pack_form.h:30-158 allocate_coefficient_storage / pack_coefficients):
 * values over the owned+ghost dofs of the form's space.  The kernels gather each entity's cstride = nd values
 * through the dofmap while they run, so no packed (n_entities, cstride) array is materialised.  The array is
 * copied (HOST) or borrowed until the next call (DEVICE). */
cfx_status cfx_form_set_coefficient(cfx_ctx* ctx, cfx_form* f, const double* values, int64_t n, int memspace);
void cfx_form_free(cfx_ctx* ctx, cfx_form* f);

/* create_sparsity_pattern + finalize + MatrixCSR(sp): assembler.h:442-592, wrappers/fem.cpp:266-276.
 * Pattern = union of cell cliques and facet macro cliques of the form's domains, plus the
 * diagonal of every owned+ghost row (insert_deactivation_diagonal, assembler.h:538-560). */
cfx_status cfx_create_sparsity(cfx_ctx* ctx, const cfx_form* a, cfx_pattern** inout);
/* la::SparsityPattern::insert(rows, cols) for entries that do not come from this rank's integration
 * domains: on a partitioned mesh SparsityPattern::finalize() sends the entries of ghost rows to the
 * owning rank, which merges them into its own rows (DOLFINx 0.11, reached from assembler.h:567-592
 * create_sparsity_pattern -> wrappers/fem.cpp:266-276).  `rows` ascending, (row, col) local indices;
 * a column >= the space's owned+ghost dof count is a NEW ghost column of the matrix (finalize()
 * extends the column index map the same way).  Used by the next cfx_create_sparsity of this form. */
cfx_status cfx_form_insert_pattern_entries(cfx_ctx* ctx, cfx_form* f, const int32_t* rows, const int32_t* cols,
                                           int64_t n, int memspace);
/* The part of the pattern finalize() ships to other ranks: rows >= row_begin only (the ghost rows,
 * numbered after the owned ones), no deactivation diagonal; other rows are left empty. */
cfx_status cfx_create_sparsity_rows(cfx_ctx* ctx, const cfx_form* a, int64_t row_begin, cfx_pattern** inout);
/* position of entry (rows[i], cols[i]) in the values array (MatrixCSR keeps the same map for
 * scatter_rev); DEVICE arrays; fails if an entry is not in the pattern */
cfx_status cfx_pattern_positions(cfx_ctx* ctx, const cfx_pattern* p, const int32_t* rows, const int32_t* cols,
                                 int64_t n, int64_t* positions);
/* adopt a pattern built elsewhere (la::MatrixCSR row_ptr int64 / cols int32, sorted per row) */
cfx_status cfx_pattern_import(cfx_ctx* ctx, int space, const int64_t* row_ptr, const int32_t* cols, int64_t n_rows,
                              int memspace, cfx_pattern** out);
cfx_status cfx_pattern_sizes(const cfx_pattern* p, int64_t* n_rows, int64_t* nnz);
int cfx_pattern_block_size(const cfx_pattern* p); /* values hold nnz * bs * bs doubles */
cfx_status cfx_pattern_fetch(cfx_ctx* ctx, const cfx_pattern* p, int64_t* row_ptr, int32_t* cols, int memspace);
const double* cfx_pattern_values_device_ptr(const cfx_pattern* p);
const int64_t* cfx_pattern_row_ptr_device_ptr(const cfx_pattern* p);
const int32_t* cfx_pattern_cols_device_ptr(const cfx_pattern* p);
cfx_status cfx_pattern_values_fetch(cfx_ctx* ctx, const cfx_pattern* p, double* values, int memspace);
void cfx_pattern_free(cfx_ctx* ctx, cfx_pattern* p);

/* assemble_matrix: assembler.h:596-703 -> assemble_matrix_impl.h:629-810 (cells :68-189, interior
 * facets :409-607) with mat_set = MatrixCSR::mat_add_values (wrappers/fem.cpp:340-385).  Like the
 * reference it ADDS into the matrix unless zero_first != 0.  Deterministic: every CSR entry is
 * accumulated by one thread in a fixed order (no floating-point atomics).
 * diag_inactive != 0 additionally writes that value on the diagonal of rows no active entity
 * touches (deactivate_outside, fem/deactivate.h:402-418).  values_out may be NULL. */
cfx_status cfx_assemble_matrix(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, int zero_first, double diag_inactive,
                               double* values_out, int memspace);
/* assemble_matrix + assemble_vector of one problem in one pass (demo_poisson.py:51-54 calls them back to
 * back on forms over the same measures): where the two forms share their integration domains the row
 * owners fill b while they walk the incidence for A, so the second sweep disappears; otherwise the call
 * is the two calls above.  b: DEVICE vector of n_dofs_total entries.  Results are bit-identical to the
 * separate calls. */
cfx_status cfx_assemble_system(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, int zero_first_A, double diag_inactive,
                               const cfx_form* L, double* b, int zero_first_b);
/* assemble_vector: assemble_vector_impl.h:62-122,259-364,573-767; b has n_dofs_total*bs entries */
cfx_status cfx_assemble_vector(cfx_ctx* ctx, const cfx_form* L, double* b, int zero_first, int memspace);
/* assemble_scalar: assemble_scalar_impl.h:26-275 (fixed-order tree reduction) */
cfx_status cfx_assemble_scalar(cfx_ctx* ctx, const cfx_form* M, double* out);

/* ------------------------------------------------------------------ Dirichlet conditions
 * Markers are int8 per blocked dof index bs*dof + k over the owned+ghost dofs (what DOLFINx builds from the
 * DirichletBC list, assembler.h:657-672); value arrays are doubles with the same indexing.
 *
 * cfx_assemble_matrix_bc: assemble_matrix(A, a, bcs), assembler.h:643-683 -- the Dirichlet rows (bc_markers0) and
 * columns (bc_markers1) of every element tensor are zeroed before mat_set (assemble_matrix_impl.h:146-185,
 * :537-603).  Either marker array may be NULL.  Like the reference it ADDS the masked contributions unless
 * zero_first != 0, and it does not touch the diagonal (cfx_set_diagonal does). */
cfx_status cfx_assemble_matrix_bc(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, int zero_first, double diag_inactive,
                                  const int8_t* bc_markers0, const int8_t* bc_markers1, int memspace_bc,
                                  double* values_out, int memspace);
/* set_diagonal, assembler.h:745-787 (called from fem.py:935-941 after assembly when test == trial space):
 * A[r][r] = diagonal for the listed OWNED Dirichlet rows; rows are blocked indices bs*dof + k. */
cfx_status cfx_set_diagonal(cfx_ctx* ctx, cfx_pattern* A, const int32_t* rows, int64_t n, double diagonal, int memspace);
/* apply_lifting, assemble_vector_impl.h:383-564 (lift_bc): b -= alpha * A_unconstrained[:, bc] (bc_values1 - x0),
 * A generated by the bilinear form a on the pattern A (its values are left as they were).  x0 may be NULL.
 * b, bc_values1, bc_markers1 and x0 live in `memspace`, n_dofs_total*bs entries each. */
cfx_status cfx_apply_lifting(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, double* b, const double* bc_values1,
                             const int8_t* bc_markers1, const double* x0, double alpha, int memspace);
/* The constrained system of demo_elasticity.py:78-84 in one assembly: A = assemble_matrix(a, bcs) with `diagonal` on
 * the owned Dirichlet rows, b = assemble_vector(L) - alpha A_unconstrained[:, bc] (bc_values - x0), then
 * b[bc] = alpha (bc_values - x0).  The unconstrained system is assembled once and one pass over its CSR rows does the
 * lifting and the masking (cfx_apply_lifting alone has to assemble the form a second time).  A and b are
 * overwritten; b, bc_markers, bc_values, x0 (may be NULL), bc_rows_owned are DEVICE arrays. */
cfx_status cfx_assemble_system_bc(cfx_ctx* ctx, const cfx_form* a, cfx_pattern* A, const cfx_form* L, double* b,
                                  const int8_t* bc_markers, const double* bc_values, const double* x0, double alpha,
                                  const int32_t* bc_rows_owned, int64_t n_bc_rows, double diagonal);
/* DirichletBC::set as fem.set_bc uses it: b[d] = alpha * (bc_values[d] - x0[d]) for the listed blocked dof
 * indices d; b, bc_values and x0 (may be NULL) have n_total entries. */
cfx_status cfx_set_bc(cfx_ctx* ctx, double* b, int64_t n_total, const int32_t* dofs, int64_t n, const double* bc_values,
                      const double* x0, double alpha, int memspace);

/* ------------------------------------------------------------------ active domain / deactivation
 * cutfemx::fem::active_domain, cpp/cutfemx/fem/deactivate.h:387-400: active cells = sorted unique OWNED
 * cells of every integral domain of the bilinear form (collect_active_cells :103-165, both cells of an
 * interior-facet row included); the indicator marks the dofs of those cells (:167-185); inactive dofs =
 * OWNED dofs the indicator leaves at zero (deactivated_dofs(.., owned_only = true) :49-64).  On more than
 * one rank the caller reverse/forward-scatters the indicator (cfx_active_indicator_device_ptr) exactly as
 * :180-181 does and takes the inactive dofs from the scattered copy with cfx_inactive_dofs. */
cfx_status cfx_active_domain(cfx_ctx* ctx, const cfx_form* a, cfx_list** active_cells, cfx_list** inactive_dofs);
/* the indicator itself: 1 byte per owned+ghost dof (DEVICE pointer, library-owned, valid until the form changes) */
const uint8_t* cfx_active_indicator_device_ptr(cfx_ctx* ctx, const cfx_form* a);
/* inactive dofs from a (scattered) indicator: owned dofs with indicator == 0 (DEVICE array of n_dofs_owned bytes) */
cfx_status cfx_inactive_dofs(cfx_ctx* ctx, const uint8_t* indicator, int64_t n_dofs_owned, cfx_list** inactive_dofs);
/* deactivate_outside, deactivate.h:402-418: dolfinx::fem::set_diagonal(A.mat_set_values, inactive_dofs,
 * diagonal) -- SETS A[r][r] -- and, if b != NULL (DEVICE), b[r] = rhs_value.  Fails if a row lacks its
 * diagonal entry (the pattern always reserves it: assembler.h:538-560, :589-590). */
cfx_status cfx_deactivate_outside(cfx_ctx* ctx, cfx_pattern* A, const int32_t* inactive_dofs, int64_t n, int memspace,
                                  double diagonal, double* b, double rhs_value);

/* ------------------------------------------------------------------ ghost exchange (one rank per GPU)
 * la::MatrixCSR::scatter_rev / la::Vector::scatter_rev(add) as called by the user after assembly
 * (python/demo/demo_poisson.py:52,54): values of ghost rows / ghost entries travel to the owning
 * rank and are added there.  The library provides the pack and unpack-add kernels around the
 * neighbour exchange (NCCL send/recv on the device buffers, cutfemx_b200/parallel.py); all pointers
 * are DEVICE pointers.  scatter_add requires distinct indices within one call (one neighbour's
 * entries are distinct matrix positions), so it needs no atomics and the result is bit-reproducible
 * when neighbours are applied in a fixed order. */
cfx_status cfx_gather_f64(cfx_ctx* ctx, const double* src, const int64_t* index, int64_t n, double* dst);
cfx_status cfx_scatter_add_f64(cfx_ctx* ctx, const double* src, const int64_t* index, int64_t n, double* dst);

/* The same exchange without a size exchange, a host round trip or host code in the step (one rank per GPU over
 * NCCL / NVLink).  DOLFINx renegotiates the ghost-row entries at every
 * SparsityPattern::finalize(); here the ranks agree ONCE per partition on a static superset: for every ghost row
 * the columns it can ever have (every cell active, every interior facet in the stabilisation band), translated by
 * the owner to its own numbering (dofs it does not know become new ghost columns, as finalize() does).  Per step
 * only messages of known size travel -- one bit per candidate ("in my pattern of this step"), then the values -- so
 * the calls below never synchronise in deferred-size mode and can be captured in the step's CUDA graph together with
 * the rest of the step.  Value messages: scalar P1 spaces send one double per candidate (0 where the bit is clear)
 * plus one per ghost vector entry, a fixed layout; every other space (P2, block size bs > 1: bs x bs doubles per
 * entry, bs per ghost vector entry) sends COMPACT messages -- only the blocks whose bit is set, in candidate order;
 * both ends know the bits and find a block's slot from them.  Compact messages are sized exactly in eager steps
 * (one read-back of the counts per side) and, in deferred-size steps, by the largest count the eager steps saw plus
 * the context's margin (the same number on both ends); more set bits than that raise the device error flag.
 *
 * cfx_comm_unique_id / cfx_comm_init: an NCCL communicator of the context's own (ncclGetUniqueId on one rank, the
 * 128 bytes broadcast by the host program, ncclCommInitRank on every rank); `nccl_path` may be NULL (the libnccl
 * the process has loaded, else libnccl.so.2).
 *
 * cfx_xplan_create (HOST arrays): neighbours in ascending rank order.
 *   send side  -- s_row_off[k..k+1]: this rank's ghost rows owned by neighbour k (indices into s_rows, local row
 *                 ids); s_ptr/s_cols: CSR of each row's candidate columns (local ids);
 *   recv side  -- arrival order = neighbour by neighbour, each neighbour's entries in ITS s_ptr/s_cols order:
 *                 r_ent_off[k..k+1] entries of neighbour k, r_row/r_col their local (row, column) here (a column
 *                 >= the space's owned+ghost dof count is a new ghost column), r_perm the arrival indices sorted by
 *                 (row, arrival); r_row_off / r_vec_row: the neighbour's ghost rows as local owned rows (vector
 *                 entries). */
typedef struct cfx_xplan cfx_xplan;
cfx_status cfx_comm_unique_id(void* out128, const char* nccl_path);
cfx_status cfx_comm_init(cfx_ctx* ctx, const void* unique_id128, int rank, int n_ranks, const char* nccl_path);
void cfx_comm_destroy(cfx_ctx* ctx);
cfx_status cfx_xplan_create(cfx_ctx* ctx, int space, int n_neigh, const int32_t* neigh_ranks, const int64_t* s_row_off,
                            const int32_t* s_rows, const int64_t* s_ptr, const int32_t* s_cols, const int64_t* r_ent_off,
                            const int32_t* r_row, const int32_t* r_col, const int32_t* r_perm, const int64_t* r_row_off,
                            const int32_t* r_vec_row, cfx_xplan** out);
void cfx_xplan_free(cfx_ctx* ctx, cfx_xplan* plan);
/* SparsityPattern::finalize(): sender half (bits of this step's ghost-row entries of form a), the exchange
 * (which = 0), owner half (the received bits become the form's inserted entries, on the device).  Call before
 * cfx_create_sparsity(a). */
cfx_status cfx_xplan_pack_pattern(cfx_ctx* ctx, cfx_xplan* plan, const cfx_form* a);
cfx_status cfx_xplan_exchange(cfx_ctx* ctx, cfx_xplan* plan, int which /* 0 pattern bits, 1 values */);
cfx_status cfx_xplan_insert_pattern(cfx_ctx* ctx, cfx_xplan* plan, cfx_form* a);
/* MatrixCSR::scatter_rev() + Vector::scatter_rev(add) after assembly (demo_poisson.py:52,54): pack the ghost-row
 * values of A and the ghost entries of b (b may be NULL), exchange (which = 1), add on the owner -- neighbour by
 * neighbour in rank order, no atomics, bit-reproducible. */
cfx_status cfx_xplan_pack_values(cfx_ctx* ctx, cfx_xplan* plan, const cfx_pattern* A, const double* b);
cfx_status cfx_xplan_unpack_add(cfx_ctx* ctx, cfx_xplan* plan, cfx_pattern* A, double* b);
/* message buffers of neighbour k for transports other than NCCL (ranks emulated on one GPU in the tests):
 * which = 0 bits to send, 1 bits received, 2 values to send, 3 values received */
cfx_status cfx_xplan_buffer(cfx_ctx* ctx, const cfx_xplan* plan, int which, int k, void** ptr, int64_t* bytes);

/* ------------------------------------------------------------------ profiling hooks
 * per-stage CUDA-event timings of the most recent calls, for bench.py's roofline block.
 * names/ms are library-owned, valid until the next call on this context. */
int cfx_stage_count(const cfx_ctx* ctx);
const char* cfx_stage_name(const cfx_ctx* ctx, int i);
cfx_status cfx_stage_timing_enable(cfx_ctx* ctx, int on);
cfx_status cfx_stage_ms(cfx_ctx* ctx, int i, double* ms, double* bytes);
cfx_status cfx_stage_reset(cfx_ctx* ctx);

/* ------------------------------------------------------------------ synthetic meshes (harness)
 * Device-side generator of the Kuhn box / right-diagonal rectangle meshes of
 * cutfemx_b200/mesh.py (same numbering), so that 256^3 never has to exist on the host.
 * All outputs are caller-allocated DEVICE arrays. f2c is dense (n_facets, 2), -1 padded. */
cfx_status cfx_meshgen_box(cfx_ctx* ctx, int nx, int ny, int nz, const double p0[3], const double p1[3],
                           double* x, int32_t* x_dofmap, int32_t* c2f);
cfx_status cfx_meshgen_rectangle(cfx_ctx* ctx, int nx, int ny, const double p0[2], const double p1[2], double* x,
                                 int32_t* x_dofmap, int32_t* c2f);
/* level-set nodal interpolation on the device: kind 0 sphere/circle (c, R), 1 torus (c, R, r) */
cfx_status cfx_meshgen_level_set(cfx_ctx* ctx, const double* x, int64_t n_nodes, int kind, const double params[5],
                                 double* values);
/* P2 Lagrange dofmap (n_cells, 10) of a cfx_meshgen_box mesh: vertex dofs, then one dof per edge in Basix edge
 * order; *n_dofs = 8 (nx+1)(ny+1)(nz+1) (sparse edge numbering, unused ids at the box boundary).  Harness only:
 * stands in for dolfinx.fem.functionspace(mesh, ("Lagrange", 2)) (python/demo/demo_elasticity.py:160). */
cfx_status cfx_meshgen_p2_tet_dofmap(cfx_ctx* ctx, int nx, int ny, int nz, const int32_t* x_dofmap, int64_t n_cells,
                                     int32_t* dofmap, int64_t* n_dofs);

#ifdef __cplusplus
}
#endif
#endif /* CUTFEMX_B200_H */

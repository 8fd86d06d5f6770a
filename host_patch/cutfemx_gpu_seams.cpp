// host_patch/cutfemx_gpu_seams.cpp -- the three seams of SURVEY.md section 8(b) re-expressed on the C ABI of
// include/cutfemx_b200.h, as a CutFEMx maintainer would add them to cpp/cutfemx (one translation unit next to
// cut/cut.cpp).  Written against the DOLFINx / CutCells member functions the reference already calls; in this
// image it is compile-checked against host_patch/stubs (tests/test_abi.py::test_host_patch_compiles) and linked
// against libcutfemx_b200.so so that every call resolves.  Reference file:line for each function is given.
#include <cutfemx_b200.h>

#include <cstdint>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef CUTFEMX_HOST_PATCH_STUBS
#include <cutcells/stub.h>
#include <dolfinx/stub.h>
#else
#include <cutcells/cutcells.h>
#include <dolfinx/fem/Function.h>
#include <dolfinx/la/MatrixCSR.h>
#include <dolfinx/mesh/Mesh.h>
#endif

namespace cutfemx::gpu
{
inline void check(cfx_ctx* c, cfx_status s)
{
  if (s != CFX_OK)
    throw std::runtime_error(cfx_last_error(c)); // what cut.cpp throws today
}

/// CutData<T>::gpu (cut.h:81-101 gains this member): one context per (rank, GPU)
struct Context
{
  explicit Context(int device)
  {
    cfx_ctx* c = nullptr;
    check(nullptr, cfx_ctx_create(device, /*stream*/ nullptr, &c));
    handle.reset(c, [](cfx_ctx* p) { cfx_ctx_destroy(p); });
  }
  cfx_ctx* get() const { return handle.get(); }
  std::shared_ptr<cfx_ctx> handle;
};

/// build_mesh_view, cut.cpp:500-538: the same spans, handed to the library instead of cutcells::MeshView
template <typename T>
void bind_mesh(const Context& ctx, const dolfinx::mesh::Mesh<T>& mesh)
{
  static_assert(sizeof(T) == sizeof(double), "the library computes in float64");
  const auto& geometry = mesh.geometry();
  auto topology = mesh.topology();
  const int tdim = topology->dim();
  auto cell_map = topology->index_map(tdim);
  auto x = geometry.x();
  auto xd = geometry.dofmap();
  check(ctx.get(),
        cfx_mesh_bind(ctx.get(), reinterpret_cast<const double*>(x.data()), static_cast<std::int64_t>(x.size() / 3),
                      xd.data_handle(), cell_map->size_local(), static_cast<std::int64_t>(xd.extent(0)),
                      static_cast<int>(xd.extent(1)), geometry.dim(), CFX_HOST));
  // ghost_penalty_facets / facet_integration_rows (cut.py:340-380, wrappers/cut.cpp:54-115) need the facet topology
  auto c2f = topology->connectivity(tdim, tdim - 1);
  auto f2c = topology->connectivity(tdim - 1, tdim);
  auto facet_map = topology->index_map(tdim - 1);
  if (c2f && f2c && facet_map)
    check(ctx.get(), cfx_topology_bind(ctx.get(), c2f->array().data(), f2c->offsets().data(), f2c->array().data(),
                                       facet_map->size_local() + facet_map->num_ghosts(), facet_map->size_local(),
                                       CFX_HOST));
}

/// build_level_set_function, cut.cpp:593-636 (+ validate_level_set :444-460): dofmap span + dof_values span
template <typename T>
void bind_level_set(const Context& ctx, int slot, const dolfinx::fem::Function<T>& phi, int degree)
{
  auto dm = phi.function_space()->dofmap()->map();
  auto values = phi.x()->array();
  check(ctx.get(), cfx_levelset_bind(ctx.get(), slot, dm.data_handle(), static_cast<int>(dm.extent(1)), degree,
                                     reinterpret_cast<const double*>(values.data()),
                                     static_cast<std::int64_t>(values.size()), CFX_HOST, /*pin_host*/ 1));
}

/// update, cut.cpp:845-868: replaces the rebind loop + cutcells::cut(mesh_view, level_sets, options)
inline void update(const Context& ctx) { check(ctx.get(), cfx_update(ctx.get())); }

/// locate_entities, cut.cpp:877-924: the selector is parsed by the host (CutCells' parser is header code) and
/// flattened into (term offsets, clause level set, clause relation); the cell loop runs on the device
inline std::vector<std::int32_t> locate_entities(const Context& ctx, std::span<const std::int32_t> term_offsets,
                                                 std::span<const std::int32_t> clause_ls,
                                                 std::span<const std::int32_t> clause_rel)
{
  cfx_list* l = nullptr;
  check(ctx.get(), cfx_locate_entities(ctx.get(), static_cast<int>(term_offsets.size()) - 1, term_offsets.data(),
                                       clause_ls.data(), clause_rel.data(), &l));
  std::vector<std::int32_t> marked(static_cast<std::size_t>(cfx_list_size(l)));
  check(ctx.get(), cfx_list_fetch(ctx.get(), l, marked.data(), CFX_HOST));
  cfx_list_free(ctx.get(), l);
  return marked;
}

/// runtime_quadrature, cut.cpp:1311-1335: replaces select_mesh_part + cutcells::output::quadrature_rules; the
/// device copy (`*keep`) stays inside RuntimeQuadrature<T> for the assembly calls
template <typename T>
cutcells::quadrature::QuadratureRules<T> runtime_quadrature(const Context& ctx, int ls_index, int relation, int order,
                                                            cfx_rules** keep)
{
  check(ctx.get(), cfx_runtime_quadrature(ctx.get(), ls_index, relation, order, keep));
  std::int64_t npts = 0, nrules = 0;
  int tdim = 0;
  check(ctx.get(), cfx_rules_sizes(*keep, &npts, &nrules, &tdim));
  cutcells::quadrature::QuadratureRules<T> rules;
  rules._tdim = tdim;
  rules._points.resize(static_cast<std::size_t>(npts * tdim));
  rules._weights.resize(static_cast<std::size_t>(npts));
  rules._offset.resize(static_cast<std::size_t>(nrules + 1));
  rules._parent_map.resize(static_cast<std::size_t>(nrules));
  check(ctx.get(), cfx_rules_fetch(ctx.get(), *keep, reinterpret_cast<double*>(rules._points.data()),
                                   reinterpret_cast<double*>(rules._weights.data()), rules._offset.data(),
                                   rules._parent_map.data(), CFX_HOST)); // AoS (npts, tdim): wrappers/cut.cpp:185-196
  return rules;
}

/// The device-resident rules of a time loop (demo_moving_poisson.py:69-107): volume rules, interface rules with
/// their normals and the selected cell list depend on cutfemx.update only, so the host issues them on lanes -- streams
/// of their own, parallel branches of a captured step -- and joins before the forms are built.  Sequential host code;
/// `keep_*` are the objects of the previous time step, refilled in place (deferred-size mode: no size reaches the host).
inline void update_rules_on_lanes(const Context& ctx, int ls_index, int order, cfx_rules** keep_volume,
                                  cfx_rules** keep_interface, std::span<const std::int32_t> term_offsets,
                                  std::span<const std::int32_t> clause_ls, std::span<const std::int32_t> clause_rel,
                                  cfx_list** keep_cells)
{
  check(ctx.get(), cfx_update(ctx.get()));
  check(ctx.get(), cfx_lane_begin(ctx.get(), 1));
  check(ctx.get(), cfx_runtime_quadrature(ctx.get(), ls_index, CFX_REL_LT, order, keep_volume));
  check(ctx.get(), cfx_lane_end(ctx.get()));
  check(ctx.get(), cfx_lane_begin(ctx.get(), 2));
  check(ctx.get(), cfx_runtime_quadrature(ctx.get(), ls_index, CFX_REL_EQ, order, keep_interface));
  check(ctx.get(), cfx_evaluate_normals(ctx.get(), ls_index, *keep_interface, 1.0, nullptr, CFX_DEVICE));
  check(ctx.get(), cfx_lane_end(ctx.get()));
  check(ctx.get(), cfx_locate_entities(ctx.get(), static_cast<int>(term_offsets.size()) - 1, term_offsets.data(),
                                       clause_ls.data(), clause_rel.data(), keep_cells)); // main stream
  check(ctx.get(), cfx_lane_join(ctx.get()));
}

/// make_surface_provenance, cut.cpp:1273-1308
struct SurfaceProvenance
{
  std::int32_t level_set_index = -1;
  std::vector<std::int32_t> cut_cell_ids, parent_cell_ids, local_zero_entity_ids, dimensions;
};
inline SurfaceProvenance surface_provenance(const Context& ctx, const cfx_rules* rules)
{
  SurfaceProvenance p;
  std::int64_t nrules = 0;
  check(ctx.get(), cfx_rules_sizes(rules, nullptr, &nrules, nullptr));
  check(ctx.get(), cfx_rules_surface_provenance(ctx.get(), rules, &p.level_set_index, nullptr, nullptr, nullptr,
                                                nullptr, CFX_HOST));
  if (p.level_set_index < 0)
    return p;
  const auto n = static_cast<std::size_t>(nrules);
  p.cut_cell_ids.resize(n);
  p.parent_cell_ids.resize(n);
  p.local_zero_entity_ids.resize(n);
  p.dimensions.resize(n);
  check(ctx.get(), cfx_rules_surface_provenance(ctx.get(), rules, &p.level_set_index, p.cut_cell_ids.data(),
                                                p.parent_cell_ids.data(), p.local_zero_entity_ids.data(),
                                                p.dimensions.data(), CFX_HOST));
  return p;
}

/// evaluate_normals, level_set/normal.h:39-188: (npts, gdim) row-major
inline std::vector<double> evaluate_normals(const Context& ctx, int ls_index, cfx_rules* rules, double sign, int gdim)
{
  std::int64_t npts = 0;
  check(ctx.get(), cfx_rules_sizes(rules, &npts, nullptr, nullptr));
  std::vector<double> n(static_cast<std::size_t>(npts * gdim));
  check(ctx.get(), cfx_evaluate_normals(ctx.get(), ls_index, rules, sign, n.data(), CFX_HOST));
  return n;
}

/// ghost_penalty_facets, cut.py:340-380 (the Python set loop becomes one bound function) + facet_integration_rows,
/// wrappers/cut.cpp:54-115
inline std::vector<std::int32_t> ghost_penalty_facets(const Context& ctx, int phi_index,
                                                      std::span<const std::int32_t> term_offsets,
                                                      std::span<const std::int32_t> clause_ls,
                                                      std::span<const std::int32_t> clause_rel, bool include_ghosts)
{
  cfx_list* l = nullptr;
  check(ctx.get(), cfx_ghost_penalty_facets(ctx.get(), phi_index, static_cast<int>(term_offsets.size()) - 1,
                                            term_offsets.data(), clause_ls.data(), clause_rel.data(),
                                            include_ghosts ? 1 : 0, &l));
  std::vector<std::int32_t> facets(static_cast<std::size_t>(cfx_list_size(l)));
  check(ctx.get(), cfx_list_fetch(ctx.get(), l, facets.data(), CFX_HOST));
  cfx_list_free(ctx.get(), l);
  return facets;
}

/// create_form_* (wrappers/fem.cpp:124-172): a form whose integrals use shipped kernel families; the mixed measure
/// subdomain_data = [cells, rules] of demo_poisson.py:165-167 is one call (entities + custom_data)
template <typename T>
cfx_form* poisson_bilinear_form(const Context& ctx, int space_slot, const dolfinx::fem::FunctionSpace<T>& V, int degree,
                                std::span<const std::int32_t> inside_cells, cfx_rules* volume_rules,
                                cfx_rules* interface_rules, std::span<const std::int32_t> facet_rows4, double gamma,
                                double gamma_g)
{
  auto dofmap = V.dofmap();
  auto dm = dofmap->map();
  auto imap = dofmap->index_map;
  check(ctx.get(), cfx_space_bind(ctx.get(), space_slot, dm.data_handle(), static_cast<int>(dm.extent(1)), dofmap->bs(),
                                  degree, imap->size_local(), imap->size_local() + imap->num_ghosts(), CFX_HOST));
  cfx_form* a = nullptr;
  check(ctx.get(), cfx_form_create(ctx.get(), space_slot, /*rank*/ 2, &a));
  const double one = 1.0;
  check(ctx.get(), cfx_form_add_cell_integral(ctx.get(), a, CFX_K_LAPLACE, inside_cells.data(),
                                              static_cast<std::int64_t>(inside_cells.size()), CFX_HOST, volume_rules,
                                              &one, 1));
  check(ctx.get(), cfx_form_add_cell_integral(ctx.get(), a, CFX_K_NITSCHE, nullptr, 0, CFX_HOST, interface_rules,
                                              &gamma, 1));
  check(ctx.get(), cfx_form_add_interior_facet_integral(ctx.get(), a, CFX_K_GHOST_GRAD_JUMP, facet_rows4.data(),
                                                        static_cast<std::int64_t>(facet_rows4.size() / 4), CFX_HOST,
                                                        &gamma_g, 1));
  return a;
}

/// create_sparsity_pattern + MatrixCSR(sp) + assemble_matrix, assembler.h:442-592, 596-703, wrappers/fem.cpp:266-276:
/// the pattern is built on the device and adopted into the host matrix, the values are ADDED (no finalise)
template <typename T>
void assemble_matrix(const Context& ctx, cfx_form* a, dolfinx::la::MatrixCSR<T>& A)
{
  cfx_pattern* P = nullptr;
  check(ctx.get(), cfx_create_sparsity(ctx.get(), a, &P));
  std::int64_t n_rows = 0, nnz = 0;
  check(ctx.get(), cfx_pattern_sizes(P, &n_rows, &nnz));
  const int bs = cfx_pattern_block_size(P);
  A.row_ptr().resize(static_cast<std::size_t>(n_rows + 1));
  A.cols().resize(static_cast<std::size_t>(nnz));
  A.values().assign(static_cast<std::size_t>(nnz) * bs * bs, T(0));
  check(ctx.get(), cfx_pattern_fetch(ctx.get(), P, A.row_ptr().data(), A.cols().data(), CFX_HOST));
  check(ctx.get(), cfx_assemble_matrix(ctx.get(), a, P, /*zero_first*/ 0, /*diag_inactive*/ 0.0,
                                       reinterpret_cast<double*>(A.values().data()), CFX_HOST));
  cfx_pattern_free(ctx.get(), P);
}

/// assemble_vector, assemble_vector_impl.h:62-122 (adds into b) and assemble_scalar, assemble_scalar_impl.h:26-59
template <typename T>
void assemble_vector(const Context& ctx, cfx_form* L, dolfinx::la::Vector<T>& b)
{
  check(ctx.get(), cfx_assemble_vector(ctx.get(), L, reinterpret_cast<double*>(b.mutable_array().data()),
                                       /*zero_first*/ 0, CFX_HOST));
}
inline double assemble_scalar(const Context& ctx, cfx_form* M)
{
  double value = 0.0;
  check(ctx.get(), cfx_assemble_scalar(ctx.get(), M, &value));
  return value;
}

// explicit instantiations: what cut.cpp:1405-1465 does for the reference's templates
template void bind_mesh<double>(const Context&, const dolfinx::mesh::Mesh<double>&);
template void bind_level_set<double>(const Context&, int, const dolfinx::fem::Function<double>&, int);
template cutcells::quadrature::QuadratureRules<double> runtime_quadrature<double>(const Context&, int, int, int,
                                                                                  cfx_rules**);
template cfx_form* poisson_bilinear_form<double>(const Context&, int, const dolfinx::fem::FunctionSpace<double>&, int,
                                                 std::span<const std::int32_t>, cfx_rules*, cfx_rules*,
                                                 std::span<const std::int32_t>, double, double);
template void assemble_matrix<double>(const Context&, cfx_form*, dolfinx::la::MatrixCSR<double>&);
template void assemble_vector<double>(const Context&, cfx_form*, dolfinx::la::Vector<double>&);
} // namespace cutfemx::gpu

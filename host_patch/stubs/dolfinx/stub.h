// Minimal stand-ins for the DOLFINx 0.11 types the CutFEMx seams touch (only the members the patch calls, with
// the signatures DOLFINx gives them), so that host_patch/cutfemx_gpu_seams.cpp can be compile-checked in an image
// without DOLFINx.  Not part of the product; nothing here is copied from DOLFINx.
#pragma once
#include <cstdint>
#include <memory>
#include <span>
#include <vector>

namespace dolfinx
{
namespace common
{
struct IndexMap
{
  std::int32_t size_local() const { return _local; }
  std::int32_t num_ghosts() const { return _ghosts; }
  std::int32_t _local = 0, _ghosts = 0;
};
} // namespace common
namespace graph
{
template <typename T>
struct AdjacencyList
{
  const std::vector<T>& array() const { return _array; }
  const std::vector<std::int32_t>& offsets() const { return _offsets; }
  std::vector<T> _array;
  std::vector<std::int32_t> _offsets;
};
} // namespace graph
// (n, width) row-major view, the shape of md::mdspan<const std::int32_t, md::dextents<std::size_t, 2>>
struct DofmapView
{
  const std::int32_t* data_handle() const { return _p; }
  std::size_t extent(int d) const { return d == 0 ? _n : _w; }
  const std::int32_t* _p = nullptr;
  std::size_t _n = 0, _w = 0;
};
namespace mesh
{
template <typename T>
struct Geometry
{
  std::span<const T> x() const { return _x; }
  DofmapView dofmap() const { return _dofmap; }
  int dim() const { return _gdim; }
  std::vector<T> _x;
  DofmapView _dofmap;
  int _gdim = 3;
};
struct Topology
{
  int dim() const { return _tdim; }
  std::shared_ptr<const common::IndexMap> index_map(int) const { return _map; }
  std::shared_ptr<const graph::AdjacencyList<std::int32_t>> connectivity(int, int) const { return _conn; }
  int _tdim = 3;
  std::shared_ptr<const common::IndexMap> _map;
  std::shared_ptr<const graph::AdjacencyList<std::int32_t>> _conn;
};
template <typename T>
struct Mesh
{
  const Geometry<T>& geometry() const { return _geometry; }
  std::shared_ptr<const Topology> topology() const { return _topology; }
  Geometry<T> _geometry;
  std::shared_ptr<const Topology> _topology;
};
} // namespace mesh
namespace la
{
template <typename T>
struct Vector
{
  std::span<const T> array() const { return _a; }
  std::span<T> mutable_array() { return _a; }
  std::vector<T> _a;
};
template <typename T>
struct MatrixCSR
{
  std::vector<std::int64_t>& row_ptr() { return _row_ptr; }
  std::vector<std::int32_t>& cols() { return _cols; }
  std::vector<T>& values() { return _values; }
  std::vector<std::int64_t> _row_ptr;
  std::vector<std::int32_t> _cols;
  std::vector<T> _values;
};
} // namespace la
namespace fem
{
struct DofMap
{
  DofmapView map() const { return _map; }
  int bs() const { return _bs; }
  std::shared_ptr<const common::IndexMap> index_map;
  DofmapView _map;
  int _bs = 1;
};
template <typename T>
struct FunctionSpace
{
  std::shared_ptr<const mesh::Mesh<T>> mesh() const { return _mesh; }
  std::shared_ptr<const DofMap> dofmap() const { return _dofmap; }
  std::shared_ptr<const mesh::Mesh<T>> _mesh;
  std::shared_ptr<const DofMap> _dofmap;
};
template <typename T, typename U = T>
struct Function
{
  std::shared_ptr<const FunctionSpace<U>> function_space() const { return _V; }
  std::shared_ptr<const la::Vector<T>> x() const { return _x; }
  std::shared_ptr<const FunctionSpace<U>> _V;
  std::shared_ptr<la::Vector<T>> _x;
};
} // namespace fem
} // namespace dolfinx

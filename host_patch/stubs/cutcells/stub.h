// Stand-in for the one CutCells container the wrappers expose (cutcells::quadrature::QuadratureRules, used at
// runtime_quadrature.h:107-137 and wrappers/cut.cpp:185-240): member names as the reference uses them.
#pragma once
#include <cstdint>
#include <vector>

namespace cutcells::quadrature
{
template <typename T>
struct QuadratureRules
{
  int _tdim = 0;
  std::vector<T> _points;                // AoS (npts, tdim), parent reference coordinates
  std::vector<T> _weights;               // physical weights
  std::vector<std::int32_t> _offset;     // (nrules + 1)
  std::vector<std::int32_t> _parent_map; // (nrules)
};
} // namespace cutcells::quadrature

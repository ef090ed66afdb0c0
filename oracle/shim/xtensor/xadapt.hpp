// Test infrastructure (oracle build only): xt::adapt stand-in (copying) for src/SMPL.cpp's OBJ export.
#pragma once
#include "xarray.hpp"

namespace xt
{
template<typename T>
xarray<T> adapt(T * ptr, const typename xarray<T>::shape_type & shape)
{
  std::size_t n = 1;
  for(std::size_t s : shape)
  {
    n *= s;
  }
  return xarray<T>(std::vector<T>(ptr, ptr + n), shape);
}
} // namespace xt

// Test infrastructure (oracle build only): xt::from_json stand-in — row-major flatten of a nested JSON
// array into the xarray stand-in, recording the shape.  See xarray.hpp in this directory.
#pragma once
#include <nlohmann/json.hpp>

#include "xarray.hpp"

namespace xt
{
namespace detail
{
template<typename T>
void flatten_json(const nlohmann::json & j, std::vector<T> & out, std::vector<std::size_t> & shape, std::size_t depth)
{
  if(j.is_array())
  {
    if(shape.size() <= depth)
    {
      shape.push_back(j.size());
    }
    for(const auto & child : j)
    {
      flatten_json<T>(child, out, shape, depth + 1);
    }
  }
  else
  {
    out.push_back(j.get<T>());
  }
}
} // namespace detail

template<typename T>
void from_json(const nlohmann::json & j, xarray<T> & arr)
{
  std::vector<T> flat;
  std::vector<std::size_t> shape;
  detail::flatten_json<T>(j, flat, shape, 0);
  arr = xarray<T>(std::move(flat), std::move(shape));
}
} // namespace xt

// Test infrastructure (oracle build only): a minimal stand-in for the parts of xtensor that the
// reference's src/SMPL.cpp and src/VPoser.cpp touch (xt::xarray<T>: data(), shape(i), dimension(),
// operator()(i, j), shape_type).  xtensor itself is not installed in this image; this header lets the
// reference sources compile UNMODIFIED from /root/reference.  Never used by the product path.
#pragma once
#include <cstddef>
#include <vector>

namespace xt
{
template<typename T>
class xarray
{
public:
  using value_type = T;
  using shape_type = std::vector<std::size_t>;

  xarray() = default;
  xarray(std::vector<T> flat, shape_type shape) : flat_(std::move(flat)), shape_(std::move(shape)) {}

  T * data() { return flat_.data(); }
  const T * data() const { return flat_.data(); }
  std::size_t dimension() const { return shape_.size(); }
  std::size_t shape(std::size_t axis) const { return shape_.at(axis); }
  const shape_type & shape() const { return shape_; }
  std::size_t size() const { return flat_.size(); }

  template<typename... Idx>
  T & operator()(Idx... idx)
  {
    return flat_[offset({static_cast<std::size_t>(idx)...})];
  }
  template<typename... Idx>
  const T & operator()(Idx... idx) const
  {
    return flat_[offset({static_cast<std::size_t>(idx)...})];
  }

  std::vector<T> & storage() { return flat_; }
  shape_type & mutable_shape() { return shape_; }

private:
  std::size_t offset(std::initializer_list<std::size_t> idx) const
  {
    std::size_t off = 0, axis = 0;
    for(std::size_t i : idx)
    {
      off = off * shape_[axis++] + i;
    }
    return off;
  }

  std::vector<T> flat_;
  shape_type shape_;
};
} // namespace xt

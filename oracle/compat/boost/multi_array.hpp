// compat shim (TEST INFRASTRUCTURE): the 2-D subset of boost::multi_array the reference uses
// (construct from boost::extents[a][b], table[i][j], shape(), resize keeping the overlap).
#ifndef RFS_COMPAT_BOOST_MULTI_ARRAY
#define RFS_COMPAT_BOOST_MULTI_ARRAY
#include <algorithm>
#include <cstddef>
#include <vector>
namespace boost {
namespace detail_ma {
template <int N> struct extent_gen {
  std::size_t e[N > 0 ? N : 1];
  extent_gen<N + 1> operator[](std::ptrdiff_t n) const {
    extent_gen<N + 1> g;
    for (int i = 0; i < N; i++) g.e[i] = e[i];
    g.e[N] = (std::size_t)n;
    return g;
  }
};
}
static const detail_ma::extent_gen<0> extents = detail_ma::extent_gen<0>();

template <typename T, int NumDims> class multi_array;
template <typename T>
class multi_array<T, 2> {
 public:
  typedef std::size_t size_type;
  multi_array() { s_[0] = s_[1] = 0; }
  multi_array(const detail_ma::extent_gen<2>& g) { s_[0] = g.e[0]; s_[1] = g.e[1]; d_.assign(s_[0] * s_[1], T()); }
  T* operator[](std::ptrdiff_t i) { return &d_[(std::size_t)i * s_[1]]; }
  const T* operator[](std::ptrdiff_t i) const { return &d_[(std::size_t)i * s_[1]]; }
  const size_type* shape() const { return s_; }
  void resize(const detail_ma::extent_gen<2>& g) {
    std::vector<T> n(g.e[0] * g.e[1], T());
    const std::size_t r = std::min(s_[0], g.e[0]), c = std::min(s_[1], g.e[1]);
    for (std::size_t i = 0; i < r; i++)
      for (std::size_t j = 0; j < c; j++) n[i * g.e[1] + j] = d_[i * s_[1] + j];
    d_.swap(n);
    s_[0] = g.e[0];
    s_[1] = g.e[1];
  }
 private:
  std::vector<T> d_;
  size_type s_[2];
};
}
#endif

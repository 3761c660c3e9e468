// compat shim (TEST INFRASTRUCTURE): boost::shared_array
#ifndef RFS_COMPAT_BOOST_SHARED_ARRAY
#define RFS_COMPAT_BOOST_SHARED_ARRAY
#include <cstddef>
#include <memory>
namespace boost {
template <class T>
class shared_array {
 public:
  shared_array() {}
  explicit shared_array(T* p) : p_(p, std::default_delete<T[]>()) {}
  T& operator[](std::ptrdiff_t i) const { return p_.get()[i]; }
  T* get() const { return p_.get(); }
  void reset() { p_.reset(); }
  void reset(T* p) { p_.reset(p, std::default_delete<T[]>()); }
  explicit operator bool() const { return (bool)p_; }
 private:
  std::shared_ptr<T> p_;
};
}
#endif

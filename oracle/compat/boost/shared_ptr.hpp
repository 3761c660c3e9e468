// compat shim (TEST INFRASTRUCTURE): boost::shared_ptr -> std::shared_ptr
#ifndef RFS_COMPAT_BOOST_SHARED_PTR
#define RFS_COMPAT_BOOST_SHARED_PTR
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
}
#endif

// compat shim (TEST INFRASTRUCTURE): boost::normal_distribution (sampling only; not on the update path)
#ifndef RFS_COMPAT_BOOST_NORMAL
#define RFS_COMPAT_BOOST_NORMAL
#include <random>
namespace boost { template <class T = double> using normal_distribution = std::normal_distribution<T>; }
#endif

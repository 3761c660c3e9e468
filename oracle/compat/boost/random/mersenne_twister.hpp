// compat shim (TEST INFRASTRUCTURE): boost::mt19937 -> std::mt19937 (same generator)
#ifndef RFS_COMPAT_BOOST_MT
#define RFS_COMPAT_BOOST_MT
#include <random>
namespace boost { typedef std::mt19937 mt19937; }
#endif

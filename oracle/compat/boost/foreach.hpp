// boost/foreach.hpp stand-in (Boost is not installed in this image): BOOST_FOREACH over any range
// that works with a range-based for.  TEST INFRASTRUCTURE for building the reference's drivers.
#ifndef COMPAT_BOOST_FOREACH_HPP
#define COMPAT_BOOST_FOREACH_HPP
#define BOOST_FOREACH(decl, coll) for (decl : coll)
#endif

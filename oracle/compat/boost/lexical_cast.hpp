// compat shim (TEST INFRASTRUCTURE): boost::lexical_cast via stringstream
#ifndef RFS_COMPAT_BOOST_LEXICAL_CAST
#define RFS_COMPAT_BOOST_LEXICAL_CAST
#include <sstream>
#include <string>
namespace boost {
template <typename To, typename From>
To lexical_cast(const From& f) {
  std::stringstream ss;
  ss << f;
  To t;
  ss >> t;
  return t;
}
template <>
inline std::string lexical_cast<std::string, std::string>(const std::string& f) { return f; }
}
#endif

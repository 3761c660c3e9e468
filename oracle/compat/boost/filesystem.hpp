// Minimal stand-in for <boost/filesystem.hpp> — TEST INFRASTRUCTURE (oracle/compat): just what the
// reference's drivers use (path, create_directories, copy_file with overwrite).  Boost is not
// installed in this image; on a normal machine the real header is found instead.
#pragma once
#include <fstream>
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
namespace boost { namespace filesystem {
class path {
 public:
  path() {}
  path(const char* s) : s_(s) {}
  path(const std::string& s) : s_(s) {}
  const std::string& string() const { return s_; }
  const char* c_str() const { return s_.c_str(); }
 private:
  std::string s_;
};
inline bool exists(const path& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool create_directories(const path& p) {
  const std::string& s = p.string();
  bool made = false;
  for (size_t i = 1; i <= s.size(); i++) {
    if (i == s.size() || s[i] == '/') {
      const std::string sub = s.substr(0, i);
      if (!sub.empty() && ::mkdir(sub.c_str(), 0777) == 0) made = true;
    }
  }
  return made;
}
namespace copy_option { enum enum_type { none, fail_if_exists = none, overwrite_if_exists }; }
inline void copy_file(const path& from, const path& to, copy_option::enum_type = copy_option::none) {
  std::ifstream in(from.c_str(), std::ios::binary);
  std::ofstream out(to.c_str(), std::ios::binary | std::ios::trunc);
  out << in.rdbuf();
}
}}  // namespace boost::filesystem

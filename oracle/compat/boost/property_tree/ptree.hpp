// Minimal stand-in for <boost/property_tree/ptree.hpp> + xml_parser — TEST INFRASTRUCTURE
// (oracle/compat): element tree with text values, dotted-path get<T>(path[, default]), read_xml for
// plain element-only XML with comments (what the reference's cfg/*.xml files are).
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
namespace boost { namespace property_tree {
class ptree {
 public:
  typedef std::vector<std::pair<std::string, ptree> > children_t;
  typedef children_t::const_iterator const_iterator;
  typedef children_t::iterator iterator;
  typedef std::pair<std::string, ptree> value_type;
  std::string data_;
  children_t kids_;
  const std::string& data() const { return data_; }
  const_iterator begin() const { return kids_.begin(); }
  const_iterator end() const { return kids_.end(); }
  iterator begin() { return kids_.begin(); }
  iterator end() { return kids_.end(); }
  const ptree* find_path(const std::string& path) const {
    const ptree* cur = this; size_t pos = 0;
    while (pos <= path.size()) {
      size_t dot = path.find('.', pos);
      const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
      const ptree* next = 0;
      for (size_t i = 0; i < cur->kids_.size(); i++) if (cur->kids_[i].first == key) { next = &cur->kids_[i].second; break; }
      if (!next) return 0;
      cur = next;
      if (dot == std::string::npos) break;
      pos = dot + 1;
    }
    return cur;
  }
  const ptree& get_child(const std::string& path) const {
    const ptree* p = find_path(path);
    if (!p) throw std::runtime_error("ptree: no such node (" + path + ")");
    return *p;
  }
  template <class T> static bool convert(const std::string& s, T& out) {
    std::istringstream is(s); is >> out; return !is.fail();
  }
  template <class T> T get_value() const { T v = T(); if (!convert(data_, v)) throw std::runtime_error("ptree: conversion failed (" + data_ + ")"); return v; }
  template <class T> T get(const std::string& path) const {
    const ptree* p = find_path(path);
    if (!p) throw std::runtime_error("ptree: no such node (" + path + ")");
    return p->get_value<T>();
  }
  template <class T> T get(const std::string& path, const T& def) const {
    const ptree* p = find_path(path);
    T v = T();
    if (!p || !convert(p->data_, v)) return def;
    return v;
  }
  std::string get(const std::string& path, const char* def) const { return get<std::string>(path, std::string(def)); }
};
template <> inline bool ptree::convert<std::string>(const std::string& s, std::string& out) { out = s; return true; }
namespace xml_parser {
inline std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline void parse_nodes(const std::string& x, size_t& pos, ptree& node, const std::string& closing) {
  std::string text;
  while (pos < x.size()) {
    if (x[pos] != '<') { text += x[pos++]; continue; }
    if (x.compare(pos, 4, "<!--") == 0) { size_t e = x.find("-->", pos); if (e == std::string::npos) throw std::runtime_error("xml: unterminated comment"); pos = e + 3; continue; }
    if (x.compare(pos, 2, "<?") == 0) { size_t e = x.find("?>", pos); if (e == std::string::npos) throw std::runtime_error("xml: unterminated declaration"); pos = e + 2; continue; }
    if (x.compare(pos, 2, "</") == 0) {
      size_t e = x.find('>', pos);
      const std::string name = trim(x.substr(pos + 2, e - pos - 2));
      if (name != closing) throw std::runtime_error("xml: mismatched </" + name + ">");
      pos = e + 1; node.data_ = trim(text); return;
    }
    size_t e = x.find('>', pos);
    if (e == std::string::npos) throw std::runtime_error("xml: unterminated tag");
    std::string tag = x.substr(pos + 1, e - pos - 1);
    const bool selfclose = !tag.empty() && tag[tag.size() - 1] == '/';
    if (selfclose) tag.erase(tag.size() - 1);
    size_t sp = tag.find_first_of(" \t\r\n");
    const std::string name = trim(sp == std::string::npos ? tag : tag.substr(0, sp));
    pos = e + 1;
    node.kids_.push_back(std::make_pair(name, ptree()));
    if (!selfclose) parse_nodes(x, pos, node.kids_.back().second, name);
  }
  if (!closing.empty()) throw std::runtime_error("xml: missing </" + closing + ">");
  node.data_ = trim(text);
}
inline void read_xml(const std::string& file, ptree& pt, int = 0) {
  std::ifstream in(file.c_str());
  if (!in) throw std::runtime_error("read_xml: cannot open " + file);
  std::stringstream ss; ss << in.rdbuf();
  const std::string x = ss.str();
  size_t pos = 0;
  pt = ptree();
  parse_nodes(x, pos, pt, "");
}
}  // namespace xml_parser
using xml_parser::read_xml;
}}  // namespace boost::property_tree

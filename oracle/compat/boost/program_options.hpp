// Minimal stand-in for <boost/program_options.hpp> — TEST INFRASTRUCTURE (oracle/compat): long and
// short options with one value ("--cfg f", "--cfg=f", "-c f"), flags, default values, variables_map.
#pragma once
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost { namespace program_options {
struct value_base {
  virtual ~value_base() {}
  virtual void parse(const std::string& s) = 0;
  virtual bool has_default() const = 0;
  virtual void apply_default() = 0;
  virtual std::string text() const = 0;
};
template <class T>
struct typed_value : value_base {
  T* dst; T cur; bool have_def; T def;
  explicit typed_value(T* d) : dst(d), cur(), have_def(false), def() {}
  typed_value* default_value(const T& v) { have_def = true; def = v; return this; }
  void parse(const std::string& s) { std::istringstream is(s); is >> cur; if (is.fail()) throw std::runtime_error("bad option value: " + s); if (dst) *dst = cur; }
  bool has_default() const { return have_def; }
  void apply_default() { cur = def; if (dst) *dst = def; }
  std::string text() const { std::ostringstream os; os << cur; return os.str(); }
};
template <> inline void typed_value<std::string>::parse(const std::string& s) { cur = s; if (dst) *dst = s; }
template <class T> typed_value<T>* value(T* d = 0) { return new typed_value<T>(d); }
struct option_def { std::string lng; char sht; value_base* val; std::string help; };
class options_description;
class options_adder {
 public:
  explicit options_adder(options_description* o) : o_(o) {}
  options_adder& operator()(const char* name, const char* help);
  options_adder& operator()(const char* name, value_base* v, const char* help);
 private:
  options_description* o_;
};
class options_description {
 public:
  explicit options_description(const std::string& caption = "") : caption_(caption) {}
  options_adder add_options() { return options_adder(this); }
  std::vector<option_def> opts;
  std::string caption_;
};
inline void split_name(const char* name, std::string& lng, char& sht) {
  std::string n(name); sht = 0;
  size_t c = n.find(',');
  if (c != std::string::npos) { lng = n.substr(0, c); if (c + 1 < n.size()) sht = n[c + 1]; } else lng = n;
}
inline options_adder& options_adder::operator()(const char* name, const char* help) {
  option_def d; split_name(name, d.lng, d.sht); d.val = 0; d.help = help; o_->opts.push_back(d); return *this;
}
inline options_adder& options_adder::operator()(const char* name, value_base* v, const char* help) {
  option_def d; split_name(name, d.lng, d.sht); d.val = v; d.help = help; o_->opts.push_back(d); return *this;
}
inline std::ostream& operator<<(std::ostream& os, const options_description& d) {
  os << d.caption_ << ":\n";
  for (size_t i = 0; i < d.opts.size(); i++) {
    os << "  ";
    if (d.opts[i].sht) os << "-" << d.opts[i].sht << " [ --" << d.opts[i].lng << " ]"; else os << "--" << d.opts[i].lng;
    if (d.opts[i].val) os << " arg";
    os << "  " << d.opts[i].help << "\n";
  }
  return os;
}
class variable_value {
 public:
  variable_value() : v_(0) {}
  explicit variable_value(value_base* v) : v_(v) {}
  template <class T> const T& as() const {
    typed_value<T>* t = dynamic_cast<typed_value<T>*>(v_);
    if (!t) throw std::runtime_error("variables_map: bad type");
    return t->cur;
  }
 private:
  value_base* v_;
};
class variables_map : public std::map<std::string, variable_value> {
 public:
  size_t count(const std::string& k) const { return std::map<std::string, variable_value>::count(k); }
  const variable_value& operator[](const std::string& k) const { return find(k)->second; }
  variable_value& ref(const std::string& k) { return std::map<std::string, variable_value>::operator[](k); }
};
struct parsed_options { std::vector<std::pair<const option_def*, std::string> > items; const options_description* desc; };
inline parsed_options parse_command_line(int argc, char** argv, const options_description& d) {
  parsed_options p; p.desc = &d;
  for (int i = 1; i < argc; i++) {
    std::string a(argv[i]), val; const option_def* o = 0; bool have_val = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      std::string n = a.substr(2); size_t e = n.find('=');
      if (e != std::string::npos) { val = n.substr(e + 1); n = n.substr(0, e); have_val = true; }
      for (size_t k = 0; k < d.opts.size(); k++) if (d.opts[k].lng == n) o = &d.opts[k];
    } else if (a.size() >= 2 && a[0] == '-') {
      for (size_t k = 0; k < d.opts.size(); k++) if (d.opts[k].sht == a[1]) o = &d.opts[k];
      if (a.size() > 2) { val = a.substr(2); have_val = true; }
    }
    if (!o) throw std::runtime_error("unrecognised option '" + a + "'");
    if (o->val && !have_val) { if (i + 1 >= argc) throw std::runtime_error("option '" + a + "' needs a value"); val = argv[++i]; }
    p.items.push_back(std::make_pair(o, val));
  }
  return p;
}
inline void store(const parsed_options& p, variables_map& vm) {
  for (size_t i = 0; i < p.items.size(); i++) {
    const option_def* o = p.items[i].first;
    if (o->val) o->val->parse(p.items[i].second);
    vm.ref(o->lng) = variable_value(o->val);
  }
  for (size_t k = 0; k < p.desc->opts.size(); k++) {
    const option_def& o = p.desc->opts[k];
    if (o.val && o.val->has_default() && !vm.count(o.lng)) { o.val->apply_default(); vm.ref(o.lng) = variable_value(o.val); }
  }
}
inline void notify(variables_map&) {}
}}  // namespace boost::program_options

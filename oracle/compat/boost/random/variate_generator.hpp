// compat shim (TEST INFRASTRUCTURE): boost::variate_generator
#ifndef RFS_COMPAT_BOOST_VARGEN
#define RFS_COMPAT_BOOST_VARGEN
namespace boost {
template <class Engine, class Dist>
class variate_generator {
 public:
  typedef typename Dist::result_type result_type;
  variate_generator(Engine e, Dist d) : e_(e), d_(d) {}
  result_type operator()() { return d_(e_); }
  Engine& engine() { return e_; }
  Dist& distribution() { return d_; }
 private:
  Engine e_;
  Dist d_;
};
}
#endif

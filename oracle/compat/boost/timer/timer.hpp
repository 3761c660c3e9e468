// compat shim (TEST INFRASTRUCTURE): boost::timer::cpu_timer (wall via steady clock, user/system
// via times(), like Boost.Timer on POSIX), nanosecond units.
#ifndef RFS_COMPAT_BOOST_TIMER
#define RFS_COMPAT_BOOST_TIMER
#include <sys/times.h>
#include <unistd.h>
#include <chrono>
#include <cstdint>
#include <string>
namespace boost { namespace timer {
typedef std::int_least64_t nanosecond_type;
struct cpu_times {
  nanosecond_type wall, user, system;
  void clear() { wall = user = system = 0; }
};
class cpu_timer {
 public:
  cpu_timer() { start(); }
  bool is_stopped() const { return stopped_; }
  cpu_times elapsed() const {
    if (stopped_) return t_;
    cpu_times c = now();
    c.wall -= t_.wall; c.user -= t_.user; c.system -= t_.system;
    return c;
  }
  void start() { stopped_ = false; t_ = now(); }
  void stop() {
    if (stopped_) return;
    stopped_ = true;
    cpu_times c = now();
    t_.wall = c.wall - t_.wall; t_.user = c.user - t_.user; t_.system = c.system - t_.system;
  }
  void resume() {
    if (!stopped_) return;
    cpu_times c = t_;
    start();
    t_.wall -= c.wall; t_.user -= c.user; t_.system -= c.system;
  }
 private:
  static cpu_times now() {
    cpu_times c;
    c.wall = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    tms tm;
    ::times(&tm);
    static const long tick = ::sysconf(_SC_CLK_TCK);
    const nanosecond_type f = 1000000000LL / (tick > 0 ? tick : 100);
    c.user = (nanosecond_type)(tm.tms_utime + tm.tms_cutime) * f;
    c.system = (nanosecond_type)(tm.tms_stime + tm.tms_cstime) * f;
    return c;
  }
  cpu_times t_;
  bool stopped_;
};
}}
#endif

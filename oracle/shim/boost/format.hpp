// Minimal stand-in for <boost/format.hpp>: a single printf-style spec applied to one
// argument with operator%, streamable (format("%.2f") % kurtosis, misc/format.cpp:1152).
#ifndef UPB_SHIM_BOOST_FORMAT_HPP
#define UPB_SHIM_BOOST_FORMAT_HPP
#include <string>
#include <ostream>
#include <cstdio>
namespace boost {
  class format {
    public:
      explicit format (const std::string& spec) : spec_(spec) {}
      format& operator% (double v) {
        char buf[512];
        std::snprintf(buf, sizeof buf, spec_.c_str(), v);
        out_ = buf;
        return *this;
      }
      std::string str () const { return out_; }
    private:
      std::string spec_;
      std::string out_;
  };
  inline std::ostream& operator<< (std::ostream& o, const format& f) { return o << f.str(); }
}
#endif

// C++98 stand-in for boost::iostreams filtering streams: the hot-path translation units hold
// only pointers to these (misc/filterstream.hpp:28,43); the real stand-in lives in oracle/shim.
#ifndef UPB_SHIM98_BOOST_IOSTREAMS_FILTERING_STREAM_HPP
#define UPB_SHIM98_BOOST_IOSTREAMS_FILTERING_STREAM_HPP
#include <istream>
#include <ostream>
namespace boost { namespace iostreams {
  class filtering_istream : public std::istream { public: filtering_istream () : std::istream(0) {} void reset () {} };
  class filtering_ostream : public std::ostream { public: filtering_ostream () : std::ostream(0) {} void reset () {} };
}}
#endif

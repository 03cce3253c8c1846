// Minimal stand-in for <boost/lexical_cast.hpp>.
// Behaviour reproduced (documented Boost behaviour the reference relies on):
//  * whole-string conversions only: leading/trailing junk or whitespace throws bad_lexical_cast;
//  * lexical_cast<std::string>(double) prints with 17 significant digits ("%.17g"), which is
//    what OutStream << double ends up using (misc/filterstream.cpp:134-137);
//  * integer -> string prints plain decimal.
#ifndef UPB_SHIM_BOOST_LEXICAL_CAST_HPP
#define UPB_SHIM_BOOST_LEXICAL_CAST_HPP
#include <string>
#include <sstream>
#include <typeinfo>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cctype>
#include <limits>
#include <regex>
#include <type_traits>

namespace boost {

  class bad_lexical_cast : public std::bad_cast {
    public:
      const char* what () const noexcept override { return "bad lexical cast"; }
  };

  namespace shim_detail {

    inline std::string to_str (const std::string& s) { return s; }
    inline std::string to_str (const char* s) { return std::string(s ? s : ""); }
    template <class It> inline std::string to_str (const std::sub_match<It>& m) { return m.str(); }

    template <class T, class Enable = void> struct from_string;

    template <class T>
    struct from_string<T, typename std::enable_if<std::is_integral<T>::value && std::is_unsigned<T>::value>::type> {
      static T apply (const std::string& s) {
        if (s.empty()) throw bad_lexical_cast();
        size_t i = 0;
        bool neg = false;
        if (s[0] == '+' || s[0] == '-') { neg = (s[0] == '-'); i = 1; }
        if (i >= s.size()) throw bad_lexical_cast();
        unsigned long long v = 0;
        for (; i < s.size(); ++i) {
          if (s[i] < '0' || s[i] > '9') throw bad_lexical_cast();
          const unsigned d = (unsigned)(s[i] - '0');
          if (v > (std::numeric_limits<unsigned long long>::max() - d) / 10ULL) throw bad_lexical_cast();
          v = v * 10ULL + d;
        }
        if (v > (unsigned long long)std::numeric_limits<T>::max()) throw bad_lexical_cast();
        T r = (T)v;
        if (neg) r = (T)(0 - r); // Boost negates in the unsigned domain
        return r;
      }
    };

    template <class T>
    struct from_string<T, typename std::enable_if<std::is_integral<T>::value && std::is_signed<T>::value>::type> {
      static T apply (const std::string& s) {
        if (s.empty()) throw bad_lexical_cast();
        size_t i = 0;
        bool neg = false;
        if (s[0] == '+' || s[0] == '-') { neg = (s[0] == '-'); i = 1; }
        if (i >= s.size()) throw bad_lexical_cast();
        unsigned long long v = 0;
        for (; i < s.size(); ++i) {
          if (s[i] < '0' || s[i] > '9') throw bad_lexical_cast();
          v = v * 10ULL + (unsigned)(s[i] - '0');
          if (v > (unsigned long long)std::numeric_limits<T>::max() + 1ULL) throw bad_lexical_cast();
        }
        if (!neg && v > (unsigned long long)std::numeric_limits<T>::max()) throw bad_lexical_cast();
        return neg ? (T)(-(long long)v) : (T)v;
      }
    };

    template <class T>
    struct from_string<T, typename std::enable_if<std::is_floating_point<T>::value>::type> {
      static T apply (const std::string& s) {
        if (s.empty() || std::isspace((unsigned char)s[0])) throw bad_lexical_cast();
        errno = 0;
        char* end = 0;
        const double v = std::strtod(s.c_str(), &end);
        if (end != s.c_str() + s.size()) throw bad_lexical_cast();
        return (T)v;
      }
    };

    template <>
    struct from_string<std::string, void> {
      static std::string apply (const std::string& s) { return s; }
    };

    template <class Target, class Source, class Enable = void> struct caster {
      static Target apply (const Source& src) { return from_string<Target>::apply(to_str(src)); }
    };

    // arithmetic -> string
    template <class Source>
    struct caster<std::string, Source, typename std::enable_if<std::is_floating_point<Source>::value>::type> {
      static std::string apply (const Source& src) {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%.17g", (double)src);
        return std::string(buf);
      }
    };
    template <class Source>
    struct caster<std::string, Source, typename std::enable_if<std::is_integral<Source>::value>::type> {
      static std::string apply (const Source& src) {
        std::ostringstream o;
        o << +src;
        return o.str();
      }
    };
  }

  template <class Target, class Source>
  inline Target lexical_cast (const Source& src) {
    return shim_detail::caster<Target, Source>::apply(src);
  }
  template <class Target>
  inline Target lexical_cast (const char* src) {
    return shim_detail::from_string<Target>::apply(shim_detail::to_str(src));
  }
}
#endif

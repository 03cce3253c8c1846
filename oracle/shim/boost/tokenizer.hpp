// Minimal stand-in for <boost/tokenizer.hpp>: tokenizer<char_separator<char> > over a
// std::string, dropping empty tokens (Boost's default for char_separator), as used for the
// comma-separated -s/-e/-z lists (src/regions.cpp:112-180).
#ifndef UPB_SHIM_BOOST_TOKENIZER_HPP
#define UPB_SHIM_BOOST_TOKENIZER_HPP
#include <string>
#include <vector>
namespace boost {
  template <class Ch> class char_separator {
    public:
      explicit char_separator (const Ch* dropped = " ") : dropped_(dropped ? dropped : "") {}
      bool is_sep (Ch c) const { return dropped_.find(c) != std::basic_string<Ch>::npos; }
    private:
      std::basic_string<Ch> dropped_;
  };
  template <class Sep> class tokenizer {
    public:
      typedef std::vector<std::string>::const_iterator const_iterator;
      typedef const_iterator iterator;
      tokenizer (const std::string& s, const Sep& sep) {
        std::string cur;
        for (size_t i = 0; i < s.size(); ++i) {
          if (sep.is_sep(s[i])) { if (!cur.empty()) { toks_.push_back(cur); cur.clear(); } }
          else cur += s[i];
        }
        if (!cur.empty()) toks_.push_back(cur);
      }
      const_iterator begin () const { return toks_.begin(); }
      const_iterator end () const { return toks_.end(); }
    private:
      std::vector<std::string> toks_;
  };
}
#endif

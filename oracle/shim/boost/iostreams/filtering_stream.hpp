// Minimal stand-in for boost::iostreams filtering streams as used by misc/filterstream.cpp:
// a std::istream/std::ostream that is given (optionally) a gzip/bzip2 marker and then a file
// source/sink, or an existing std stream (stdin/stdout). gzip goes through zlib's gz* API;
// bzip2 is not available in this image (no bzlib.h) and aborts with a clear message.
#ifndef UPB_SHIM_BOOST_IOSTREAMS_FILTERING_STREAM_HPP
#define UPB_SHIM_BOOST_IOSTREAMS_FILTERING_STREAM_HPP
#include <istream>
#include <ostream>
#include <streambuf>
#include <fstream>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <zlib.h>

namespace boost { namespace iostreams {

  struct gzip_decompressor {};
  struct gzip_compressor {};
  struct bzip2_decompressor {};
  struct bzip2_compressor {};

  class file_source {
    public:
      explicit file_source (const char* path) : path_(path) { std::ifstream f(path); open_ = f.good(); }
      bool is_open () const { return open_; }
      const std::string& path () const { return path_; }
    private:
      std::string path_;
      bool open_;
  };

  class file_sink {
    public:
      explicit file_sink (const char* path) : path_(path) { std::ofstream f(path, std::ios::out | std::ios::trunc); open_ = f.good(); }
      bool is_open () const { return open_; }
      const std::string& path () const { return path_; }
    private:
      std::string path_;
      bool open_;
  };

  namespace shim_detail {
    class gz_streambuf : public std::streambuf {
      public:
        gz_streambuf () : file_(0), writing_(false) {}
        ~gz_streambuf () { close(); }
        bool open (const std::string& path, bool writing) {
          close();
          writing_ = writing;
          file_ = gzopen(path.c_str(), writing ? "wb" : "rb");
          if (writing_) setp(obuf_, obuf_ + sizeof obuf_);
          return file_ != 0;
        }
        void close () {
          if (file_) {
            if (writing_) sync();
            gzclose(file_);
            file_ = 0;
          }
        }
      protected:
        int_type underflow () override {
          if (!file_ || writing_) return traits_type::eof();
          const int n = gzread(file_, ibuf_, sizeof ibuf_);
          if (n <= 0) return traits_type::eof();
          setg(ibuf_, ibuf_, ibuf_ + n);
          return traits_type::to_int_type(ibuf_[0]);
        }
        int_type overflow (int_type c) override {
          if (!file_ || !writing_) return traits_type::eof();
          if (sync() != 0) return traits_type::eof();
          if (!traits_type::eq_int_type(c, traits_type::eof())) { *pptr() = traits_type::to_char_type(c); pbump(1); }
          return traits_type::not_eof(c);
        }
        int sync () override {
          if (file_ && writing_ && pptr() > pbase()) {
            const int n = (int)(pptr() - pbase());
            if (gzwrite(file_, pbase(), n) != n) return -1;
            setp(obuf_, obuf_ + sizeof obuf_);
          }
          return 0;
        }
      private:
        gzFile file_;
        bool writing_;
        char ibuf_[1 << 16];
        char obuf_[1 << 16];
    };
    inline void no_bzip2 () {
      std::fprintf(stderr, "error: bzip2 streams are not supported by the oracle shim (no bzlib.h in this image)\n\n");
      std::exit(1);
    }
  }

  class filtering_istream : public std::istream {
    public:
      filtering_istream () : std::istream(0), gz_(false) {}
      void push (const gzip_decompressor&) { gz_ = true; }
      void push (const bzip2_decompressor&) { shim_detail::no_bzip2(); }
      void push (std::istream& s) { rdbuf(s.rdbuf()); clear(); }
      void push (const file_source& src) {
        if (gz_) { if (gzbuf_.open(src.path(), false)) { rdbuf(&gzbuf_); clear(); } else setstate(std::ios::failbit); }
        else { if (fbuf_.open(src.path().c_str(), std::ios::in)) { rdbuf(&fbuf_); clear(); } else setstate(std::ios::failbit); }
      }
      void reset () { rdbuf(0); gzbuf_.close(); if (fbuf_.is_open()) fbuf_.close(); gz_ = false; }
    private:
      bool gz_;
      std::filebuf fbuf_;
      shim_detail::gz_streambuf gzbuf_;
  };

  class filtering_ostream : public std::ostream {
    public:
      filtering_ostream () : std::ostream(0), gz_(false), borrowed_(false) {}
      ~filtering_ostream () { reset(); }
      void push (const gzip_compressor&) { gz_ = true; }
      void push (const bzip2_compressor&) { shim_detail::no_bzip2(); }
      void push (std::ostream& s) { rdbuf(s.rdbuf()); borrowed_ = true; clear(); }
      void push (const file_sink& dst) {
        if (gz_) { if (gzbuf_.open(dst.path(), true)) { rdbuf(&gzbuf_); clear(); } else setstate(std::ios::failbit); }
        else { if (fbuf_.open(dst.path().c_str(), std::ios::out | std::ios::trunc)) { rdbuf(&fbuf_); clear(); } else setstate(std::ios::failbit); }
      }
      void reset () {
        if (rdbuf()) flush();
        rdbuf(0);
        gzbuf_.close();
        if (fbuf_.is_open()) fbuf_.close();
        gz_ = false;
        borrowed_ = false;
      }
    private:
      bool gz_;
      bool borrowed_;
      std::filebuf fbuf_;
      shim_detail::gz_streambuf gzbuf_;
  };

}}
#endif

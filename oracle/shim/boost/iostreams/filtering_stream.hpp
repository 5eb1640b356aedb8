// Test scaffolding (see oracle/shim/Eigen/Core): boost::iostreams::filtering_istream / filtering_ostream with an
// optional gzip filter in front of a std::ifstream / std::ofstream, which is all the reference's FileUtils does with
// them (ref: ASMC_SRC/SRC/FileUtils.cpp:143-216).  Implemented over zlib.
#pragma once
#include <zlib.h>

#include <cstring>
#include <istream>
#include <memory>
#include <ostream>
#include <streambuf>
#include <vector>

#include "filter/gzip.hpp"

namespace boost
{
namespace iostreams
{
namespace shim
{
class InBuf : public std::streambuf
{
  std::istream* src = nullptr;
  bool gz = false, zInit = false, eof = false;
  z_stream z{};
  std::vector<char> in, out;

public:
  InBuf() : in(1 << 16), out(1 << 18) {}
  ~InBuf() override { close(); }
  void setGz() { gz = true; }
  void setSource(std::istream* s)
  {
    src = s;
    setg(out.data(), out.data(), out.data());
  }
  void close()
  {
    if (zInit) {
      inflateEnd(&z);
      zInit = false;
    }
    src = nullptr;
    gz = false;
    eof = false;
    setg(nullptr, nullptr, nullptr);
  }

protected:
  int_type underflow() override
  {
    if (!src || eof) {
      return traits_type::eof();
    }
    if (!gz) {
      src->read(out.data(), static_cast<std::streamsize>(out.size()));
      const std::streamsize n = src->gcount();
      if (n <= 0) {
        eof = true;
        return traits_type::eof();
      }
      setg(out.data(), out.data(), out.data() + n);
      return traits_type::to_int_type(*gptr());
    }
    if (!zInit) {
      std::memset(&z, 0, sizeof z);
      inflateInit2(&z, 16 + MAX_WBITS);
      zInit = true;
    }
    z.next_out = reinterpret_cast<Bytef*>(out.data());
    z.avail_out = static_cast<uInt>(out.size());
    while (z.avail_out == out.size()) {
      if (z.avail_in == 0) {
        src->read(in.data(), static_cast<std::streamsize>(in.size()));
        const std::streamsize n = src->gcount();
        if (n <= 0) {
          eof = true;
          break;
        }
        z.next_in = reinterpret_cast<Bytef*>(in.data());
        z.avail_in = static_cast<uInt>(n);
      }
      const int rc = inflate(&z, Z_NO_FLUSH);
      if (rc == Z_STREAM_END) {
        inflateReset(&z);  // concatenated gzip members
      } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
        eof = true;
        break;
      }
    }
    const std::size_t got = out.size() - z.avail_out;
    if (!got) {
      return traits_type::eof();
    }
    setg(out.data(), out.data(), out.data() + got);
    return traits_type::to_int_type(*gptr());
  }
};

class OutBuf : public std::streambuf
{
  std::ostream* dst = nullptr;
  bool gz = false, zInit = false;
  z_stream z{};
  std::vector<char> in, out;

  void pump(const int flush)
  {
    if (!dst) {
      return;
    }
    const std::size_t n = static_cast<std::size_t>(pptr() - pbase());
    if (!gz) {
      dst->write(pbase(), static_cast<std::streamsize>(n));
    } else {
      if (!zInit) {
        std::memset(&z, 0, sizeof z);
        deflateInit2(&z, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 16 + MAX_WBITS, 8, Z_DEFAULT_STRATEGY);
        zInit = true;
      }
      z.next_in = reinterpret_cast<Bytef*>(pbase());
      z.avail_in = static_cast<uInt>(n);
      do {
        z.next_out = reinterpret_cast<Bytef*>(out.data());
        z.avail_out = static_cast<uInt>(out.size());
        deflate(&z, flush);
        dst->write(out.data(), static_cast<std::streamsize>(out.size() - z.avail_out));
      } while (z.avail_out == 0);
    }
    setp(in.data(), in.data() + in.size());
  }

public:
  OutBuf() : in(1 << 16), out(1 << 16) { setp(in.data(), in.data() + in.size()); }
  ~OutBuf() override { close(); }
  void setGz() { gz = true; }
  void setSink(std::ostream* s) { dst = s; }
  void close()
  {
    if (dst) {
      pump(Z_FINISH);
      dst->flush();
    }
    if (zInit) {
      deflateEnd(&z);
      zInit = false;
    }
    dst = nullptr;
    gz = false;
  }

protected:
  int_type overflow(int_type ch) override
  {
    pump(Z_NO_FLUSH);
    if (!traits_type::eq_int_type(ch, traits_type::eof())) {
      *pptr() = traits_type::to_char_type(ch);
      pbump(1);
    }
    return traits_type::not_eof(ch);
  }
  int sync() override
  {
    pump(Z_NO_FLUSH);
    return 0;
  }
};
}  // namespace shim

class filtering_istream : public std::istream
{
  shim::InBuf buf;

public:
  filtering_istream() : std::istream(nullptr) { rdbuf(&buf); }
  void push(const gzip_decompressor&) { buf.setGz(); }
  void push(std::istream& s)
  {
    buf.setSource(&s);
    clear();
  }
  void reset()
  {
    buf.close();
    clear();
  }
};

class filtering_ostream : public std::ostream
{
  shim::OutBuf buf;

public:
  filtering_ostream() : std::ostream(nullptr) { rdbuf(&buf); }
  ~filtering_ostream() override { buf.close(); }
  void push(const gzip_compressor&) { buf.setGz(); }
  void push(std::ostream& s)
  {
    buf.setSink(&s);
    clear();
  }
  void reset()
  {
    buf.close();
    clear();
  }
};
}  // namespace iostreams
}  // namespace boost

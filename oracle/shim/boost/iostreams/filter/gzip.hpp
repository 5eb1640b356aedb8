// Test scaffolding (see oracle/shim/Eigen/Core): tag types; the work is in filtering_stream.hpp.
#pragma once
namespace boost
{
namespace iostreams
{
struct gzip_decompressor {
};
struct gzip_compressor {
};
}  // namespace iostreams
}  // namespace boost

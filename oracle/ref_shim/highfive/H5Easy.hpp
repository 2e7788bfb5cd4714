// TEST INFRASTRUCTURE: see H5File.hpp
#ifndef MDG_H5EASY_STUB
#define MDG_H5EASY_STUB
#include "H5File.hpp"
namespace H5Easy {
using File = HighFive::File;
inline std::size_t getSize(const File&, const std::string&) { throw HighFive::NoHdf(); }
inline std::vector<std::size_t> getShape(const File&, const std::string&) { throw HighFive::NoHdf(); }
}  // namespace H5Easy
#endif

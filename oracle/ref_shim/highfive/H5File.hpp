// TEST INFRASTRUCTURE: throwing stand-ins for the HighFive/HDF5 API that HDFSourceSingle (out of scope:
// file market data) uses.  They exist only so that the reference's DataSource.cpp
// compiles unmodified; every entry point throws.
#ifndef MDG_HIGHFIVE_STUB
#define MDG_HIGHFIVE_STUB
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>
namespace HighFive {
struct NoHdf : std::runtime_error { NoHdf() : std::runtime_error("HDF5 is not available in the oracle build") {} };
struct DataSpace { std::vector<std::size_t> getDimensions() const { throw NoHdf(); } };
struct Selection { template <class T> void read(T) const { throw NoHdf(); } };
struct Attribute {
  DataSpace getSpace() const { throw NoHdf(); }
  template <class T> void read(T&) const { throw NoHdf(); }
};
struct DataSet {
  DataSpace getSpace() const { throw NoHdf(); }
  Selection select(std::vector<std::size_t>, std::vector<std::size_t>) const { throw NoHdf(); }
  template <class T> void read(T) const { throw NoHdf(); }
};
struct Group {
  std::vector<std::string> listAttributeNames() const { throw NoHdf(); }
  std::vector<std::string> listObjectNames() const { throw NoHdf(); }
  Attribute getAttribute(const std::string&) const { throw NoHdf(); }
  DataSet getDataSet(const std::string&) const { throw NoHdf(); }
};
struct File {
  enum AccessMode { ReadOnly = 0 };
  File(const std::string&, int = 0) { throw NoHdf(); }
  Group getGroup(const std::string&) const { throw NoHdf(); }
  DataSet getDataSet(const std::string&) const { throw NoHdf(); }
};
}  // namespace HighFive
#endif

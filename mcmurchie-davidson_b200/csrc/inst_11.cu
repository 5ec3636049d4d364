// explicit instantiations of the class-specialised shell-quartet kernels (split for parallel compilation)
#include "kernels_a.cuh"
namespace mmdb {
MMDB_INSTANTIATE_CLASS(2, 2, 2, 2)
}

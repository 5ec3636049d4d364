// explicit instantiations of the class kernels with an S2 pseudo-shell (shell type code 3, core.cuh)
#include "kernels_a.cuh"
namespace mmdb {
MMDB_INSTANTIATE_CLASS(3, 3, 3, 0)
MMDB_INSTANTIATE_CLASS(3, 3, 3, 3)
}

// (ps|dp): the narrow pair as the bra (see SWAPPED in lib.cu)
#include "kernels_a.cuh"
namespace mmdb {
MMDB_INSTANTIATE_CLASS(1, 0, 2, 1)
}

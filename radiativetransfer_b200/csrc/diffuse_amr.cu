// placeholder: AMR diffuse sweep (implemented next)
#include "rtb200_internal.h"
namespace rtb {
int diffuse_amr(Context&, int, const double*, const std::vector<Direction>&, double*, cudaStream_t, int64_t*) { return RTB200_ERR_ARG; }
int amr_neighbours(Context&, const Direction&, int32_t*) { return RTB200_ERR_ARG; }
}

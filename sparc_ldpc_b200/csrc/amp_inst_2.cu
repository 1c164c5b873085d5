// explicit instantiations of the AMP / operator kernels for M = 2^{9}
#include "amp_impl.cuh"

namespace sb {
template int launch_amp<9>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
}  // namespace sb

// explicit instantiations of the AMP / operator kernels for M = 2^{10}
#include "amp_impl.cuh"

namespace sb {
template int launch_amp<10>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
}  // namespace sb

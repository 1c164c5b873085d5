// explicit instantiations of the AMP / operator kernels for M = 2^{6,7,8}
#include "amp_impl.cuh"

namespace sb {
template int launch_amp<6>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
template int launch_amp<7>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
template int launch_amp<8>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
}  // namespace sb

// explicit instantiations of the AMP / operator kernels for M = 2^{1,2,3,4,5}
#include "amp_impl.cuh"

namespace sb {
template int launch_amp<1>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
template int launch_amp<2>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
template int launch_amp<3>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
template int launch_amp<4>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
template int launch_amp<5>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
}  // namespace sb

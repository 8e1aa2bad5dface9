// f64 instantiations of the launch orchestration (see run_impl.cuh)
#include "run_impl.cuh"
namespace sgmhost {
template int run_model<double>(const sgm_pf_desc* d, cudaStream_t s);
template int run_sgld_persistent<double>(const sgm_pf_desc* d, const sgm::SgldArgs& sa, int K, cudaStream_t s);
}

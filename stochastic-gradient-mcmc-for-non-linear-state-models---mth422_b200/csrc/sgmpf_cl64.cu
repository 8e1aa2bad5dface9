// f64 instantiations of the thread-block-cluster kernels (see cluster_impl.cuh)
#include "cluster_impl.cuh"
namespace sgmhost {
template bool run_cluster<double>(const sgm_pf_desc* d, const KArgs& a, cudaStream_t s);
template int run_sgld_cluster<double>(const sgm_pf_desc* d, const sgm::SgldArgs& sa, const KArgs& a, int K, cudaStream_t s);
}

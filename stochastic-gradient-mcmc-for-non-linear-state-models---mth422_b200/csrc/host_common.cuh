// Host-side pieces shared by the translation units of libsgmpf.so: the error string, the workspace layout and the
// per-dtype launch entry (run_model<R>), which is explicitly instantiated in sgmpf_f32.cu / sgmpf_f64.cu so the two
// halves of the template instantiations compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sgmpf.h"
#include "pf_kernels.cuh"
#include "backward_kernels.cuh"
#include "small_kernels.cuh"
#include "sgld_kernels.cuh"

namespace sgmhost {
using namespace sgm;

int fail(int code, const char* fmt, const char* extra = "");
void set_launch_count(int64_t n);

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Layout {
    size_t rec[2], tail[2], fine[2], lw[2], sub[2], hdr, acc, thc, yw, Jidx, Llist[2], counters, pcdf, pguide, pkey, n2part, total;
};

inline bool backward_pf(int pf) { return pf == SGM_PF_POY_N2 || pf == SGM_PF_PARIS; }
inline int state_dim(int model) { return model == SGM_MODEL_GARCH ? 2 : 1; }
inline int score_dim(int model) { return model == SGM_MODEL_SVM ? 3 : 4; }
Layout make_layout(const sgm_pf_desc* d);

template <class R> int run_model(const sgm_pf_desc* d, cudaStream_t s);
extern template int run_model<float>(const sgm_pf_desc* d, cudaStream_t s);
extern template int run_model<double>(const sgm_pf_desc* d, cudaStream_t s);
// thread-block-cluster kernels (cluster_impl.cuh, own translation units)
template <class R> bool run_cluster(const sgm_pf_desc* d, const sgm::KArgs& a, cudaStream_t s);
extern template bool run_cluster<float>(const sgm_pf_desc* d, const sgm::KArgs& a, cudaStream_t s);
extern template bool run_cluster<double>(const sgm_pf_desc* d, const sgm::KArgs& a, cudaStream_t s);
template <class R> int run_sgld_cluster(const sgm_pf_desc* d, const sgm::SgldArgs& sa, const sgm::KArgs& a, int K, cudaStream_t s);
extern template int run_sgld_cluster<float>(const sgm_pf_desc* d, const sgm::SgldArgs& sa, const sgm::KArgs& a, int K, cudaStream_t s);
extern template int run_sgld_cluster<double>(const sgm_pf_desc* d, const sgm::SgldArgs& sa, const sgm::KArgs& a, int K, cudaStream_t s);
template <class R> int run_sgld_persistent(const sgm_pf_desc* d, const sgm::SgldArgs& sa, int K, cudaStream_t s);
extern template int run_sgld_persistent<float>(const sgm_pf_desc* d, const sgm::SgldArgs& sa, int K, cudaStream_t s);
extern template int run_sgld_persistent<double>(const sgm_pf_desc* d, const sgm::SgldArgs& sa, int K, cudaStream_t s);

}  // namespace sgmhost

// ibss_kernels.cu -- the instantiations of the persistent IBSS kernel for ONE ancestry count (SB_K = 1..6).
//
// Compiled once per ancestry count (nvcc -DSB_K=k), in parallel, and linked with capi.cu into the one shared
// library: every (storage type, summary / individual level, traversal variant) combination is a kernel of its
// own so that each gets its own register allocation (ibss_kernel.cuh).  The only thing exported to capi.cu is
// the lookup below; it returns the host-side kernel handle used for cudaFuncSetAttribute / occupancy queries /
// cudaLaunchKernelExC.
#include "../../include/sushie_b200.h"
#include "ibss_kernel.cuh"

#ifndef SB_K
#error "compile with -DSB_K=<ancestries>, 1..6"
#endif

#define SB_CAT2(a, b) a##b
#define SB_CAT(a, b) SB_CAT2(a, b)

template <typename XT, bool SS, int V = 0>
static const void* kfn() {
  return reinterpret_cast<const void*>(&sb::ibss_kernel<SB_K, XT, SS, V>);
}

// variant: traversal variant compiled into the kernel (capi.cu: launch_variant) -- hard-call storage 1..5; dense fp64
// individual level 1 / 3 / 5 / 6 for the tiles that live in kernels of their own; everything else 0.
extern "C" __attribute__((visibility("hidden"))) const void* SB_CAT(sb_kernel_table_k, SB_K)(int storage, int ss,
                                                                                            int variant) {
  const bool f32 = storage == SB_F32;
  if (storage == SB_B2) return ss ? nullptr : kfn<uint8_t, false>();
  if (storage == SB_I8) {
    if (ss) return nullptr;
    return variant == 1   ? kfn<int8_t, false, 1>()
           : variant == 2 ? kfn<int8_t, false, 2>()
           : variant == 3 ? kfn<int8_t, false, 3>()
           : variant == 5 ? kfn<int8_t, false, 5>()
                          : kfn<int8_t, false, 4>();
  }
  if (!ss && !f32) {
    if (variant == 1) return kfn<double, false, 1>();
    if (variant == 3) return kfn<double, false, 3>();
    if (variant == 5) return kfn<double, false, 5>();
    if (variant == 6) return kfn<double, false, 6>();
  }
  return ss ? (f32 ? kfn<float, true>() : kfn<double, true>()) : (f32 ? kfn<float, false>() : kfn<double, false>());
}

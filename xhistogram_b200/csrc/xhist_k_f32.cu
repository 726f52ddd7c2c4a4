// float instantiations of the histogram kernels (see xhist_kernels_impl.cuh / xhist_kernels.cu)
#include "xhist_kernels_impl.cuh"
XHK_DEFINE_PICKERS(f32, float, true)
